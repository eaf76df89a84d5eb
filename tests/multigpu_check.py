"""Launched by torchrun on N GPUs (see tests/test_multigpu_gpu.py): the task-sharded CUDA update must reproduce
the unsharded fp64 oracle -- logs combined over ranks, replicated trunk identical on every rank, each rank's heads."""
import dataclasses
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch
import torch.distributed as dist

import sac_util as SU
from oracle import mtsac_oracle as O


def main():
    import faulthandler

    faulthandler.dump_traceback_later(int(os.environ.get("MG_HANG_DUMP_S", "150")), exit=True)   # a hang must not eat the GPU budget
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    from mtrl_b200.rl.algorithms.mtsac import task_partition

    T, W, per_task = int(os.environ.get("MG_T", "10")), int(os.environ.get("MG_W", "256")), 32
    cfg = O.OracleConfig(num_tasks=T, obs_dim=39 + T, action_dim=4, width=W)
    st = O.init_state(cfg, seed=1, dtype=torch.float32)
    t0, t1 = task_partition(T, world)[rank]
    exchange = os.environ.get("MG_EXCHANGE", "p2p")
    agent = SU.make_agent(cfg, per_task, seed=1, max_batch=per_task * (t1 - t0), rank=rank, world_size=world,
                          process_group=dist.group.WORLD, exchange=exchange)
    SU.load_oracle_state(agent, st, task_slice=slice(t0, t1))
    st64 = st.to(torch.float64)
    steps = 2
    for step in range(steps):
        batch, ec, ea = O.synthetic_batch(cfg, per_task, seed=50 + step, dtype=torch.float32)
        B = batch[0].shape[0]
        task = batch[0][:, -T:].argmax(1)
        rows = (task >= t0) & (task < t1)
        st64, logs64 = O.mtsac_update(st64, tuple(b.double() for b in batch), ec.double(), ea.double(), cfg)
        _, logs = agent.update(tuple(b[rows].cuda() for b in batch), eps_c=ec[rows].cuda(), eps_a=ea[rows].cuda(),
                               global_batch=B, check=True)
        for k in O.LOG_KEYS:
            ref, got = float(logs64[k]), float(logs[k])
            assert abs(got - ref) <= 1e-3 * (1 + step) * abs(ref) + 1e-5, f"rank {rank} step {step} {k}: {got} vs {ref}"
    sl = slice(t0, t1)
    for name, new_t, tree, ens in (("actor", st64.actor, agent.actor.params, False), ("critic", st64.critic, agent.critic.params, True),
                                   ("target", st64.critic_target, agent.critic.target_params, True)):
        for leaf, e in SU.compare_trees(new_t, tree, ens, task_slice=sl).items():
            if leaf.startswith("layer_") and leaf.endswith("kernel"):
                assert e <= 2e-3, f"rank {rank} {name}/{leaf}: {e}"
    la = agent.alpha.params["params"]["log_alpha"]
    assert SU.rel(la, st64.log_alpha[sl]) <= 1e-3
    # replicated trunk must be bit-identical across ranks (same all-reduced gradients, same Adam)
    trunk = agent._flat["critic_params"][: agent._lay.critic.trunk_total].clone()
    ref = trunk.clone()
    dist.broadcast(ref, src=0)
    assert torch.equal(trunk, ref), f"rank {rank}: critic trunk diverged from rank 0"
    assert agent.exchange_error() == 0, f"rank {rank}: peer exchange timed out (code {agent.exchange_error()})"
    dist.barrier()
    if rank == 0:
        print(f"multigpu_check ok: world={world} T={T} W={W} exchange={exchange} multicast={getattr(agent, 'multicast', False)}", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
