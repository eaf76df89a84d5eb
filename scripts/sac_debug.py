"""Prints every log / gradient-leaf / update-leaf error of the CUDA update against the fp64 oracle."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import torch

import sac_util as SU
from oracle import mtsac_oracle as O


def run(cfg, per_task, seed=1):
    import dataclasses
    for mode in ("exact", "tf32"):
        _run(dataclasses.replace(cfg, matmul_operands=mode), per_task, seed)


def _run(cfg, per_task, seed=1):
    st = O.init_state(cfg, seed=seed, dtype=torch.float32)
    agent = SU.make_agent(cfg, per_task, seed=seed)
    SU.load_oracle_state(agent, st)
    st64 = st.to(torch.float64)
    batch, ec, ea = O.synthetic_batch(cfg, per_task, seed=100, dtype=torch.float32)
    b64 = tuple(b.double() for b in batch)
    new64, logs64, grads64, aux = O.mtsac_update(st64, b64, ec.double(), ea.double(), cfg, return_grads=True)
    _, logs = agent.update(tuple(b.cuda() for b in batch), eps_c=ec.cuda(), eps_a=ea.cuda(), check=True)
    print(f"--- oracle operands={cfg.matmul_operands} T={cfg.num_tasks} W={cfg.width} B={per_task*cfg.num_tasks} launches={agent.launches_per_update()}")
    for k in O.LOG_KEYS:
        ref, got = float(logs64[k]), float(logs[k])
        print(f"  log {k:34s} gpu={got:+.6e} ref={ref:+.6e} rel={abs(got-ref)/max(abs(ref),1e-30):.2e}")
    for name, g64, tree, ens in (("actor", grads64["actor"], agent.actor.grads, False), ("critic", grads64["critic"], agent.critic.grads, True)):
        for leaf, e in SU.compare_trees(g64, tree, ens).items():
            print(f"  grad  {name}/{leaf:22s} rel={e:.2e}")
    for name, new_t, old_t, tree, ens in (("actor", new64.actor, st64.actor, agent.actor.params, False),
                                          ("critic", new64.critic, st64.critic, agent.critic.params, True),
                                          ("target", new64.critic_target, st64.critic_target, agent.critic.target_params, True)):
        for leaf, e in SU.compare_deltas(old_t, new_t, tree, ens).items():
            print(f"  delta {name}/{leaf:22s} rel={e:.2e}")
    la = agent.alpha.params["params"]["log_alpha"]
    print("  delta log_alpha rel=%.2e" % SU.rel(la - st64.log_alpha.cuda().float(), new64.log_alpha - st64.log_alpha))


run(O.OracleConfig(num_tasks=10, obs_dim=49, action_dim=4, width=256), 128)
run(O.OracleConfig(num_tasks=10, obs_dim=49, action_dim=4, width=400, clip=True), 128)
