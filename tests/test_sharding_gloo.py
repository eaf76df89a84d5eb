"""CPU, world_size 2, gloo: the task-sharding scheme of SURVEY 8(e) / DESIGN.md section 6 reproduces the
unsharded update.  Each rank owns a contiguous block of tasks (rows, heads, log_alpha), computes local
gradients with losses normalised by the GLOBAL batch, all-reduces the trunk gradients with its head-gradient
squared norm riding along as one extra element, and applies the same clip + Adam.  The arithmetic here is the
oracle's (this is a test of the host-side protocol, the kernels are covered by the gpu tests)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import mtsac_oracle as O


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _slice_heads(tree, sl, ensemble):
    out = {k: dict(v) for k, v in tree.items()}
    out["heads"] = {k: (v[:, sl] if ensemble else v[sl]) for k, v in tree["heads"].items()}
    return out


def _local_update(rank, world, cfg, st, batch, ec, ea, B_global):
    """One rank's share of the critic step, written against the phase protocol."""
    from mtrl_b200.rl.algorithms.mtsac import task_partition

    t0, t1 = task_partition(cfg.num_tasks, world)[rank]
    obs = batch[0]
    task = obs[:, -cfg.num_tasks:].argmax(1)
    rows = (task >= t0) & (task < t1)
    lb = tuple(b[rows] for b in batch)
    # full-width oracle on the local rows only, loss scaled so that the mean divides by the global batch
    scale = lb[0].shape[0] / B_global
    cp = O._with_grad(st.critic)
    with torch.no_grad():
        na, nlp = O.actor_sample_and_log_prob(st.actor, lb[2], ec[rows], cfg)
        qt = O.critic_forward(st.critic_target, lb[2], na, cfg)
        alpha = torch.exp(lb[0][:, -cfg.num_tasks:] @ st.log_alpha.reshape(-1, 1))
        y = lb[4] + (1 - lb[3]) * cfg.gamma * (qt.min(0).values - alpha * nlp.reshape(-1, 1))
    q = O.critic_forward(cp, lb[0], lb[1], cfg)
    loss = ((q - y) ** 2).mean() * scale
    loss.backward()
    g = O._grads_of(cp)
    # trunk gradients: all-reduce; head gradients: only this rank's tasks are non-zero -> keep local slice
    trunk = torch.cat([g[k][leaf].flatten() for k in g if k != "heads" for leaf in ("kernel", "bias")])
    heads_local = _slice_heads(g, slice(t0, t1), True)["heads"]
    head_g2 = sum((x.double() ** 2).sum() for x in heads_local.values())
    payload = torch.cat([trunk, head_g2.reshape(1).to(trunk.dtype)])
    dist.all_reduce(payload)
    trunk_sum, head_g2_all = payload[:-1], payload[-1]
    gnorm = torch.sqrt((trunk_sum.double() ** 2).sum() + head_g2_all.double())
    return trunk_sum, heads_local, gnorm, loss.detach(), (t0, t1)


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        cfg = O.OracleConfig(num_tasks=5, obs_dim=11 + 5, action_dim=3, width=24)
        st = O.init_state(cfg, seed=4, dtype=torch.float64)
        batch, ec, ea = O.synthetic_batch(cfg, per_task=6, seed=2, dtype=torch.float64)
        B = batch[0].shape[0]
        trunk_sum, heads_local, gnorm, loss_part, (t0, t1) = _local_update(rank, world, cfg, st, batch, ec, ea, B)
        # reference: the unsharded oracle
        _, logs, grads, _ = O.mtsac_update(st.clone(), batch, ec, ea, cfg, return_grads=True)
        g = grads["critic"]
        ref_trunk = torch.cat([g[k][leaf].flatten() for k in g if k != "heads" for leaf in ("kernel", "bias")])
        assert torch.allclose(trunk_sum, ref_trunk, rtol=1e-9, atol=1e-12), "all-reduced trunk gradient"
        for leaf in ("kernel", "bias"):
            assert torch.allclose(heads_local[leaf], g["heads"][leaf][:, t0:t1], rtol=1e-9, atol=1e-12), "local head gradient"
            other = torch.ones(cfg.num_tasks, dtype=torch.bool)
            other[t0:t1] = False
        assert torch.allclose(gnorm, logs["metrics/critic_grad_magnitude"].double(), rtol=1e-9), "global norm from one all-reduce"
        tot = loss_part.clone()
        dist.all_reduce(tot)
        assert torch.allclose(tot, logs["losses/qf_loss"], rtol=1e-9), "loss partial sums"
        ret[rank] = "ok"
    finally:
        dist.destroy_process_group()


def test_task_sharded_critic_step_matches_unsharded():
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert dict(ret) == {0: "ok", 1: "ok"}


def test_combine_rank_logs():
    from mtrl_b200.rl.algorithms.mtsac import LOG_KEYS, combine_rank_logs

    a = torch.arange(16, dtype=torch.float32) + 1
    b = torch.arange(16, dtype=torch.float32) * 2 + 1
    out = combine_rank_logs(a, a + b)
    assert set(out) == set(LOG_KEYS)
    assert out["losses/qf_loss"] == (a + b)[1] and out["metrics/critic_grad_magnitude"] == a[2]
    assert torch.isclose(out["metrics/critic_params_norm"], torch.sqrt(a[10] + (a + b)[11]))
    assert torch.isclose(out["metrics/actor_params_norm"], torch.sqrt(a[12] + (a + b)[13]))
