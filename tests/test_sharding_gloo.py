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


# ---------------------------------------------------------------------------------------------------------------------
# The peer-memory exchange protocol (csrc/comm.cuh trunk_step_kernel + the per-owner dW reduce-adds), emulated with gloo:
# every trunk element is stepped by exactly one owner (table from the library's host-only mtrl_trunk_segments), the
# global norm is assembled from per-owner partial norms plus every rank's head norm, and the owners' new parameters
# are broadcast.  Must equal clip_by_global_norm + adam on the summed gradient, and leave all replicas identical.
# ---------------------------------------------------------------------------------------------------------------------
def _segments(world, T=5, W=64, depth=3, E=2):
    import ctypes as C

    from mtrl_b200 import _lib as L
    from mtrl_b200.rl.algorithms.mtsac import SacConfigC, SacLayoutC

    t_local = -(-T // world)
    cfg = SacConfigC(num_tasks=T, task_begin=0, num_local_tasks=t_local, obs_dim=11 + T, action_dim=3, width=W, depth=depth,
                     num_critics=E, max_rows=128 * t_local, max_batch=128 * t_local, gamma=0.99, tau=0.005, actor_lr=3e-4,
                     critic_lr=3e-4, alpha_lr=3e-4, adam_b1=0.9, adam_b2=0.999, adam_eps=1e-5, actor_max_grad_norm=1.0,
                     critic_max_grad_norm=1.0, alpha_max_grad_norm=-1.0, log_std_min=-20.0, log_std_max=2.0,
                     target_entropy=-3.0, clip_q=0, use_task_weights=0, noise_seed=1, variant=0)
    lay = SacLayoutC()
    L.check(L.lib().mtrl_sac_query_layout(C.byref(cfg), C.byref(lay)))
    fn = L.lib().mtrl_trunk_segments
    fn.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_longlong), C.c_int, C.POINTER(C.c_int)]
    out = (C.c_longlong * (4 * 256))()
    n = C.c_int()
    L.check(fn(C.byref(lay.critic), world, out, 256, C.byref(n)))
    return [(out[4 * i], out[4 * i + 1], out[4 * i + 2], out[4 * i + 3]) for i in range(n.value)], int(lay.critic.trunk_total)


def _exchange_worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        segs, n = _segments(world)
        gen = torch.Generator().manual_seed(100)           # same stream on every rank: the replicated state
        p = torch.randn(n, generator=gen, dtype=torch.float64)
        m = torch.randn(n, generator=gen, dtype=torch.float64) * 1e-3
        v = torch.rand(n, generator=gen, dtype=torch.float64) * 1e-4
        g_all = [torch.randn(n, generator=torch.Generator().manual_seed(200 + r), dtype=torch.float64) * 3 for r in range(world)]
        head_g2 = [float(r + 1) * 0.37 for r in range(world)]            # every rank's local head-gradient squared norm
        lr, b1, b2, eps, max_norm, t = 3e-4, 0.9, 0.999, 1e-5, 1.0, 7
        # --- reference: one process, summed gradient ---
        gs = sum(g_all)
        gn = torch.sqrt((gs ** 2).sum() + sum(head_g2))
        gc = gs if gn < max_norm else gs / gn * max_norm
        m_ref, v_ref = b1 * m + (1 - b1) * gc, b2 * v + (1 - b2) * gc ** 2
        p_ref = p - lr * (m_ref / (1 - b1 ** t)) / (torch.sqrt(v_ref / (1 - b2 ** t)) + eps)
        # --- protocol on this rank ---
        g = g_all[rank].clone()
        owned = [(b, e) for b, e, o, _ in segs if o == rank]
        # reduce-scatter: GEMM epilogue reduce-adds (pre_reduced) or owner peer loads -- either way the owner ends up with the sum
        for b, e, o, _ in segs:
            piece = g[b:e].clone()
            dist.reduce(piece, dst=o)
            if o == rank:
                g[b:e] = piece
        part = torch.tensor([sum(float((g[b:e] ** 2).sum()) for b, e in owned), head_g2[rank]], dtype=torch.float64)
        parts = [torch.zeros(2, dtype=torch.float64) for _ in range(world)]
        dist.all_gather(parts, part)                                     # the inbox_g2 / inbox_head_g2 exchange
        gnorm = torch.sqrt(sum(x[0] for x in parts) + sum(x[1] for x in parts))
        scale = 1.0 if gnorm < max_norm else max_norm / gnorm
        p_new = p.clone()
        for b, e in owned:
            gg = g[b:e] * scale
            m[b:e] = b1 * m[b:e] + (1 - b1) * gg
            v[b:e] = b2 * v[b:e] + (1 - b2) * gg ** 2
            p_new[b:e] = p[b:e] - lr * (m[b:e] / (1 - b1 ** t)) / (torch.sqrt(v[b:e] / (1 - b2 ** t)) + eps)
        for b, e, o, _ in segs:                                          # all-gather: the owner's stores into every replica
            piece = p_new[b:e].clone()
            dist.broadcast(piece, src=o)
            p_new[b:e] = piece
        assert torch.allclose(gnorm, gn, rtol=1e-12)
        assert torch.allclose(p_new, p_ref, rtol=1e-12, atol=1e-15), "replica differs from the unsharded Adam step"
        for b, e in owned:
            assert torch.allclose(m[b:e], m_ref[b:e], rtol=1e-12) and torch.allclose(v[b:e], v_ref[b:e], rtol=1e-12)
        assert sum(e - b for b, e in owned) > 0
        chk = p_new.clone()
        dist.broadcast(chk, src=0)
        assert torch.equal(chk, p_new), "replicas must be bit-identical"
        ret[rank] = "ok"
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_adam_exchange_protocol(world):
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_exchange_worker, args=(world, port, ret), nprocs=world, join=True)
    assert dict(ret) == {r: "ok" for r in range(world)}
