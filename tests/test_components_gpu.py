"""The component-level C entries SURVEY.md 8(b) lists -- `mtrl_mlp_forward` (MultiHeadNetwork.__call__ alone),
`mtrl_adam_polyak_step`, `mtrl_sac_losses_fwd_bwd` -- each against the oracle."""
import pytest
import torch

import sac_util as SU
from oracle import mtsac_oracle as O

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("precision,tol", [("tf32", 2e-3), ("fp32x3", 2e-6)])
def test_network_forward_matches_multihead_network(cuda, precision, tol):
    """multi_head.py:21-68 (trunk on the full input incl. the one-hot, own-task head) for actor, critics and targets."""
    T, W = 7, 192
    cfg = O.OracleConfig(num_tasks=T, obs_dim=39 + T, action_dim=4, width=W)
    st = O.init_state(cfg, seed=4, dtype=torch.float32)
    for net, scale in ((st.actor, 100.0), (st.critic, 30.0)):
        for k in ("kernel", "bias"):
            net["heads"][k] = net["heads"][k] * scale
    st.critic_target = O.tree_map(lambda x: x * 1.01, st.critic)
    agent = SU.make_agent(cfg, 32, seed=4, precision=precision)
    SU.load_oracle_state(agent, st)
    batch, _, _ = O.synthetic_batch(cfg, 11, seed=3, dtype=torch.float32)      # 77 rows, tasks interleaved
    perm = torch.randperm(77, generator=torch.Generator().manual_seed(0))[:50]   # any subset, any order
    obs, act = batch[0][perm], batch[1][perm]
    s64 = st.to(torch.float64)
    ref_a = O.multihead_forward(s64.actor, obs.double(), T, cfg.depth)
    ref_c = O.critic_forward(s64.critic, obs.double(), act.double(), cfg)
    ref_t = O.critic_forward(s64.critic_target, obs.double(), act.double(), cfg)
    got_a = agent.network_forward("actor", obs.cuda())
    got_c = agent.network_forward("critic", obs.cuda(), act.cuda())
    got_t = agent.network_forward("target", obs.numpy(), act.numpy())
    agent._check_status()
    assert got_a.shape == (50, 8) and got_c.shape == (2, 50, 1)
    assert SU.rel(got_a, ref_a) < tol and SU.rel(got_c, ref_c) < tol and SU.rel(got_t, ref_t) < tol
    # the reference evaluates every head and gathers (multi_head.py:50-66): same numbers
    assert SU.rel(got_a, O.multihead_forward(s64.actor, obs.double(), T, cfg.depth, all_heads=True)) < tol


@pytest.mark.parametrize("with_target,clip", [(True, 1.0), (False, None), (True, 1e9)])
def test_adam_polyak_step_matches_optax_restatement(cuda, with_target, clip):
    from mtrl_b200.ops import adam_polyak_step

    g = torch.Generator().manual_seed(1)
    n = 4 * 12345
    p, gr = torch.randn(n, generator=g), torch.randn(n, generator=g) * 0.01
    m, v = torch.randn(n, generator=g) * 1e-3, torch.rand(n, generator=g) * 1e-5
    tgt = torch.randn(n, generator=g)
    opt = {"m": m.double(), "v": v.double(), "count": 6}
    ref_p, ref_opt = O.adam_step(p.double(), gr.double(), opt, 3e-4, 1e-5, 0.9, 0.999, clip)
    ref_t = 0.005 * ref_p + 0.995 * tgt.double()
    dp, dg, dm, dv, dt = (x.cuda().clone() for x in (p, gr, m, v, tgt))
    step = torch.tensor([6], dtype=torch.int32, device="cuda")
    out = adam_polyak_step(dp, dg, dm, dv, step, max_grad_norm=clip, target=dt if with_target else None)
    assert int(step) == 7
    assert SU.rel(dp, ref_p) < 1e-6 and SU.rel(dm, ref_opt["m"]) < 1e-6 and SU.rel(dv, ref_opt["v"]) < 1e-6
    # the step itself: |p| ~ 1 stored in fp32 carries 6e-8 absolute, i.e. ~2e-4 of a 3e-4 step
    assert SU.rel(dp - p.cuda(), ref_p - p.double()) < 2e-3
    assert abs(float(out["grad_norm"]) - float(gr.double().norm())) < 1e-6 * float(gr.double().norm())
    assert abs(float(out["params_norm"]) - float(ref_p.norm())) < 1e-6 * float(ref_p.norm())
    if with_target:
        assert SU.rel(dt, ref_t) < 1e-6
    else:
        assert torch.equal(dt, tgt.cuda())


def test_sac_losses_match_the_reference_formulas(cuda):
    """mtsac.py:547-566 (critic) and :659-666 (actor) on packed rows: 3 tasks x one 128-row tile, 100 / 128 / 7 real rows."""
    from mtrl_b200.ops import sac_losses

    g = torch.Generator().manual_seed(2)
    T, W, E, rows = 3, 64, 2, 384
    valid = torch.full((rows,), -1, dtype=torch.int32)
    for t, n in enumerate((100, 128, 7)):
        valid[128 * t: 128 * t + n] = 1
    ok = valid >= 0
    task = torch.arange(rows) // 128
    Ht = [torch.randn(rows, W, generator=g) for _ in range(E)]
    Ho = [torch.randn(rows, W, generator=g) for _ in range(E)]
    wt = [torch.randn(T, W, 1, generator=g) * 0.1 for _ in range(E)]
    wo = [torch.randn(T, W, 1, generator=g) * 0.1 for _ in range(E)]
    bt = [torch.randn(T, 1, generator=g) for _ in range(E)]
    bo = [torch.randn(T, 1, generator=g) for _ in range(E)]
    rew, done = torch.rand(rows, generator=g) * 10, (torch.rand(rows, generator=g) < 0.1).float()
    logp_next, logp = torch.randn(rows, generator=g), torch.randn(rows, generator=g)
    alpha, tw = torch.tensor([0.5, 1.0, 2.0]), torch.tensor([1.2, 0.8, 1.0])
    B = int(ok.sum())
    c = lambda xs: [x.cuda() for x in xs]  # noqa: E731
    common = dict(H_online=c(Ho), w_online=c(wo), b_online=c(bo), tile_task=torch.arange(3, dtype=torch.int32).cuda(), row_valid=valid.cuda(),
                  alpha=alpha.cuda(), task_weights=tw.cuda(), global_batch=B)
    q = lambda H, w, b: torch.stack([(H[e].double() * w[e][task, :, 0].double()).sum(1) + b[e][task, 0].double() for e in range(E)])  # noqa: E731
    qt, qo = q(Ht, wt, bt), q(Ho, wo, bo)
    y = rew.double() + (1 - done.double()) * 0.99 * (qt.min(0).values - alpha[task].double() * logp_next.double())
    w_row = tw[task].double()
    out = sac_losses("critic", H_target=c(Ht), w_target=c(wt), b_target=c(bt), rewards=rew.cuda(), dones=done.cuda(),
                     logp_next=logp_next.cuda(), gamma=0.99, **common)
    ref_loss = (w_row * (qo - y) ** 2)[:, ok].sum()
    assert abs(float(out["loss_sum"]) - float(ref_loss)) < 1e-5 * float(ref_loss)
    assert abs(float(out["q_sum"]) - float(qo[:, ok].sum())) < 1e-4 * float(qo[:, ok].abs().sum())
    ref_dq = 2.0 / (E * B) * w_row * (qo - y) * ok
    assert SU.rel(out["dq"], ref_dq) < 1e-5
    out = sac_losses("actor", logp=logp.cuda(), **common)
    ref = (w_row * (alpha[task].double() * logp.double() - qo.min(0).values))[ok].sum()
    assert abs(float(out["loss_sum"]) - float(ref)) < 1e-5 * abs(float(ref))
    ref_dq = torch.where(qo == qo.min(0).values, -w_row / B, torch.zeros_like(qo)) * ok
    assert SU.rel(out["dq"], ref_dq) < 1e-6
