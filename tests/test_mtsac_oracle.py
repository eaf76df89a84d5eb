"""CPU checks of the update oracle: known parameter count, head-selection equivalence, fp32 vs
fp64 agreement, closed-form VJPs of SURVEY Appendix A, optimiser algebra, ordering facts."""
import math

import pytest
import torch

from oracle import mtsac_oracle as O


def small_cfg(**kw):
    base = dict(num_tasks=3, obs_dim=7 + 3, action_dim=2, width=16, depth=3)
    base.update(kw)
    return O.OracleConfig(**base)


def test_param_count_kat():
    """plots/get_data.py:51-53 hard-codes 370_000 for mt10_mtmhsac; the exact architecture count is
    372 880 = 49*400+400 + 2*(400^2+400) + 10*(400*8+8)."""
    cfg = O.OracleConfig(num_tasks=10, obs_dim=49, action_dim=4, width=400)
    st = O.init_state(cfg)
    assert O.num_params(st.actor) == 372_880
    assert O.num_params(st.critic) == 692_820  # SURVEY 8d table, two members


def test_all_heads_equals_selected_head():
    cfg = small_cfg()
    st = O.init_state(cfg, dtype=torch.float64)
    batch, ec, ea = O.synthetic_batch(cfg, per_task=5, dtype=torch.float64)
    s1, l1 = O.mtsac_update(st.clone(), batch, ec, ea, cfg, all_heads=False)
    s2, l2 = O.mtsac_update(st.clone(), batch, ec, ea, cfg, all_heads=True)
    for k in O.LOG_KEYS:
        assert torch.allclose(l1[k], l2[k], rtol=1e-12, atol=1e-14), k
    for a, b in zip(O.tree_leaves(s1.critic), O.tree_leaves(s2.critic)):
        assert torch.allclose(a, b, rtol=1e-12, atol=1e-14)


def test_fp32_tracks_fp64():
    cfg = small_cfg(width=32)
    st64 = O.init_state(cfg, dtype=torch.float64)
    # make the fp64 state exactly the fp32 one
    st32 = O.init_state(cfg, dtype=torch.float32)
    st64 = st32.to(torch.float64)
    b32, ec32, ea32 = O.synthetic_batch(cfg, per_task=8, dtype=torch.float32)
    b64 = tuple(t.double() for t in b32)
    n32, l32 = O.mtsac_update(st32, b32, ec32, ea32, cfg)
    n64, l64 = O.mtsac_update(st64, b64, ec32.double(), ea32.double(), cfg)
    for k in O.LOG_KEYS:
        assert abs(float(l32[k]) - float(l64[k])) <= 1e-4 * max(1e-6, abs(float(l64[k]))), k
    for a, b in zip(O.tree_leaves(n32.actor), O.tree_leaves(n64.actor)):
        assert (a.double() - b).norm() <= 1e-4 * b.norm()


def test_tanh_gaussian_vjp_closed_form():
    """dlogp/dx = 2 tanh(x), dx/dmu = 1, dx/dl = sigma*eps, dlogp/dl|direct = -1,
    da/dx = 1 - a^2.  These are the formulas the CUDA loss kernel implements."""
    torch.manual_seed(0)
    mu = torch.randn(6, 4, dtype=torch.float64, requires_grad=True)
    ls = (torch.randn(6, 4, dtype=torch.float64) * 0.5).requires_grad_(True)
    eps = torch.randn(6, 4, dtype=torch.float64)
    std = torch.exp(ls)
    x = mu + std * eps
    a = torch.tanh(x)
    logp = (-0.5 * eps**2 - 0.5 * math.log(2 * math.pi) - ls).sum(-1) - (2 * (math.log(2) - x - torch.nn.functional.softplus(-2 * x))).sum(-1)
    ga = torch.randn(6, 4, dtype=torch.float64)
    gl = torch.randn(6, dtype=torch.float64)
    (a * ga).sum().add((logp * gl).sum()).backward()
    gx = ga * (1 - a.detach() ** 2) + gl[:, None] * 2 * a.detach()
    assert torch.allclose(mu.grad, gx, rtol=1e-10, atol=1e-12)
    assert torch.allclose(ls.grad, gx * std.detach() * eps - gl[:, None], rtol=1e-10, atol=1e-12)


def test_adam_first_step_and_clip():
    p = {"w": torch.tensor([1.0, -2.0, 3.0], dtype=torch.float64)}
    g = {"w": torch.tensor([3.0, 4.0, 0.0], dtype=torch.float64)}  # norm 5 -> clipped to 1
    opt = {"m": {"w": torch.zeros(3, dtype=torch.float64)}, "v": {"w": torch.zeros(3, dtype=torch.float64)}, "count": 0}
    newp, o2 = O.adam_step(p, g, opt, lr=3e-4, eps=1e-5, b1=0.9, b2=0.999, max_norm=1.0)
    gc = torch.tensor([0.6, 0.8, 0.0], dtype=torch.float64)
    exp = p["w"] - 3e-4 * gc / (gc.abs() + 1e-5)
    assert torch.allclose(newp["w"], exp, rtol=1e-12)
    assert o2["count"] == 1
    # below the threshold the gradient is untouched
    newp2, _ = O.adam_step(p, {"w": gc * 0.5}, opt, 3e-4, 1e-5, 0.9, 0.999, 1.0)
    assert torch.allclose(newp2["w"], p["w"] - 3e-4 * (gc * 0.5) / ((gc * 0.5).abs() + 1e-5), rtol=1e-12)


def test_update_ordering_facts():
    """mtsac.py: target uses the NEW critic (:607-613); alpha step uses log-probs of the OLD actor."""
    cfg = small_cfg()
    st = O.init_state(cfg, dtype=torch.float64)
    batch, ec, ea = O.synthetic_batch(cfg, per_task=4, dtype=torch.float64)
    new, logs, grads, aux = O.mtsac_update(st.clone(), batch, ec, ea, cfg, return_grads=True)
    for n, t_old, t_new in zip(O.tree_leaves(new.critic), O.tree_leaves(st.critic_target), O.tree_leaves(new.critic_target)):
        assert torch.allclose(t_new, cfg.tau * n + (1 - cfg.tau) * t_old, rtol=1e-12)
    obs = batch[0]
    task = obs[:, -cfg.num_tasks:].argmax(1)
    _, logp_old = O.actor_sample_and_log_prob(st.actor, obs, ea, cfg)
    g = torch.zeros(cfg.num_tasks, dtype=torch.float64)
    g.index_add_(0, task, -(logp_old + cfg.target_entropy) / obs.shape[0])
    assert torch.allclose(grads["alpha"], g, rtol=1e-10)
    assert set(logs) == set(O.LOG_KEYS)


def test_input_state_not_mutated():
    cfg = small_cfg()
    st = O.init_state(cfg, dtype=torch.float64)
    ref = st.clone()
    batch, ec, ea = O.synthetic_batch(cfg, per_task=4, dtype=torch.float64)
    O.mtsac_update(st, batch, ec, ea, cfg)
    assert st.opt["critic"]["count"] == 0
    for a, b in zip(O.tree_leaves(st.critic), O.tree_leaves(ref.critic)):
        assert torch.equal(a, b)


def test_per_task_gradients_average_to_the_batch_gradient():
    """oracle/taskgrad_oracle.py: with equally many rows per task the full-batch loss is the mean of the per-task
    losses, so the per-task gradients must average to the gradients mtsac_update differentiates."""
    from oracle import taskgrad_oracle as TG

    cfg = O.OracleConfig(num_tasks=3, obs_dim=42, action_dim=4, width=32)
    st = O.init_state(cfg, seed=4, dtype=torch.float64)
    batch, ec, ea = O.synthetic_batch(cfg, 8, seed=9, dtype=torch.float64)
    _, _, grads, _ = O.mtsac_update(st, batch, ec, ea, cfg, return_grads=True)
    # with the UN-split branch's targets (a' sampled on next_observations, mtsac.py:526-536) the per-task losses are the
    # batch loss cut by task ...
    _, target = TG._targets(st, batch, ec, cfg, split=False)
    mean_c = O.tree_map(lambda *xs: sum(xs) / len(xs), *TG.critic_task_grads(st.critic, batch, target, cfg))
    for a, b in zip(O.tree_leaves(mean_c), O.tree_leaves(grads["critic"])):
        assert torch.allclose(a, b, rtol=1e-9, atol=1e-12)
    # ... whereas the reference's split / compute_weights branches sample a' on data.observations (mtsac.py:515-520,
    # 995-999): restated literally, so their mean is NOT the un-split gradient
    per = TG.per_task_grads(st, batch, ec, ea, cfg)
    mean_q = O.tree_map(lambda *xs: sum(xs) / len(xs), *per["critic"])
    assert any(not torch.allclose(a, b, rtol=1e-6, atol=1e-12) for a, b in zip(O.tree_leaves(mean_q), O.tree_leaves(grads["critic"])))
    g = TG.flatten(per["critic"])
    avg, cos = TG.vmap_cos_sim(g)
    assert cos.shape == (3, 3) and torch.allclose(torch.diagonal(cos), torch.ones(3, dtype=torch.float64), atol=1e-6)
    assert -1.0 <= float(avg) <= 1.0
    # the actor's per-task gradients use the CURRENT critic (compute_weights, mtsac.py:1060), the update's the NEW one,
    # so they are only compared in shape
    assert TG.flatten(per["actor"]).shape[0] == 3


def test_split_update_without_surgery_equals_plain_update():
    """oracle/taskgrad_oracle.py: averaging the per-task gradients (what the split-loss branch does without pcgrad,
    mtsac.py:581-585) must reproduce mtsac_update exactly; with pcgrad the conflict count is symmetric (pairs / 2)."""
    from oracle import taskgrad_oracle as TG

    cfg = O.OracleConfig(num_tasks=3, obs_dim=42, action_dim=4, width=32)
    st = O.init_state(cfg, seed=4, dtype=torch.float64)
    batch, ec, ea = O.synthetic_batch(cfg, 8, seed=9, dtype=torch.float64)
    ns, _ = TG.mtsac_update_pcgrad(st, batch, ec, ea, cfg, critic=False, actor=False)
    ref, _ = O.mtsac_update(st, batch, ec, ea, cfg)
    for a, b in zip(O.tree_leaves(ns.critic) + O.tree_leaves(ns.actor), O.tree_leaves(ref.critic) + O.tree_leaves(ref.actor)):
        assert torch.allclose(a, b, rtol=1e-9, atol=1e-12)
    assert torch.allclose(ns.log_alpha, ref.log_alpha)
    g = torch.tensor([[1.0, 0.0], [-1.0, 1.0], [0.0, 2.0]], dtype=torch.float64)
    avg, stats = TG.pcgrad(g)
    # g0 vs g1 conflict (dot = -1): g0 -> g0 + 0.5 g1 = (0.5, 0.5); g1 -> g1 + g0 = (0, 1); g2 untouched
    assert stats["n_grad_conflicts"] == 1.0
    assert torch.allclose(avg, torch.tensor([0.5 / 3, 3.5 / 3], dtype=torch.float64), atol=1e-7)


def test_tanh_gaussian_log_prob_agrees_with_torch_distributions():
    """Row a7: the oracle restates distrax's Transformed(MultivariateNormalDiag, Block(Tanh, 1)).sample_and_log_prob
    (mtrl/nn/distributions.py:6-16).  PyTorch ships an independent implementation of the same distribution
    (TransformedDistribution(Independent(Normal), TanhTransform)); the two must agree on the log-density of the sample."""
    import torch.distributions as D

    cfg = O.OracleConfig(num_tasks=4, obs_dim=43, action_dim=4, width=32)
    st = O.init_state(cfg, seed=3, dtype=torch.float64)
    for k in ("kernel", "bias"):
        st.actor["heads"][k] = st.actor["heads"][k] * 200.0          # non-trivial means / log-stds
    batch, _, ea = O.synthetic_batch(cfg, 16, seed=12, dtype=torch.float64)
    obs = batch[0]
    a, logp = O.actor_sample_and_log_prob(st.actor, obs, ea, cfg)
    out = O.multihead_forward(st.actor, obs, cfg.num_tasks, cfg.depth)
    mean, log_std = out[..., :4], torch.clamp(out[..., 4:], cfg.log_std_min, cfg.log_std_max)
    dist = D.TransformedDistribution(D.Independent(D.Normal(mean, torch.exp(log_std)), 1), [D.TanhTransform(cache_size=1)])
    x = mean + torch.exp(log_std) * ea
    y = torch.tanh(x)
    assert torch.allclose(a, y)
    # evaluate the density through the pre-image to avoid atanh round-off near +-1
    ref = dist.base_dist.log_prob(x) - D.TanhTransform().log_abs_det_jacobian(x, y).sum(-1)
    assert torch.allclose(logp, ref, rtol=1e-10, atol=1e-10)
    inside = y.abs().max(dim=-1).values < 0.999
    assert inside.any() and torch.allclose(dist.log_prob(y)[inside], logp[inside], rtol=1e-6, atol=1e-6)


def test_adam_restatement_agrees_with_torch_optim_adam():
    """Rows a14: the oracle restates optax.adam(lr, b1, b2, eps, eps_root=0) (mtrl/config/optim.py:26-43).  torch.optim.Adam
    is an independent implementation of the same update (eps added to sqrt of the bias-corrected second moment): three
    unclipped steps on the same gradients must coincide.  The global-norm clip is checked against its definition."""
    g = torch.Generator().manual_seed(0)
    p0 = {"a": {"kernel": torch.randn(7, 5, generator=g, dtype=torch.float64), "bias": torch.randn(5, generator=g, dtype=torch.float64)}}
    grads = [O.tree_map(lambda x: torch.randn(x.shape, generator=g, dtype=torch.float64) * (10.0 ** (i - 1)), p0) for i in range(3)]
    opt = {"count": 0, "m": O.tree_map(torch.zeros_like, p0), "v": O.tree_map(torch.zeros_like, p0)}
    leaves = [x.clone().requires_grad_(True) for x in O.tree_leaves(p0)]
    ref = torch.optim.Adam(leaves, lr=3e-4, betas=(0.9, 0.999), eps=1e-5)
    p = p0
    for gr in grads:
        p, opt = O.adam_step(p, gr, opt, 3e-4, 1e-5, 0.9, 0.999, None)
        for leaf, gl in zip(leaves, O.tree_leaves(gr)):
            leaf.grad = gl.clone()
        ref.step()
        for a, b in zip(O.tree_leaves(p), leaves):
            assert torch.allclose(a, b.detach(), rtol=1e-12, atol=1e-15)
    # optax.clip_by_global_norm(1.0): untouched below the threshold, rescaled to exactly the threshold above it
    big = O.tree_map(lambda x: x * 100, grads[1])
    opt0 = {"count": 0, "m": O.tree_map(torch.zeros_like, p0), "v": O.tree_map(torch.zeros_like, p0)}
    _, o_big = O.adam_step(p0, big, opt0, 3e-4, 1e-5, 0.9, 0.999, 1.0)
    m_norm = O.global_norm(o_big["m"]) / 0.1            # m = (1 - b1) * clipped gradient
    assert abs(float(m_norm) - 1.0) < 1e-9
    small = O.tree_map(lambda x: x * 1e-3, grads[1])
    _, o_small = O.adam_step(p0, small, opt0, 3e-4, 1e-5, 0.9, 0.999, 1.0)
    assert torch.allclose(O.tree_leaves(o_small["m"])[0], 0.1 * O.tree_leaves(small)[0])
