"""`mtrl_b200.optim`: the reference's multi-task gradient transformations behind the optax protocol
(init(params) -> state; update(updates, state, params, **extra) -> (updates, state); mtrl/optim/dummy.py:5-20,
pcgrad.py:133-136) against literal restatements of the reference loops (oracle/taskgrad_oracle.py)."""
import pytest
import torch

from oracle import taskgrad_oracle as TG

pytestmark = pytest.mark.gpu


def make_updates(T, seed, conflict=True):
    g = torch.Generator().manual_seed(seed)
    tree = {"params": {"layer_0": {"kernel": torch.randn(T, 37, 24, generator=g), "bias": torch.randn(T, 24, generator=g)},
                       "heads": {"kernel": torch.randn(T, 5, 24, 3, generator=g) * 3.0, "bias": torch.randn(T, 5, 3, generator=g)}}}
    if conflict:   # odd tasks pull the other way on one leaf, so projections actually happen
        tree["params"]["layer_0"]["kernel"][1::2] *= -1.0
    return tree


def ravel(tree):
    return torch.stack([torch.cat([x[t].flatten() for x in TG.O.tree_leaves(tree)]) for t in range(TG.O.tree_leaves(tree)[0].shape[0])])


def to_cuda(tree):
    return TG.O.tree_map(lambda x: x.cuda(), tree)


def flat_out(tree):
    return torch.cat([x.flatten() for x in TG.O.tree_leaves(tree)]).double().cpu()


def rel(a, b):
    return float((a - b).norm() / b.norm())


def test_dummy_is_the_mean_over_tasks(cuda):
    from mtrl_b200.optim import dummy_multitask_optimizer

    up = make_updates(6, 1)
    tx = dummy_multitask_optimizer()
    state = tx.init(to_cuda(up))
    assert state == {}
    new, state = tx.update(to_cuda(up), state)
    assert state == {}
    assert new["params"]["heads"]["kernel"].shape == (5, 24, 3)
    assert rel(flat_out(new), ravel(up).double().mean(0)) < 1e-6


def test_pcgrad_matches_the_literal_loop(cuda):
    from mtrl_b200.optim import PCGradState, pcgrad

    T = 6
    up = make_updates(T, 2)
    tx = pcgrad(T)
    state = tx.init(to_cuda(up))
    assert isinstance(state, PCGradState) and float(state.n_grad_conflicts) == 0
    perm = torch.tensor([3, 0, 5, 1, 4, 2])
    new, state = tx.update(to_cuda(up), state, to_cuda(up), key=7, perm=perm)
    ref, stats = TG.pcgrad(ravel(up).double(), perm)
    assert stats["n_grad_conflicts"] > 0
    assert rel(flat_out(new), ref) < 1e-5
    assert float(state.n_grad_conflicts) == stats["n_grad_conflicts"]
    assert abs(float(state.avg_grad_magnitude) - float(stats["avg_grad_magnitude"])) < 1e-4 * float(stats["avg_grad_magnitude"])
    assert abs(float(state.avg_grad_magnitude_before_surgery) - float(stats["avg_grad_magnitude_before_surgery"])) < \
        1e-4 * float(stats["avg_grad_magnitude_before_surgery"])
    # the key drives the permutation (pcgrad.py:79): same key, same result; no key is an error as in the reference
    a, _ = tx.update(to_cuda(up), state, to_cuda(up), key=11)
    b, _ = tx.update(to_cuda(up), state, to_cuda(up), key=11)
    assert torch.equal(flat_out(a), flat_out(b))
    with pytest.raises(AssertionError):
        tx.update(to_cuda(up), state, to_cuda(up))
    with pytest.raises(ValueError):
        tx.update(to_cuda(make_updates(T - 1, 2)), state, to_cuda(up), key=1)


def test_cagrad_matches_the_literal_restatement(cuda):
    from mtrl_b200.optim import cagrad

    T = 5
    up = make_updates(T, 3)
    tx = cagrad(T)
    new, state = tx.update(to_cuda(up), tx.init(to_cuda(up)), to_cuda(up))
    ref, stats = TG.cagrad(ravel(up).double())
    assert rel(flat_out(new), ref) < 1e-4
    assert (state.task_weights.double().cpu() - stats["task_weights"]).abs().max() < 1e-4
    assert abs(float(state.cagrad_objective) - float(stats["cagrad_objective"])) < 1e-4 * abs(float(stats["cagrad_objective"]))


@pytest.mark.parametrize("clip", [None, 1.0])
def test_gradnorm_matches_the_literal_restatement(cuda, clip):
    from mtrl_b200.optim import gradnorm

    T = 4
    up = make_updates(T, 4, conflict=False)
    tx = gradnorm(None, T, max_grad_norm=clip)
    new, state = tx.update(to_cuda(up), tx.init(to_cuda(up)), to_cuda(up), task_losses=torch.ones(T))
    ref, stats = TG.gradnorm(ravel(up).double(), torch.ones(T, dtype=torch.float64), max_grad_norm=clip)
    assert rel(flat_out(new), ref) < 1e-5
    assert torch.allclose(state.task_weights.cpu().double(), stats["task_weights"])
    assert abs(float(state.grad_magnitude) - float(stats["grad_magnitude"])) < 1e-4 * float(stats["grad_magnitude"])


def test_dummy_config_update_takes_the_split_path(cuda):
    """DummyMultiTaskConfig (mtrl/config/optim.py:46-59) = mean of the per-task gradients of the reference's SPLIT losses:
    a' sampled on data.observations in the critic target, explore term in the actor loss -- not the un-split update."""
    import dataclasses

    import sac_util as SU
    from mtrl_b200.config.networks import ContinuousActionPolicyConfig, QValueFunctionConfig
    from mtrl_b200.config.nn import MultiHeadConfig
    from mtrl_b200.config.optim import DummyMultiTaskConfig, OptimizerConfig
    from mtrl_b200.rl.algorithms import MTSAC, MTSACConfig
    from oracle import mtsac_oracle as O

    T, W, per_task = 5, 128, 32
    cfg = O.OracleConfig(num_tasks=T, obs_dim=39 + T, action_dim=4, width=W)
    st = O.init_state(cfg, seed=9, dtype=torch.float32)
    for net, scale in ((st.actor, 100.0), (st.critic, 30.0), (st.critic_target, 30.0)):
        for k in ("kernel", "bias"):
            net["heads"][k] = net["heads"][k] * scale
    opt = DummyMultiTaskConfig(lr=cfg.lr, max_grad_norm=cfg.max_grad_norm, eps=cfg.adam_eps)
    assert opt.requires_split_task_losses
    netc = MultiHeadConfig(width=W, depth=cfg.depth, num_tasks=T, optimizer=opt)
    mc = MTSACConfig(num_tasks=T, gamma=cfg.gamma, actor_config=ContinuousActionPolicyConfig(network_config=netc),
                     critic_config=QValueFunctionConfig(network_config=netc),
                     temperature_optimizer_config=OptimizerConfig(lr=cfg.alpha_lr, max_grad_norm=None, eps=cfg.adam_eps))
    agent = MTSAC.initialize(mc, SU.EnvSpec(cfg.obs_dim, 4), seed=9, max_batch=per_task * T, precision="fp32x3")
    assert agent.split_actor_losses and agent.split_critic_losses
    SU.load_oracle_state(agent, st)
    batch, ec, ea = O.synthetic_batch(cfg, per_task, seed=80, dtype=torch.float32)
    st64 = st.to(torch.float64)
    b64 = tuple(b.double() for b in batch)
    new, stats = TG.mtsac_update_pcgrad(st64, b64, ec.double(), ea.double(), cfg, surgery="dummy")
    plain, _ = O.mtsac_update(st64, b64, ec.double(), ea.double(), cfg)
    _, logs = agent.update(tuple(b.cuda() for b in batch), eps_c=ec.cuda(), eps_a=ea.cuda(), check=True)
    for name, ens, tree in (("critic", True, agent.critic.grads), ("actor", False, agent.actor.grads)):
        for leaf, e in SU.compare_trees(stats[name]["grad_tree"], tree, ens).items():
            assert e <= 1e-3, (name, leaf, e)
    for name, new_t, plain_t, tree, ens in (("critic", new.critic, plain.critic, agent.critic.params, True),
                                            ("actor", new.actor, plain.actor, agent.actor.params, False)):
        errs = SU.compare_trees(new_t, tree, ens)
        assert max(errs.values()) <= 1e-3, (name, errs)
        # ... and it is NOT the un-split update
        assert max(SU.compare_trees(plain_t, tree, ens).values()) > 1e-2, name
    assert abs(float(logs["metrics/explore_loss"]) - float(stats["logs"]["metrics/explore_loss"])) <= 1e-3 * float(stats["logs"]["metrics/explore_loss"])
    del dataclasses
