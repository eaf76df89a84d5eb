"""Per-task gradients (MTSAC.compute_weights, mtsac.py:870-1170) from the CUDA backward + per-task dW GEMMs vs PyTorch
autograd run one task at a time (oracle/taskgrad_oracle.py)."""
import dataclasses

import pytest
import torch

import sac_util as SU
from oracle import mtsac_oracle as O
from oracle import taskgrad_oracle as TG

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("T,W,per_task", [(4, 128, 32), (10, 256, 128)])
def test_per_task_gradients_and_cos_sim(cuda, T, W, per_task):
    cfg = O.OracleConfig(num_tasks=T, obs_dim=39 + T, action_dim=4, width=W)
    st = O.init_state(cfg, seed=2, dtype=torch.float32)
    # make the heads non-trivial so the tasks' gradients differ in direction
    for net, scale in ((st.actor, 100.0), (st.critic, 30.0), (st.critic_target, 30.0)):
        for k in ("kernel", "bias"):
            net["heads"][k] = net["heads"][k] * scale
    agent = SU.make_agent(cfg, per_task, seed=2)
    SU.load_oracle_state(agent, st)
    batch, ec, ea = O.synthetic_batch(cfg, per_task, seed=31, dtype=torch.float32)
    before = agent._flat["critic_params"].clone()
    got = agent.per_task_gradients(tuple(b.cuda() for b in batch), eps_c=ec.cuda(), eps_a=ea.cuda())
    assert torch.equal(before, agent._flat["critic_params"]), "per_task_gradients must not update parameters"
    tcfg = dataclasses.replace(cfg, matmul_operands="tf32")
    st64 = st.to(torch.float64)
    b64 = tuple(b.double() for b in batch)
    ref = TG.per_task_grads(st64, b64, ec.double(), ea.double(), tcfg)
    for name, ens in (("critic", True), ("actor", False)):
        assert got[name].shape[0] == T
        worst = 0.0
        for t in range(T):
            tree = agent.task_gradient_view(got[name][t], critic=ens)
            for leaf, o, a in SU._pairs(ref[name][t], SU._net(tree, ens)):
                if o.abs().max() == 0:
                    assert float(a.abs().max()) == 0.0, (name, t, leaf)   # other tasks' heads
                    continue
                e = SU.rel(a, o)
                worst = max(worst, e)
                assert e <= 2e-2, (name, t, leaf, e)
        # metrics: cosine-similarity matrix and its summaries (utils.py:49-72, 118-174)
        g_ref = TG.flatten(ref[name])
        avg_ref, cos_ref = TG.vmap_cos_sim(g_ref)
        cm_ref = TG.conflict_metrics(cos_ref, g_ref)
        _, logs = agent.compute_weights(tuple(b.cuda() for b in batch), eps_c=ec.cuda(), eps_a=ea.cuda())
        import json
        import os

        ref_keys = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "compute_weights_keys.json")))
        assert set(logs) == set(ref_keys), (set(ref_keys) - set(logs), set(logs) - set(ref_keys))   # mtsac.py:1085-1170
        cos = logs[f"{name}_pairwise_cos_sim"].double().cpu()
        assert (cos - cos_ref).abs().max() <= 2e-2, (name, float((cos - cos_ref).abs().max()))
        assert abs(float(logs[f"{name}_avg_cos_sim"]) - float(avg_ref)) <= 1e-2
        mag = logs[f"{name}_per_task_grad_magnitude"].double().cpu()
        assert SU.rel(mag, cm_ref["per_task_grad_magnitude"]) <= 1e-2
        assert abs(float(logs[f"{name}_avg_grad_magnitude"]) - float(g_ref.norm(dim=1).mean())) <= 1e-2 * float(g_ref.norm(dim=1).mean())
        # compute_gram_metrics (mtsac.py:733-771) and compute_effective_rank (utils.py:104-115)
        gram_ref = g_ref @ g_ref.T
        assert SU.rel(logs[f"{name}_pairwise_gram"], gram_ref) <= 2e-2
        sv = torch.linalg.svdvals(gram_ref)
        sd = sv / sv.sum()
        assert abs(float(logs[f"{name}_effective_rank"]) - float(torch.exp(-(sd * torch.log(sd + 1e-10)).sum()))) <= 5e-2
        # element-wise metrics (utils.py:75-101, 146-156): thresholds at 1e-3 / 1.0, so tf32 noise moves a few elements
        em = TG.elementwise_metrics(g_ref)
        assert SU.rel(logs[f"{name}_per_task_participation_ratio"], em["per_task_participation_ratio"]) <= 1e-2
        rate = logs[f"{name}_pairwise_interference_rate"].double().cpu()
        assert (rate - em["pairwise_interference_rate"]).abs().max() <= 2e-3 + 2e-2 * float(em["pairwise_interference_rate"].max())
        assert abs(float(logs[f"{name}_avg_interference_rate"]) - float(em["avg_interference_rate"])) <= \
            1e-3 + 2e-2 * float(em["avg_interference_rate"])
        # support metrics (mtsac.py:774-860): the 0.8-quantile threshold by radix select, then pairwise counts
        sm = TG.support_metrics(g_ref)
        assert SU.rel(logs[f"{name}_per_task_support_size"], sm["per_task_support_size"]) <= 1e-3
        assert (logs[f"{name}_pairwise_jaccard"].double().cpu() - sm["pairwise_jaccard"]).abs().max() <= 2e-2
        assert (logs[f"{name}_pairwise_genuine_conflict_rate"].double().cpu() - sm["genuine_conflict_rate"]).abs().max() <= 2e-2
        assert abs(float(logs[f"{name}_ghost_to_genuine_ratio"]) - float(sm["ghost_to_genuine_ratio"])) <= \
            3e-2 * float(sm["ghost_to_genuine_ratio"]) + 1e-3
        assert abs(float(logs[f"{name}_avg_jaccard"]) - float(sm["avg_jaccard"])) <= 1e-2
        off = 1 - torch.eye(T, dtype=torch.float64)
        assert abs(float(logs[f"{name}_gram_off_diag_mean"]) - float((gram_ref * off).sum() / (T * (T - 1)))) <= \
            2e-2 * float(gram_ref.abs().max())


def test_uneven_split_is_rejected(cuda):
    cfg = O.OracleConfig(num_tasks=4, obs_dim=43, action_dim=4, width=64)
    agent = SU.make_agent(cfg, 16, seed=1)
    batch, ec, ea = O.synthetic_batch(cfg, 16, seed=3)
    task = batch[0][:, -4:].argmax(1)
    keep = torch.ones(task.shape[0], dtype=torch.bool)
    keep[(task == 1).nonzero().flatten()[:4]] = False
    keep[(task == 2).nonzero().flatten()[:4]] = False      # 56 rows: 14 per task on average, but 16/12/12/16
    with pytest.raises(ValueError):
        agent.per_task_gradients(tuple(b[keep].cuda() for b in batch))


def test_pcgrad_update_matches_oracle(cuda):
    """PCGradConfig on both networks (mtrl/config/optim.py:62-76): split losses -> pcgrad (coefficient space over the
    Gram matrix on the GPU) -> clip + adam, vs the literal pcgrad.py loop on autograd per-task gradients."""
    from mtrl_b200.config.networks import ContinuousActionPolicyConfig, QValueFunctionConfig
    from mtrl_b200.config.nn import MultiHeadConfig
    from mtrl_b200.config.optim import OptimizerConfig, PCGradConfig
    from mtrl_b200.rl.algorithms import MTSAC, MTSACConfig

    T, W, per_task = 6, 128, 32
    cfg = O.OracleConfig(num_tasks=T, obs_dim=39 + T, action_dim=4, width=W)
    st = O.init_state(cfg, seed=5, dtype=torch.float32)
    for net, scale in ((st.actor, 100.0), (st.critic, 30.0), (st.critic_target, 30.0)):
        for k in ("kernel", "bias"):
            net["heads"][k] = net["heads"][k] * scale
    opt = PCGradConfig(lr=cfg.lr, max_grad_norm=cfg.max_grad_norm, eps=cfg.adam_eps, num_tasks=T)
    netc = MultiHeadConfig(width=W, depth=cfg.depth, num_tasks=T, optimizer=opt)
    mc = MTSACConfig(num_tasks=T, gamma=cfg.gamma, actor_config=ContinuousActionPolicyConfig(network_config=netc),
                     critic_config=QValueFunctionConfig(network_config=netc),
                     temperature_optimizer_config=OptimizerConfig(lr=cfg.alpha_lr, max_grad_norm=None, eps=cfg.adam_eps))
    agent = MTSAC.initialize(mc, SU.EnvSpec(cfg.obs_dim, 4), seed=5, max_batch=per_task * T)
    SU.load_oracle_state(agent, st)
    batch, ec, ea = O.synthetic_batch(cfg, per_task, seed=77, dtype=torch.float32)
    # make tasks pull in different directions: task-dependent reward sign
    task = batch[0][:, -T:].argmax(1)
    rew = batch[4] * torch.where(task % 2 == 0, 1.0, -1.0).reshape(-1, 1)
    batch = (batch[0], batch[1], batch[2], batch[3], rew)
    perm_c, perm_a = torch.tensor([3, 0, 5, 1, 4, 2]), torch.tensor([1, 2, 0, 5, 3, 4])
    tcfg = dataclasses.replace(cfg, matmul_operands="tf32")
    st64 = st.to(torch.float64)
    new, stats = TG.mtsac_update_pcgrad(st64, tuple(b.double() for b in batch), ec.double(), ea.double(), tcfg, perm_c, perm_a)
    _, logs = agent.update(tuple(b.cuda() for b in batch), eps_c=ec.cuda(), eps_a=ea.cuda(), check=True, pcgrad_perm=(perm_c, perm_a))
    got = agent.pcgrad_stats()
    assert stats["critic"]["n_grad_conflicts"] > 0, "degenerate test: no conflicting task gradients"
    # the split actor branch carries the explore term (mtsac.py:631-637, 668-682): logged, and part of the actor loss
    ex_ref, al_ref = float(stats["logs"]["metrics/explore_loss"]), float(stats["logs"]["losses/actor_loss"])
    assert ex_ref > 1e-2, "degenerate test: no explore term"
    assert abs(float(logs["metrics/explore_loss"]) - ex_ref) <= 1e-3 * ex_ref
    assert abs(float(logs["losses/actor_loss"]) - al_ref) <= 1e-3 * abs(al_ref)
    for net in ("critic", "actor"):
        assert float(got[net]["n_grad_conflicts"]) == stats[net]["n_grad_conflicts"], net
        for k in ("avg_grad_magnitude", "avg_grad_magnitude_before_surgery"):
            assert abs(float(got[net][k]) - float(stats[net][k])) <= 1e-2 * float(stats[net][k]), (net, k)
    # the Adam step (new - old) of the trunk kernels, the part pcgrad changes
    # the combined gradient the surgery hands to clip + adam, leaf by leaf
    for name, ens, tree in (("critic", True, agent.critic.grads), ("actor", False, agent.actor.grads)):
        for leaf, e in SU.compare_trees(stats[name]["grad_tree"], tree, ens).items():
            assert e <= 1e-2, (name, leaf, e)
    for name, old_t, new_t, tree, ens in (("critic", st64.critic, new.critic, agent.critic.params, True),
                                          ("actor", st64.actor, new.actor, agent.actor.params, False)):
        for leaf, e in SU.compare_deltas(old_t, new_t, tree, ens).items():
            if leaf.startswith("layer_") and leaf.endswith("kernel"):
                assert e <= 3e-2, (name, leaf, e)
    assert SU.rel(agent.alpha.params["params"]["log_alpha"], new.log_alpha) <= 1e-3


def test_cagrad_update_matches_oracle(cuda):
    """CAGradConfig on both networks (mtrl/config/optim.py:104-124): per-task clip, weight SGD and combination of
    cagrad.py computed from the Gram matrix on the GPU vs the literal restatement on autograd per-task gradients."""
    from mtrl_b200.config.networks import ContinuousActionPolicyConfig, QValueFunctionConfig
    from mtrl_b200.config.nn import MultiHeadConfig
    from mtrl_b200.config.optim import CAGradConfig, OptimizerConfig
    from mtrl_b200.rl.algorithms import MTSAC, MTSACConfig

    T, W, per_task = 6, 128, 32
    cfg = O.OracleConfig(num_tasks=T, obs_dim=39 + T, action_dim=4, width=W)
    st = O.init_state(cfg, seed=6, dtype=torch.float32)
    for net, scale in ((st.actor, 100.0), (st.critic, 30.0), (st.critic_target, 30.0)):
        for k in ("kernel", "bias"):
            net["heads"][k] = net["heads"][k] * scale
    opt = CAGradConfig(lr=cfg.lr, max_grad_norm=cfg.max_grad_norm, eps=cfg.adam_eps, num_tasks=T)
    netc = MultiHeadConfig(width=W, depth=cfg.depth, num_tasks=T, optimizer=opt)
    mc = MTSACConfig(num_tasks=T, gamma=cfg.gamma, actor_config=ContinuousActionPolicyConfig(network_config=netc),
                     critic_config=QValueFunctionConfig(network_config=netc),
                     temperature_optimizer_config=OptimizerConfig(lr=cfg.alpha_lr, max_grad_norm=None, eps=cfg.adam_eps))
    agent = MTSAC.initialize(mc, SU.EnvSpec(cfg.obs_dim, 4), seed=6, max_batch=per_task * T)
    SU.load_oracle_state(agent, st)
    batch, ec, ea = O.synthetic_batch(cfg, per_task, seed=78, dtype=torch.float32)
    task = batch[0][:, -T:].argmax(1)
    rew = batch[4] * torch.where(task % 2 == 0, 1.0, -1.0).reshape(-1, 1)
    batch = (batch[0], batch[1], batch[2], batch[3], rew)
    tcfg = dataclasses.replace(cfg, matmul_operands="tf32")
    st64 = st.to(torch.float64)
    new, stats = TG.mtsac_update_pcgrad(st64, tuple(b.double() for b in batch), ec.double(), ea.double(), tcfg, surgery="cagrad")
    agent.update(tuple(b.cuda() for b in batch), eps_c=ec.cuda(), eps_a=ea.cuda(), check=True)
    got = agent.pcgrad_stats()
    for net in ("critic", "actor"):
        assert (got[net]["task_weights"].double().cpu() - stats[net]["task_weights"]).abs().max() <= 1e-3, net
        for k in ("avg_grad_magnitude", "avg_grad_magnitude_before_surgery", "cagrad_objective"):
            assert abs(float(got[net][k]) - float(stats[net][k])) <= 1e-2 * abs(float(stats[net][k])) + 1e-6, (net, k)
    # the combined gradient the surgery hands to clip + adam, leaf by leaf
    for name, ens, tree in (("critic", True, agent.critic.grads), ("actor", False, agent.actor.grads)):
        for leaf, e in SU.compare_trees(stats[name]["grad_tree"], tree, ens).items():
            assert e <= 1e-2, (name, leaf, e)
    for name, old_t, new_t, tree, ens in (("critic", st64.critic, new.critic, agent.critic.params, True),
                                          ("actor", st64.actor, new.actor, agent.actor.params, False)):
        for leaf, e in SU.compare_deltas(old_t, new_t, tree, ens).items():
            if leaf.startswith("layer_") and leaf.endswith("kernel"):
                assert e <= 3e-2, (name, leaf, e)


@pytest.mark.parametrize("clip", [False, True])
def test_gradnorm_update_matches_oracle(cuda, clip):
    """GradNormConfig (mtrl/config/optim.py:79-102).  The reference's gradnorm loss ignores the task weights
    (gradnorm.py:134-142): the literal restatement (autograd gives a zero weight gradient) reduces to the sum of the
    per-task gradients, per-task clipped when max_grad_norm is set -- which is what the GPU path must hand to clip + adam."""
    from mtrl_b200.config.networks import ContinuousActionPolicyConfig, QValueFunctionConfig
    from mtrl_b200.config.nn import MultiHeadConfig
    from mtrl_b200.config.optim import GradNormConfig, OptimizerConfig
    from mtrl_b200.rl.algorithms import MTSAC, MTSACConfig

    T, W, per_task = 5, 128, 32
    cfg = O.OracleConfig(num_tasks=T, obs_dim=39 + T, action_dim=4, width=W, max_grad_norm=1.0 if clip else None)
    st = O.init_state(cfg, seed=8, dtype=torch.float32)
    for net, scale in ((st.actor, 100.0), (st.critic, 30.0), (st.critic_target, 30.0)):
        for k in ("kernel", "bias"):
            net["heads"][k] = net["heads"][k] * scale
    opt = GradNormConfig(lr=cfg.lr, max_grad_norm=cfg.max_grad_norm, eps=cfg.adam_eps, num_tasks=T, gradnorm_optimizer=OptimizerConfig())
    netc = MultiHeadConfig(width=W, depth=cfg.depth, num_tasks=T, optimizer=opt)
    mc = MTSACConfig(num_tasks=T, gamma=cfg.gamma, actor_config=ContinuousActionPolicyConfig(network_config=netc),
                     critic_config=QValueFunctionConfig(network_config=netc),
                     temperature_optimizer_config=OptimizerConfig(lr=cfg.alpha_lr, max_grad_norm=None, eps=cfg.adam_eps))
    agent = MTSAC.initialize(mc, SU.EnvSpec(cfg.obs_dim, 4), seed=8, max_batch=per_task * T)
    SU.load_oracle_state(agent, st)
    batch, ec, ea = O.synthetic_batch(cfg, per_task, seed=79, dtype=torch.float32)
    tcfg = dataclasses.replace(cfg, matmul_operands="tf32")
    st64 = st.to(torch.float64)
    new, stats = TG.mtsac_update_pcgrad(st64, tuple(b.double() for b in batch), ec.double(), ea.double(), tcfg, surgery="gradnorm",
                                        gradnorm_clip=clip)
    agent.update(tuple(b.cuda() for b in batch), eps_c=ec.cuda(), eps_a=ea.cuda(), check=True)
    got = agent.pcgrad_stats()
    for net in ("critic", "actor"):
        assert torch.allclose(stats[net]["task_weights"], torch.ones(T, dtype=torch.float64))   # the weights cannot move
        assert abs(float(got[net]["grad_magnitude"]) - float(stats[net]["grad_magnitude"])) <= 1e-2 * float(stats[net]["grad_magnitude"])
    for name, ens, tree in (("critic", True, agent.critic.grads), ("actor", False, agent.actor.grads)):
        for leaf, e in SU.compare_trees(stats[name]["grad_tree"], tree, ens).items():
            assert e <= 1e-2, (name, leaf, e)
