"""Fused CUDA MT-PPO update (mtppo.py:196-317) vs the fp64 oracle on identical rollouts, weights and noise."""
import dataclasses

import pytest
import torch

import sac_util as SU
from oracle import mtsac_oracle as O
from oracle import ppo_oracle as P

pytestmark = pytest.mark.gpu


def make_ppo(cfg: P.PPOConfig, steps: int, multihead: bool, seed=1):
    from mtrl_b200.config.networks import ContinuousActionPolicyConfig, ValueFunctionConfig
    from mtrl_b200.config.nn import MultiHeadConfig, VanillaNetworkConfig
    from mtrl_b200.config.optim import OptimizerConfig
    from mtrl_b200.rl.algorithms import MTPPO, MTPPOConfig

    n = cfg.net
    opt = OptimizerConfig(lr=n.lr, max_grad_norm=n.max_grad_norm, eps=n.adam_eps)
    net = (MultiHeadConfig(width=n.width, depth=n.depth, num_tasks=n.num_tasks, optimizer=opt) if multihead
           else VanillaNetworkConfig(width=n.width, depth=n.depth, optimizer=opt, use_layer_norm=n.use_layer_norm,
                                     use_skip_connections=n.use_skip_connections))
    pc = MTPPOConfig(num_tasks=n.num_tasks if multihead else 5,
                     policy_config=ContinuousActionPolicyConfig(network_config=net, squash_tanh=False),
                     vf_config=ValueFunctionConfig(network_config=net), clip_eps=cfg.clip_eps, clip_vf_loss=cfg.clip_vf_loss,
                     entropy_coefficient=cfg.entropy_coefficient, vf_coefficient=cfg.vf_coefficient,
                     normalize_advantages=cfg.normalize_advantages)
    return MTPPO.initialize(pc, SU.EnvSpec(n.obs_dim, n.action_dim), seed=seed, rollout_steps=steps)


def pairs(otree, agent, ts):
    inner = agent._inner(ts)
    out = []
    for k, v in otree.items():
        ak = ("VmapDense_0" if agent._multihead else f"layer_{agent._cfg.depth}") if k == "heads" else k
        for leaf in v:
            o = v[leaf]
            if k == "heads" and not agent._multihead:
                o = o[0]
            out.append((f"{ak}/{leaf}", o, inner[ak][leaf]))
    return out


def run(cfg, steps, multihead, perturb_ln=False):
    st = P.init_state(cfg, seed=3)
    if perturb_ln:   # non-trivial LayerNorm scales / biases
        g = torch.Generator().manual_seed(11)
        for net in (st.policy, st.vf):
            for k in range(cfg.net.depth):
                net[f"LayerNorm_{k}"]["scale"] = 1.0 + 0.3 * torch.randn(cfg.net.width, generator=g)
                net[f"LayerNorm_{k}"]["bias"] = 0.2 * torch.randn(cfg.net.width, generator=g)
    agent = make_ppo(cfg, steps if multihead else steps // 5, multihead, seed=3)
    for o, ts in ((st.policy, agent.policy.params), (st.vf, agent.value_function.params)):
        for _, src, dst in pairs(o, agent, ts):
            dst.copy_(src.float())
    agent.refresh()
    from mtrl_b200.types import Rollout

    st64 = st.to(torch.float64)
    tcfg = dataclasses.replace(cfg, net=dataclasses.replace(cfg.net, matmul_operands="tf32"))
    for step in range(2):
        r, eps = P.synthetic_rollout(cfg, steps, seed=20 + step)
        r64 = tuple(x.double() for x in r)
        if step == 0:
            _, _, tg = P.ppo_update(st64, r64, eps.double(), tcfg, return_grads=True)
        st64, logs64, g = P.ppo_update(st64, r64, eps.double(), cfg, return_grads=True)
        roll = Rollout(observations=r[0].cuda(), actions=None, rewards=None, dones=None, log_probs=r[1].cuda(),
                       advantages=r[2].cuda(), returns=r[3].cuda(), values=r[4].cuda())
        _, logs = agent.update(roll, eps=eps.cuda())
        for k in P.PPO_LOG_KEYS:
            ref, got = float(logs64[k]), float(logs[k])
            assert abs(got - ref) <= 1e-3 * (1 + step) * abs(ref) + 2e-5, f"step {step} {k}: {got} vs {ref}"
        if step == 0:
            for name, ts in (("policy", agent.policy.grads), ("vf", agent.value_function.grads)):
                for leaf, o, a in pairs(tg[name], agent, ts):
                    if float(o.norm()) > 0:
                        assert SU.rel(a, o) <= 1e-2, f"grad vs tf32-operand oracle {name}/{leaf}: {SU.rel(a, o)}"
                    else:
                        assert float(a.abs().max()) == 0.0, f"{name}/{leaf} must have zero gradient"
        for name, new_t, ts in (("policy", st64.policy, agent.policy.params), ("vf", st64.vf, agent.value_function.params)):
            ps = pairs(new_t, agent, ts)
            fa = torch.cat([a.detach().double().flatten().cpu() for _, _, a in ps])
            fo = torch.cat([o.flatten() for _, o, _ in ps])
            assert float((fa - fo).norm() / fo.norm()) <= 1e-3, f"{name} parameters"
    assert int(agent.policy.step) == 2 and int(agent.value_function.step) == 2


def test_ppo_multihead_small(cuda):
    cfg = P.PPOConfig(net=O.OracleConfig(num_tasks=5, obs_dim=20 + 5, action_dim=4, width=96))
    run(cfg, steps=70, multihead=True)   # 70 rows per task: padded to 128-row tiles


def test_ppo_multihead_mt10_w256_no_clip_vf_no_norm(cuda):
    cfg = P.PPOConfig(net=O.OracleConfig(num_tasks=10, obs_dim=49, action_dim=4, width=256), clip_vf_loss=False,
                      normalize_advantages=False)
    run(cfg, steps=256, multihead=True)


def test_ppo_vanilla_mlp(cuda):
    cfg = P.PPOConfig(net=O.OracleConfig(num_tasks=1, obs_dim=49, action_dim=4, width=128, depth=2))
    run(cfg, steps=640, multihead=False)


@pytest.mark.parametrize("ln,skip", [(True, False), (False, True), (True, True)])
def test_ppo_vanilla_mlp_layer_norm_and_skip_connections(cuda, ln, skip):
    """MT-PPO on VanillaNetworkConfig(use_layer_norm / use_skip_connections) (mtrl/nn/base.py:32-63): the junction kernels of
    csrc/ln_kernels.cuh in the PPO update, LayerNorm scale / bias gradients included."""
    cfg = P.PPOConfig(net=O.OracleConfig(num_tasks=1, obs_dim=49, action_dim=4, width=128, depth=3, use_layer_norm=ln,
                                         use_skip_connections=skip))
    run(cfg, steps=640, multihead=False, perturb_ln=ln)


def test_ppo_requires_unsquashed_policy(cuda):
    from mtrl_b200.config.networks import ContinuousActionPolicyConfig, ValueFunctionConfig
    from mtrl_b200.config.nn import MultiHeadConfig
    from mtrl_b200.rl.algorithms import MTPPO, MTPPOConfig

    net = MultiHeadConfig(width=64, num_tasks=3)
    with pytest.raises(ValueError):
        MTPPO.initialize(MTPPOConfig(num_tasks=3, policy_config=ContinuousActionPolicyConfig(network_config=net),
                                     vf_config=ValueFunctionConfig(network_config=net)), SU.EnvSpec(20, 4), rollout_steps=64)


def test_ppo_config5_shape_runs(cuda):
    """BASELINE configs[4]: MT50, width 4096, large batch.  50 x 2048 rows here (the full 50 x 10 000 needs ~90 GB of
    activations); checks the large-M grouped GEMM path end to end: finite logs, mean head untouched, parameters move."""
    cfg = P.PPOConfig(net=O.OracleConfig(num_tasks=50, obs_dim=89, action_dim=4, width=4096))
    agent = make_ppo(cfg, 2048, True, seed=1)
    from mtrl_b200.types import Rollout

    r, eps = P.synthetic_rollout(cfg, 2048, seed=5)
    before = agent._flat["policy_params"].clone()
    roll = Rollout(observations=r[0].cuda(), actions=None, rewards=None, dones=None, log_probs=r[1].cuda(),
                   advantages=r[2].cuda(), returns=r[3].cuda(), values=r[4].cuda())
    _, logs = agent.update(roll, eps=eps.cuda())
    vals = torch.stack([logs[k] for k in P.PPO_LOG_KEYS])
    assert torch.isfinite(vals).all()
    assert not torch.equal(before, agent._flat["policy_params"])
    hk = agent._inner(agent.policy.grads)["VmapDense_0"]["kernel"]
    assert float(hk[..., :4].abs().max()) == 0.0 and float(hk[..., 4:].abs().max()) > 0.0
