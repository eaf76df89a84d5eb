"""Single-task SAC (sac.py:262-383) on the fused CUDA path vs the fp64 oracle, and the single-task ReplayBuffer
mirror against the reference's own two tests (tests/test_rl_buffers.py:21-62)."""
import dataclasses

import numpy as np
import pytest
import torch

import sac_util as SU
from oracle import mtsac_oracle as O
from oracle import sac_oracle as S

pytestmark = pytest.mark.gpu


def make_sac(cfg, max_batch, seed=1, **kw):
    from mtrl_b200.config.networks import ContinuousActionPolicyConfig, QValueFunctionConfig
    from mtrl_b200.config.nn import VanillaNetworkConfig
    from mtrl_b200.config.optim import OptimizerConfig
    from mtrl_b200.rl.algorithms import SAC, SACConfig

    opt = OptimizerConfig(lr=cfg.lr, max_grad_norm=cfg.max_grad_norm, eps=cfg.adam_eps)
    net = VanillaNetworkConfig(width=cfg.width, depth=cfg.depth, optimizer=opt, use_layer_norm=cfg.use_layer_norm,
                               use_skip_connections=cfg.use_skip_connections)
    sc = SACConfig(num_tasks=10, gamma=cfg.gamma, actor_config=ContinuousActionPolicyConfig(network_config=net),
                   critic_config=QValueFunctionConfig(network_config=net), num_critics=cfg.num_critics, tau=cfg.tau,
                   initial_temperature=cfg.initial_temperature)
    return SAC.initialize(sc, SU.EnvSpec(cfg.obs_dim, cfg.action_dim), seed=seed, max_batch=max_batch, **kw)


def mlp_pairs(otree, atree, ens):
    p = atree["params"]
    p = (p["VmapQValueFunction_0"] if ens else p)["VanillaNetwork_0"]["MLP_0"]
    return [(f"{k}/{leaf}", otree[k][leaf], p[k][leaf]) for k in otree for leaf in otree[k]]


def load(agent, st):
    for o, a, ens in ((st.actor, agent.actor.params, False), (st.critic, agent.critic.params, True),
                      (st.critic_target, agent.critic.target_params, True)):
        for _, src, dst in mlp_pairs(o, a, ens):
            dst.copy_(src.float())
    agent.alpha.params["params"]["log_alpha"].copy_(st.log_alpha.float())
    agent.refresh()


@pytest.mark.parametrize("width,batch", [(64, 96), (400, 1280)])
def test_sac_update_matches_oracle(cuda, width, batch):
    """(400, 1280) is experiments/baselines/mt10_sac_v2.py: MT10 observations (49 incl. one-hot) through a plain MLP."""
    cfg = O.OracleConfig(num_tasks=1, obs_dim=49, action_dim=4, width=width, initial_temperature=0.8)
    st = S.init_state(cfg, seed=2)
    agent = make_sac(cfg, batch, seed=2)
    load(agent, st)
    st64 = st.to(torch.float64)
    tcfg = dataclasses.replace(cfg, matmul_operands="tf32")
    for step in range(2):
        b, ec, ea = S.synthetic_batch(cfg, batch, seed=10 + step)
        b64 = tuple(x.double() for x in b)
        if step == 0:
            _, _, tgr = S.sac_update(st64, b64, ec.double(), ea.double(), tcfg, return_grads=True)
        old = st64
        st64, logs64, gr = S.sac_update(st64, b64, ec.double(), ea.double(), cfg, return_grads=True)
        _, logs = agent.update(tuple(x.cuda() for x in b), eps_c=ec.cuda(), eps_a=ea.cuda(), check=True)
        assert set(logs) == set(S.SAC_LOG_KEYS)
        for k in S.SAC_LOG_KEYS:
            ref, got = float(logs64[k]), float(logs[k])
            assert abs(got - ref) <= 1e-3 * (1 + step) * abs(ref) + 1e-5, f"step {step} {k}: {got} vs {ref}"
        if step == 0:
            for name, tree, ens in (("actor", agent.actor.grads, False), ("critic", agent.critic.grads, True)):
                for leaf, o, a in mlp_pairs(tgr[name], tree, ens):
                    assert SU.rel(a, o) <= 1e-2, f"grad vs tf32-operand oracle {name}/{leaf}: {SU.rel(a, o)}"
                for leaf, o, a in mlp_pairs(gr[name], tree, ens):
                    assert SU.rel(a, o) <= 8e-2, f"grad vs exact oracle {name}/{leaf}: {SU.rel(a, o)}"
        for name, new_t, tree, ens in (("actor", st64.actor, agent.actor.params, False), ("critic", st64.critic, agent.critic.params, True),
                                       ("target", st64.critic_target, agent.critic.target_params, True)):
            pairs = mlp_pairs(new_t, tree, ens)
            fa = torch.cat([a.detach().double().flatten().cpu() for _, _, a in pairs])
            fo = torch.cat([o.flatten() for _, o, _ in pairs])
            assert float((fa - fo).norm() / fo.norm()) <= 1e-3, f"{name} parameters"
        la = agent.alpha.params["params"]["log_alpha"]
        assert SU.rel(la - old.log_alpha.cuda().float(), st64.log_alpha - old.log_alpha) <= 1e-2
    assert agent.get_num_params()["actor_num_params"] == O.num_params(st.actor)


@pytest.mark.parametrize("ln,skip", [(True, False), (False, True), (True, True)])
@pytest.mark.parametrize("width,batch,depth", [(64, 96, 3), (256, 640, 4)])
def test_sac_update_with_layer_norm_and_skip_connections(cuda, ln, skip, width, batch, depth):
    """MLP(use_layer_norm, use_skip_connections) (mtrl/nn/base.py:32-63; flax LayerNorm eps 1e-6, scale + bias): the
    junction kernels (csrc/ln_kernels.cuh) between the Dense GEMMs, in the parity precision: three updates, every leaf --
    LayerNorm scales and biases included -- and every log scalar to 1e-3 of the fp64 oracle."""
    cfg = O.OracleConfig(num_tasks=1, obs_dim=49, action_dim=4, width=width, depth=depth, initial_temperature=0.8, use_layer_norm=ln,
                         use_skip_connections=skip)
    st = S.init_state(cfg, seed=3)
    if ln:   # non-trivial scales / biases, different per member
        g = torch.Generator().manual_seed(5)
        for net in (st.actor, st.critic):
            for k in range(depth):
                net[f"LayerNorm_{k}"]["scale"] = 1.0 + 0.3 * torch.randn(net[f"LayerNorm_{k}"]["scale"].shape, generator=g)
                net[f"LayerNorm_{k}"]["bias"] = 0.2 * torch.randn(net[f"LayerNorm_{k}"]["bias"].shape, generator=g)
        st.critic_target = O.tree_map(lambda x: x.clone(), st.critic)
    agent = make_sac(cfg, batch, seed=3, precision="fp32x3")
    load(agent, st)
    assert agent.get_num_params()["actor_num_params"] == O.num_params(st.actor)
    assert agent.get_num_params()["critic_num_params"] == O.num_params(st.critic)
    st64 = st.to(torch.float64)
    for step in range(3):
        b, ec, ea = S.synthetic_batch(cfg, batch, seed=20 + step)
        st64, logs64, gr = S.sac_update(st64, tuple(x.double() for x in b), ec.double(), ea.double(), cfg, return_grads=True)
        _, logs = agent.update(tuple(x.cuda() for x in b), eps_c=ec.cuda(), eps_a=ea.cuda(), check=True)
        for k in S.SAC_LOG_KEYS:
            ref, got = float(logs64[k]), float(logs[k])
            assert abs(got - ref) <= 1e-3 * abs(ref) + 1e-7, f"step {step} {k}: {got} vs {ref}"
        if step == 0:
            for name, tree, ens in (("actor", agent.actor.grads, False), ("critic", agent.critic.grads, True)):
                for leaf, o, a in mlp_pairs(gr[name], tree, ens):
                    assert SU.rel(a, o) <= 1e-3, f"grad {name}/{leaf}: {SU.rel(a, o)}"
        for name, new_t, tree, ens in (("actor", st64.actor, agent.actor.params, False), ("critic", st64.critic, agent.critic.params, True),
                                       ("target", st64.critic_target, agent.critic.target_params, True)):
            for leaf, o, a in mlp_pairs(new_t, tree, ens):
                assert SU.rel(a, o) <= 1e-3, f"step {step} param {name}/{leaf}: {SU.rel(a, o)}"
    # the policy (action sampling) goes through the same junctions
    obs = S.synthetic_batch(cfg, 8, seed=99)[0][0]
    got = agent.eval_action(obs.numpy())
    ref = torch.tanh(S.mlp_forward(st64.actor, obs.double(), depth, "exact", ln, skip)[:, :4])
    assert SU.rel(torch.from_numpy(got), ref) <= 1e-4


def test_sac_layer_norm_tf32_precision(cuda):
    """The same path with tf32 operands: log scalars to 1e-3, networks to 1e-3 (relative l2)."""
    cfg = O.OracleConfig(num_tasks=1, obs_dim=49, action_dim=4, width=256, initial_temperature=0.8, use_layer_norm=True,
                         use_skip_connections=True)
    st = S.init_state(cfg, seed=4)
    agent = make_sac(cfg, 640, seed=4)
    load(agent, st)
    st64 = st.to(torch.float64)
    b, ec, ea = S.synthetic_batch(cfg, 640, seed=30)
    st64, logs64 = S.sac_update(st64, tuple(x.double() for x in b), ec.double(), ea.double(), cfg)
    _, logs = agent.update(tuple(x.cuda() for x in b), eps_c=ec.cuda(), eps_a=ea.cuda(), check=True)
    for k in S.SAC_LOG_KEYS:
        ref, got = float(logs64[k]), float(logs[k])
        assert abs(got - ref) <= 1e-3 * abs(ref) + 1e-5, f"{k}: {got} vs {ref}"
    for name, new_t, tree, ens in (("actor", st64.actor, agent.actor.params, False), ("critic", st64.critic, agent.critic.params, True)):
        pairs = mlp_pairs(new_t, tree, ens)
        fa = torch.cat([a.detach().double().flatten().cpu() for _, _, a in pairs])
        fo = torch.cat([o.flatten() for _, o, _ in pairs])
        assert float((fa - fo).norm() / fo.norm()) <= 1e-3, name


def _buffer(capacity):
    from mtrl_b200.rl.buffers import ReplayBuffer

    return ReplayBuffer(capacity=capacity, env_obs_space=SU._Space((3,)), env_action_space=SU._Space((2,)), seed=0)


def test_replay_buffer_sets_full_flag_after_single_transitions(cuda):
    """/root/reference/tests/test_rl_buffers.py:21-35."""
    buffer = _buffer(4)
    for idx in range(buffer.capacity):
        value = float(idx)
        buffer.add(obs=np.full((3,), value, dtype=np.float32), next_obs=np.full((3,), value + 1.0, dtype=np.float32),
                   action=np.full((2,), -value, dtype=np.float32), reward=np.array([value], dtype=np.float32),
                   done=np.array([idx % 2], dtype=np.float32))
    assert buffer.full is True
    assert buffer.pos == 0
    assert torch.equal(buffer.obs[:, 0].cpu(), torch.arange(4.0))


def test_replay_buffer_sets_full_flag_after_batched_transitions(cuda):
    """/root/reference/tests/test_rl_buffers.py:38-62."""
    buffer = _buffer(5)
    n = buffer.capacity
    obs = np.arange(n * 3, dtype=np.float32).reshape(n, 3)
    next_obs = obs + 1.0
    action = np.arange(n * 2, dtype=np.float32).reshape(n, 2)
    reward = np.arange(n, dtype=np.float32)
    done = np.zeros(n, dtype=np.float32)
    buffer.add(obs=obs, next_obs=next_obs, action=action, reward=reward, done=done)
    assert buffer.full is True
    assert buffer.pos == 0
    buffer.add(obs=obs[:2], next_obs=next_obs[:2], action=action[:2], reward=reward[:2], done=done[:2])
    assert buffer.full is True
    assert buffer.pos == 2
    s = buffer.sample(4)
    g = np.random.default_rng(0)
    idx = g.integers(0, 5, size=(4,))
    assert np.array_equal(s.observations.cpu().numpy(), buffer.obs.cpu().numpy()[idx])
    assert s.rewards.shape == (4, 1) and s.actions.shape == (4, 2)
