"""Glue between the update oracle (oracle/mtsac_oracle.py) and the CUDA MTSAC mirror, for tests,
smoke() and bench.py's checker legs."""
from __future__ import annotations

import torch

from oracle import mtsac_oracle as O


class _Space:
    def __init__(self, shape):
        self.shape = shape


class EnvSpec:
    """Minimal stand-in for mtrl.envs.EnvConfig: only the two spaces' shapes are read (mtsac.py:157-196)."""

    def __init__(self, obs_dim: int, action_dim: int):
        self.observation_space = _Space((obs_dim,))
        self.action_space = _Space((action_dim,))


def make_agent(cfg: O.OracleConfig, per_task: int, seed: int = 1, **kw):
    from mtrl_b200.config.networks import ContinuousActionPolicyConfig, QValueFunctionConfig
    from mtrl_b200.config.nn import MultiHeadConfig
    from mtrl_b200.config.optim import OptimizerConfig
    from mtrl_b200.rl.algorithms import MTSAC, MTSACConfig

    opt = OptimizerConfig(lr=cfg.lr, max_grad_norm=cfg.max_grad_norm, eps=cfg.adam_eps)
    net = MultiHeadConfig(width=cfg.width, depth=cfg.depth, num_tasks=cfg.num_tasks, optimizer=opt)
    mc = MTSACConfig(
        num_tasks=cfg.num_tasks, gamma=cfg.gamma, clip=cfg.clip,
        actor_config=ContinuousActionPolicyConfig(network_config=net, log_std_min=cfg.log_std_min, log_std_max=cfg.log_std_max),
        critic_config=QValueFunctionConfig(network_config=net),
        temperature_optimizer_config=OptimizerConfig(lr=cfg.alpha_lr, max_grad_norm=cfg.alpha_max_grad_norm, eps=cfg.adam_eps),
        initial_temperature=cfg.initial_temperature, num_critics=cfg.num_critics, tau=cfg.tau,
        use_task_weights=cfg.use_task_weights)
    kw.setdefault("max_batch", per_task * cfg.num_tasks)
    return MTSAC.initialize(mc, EnvSpec(cfg.obs_dim, cfg.action_dim), seed=seed, **kw)


def _net(tree, ensemble):
    p = tree["params"]
    return (p["VmapQValueFunction_0"] if ensemble else p)["MultiHeadNetwork_0"]


def _pairs(oracle_tree, agent_tree):
    """(oracle leaf, agent view) pairs; oracle 'heads' <-> Flax 'VmapDense_0'."""
    out = []
    for k, v in oracle_tree.items():
        ak = "VmapDense_0" if k == "heads" else k
        for leaf in ("kernel", "bias"):
            out.append((f"{ak}/{leaf}", v[leaf], agent_tree[ak][leaf]))
    return out


def load_oracle_state(agent, st: O.OracleState, task_slice=None) -> None:
    """Copy an oracle state (params, target, Adam moments, counts) into the agent's device buffers."""
    def put(otree, atree, ensemble):
        for name, o, a in _pairs(otree, _net(atree, ensemble)):
            src = o
            if task_slice is not None and name.startswith("VmapDense_0"):
                src = o[:, task_slice] if ensemble else o[task_slice]
            a.copy_(src.to(torch.float32))
    put(st.actor, agent.actor.params, False)
    put(st.critic, agent.critic.params, True)
    put(st.critic_target, agent.critic.target_params, True)
    put(st.opt["actor"]["m"], agent.actor.opt_state["mu"], False)
    put(st.opt["actor"]["v"], agent.actor.opt_state["nu"], False)
    put(st.opt["critic"]["m"], agent.critic.opt_state["mu"], True)
    put(st.opt["critic"]["v"], agent.critic.opt_state["nu"], True)
    sl = task_slice if task_slice is not None else slice(None)
    agent.alpha.params["params"]["log_alpha"].copy_(st.log_alpha[sl].float())
    agent.alpha.opt_state["mu"]["params"]["log_alpha"].copy_(st.opt["alpha"]["m"][sl].float())
    agent.alpha.opt_state["nu"]["params"]["log_alpha"].copy_(st.opt["alpha"]["v"][sl].float())
    agent._steps[0] = st.opt["actor"]["count"]
    agent._steps[1] = st.opt["critic"]["count"]
    agent._steps[2] = st.opt["alpha"]["count"]
    agent.refresh()


def rel(a: torch.Tensor, b: torch.Tensor) -> float:
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-300))


def compare_trees(oracle_tree, agent_tree, ensemble, task_slice=None):
    """{leaf name: relative l2 error} of agent views against oracle leaves."""
    out = {}
    for name, o, a in _pairs(oracle_tree, _net(agent_tree, ensemble)):
        if task_slice is not None and name.startswith("VmapDense_0"):
            o = o[:, task_slice] if ensemble else o[task_slice]
        out[name] = rel(a, o)
    return out


def compare_deltas(old_tree, new_tree, agent_tree, ensemble):
    """Relative error of the parameter UPDATE (new - old), the part the kernels actually compute."""
    out = {}
    for (name, o_old, a), (_, o_new, _) in zip(_pairs(old_tree, _net(agent_tree, ensemble)),
                                               _pairs(new_tree, _net(agent_tree, ensemble))):
        d_ref = (o_new - o_old).double().cpu()
        d_gpu = a.detach().double().cpu() - o_old.double().cpu()
        out[name] = float((d_gpu - d_ref).norm() / d_ref.norm().clamp_min(1e-300))
    return out
