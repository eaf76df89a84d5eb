"""Oracle for the single-task SAC update: PyTorch-CPU restatement of `SAC._update_inner`
(/root/reference/mtrl/rl/algorithms/sac.py:262-383) on `VanillaNetwork` / `MLP`
(/root/reference/mtrl/nn/base.py:11-89), including its optional pre-layer LayerNorm and residual connections
(`VanillaNetworkConfig.use_layer_norm / use_skip_connections`, off in every SAC experiment of the reference, e.g.
experiments/baselines/mt10_sac_v2.py:36-50, but part of the MLP).

TEST INFRASTRUCTURE (see oracle/__init__.py).  Parity status: PARITY UNPINNED, for the same reasons as
oracle/mtsac_oracle.py (no reference test or golden vector for the update; JAX stack absent).
"""
from __future__ import annotations

import math

import torch

from .mtsac_oracle import (OracleConfig, OracleState, _grads_of, _rt, _with_grad, adam_step, global_norm, tree_map)

SAC_LOG_KEYS = (
    "losses/alpha_loss", "alpha", "losses/qf_values", "losses/qf_loss", "metrics/critic_grad_magnitude",
    "metrics/actor_grad_magnitude", "metrics/actor_params_norm", "metrics/critic_params_norm", "losses/actor_loss",
)  # sac.py:300-304, 326-329, 359-364, 383


def init_mlp(gen: torch.Generator, in_dim: int, cfg: OracleConfig, head_dim: int, head_bound: float,
             ensemble: int | None = None, dtype=torch.float32) -> dict:
    """MLP parameters (nn/base.py:32-63): layer_0..layer_{depth-1} hidden, layer_{depth} the output Dense.
    he_uniform / zero bias from VanillaNetworkConfig (config/nn.py:15-19), output layer uniform(+-bound)
    (networks.py:33-34, 65-66)."""
    lead = () if ensemble is None else (ensemble,)

    def u(shape, bound):
        return ((torch.rand(*shape, generator=gen, dtype=torch.float64) * 2 - 1) * bound).to(dtype)

    p = {}
    d = in_dim
    for i in range(cfg.depth):
        p[f"layer_{i}"] = {"kernel": u(lead + (d, cfg.width), math.sqrt(6.0 / d)), "bias": torch.zeros(lead + (cfg.width,), dtype=dtype)}
        d = cfg.width
    p[f"layer_{cfg.depth}"] = {"kernel": u(lead + (cfg.width, head_dim), head_bound), "bias": u(lead + (head_dim,), head_bound)}
    if cfg.use_layer_norm:   # nn.LayerNorm() before layers 1..depth-1 and before the output Dense (base.py:35-37, 52-53):
        for k in range(cfg.depth):   # Flax names them LayerNorm_0.. in creation order; scale = ones, bias = zeros
            p[f"LayerNorm_{k}"] = {"scale": torch.ones(lead + (cfg.width,), dtype=dtype), "bias": torch.zeros(lead + (cfg.width,), dtype=dtype)}
    return p


def init_state(cfg: OracleConfig, seed: int = 1, dtype=torch.float32) -> OracleState:
    """SAC.initialize (sac.py:120-200): scalar temperature log_alpha of shape (1,) (sac.py:46-56)."""
    gen = torch.Generator().manual_seed(seed)
    actor = init_mlp(gen, cfg.obs_dim, cfg, 2 * cfg.action_dim, 1e-3, None, dtype)
    critic = init_mlp(gen, cfg.action_dim + cfg.obs_dim, cfg, 1, 3e-3, cfg.num_critics, dtype)
    log_alpha = torch.full((1,), math.log(cfg.initial_temperature), dtype=dtype)
    zeros = lambda t: tree_map(torch.zeros_like, t)  # noqa: E731
    opt = {"actor": {"m": zeros(actor), "v": zeros(actor), "count": 0},
           "critic": {"m": zeros(critic), "v": zeros(critic), "count": 0},
           "alpha": {"m": torch.zeros_like(log_alpha), "v": torch.zeros_like(log_alpha), "count": 0}}
    return OracleState(actor, critic, tree_map(lambda x: x.clone(), critic), log_alpha, opt)


def layer_norm(x: torch.Tensor, scale: torch.Tensor, bias: torch.Tensor, eps: float = 1e-6) -> torch.Tensor:
    """flax.linen.LayerNorm defaults (flax 0.10.4): over the last axis, epsilon 1e-6, use_fast_variance=True, i.e.
    var = max(E[x^2] - E[x]^2, 0); y = (x - mean) * rsqrt(var + eps) * scale + bias."""
    mean = x.mean(dim=-1, keepdim=True)
    var = torch.clamp((x * x).mean(dim=-1, keepdim=True) - mean * mean, min=0.0)
    return (x - mean) * torch.rsqrt(var + eps) * scale + bias


def mlp_forward(p: dict, x: torch.Tensor, depth: int, operands: str = "exact", use_layer_norm: bool = False,
                use_skip_connections: bool = False) -> torch.Tensor:
    """MLP.__call__ (mtrl/nn/base.py:32-63).  With use_layer_norm a LayerNorm precedes every Dense except the first
    (:35-37, 52-53); with use_skip_connections a layer whose input already has the hidden width adds that (normalised)
    input to its activation (:45-48)."""
    width = p["layer_0"]["kernel"].shape[-1]
    h = x
    for i in range(depth):
        if use_layer_norm and i != 0:
            h = layer_norm(h, p[f"LayerNorm_{i - 1}"]["scale"], p[f"LayerNorm_{i - 1}"]["bias"])
        h = _rt(h, operands)
        d = _rt(torch.relu(h @ _rt(p[f"layer_{i}"]["kernel"], operands) + p[f"layer_{i}"]["bias"]), operands)
        h = h + d if (use_skip_connections and h.shape[-1] == width) else d
    if use_layer_norm and depth != 0:
        h = layer_norm(h, p[f"LayerNorm_{depth - 1}"]["scale"], p[f"LayerNorm_{depth - 1}"]["bias"])
    return h @ p[f"layer_{depth}"]["kernel"] + p[f"layer_{depth}"]["bias"]


def critic_forward(p: dict, obs, act, cfg: OracleConfig) -> torch.Tensor:
    x = torch.cat((act, obs), dim=-1)  # networks.py:61
    E = p["layer_0"]["kernel"].shape[0]
    return torch.stack([mlp_forward(tree_map(lambda t: t[e], p), x, cfg.depth, cfg.matmul_operands, cfg.use_layer_norm,
                                    cfg.use_skip_connections) for e in range(E)], 0)


def actor_sample_and_log_prob(p: dict, obs, eps, cfg: OracleConfig):
    out = mlp_forward(p, obs, cfg.depth, cfg.matmul_operands, cfg.use_layer_norm, cfg.use_skip_connections)
    mean, log_std = out[..., : cfg.action_dim], out[..., cfg.action_dim:]
    log_std = torch.clamp(log_std, cfg.log_std_min, cfg.log_std_max)
    std = torch.exp(log_std)
    x = mean + std * eps
    base_lp = (-0.5 * eps**2 - 0.5 * math.log(2 * math.pi) - log_std).sum(-1)
    fldj = (2.0 * (math.log(2.0) - x - torch.nn.functional.softplus(-2.0 * x))).sum(-1)
    return torch.tanh(x), base_lp - fldj


def sac_update(state: OracleState, batch, eps_c, eps_a, cfg: OracleConfig, return_grads: bool = False):
    """One `SAC.update` (sac.py:262-386): alpha first, critic with the new alpha, actor with both new."""
    obs, actions, next_obs, dones, rewards = batch
    logs, grads_out = {}, {}
    opt = dict(state.opt)
    # actor sample (differentiated in the actor loss below)                                        sac.py:335-337
    ap = _with_grad(state.actor)
    a_samples, logp = actor_sample_and_log_prob(ap, obs, eps_a, cfg)
    logp_col = logp.reshape(-1, 1)
    # ---- alpha (sac.py:308-331) ----
    la = state.log_alpha.detach().clone().requires_grad_(True)
    alpha_loss = (-la * (logp_col.detach() + cfg.target_entropy)).mean()
    alpha_loss.backward()
    new_la, opt["alpha"] = adam_step(state.log_alpha, la.grad, opt["alpha"], cfg.alpha_lr, cfg.adam_eps, cfg.b1, cfg.b2,
                                     cfg.alpha_max_grad_norm)
    alpha_val = torch.exp(new_la).detach()                                                          # sac.py:324, 342
    logs["losses/alpha_loss"] = alpha_loss.detach()
    logs["alpha"] = torch.exp(new_la).sum()
    # ---- critic (sac.py:267-304) ----
    with torch.no_grad():
        na, nlp = actor_sample_and_log_prob(state.actor, next_obs, eps_c, cfg)
        qt = critic_forward(state.critic_target, next_obs, na, cfg)
        y = rewards + (1 - dones) * cfg.gamma * (qt.min(0).values - alpha_val * nlp.reshape(-1, 1))
    cp = _with_grad(state.critic)
    q_pred = critic_forward(cp, obs, actions, cfg)
    critic_loss = 0.5 * ((q_pred - y) ** 2).mean(dim=1).sum()                                       # sac.py:292
    critic_loss.backward()
    cgrads = _grads_of(cp)
    new_critic, opt["critic"] = adam_step(state.critic, cgrads, opt["critic"], cfg.lr, cfg.adam_eps, cfg.b1, cfg.b2,
                                          cfg.max_grad_norm)
    logs["losses/qf_values"] = q_pred.mean().detach()
    logs["losses/qf_loss"] = critic_loss.detach()
    logs["metrics/critic_grad_magnitude"] = global_norm(cgrads)
    # ---- actor (sac.py:344-356): new critic, new alpha ----
    q_pi = critic_forward(new_critic, obs, a_samples, cfg)
    actor_loss = (alpha_val * logp_col - q_pi.min(0).values).mean()
    actor_loss.backward()
    agrads = _grads_of(ap)
    new_actor, opt["actor"] = adam_step(state.actor, agrads, opt["actor"], cfg.lr, cfg.adam_eps, cfg.b1, cfg.b2,
                                        cfg.max_grad_norm)
    logs["metrics/actor_grad_magnitude"] = global_norm(agrads)
    logs["metrics/actor_params_norm"] = global_norm(state.actor)      # sac.py:360-361: self.actor.params, i.e. PRE-update
    logs["metrics/critic_params_norm"] = global_norm(state.critic)    # sac.py:363-364
    logs["losses/actor_loss"] = actor_loss.detach()
    new_target = tree_map(lambda n, t: cfg.tau * n + (1 - cfg.tau) * t, new_critic, state.critic_target)  # sac.py:367-374
    new_state = OracleState(new_actor, new_critic, new_target, new_la, opt)
    grads_out = {"actor": agrads, "critic": cgrads, "alpha": la.grad}
    if return_grads:
        return new_state, logs, grads_out
    return new_state, logs


def synthetic_batch(cfg: OracleConfig, batch: int, seed: int = 1234, dtype=torch.float32):
    g = torch.Generator().manual_seed(seed)
    o = torch.randn(batch, cfg.obs_dim, generator=g, dtype=torch.float64)
    nxt = o + 0.01 * torch.randn(batch, cfg.obs_dim, generator=g, dtype=torch.float64)
    act = torch.rand(batch, cfg.action_dim, generator=g, dtype=torch.float64) * 2 - 1
    rew = torch.rand(batch, 1, generator=g, dtype=torch.float64) * 10
    done = (torch.rand(batch, 1, generator=g, dtype=torch.float64) < 0.002).to(torch.float64)
    ec = torch.randn(batch, cfg.action_dim, generator=g, dtype=torch.float64)
    ea = torch.randn(batch, cfg.action_dim, generator=g, dtype=torch.float64)
    cast = lambda t: t.to(torch.float32).to(dtype)  # noqa: E731
    return tuple(cast(t) for t in (o, act, nxt, done, rew)), cast(ec), cast(ea)
