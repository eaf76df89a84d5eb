"""Oracle for the multi-task PPO update: PyTorch-CPU restatement of `MTPPO._update_inner`
(/root/reference/mtrl/rl/algorithms/mtppo.py:196-317): one clipped-surrogate policy step and one clipped value
step on the whole rollout (no epochs / minibatches -- `num_epochs`, `num_gradient_steps`, `target_kl` in
mtrl/config/rl.py:85-92 are never read by the reference).

TEST INFRASTRUCTURE (see oracle/__init__.py).  Parity status: PARITY UNPINNED (no reference test, JAX stack absent).

Restated literally, including two things a reader should know:
  * `new_log_probs` is the log-prob of a FRESH sample of the current policy (mtppo.py:205-207 calls
    `sample_and_log_prob`), not of `data.actions`.  For the diagonal Gaussian that is
    sum_d(-eps^2/2 - log sigma_d - log(2 pi)/2): it does not depend on the mean, so the surrogate only trains log-std.
  * `action_dist.entropy()` (mtppo.py:232) is only defined by distrax for the un-squashed Gaussian, so the policy is
    `ContinuousActionPolicyConfig(squash_tanh=False)` (networks.py:43-45): entropy = sum_d(log sigma_d + log(2 pi e)/2).
Networks: MultiHeadNetwork (own-task head) when num_tasks > 1, plain MLP semantics when num_tasks == 1; rows are the
rollout flattened task-major, (task, timestep) -> task * steps + timestep (mtrl/types.py:48-63).
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import torch

from .mtsac_oracle import OracleConfig, _grads_of, _with_grad, adam_step, global_norm, init_multihead, multihead_forward, tree_map

PPO_LOG_KEYS = ("losses/entropy_loss", "losses/policy_loss", "losses/approx_kl", "losses/clip_fracs",
                "losses/value_function", "losses/values")  # mtppo.py:234-239, 274-277


@dataclass
class PPOConfig:
    net: OracleConfig                 # widths / lr / clip norm / log-std bounds are read from here
    clip_eps: float = 0.2             # MTPPOConfig, mtppo.py:85-90
    clip_vf_loss: bool = True
    entropy_coefficient: float = 5e-3
    vf_coefficient: float = 0.001
    normalize_advantages: bool = True


@dataclass
class PPOState:
    policy: dict
    vf: dict
    opt: dict

    def to(self, dtype):
        c = lambda t: tree_map(lambda x: x.to(dtype), t)  # noqa: E731
        return PPOState(c(self.policy), c(self.vf), {k: {"m": c(v["m"]), "v": c(v["v"]), "count": v["count"]} for k, v in self.opt.items()})


def init_state(cfg: PPOConfig, seed: int = 1, dtype=torch.float32) -> PPOState:
    """mtppo.py:100-150: policy head uniform(1e-3) (networks.py:33-34), value head uniform(3e-3) (networks.py:199-200)."""
    gen = torch.Generator().manual_seed(seed)
    n = cfg.net
    policy = init_multihead(gen, n.obs_dim, n, 2 * n.action_dim, 1e-3, None, dtype)
    vf = init_multihead(gen, n.obs_dim, n, 1, 3e-3, None, dtype)
    if n.use_layer_norm:   # plain MLP (VanillaNetworkConfig): flax LayerNorm params, scale = ones, bias = zeros
        for net in (policy, vf):
            for k in range(n.depth):
                net[f"LayerNorm_{k}"] = {"scale": torch.ones(n.width, dtype=dtype), "bias": torch.zeros(n.width, dtype=dtype)}
    zeros = lambda t: tree_map(torch.zeros_like, t)  # noqa: E731
    return PPOState(policy, vf, {"policy": {"m": zeros(policy), "v": zeros(policy), "count": 0},
                                 "vf": {"m": zeros(vf), "v": zeros(vf), "count": 0}})


def net_forward(p: dict, obs: torch.Tensor, n: OracleConfig) -> torch.Tensor:
    """MultiHeadNetwork (num_tasks > 1), or the plain MLP of VanillaNetwork for num_tasks == 1 -- then with its optional
    LayerNorm / skip connections (mtrl/nn/base.py:32-63); the MLP's output Dense is the single 'head'."""
    if not (n.use_layer_norm or n.use_skip_connections):
        return multihead_forward(p, obs, n.num_tasks, n.depth, False, n.matmul_operands)
    assert n.num_tasks == 1, "LayerNorm / skip connections belong to the plain MLP"
    from .sac_oracle import mlp_forward

    q = {k: v for k, v in p.items() if k != "heads"}
    q[f"layer_{n.depth}"] = {"kernel": p["heads"]["kernel"][0], "bias": p["heads"]["bias"][0]}
    return mlp_forward(q, obs, n.depth, n.matmul_operands, n.use_layer_norm, n.use_skip_connections)


def ppo_update(state: PPOState, rollout, eps: torch.Tensor, cfg: PPOConfig, return_grads: bool = False):
    """rollout = (observations (B, obs), log_probs (B, 1), advantages (B, 1), returns (B, 1), values (B, 1))."""
    obs, old_logp, adv, returns, old_values = rollout
    n = cfg.net
    logs, opt = {}, dict(state.opt)
    # ---- policy (mtppo.py:196-254) ----
    pp = _with_grad(state.policy)
    out = net_forward(pp, obs, n)
    log_std = torch.clamp(out[..., n.action_dim:], n.log_std_min, n.log_std_max)          # networks.py:37-41
    new_logp = (-0.5 * eps**2 - 0.5 * math.log(2 * math.pi) - log_std).sum(-1)               # fresh sample, :205-207
    log_ratio = new_logp.reshape(-1, 1) - old_logp                                          # :208
    ratio = torch.exp(log_ratio)
    approx_kl = ((ratio - 1) - log_ratio).mean().detach()                                   # :212
    clip_fracs = ((ratio - 1.0).abs() > cfg.clip_eps).to(obs.dtype).mean().detach()         # :213-218
    a = (adv - adv.mean()) / (adv.std(unbiased=False) + 1e-8) if cfg.normalize_advantages else adv   # :220-225
    pg_loss = torch.maximum(-a * ratio, -a * torch.clamp(ratio, 1 - cfg.clip_eps, 1 + cfg.clip_eps)).mean()  # :227-231
    entropy = (log_std + 0.5 * math.log(2 * math.pi * math.e)).sum(-1).mean()               # :232
    (pg_loss - cfg.entropy_coefficient * entropy).backward()
    pgrads = _grads_of(pp)
    new_policy, opt["policy"] = adam_step(state.policy, pgrads, opt["policy"], n.lr, n.adam_eps, n.b1, n.b2, n.max_grad_norm)
    logs.update({"losses/entropy_loss": entropy.detach(), "losses/policy_loss": pg_loss.detach(),
                 "losses/approx_kl": approx_kl, "losses/clip_fracs": clip_fracs})
    # ---- value function (mtppo.py:256-290) ----
    vp = _with_grad(state.vf)
    v = net_forward(vp, obs, n)
    if cfg.clip_vf_loss:
        unclipped = (v - returns) ** 2
        v_clipped = old_values + torch.clamp(v - old_values, -cfg.clip_eps, cfg.clip_eps)
        vf_loss = 0.5 * torch.maximum(unclipped, (v_clipped - returns) ** 2).mean()          # :262-270
    else:
        vf_loss = 0.5 * ((v - returns) ** 2).mean()
    (cfg.vf_coefficient * vf_loss).backward()                                                # :274
    vgrads = _grads_of(vp)
    new_vf, opt["vf"] = adam_step(state.vf, vgrads, opt["vf"], n.lr, n.adam_eps, n.b1, n.b2, n.max_grad_norm)
    logs.update({"losses/value_function": vf_loss.detach(), "losses/values": v.mean().detach()})
    new_state = PPOState(new_policy, new_vf, opt)
    if return_grads:
        return new_state, logs, {"policy": pgrads, "vf": vgrads}
    return new_state, logs


def synthetic_rollout(cfg: PPOConfig, steps: int, seed: int = 1, dtype=torch.float32):
    """Task-major flattened rollout of `steps` timesteps per task (one-hot task id in the observation)."""
    g = torch.Generator().manual_seed(seed)
    n = cfg.net
    T, B = n.num_tasks, n.num_tasks * steps
    feat = n.obs_dim - (T if T > 1 else 0)
    o = torch.randn(B, feat, generator=g, dtype=torch.float64)
    if T > 1:
        task = torch.arange(B) // steps
        o = torch.cat((o, torch.nn.functional.one_hot(task, T).to(torch.float64)), 1)
    eps = torch.randn(B, n.action_dim, generator=g, dtype=torch.float64)
    # old log-probs near the new ones (log_std ~ 0 at init) so that ratios straddle the clip range
    old_logp = (-0.5 * eps**2 - 0.5 * math.log(2 * math.pi)).sum(-1, keepdim=True) + 0.15 * torch.randn(B, 1, generator=g, dtype=torch.float64)
    adv = torch.randn(B, 1, generator=g, dtype=torch.float64) * 3 + 0.5
    ret = torch.randn(B, 1, generator=g, dtype=torch.float64) * 2
    val = ret + 0.3 * torch.randn(B, 1, generator=g, dtype=torch.float64)
    cast = lambda t: t.to(torch.float32).to(dtype)  # noqa: E731
    return tuple(cast(t) for t in (o, old_logp, adv, ret, val)), cast(eps)
