"""TEST INFRASTRUCTURE ONLY (never imported by the product path).  PARITY UNPINNED: like the update itself, the
reference's per-task gradient analysis cannot run here (no jax) and has no test or golden vector.

CPU restatement of the first half of `MTSAC.compute_weights` (/root/reference/mtrl/rl/algorithms/mtsac.py:870-1170,
MSE / vanilla branch): the batch is split by task (`split_data_by_tasks`, :313-327), the critic loss (:1009-1026) and
the actor loss (:1049-1069) are differentiated per task with `jax.vmap(jax.value_and_grad(...))`, the per-task gradient
pytrees are flattened to (num_tasks, num_params) (:1039-1041, :1077-1079) and summarised by `vmap_cos_sim` and
`compute_conflict_metrics` (mtrl/rl/algorithms/utils.py:49-174).  Plain PyTorch autograd, one task at a time.
"""
from __future__ import annotations

import torch

from . import mtsac_oracle as O


def per_task_grads(state: O.OracleState, batch, eps_c, eps_a, cfg: O.OracleConfig):
    """({'critic': [tree per task], 'actor': [tree per task]}): gradients of the per-task mean losses."""
    obs, actions, next_obs, dones, rewards = batch
    T = cfg.num_tasks
    task = obs[..., -T:].argmax(dim=-1)
    alpha_vals = torch.exp(obs[..., -T:] @ state.log_alpha.reshape(-1, 1))
    with torch.no_grad():   # mtsac.py:995-1006: target computed once from the current actor and target critic
        next_actions, next_logp = O.actor_sample_and_log_prob(state.actor, next_obs, eps_c, cfg)
        q_t = O.critic_forward(state.critic_target, next_obs, next_actions, cfg)
        target = rewards + (1 - dones) * cfg.gamma * (q_t.min(dim=0).values - alpha_vals * next_logp.reshape(-1, 1))
        if cfg.clip:
            target = torch.clamp(target, -5000, 5000)
    out = {"critic": [], "actor": []}
    for t in range(T):
        rows = task == t
        cp = O._with_grad(state.critic)
        q_pred = O.critic_forward(cp, obs[rows], actions[rows], cfg)
        if cfg.clip:
            q_pred = torch.clamp(q_pred, -5000, 5000)
        ((q_pred - target[rows]) ** 2).mean().backward()                      # :1022-1025 (mean over ensemble and rows)
        out["critic"].append(O._grads_of(cp))
        ap = O._with_grad(state.actor)
        a, logp = O.actor_sample_and_log_prob(ap, obs[rows], eps_a[rows], cfg)
        q_pi = O.critic_forward(state.critic, obs[rows], a, cfg)              # :1060-1062: the CURRENT critic params
        (alpha_vals[rows] * logp.reshape(-1, 1) - q_pi.min(dim=0).values).mean().backward()   # :1067
        out["actor"].append(O._grads_of(ap))
    return out


def flatten(trees) -> torch.Tensor:
    """ravel_pytree per task -> (num_tasks, num_params) (:1039-1041); leaf order is irrelevant to every metric below."""
    return torch.stack([torch.cat([x.flatten() for x in O.tree_leaves(tr)]) for tr in trees])


def vmap_cos_sim(g: torch.Tensor):
    """utils.py:49-72: cos_ij = <g_i, g_j> / (|g_i| |g_j| + 1e-8); average over the strict upper triangle."""
    T = g.shape[0]
    n = g.norm(dim=1)
    cos = (g @ g.T) / (n[:, None] * n[None, :] + 1e-8)
    mask = torch.triu(torch.ones(T, T, dtype=g.dtype), diagonal=1)
    return (mask * cos).sum() / (mask.sum() + 1e-8), cos


def conflict_metrics(cos: torch.Tensor, g: torch.Tensor) -> dict:
    """The Gram-derived part of compute_conflict_metrics (utils.py:118-174)."""
    T = g.shape[0]
    off = 1 - torch.eye(T, dtype=g.dtype)
    conflict = (cos < 0).to(g.dtype)
    n_off = T * (T - 1)
    mag = g.norm(dim=1)
    cm = torch.where((conflict * off).bool(), cos.abs() * (mag[:, None] * mag[None, :]), torch.zeros_like(cos))
    angles = torch.rad2deg(torch.arccos(torch.clamp(cos, -1.0, 1.0)))
    return {"conflict_rate": (conflict * off).sum() / n_off, "mean_conflict_magnitude": (cm * off).sum() / n_off,
            "mean_conflict_angle": (angles * off).sum() / n_off,
            "per_task_conflict_rate": (conflict * off).sum(dim=1) / (T - 1), "per_task_grad_magnitude": mag}
