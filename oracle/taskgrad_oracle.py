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


def _targets(state: O.OracleState, batch, eps_c, cfg: O.OracleConfig, split: bool = True):
    """Bellman targets.  `split=True` restates the reference's split / compute_weights branches LITERALLY: there a' and
    log pi(a') are sampled from the actor on `data.observations` (mtsac.py:515-520, 995-999) although the target critics
    are then evaluated on `data.next_observations` (:521-523, 1000-1002); the un-split branch (:526-536) uses
    next_observations for both.  (The reference draws that noise with one key under jax.vmap, i.e. the same draw for
    every task; here the noise is the caller's `eps_c`.)"""
    obs, actions, next_obs, dones, rewards = batch
    T = cfg.num_tasks
    alpha_vals = torch.exp(obs[..., -T:] @ state.log_alpha.reshape(-1, 1))
    with torch.no_grad():   # mtsac.py:995-1006 / :515-553: target computed once from the current actor and target critic
        next_actions, next_logp = O.actor_sample_and_log_prob(state.actor, obs if split else next_obs, eps_c, cfg)
        q_t = O.critic_forward(state.critic_target, next_obs, next_actions, cfg)
        target = rewards + (1 - dones) * cfg.gamma * (q_t.min(dim=0).values - alpha_vals * next_logp.reshape(-1, 1))
        if cfg.clip:
            target = torch.clamp(target, -5000, 5000)
    return alpha_vals, target


def critic_task_grads(critic: dict, batch, target, cfg: O.OracleConfig) -> list:
    """Per-task gradients of the critic loss (mean over the task's rows and the ensemble, :1022-1025 / :562-566)."""
    obs, actions = batch[0], batch[1]
    task = obs[..., -cfg.num_tasks:].argmax(dim=-1)
    out = []
    for t in range(cfg.num_tasks):
        rows = task == t
        cp = O._with_grad(critic)
        q_pred = O.critic_forward(cp, obs[rows], actions[rows], cfg)
        if cfg.clip:
            q_pred = torch.clamp(q_pred, -5000, 5000)
        ((q_pred - target[rows]) ** 2).mean().backward()
        out.append(O._grads_of(cp))
    return out


def actor_task_grads(actor: dict, critic: dict, batch, alpha_vals, eps_a, cfg: O.OracleConfig, explore: bool = False,
                     return_losses: bool = False):
    """Per-task gradients of the actor loss (:1049-1069 / :631-666) against the given critic parameters; also returns
    the (detached) log-probs in batch order, which the temperature step consumes.  `explore=True` is the split branch of
    `update_actor`: its vmapped call passes four arguments, so `_explore` keeps its default True (:631-637, 676-682) and
    each task's loss is reduced by mean((data.actions - action_samples)^2) (:668-673).  compute_weights' own actor loss
    (:1049-1069) has no such term."""
    obs, actions = batch[0], batch[1]
    task = obs[..., -cfg.num_tasks:].argmax(dim=-1)
    out, logp_all = [], torch.zeros(obs.shape[0], dtype=obs.dtype)
    losses, explores = [], []
    for t in range(cfg.num_tasks):
        rows = task == t
        ap = O._with_grad(actor)
        a, logp = O.actor_sample_and_log_prob(ap, obs[rows], eps_a[rows], cfg)
        q_pi = O.critic_forward(critic, obs[rows], a, cfg)
        loss = (alpha_vals[rows] * logp.reshape(-1, 1) - q_pi.min(dim=0).values).mean()
        exp_loss = ((actions[rows] - a) ** 2).mean() if explore else torch.zeros((), dtype=obs.dtype)
        loss = loss - exp_loss
        loss.backward()
        out.append(O._grads_of(ap))
        logp_all[rows] = logp.detach()
        losses.append(loss.detach())
        explores.append(exp_loss.detach())
    if return_losses:
        return out, logp_all, torch.stack(losses), torch.stack(explores)
    return out, logp_all


def per_task_grads(state: O.OracleState, batch, eps_c, eps_a, cfg: O.OracleConfig):
    """compute_weights (:870-1170): {'critic': [tree per task], 'actor': [tree per task]}; the actor loss is taken
    against the CURRENT critic parameters (:1060-1062)."""
    alpha_vals, target = _targets(state, batch, eps_c, cfg, split=True)
    return {"critic": critic_task_grads(state.critic, batch, target, cfg),
            "actor": actor_task_grads(state.actor, state.critic, batch, alpha_vals, eps_a, cfg)[0]}


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


def elementwise_metrics(g: torch.Tensor, eps: float = 1e-3, tau: float = 1.0) -> dict:
    """compute_sparsity_mismatch, compute_participation_ratio and their summaries in compute_conflict_metrics
    (utils.py:75-101, 146-156)."""
    T, n = g.shape
    off = 1 - torch.eye(T, dtype=g.dtype)
    near_zero, large = g.abs() < eps, g.abs() > tau
    mismatch = (near_zero[:, None, :] & large[None, :, :]).sum(dim=-1).to(g.dtype)
    rate = mismatch / near_zero.sum(dim=1).clamp(min=1).to(g.dtype)[:, None] * off
    n_off = T * (T - 1)
    pr = g.abs().sum(dim=1) ** 2 / (n * (g ** 2).sum(dim=1).clamp(min=1e-10))
    return {"avg_interference_rate": (rate * off).sum() / n_off, "interference_asymmetry": ((rate - rate.T).abs() * off).sum() / n_off,
            "per_task_interference_in": (rate * off).sum(dim=0) / (T - 1), "per_task_interference_out": (rate * off).sum(dim=1) / (T - 1),
            "pairwise_interference_rate": rate, "avg_participation_ratio": pr.mean(), "per_task_participation_ratio": pr}


def support_metrics(g: torch.Tensor, support_percentile: float = 0.8) -> dict:
    """compute_support_metrics (mtsac.py:774-860)."""
    T = g.shape[0]
    thr = torch.quantile(g.abs(), support_percentile, dim=1, keepdim=True)          # :804-806 (linear interpolation)
    sup = g.abs() >= thr
    si, sj = sup[:, None, :], sup[None, :, :]
    inter = (si & sj).sum(dim=-1).to(g.dtype)
    union = (si | sj).sum(dim=-1).to(g.dtype)
    jacc = inter / (union + 1e-8)
    off = 1 - torch.eye(T, dtype=g.dtype)
    n_pairs = T * (T - 1)
    conflict = (g[:, None, :] * g[None, :, :]) < 0
    genuine = ((si & sj) & conflict).sum(dim=-1).to(g.dtype)
    ghost = (~(si & sj) & conflict).sum(dim=-1).to(g.dtype)
    tot = genuine + ghost + 1e-8
    return {"pairwise_jaccard": jacc, "avg_jaccard": (jacc * off).sum() / n_pairs,
            "genuine_conflict_rate": genuine / tot, "ghost_conflict_rate": ghost / tot,
            "avg_genuine_conflict_rate": ((genuine / tot) * off).sum() / n_pairs,
            "avg_ghost_conflict_rate": ((ghost / tot) * off).sum() / n_pairs,
            "ghost_to_genuine_ratio": ghost.sum() / (genuine.sum() + 1e-8),
            "per_task_support_size": sup.sum(dim=-1).to(g.dtype), "avg_support_size": sup.sum(dim=-1).to(g.dtype).mean()}


# --------------------------------------------------------------------------------------------
# pcgrad (mtrl/optim/pcgrad.py:22-136) and the update that uses it (PCGradConfig, mtrl/config/optim.py:62-76)
# --------------------------------------------------------------------------------------------
def pcgrad(flat: torch.Tensor, perm: torch.Tensor | None = None):
    """pcgrad.py:58-83 literally: rows permuted (:79; the reference draws the permutation from a jax key, here it is an
    argument), every row projected sequentially against ALL rows in that order (:60-69), result averaged over tasks (:81).
    Returns (avg_grad, stats)."""
    g = flat if perm is None else flat[perm]
    T = g.shape[0]
    out, total = [], 0
    for i in range(T):
        gi = g[i].clone()
        for j in range(T):
            proj = torch.dot(gi, g[j]) / ((g[j] ** 2).sum() + 1e-8)
            gi = gi - torch.minimum(proj, torch.zeros_like(proj)) * g[j]
            total += int(proj < 0)
        out.append(gi)
    final = torch.stack(out)
    stats = {"n_grad_conflicts": total / 2, "avg_grad_magnitude": final.norm(dim=1).mean(),
             "avg_grad_magnitude_before_surgery": g.norm(dim=1).mean()}
    return final.mean(dim=0), stats


def cagrad(flat: torch.Tensor, c: float = 0.5, num_iterations: int = 21, learning_rate: float | None = None,
           momentum: float = 0.5):
    """mtrl/optim/cagrad.py:165-203 with its defaults (:20-41): per-task clipping to unit norm (:181-187), 21 momentum-SGD
    steps on the task weights starting from zero with the gradient of the objective by autodiff (:56-123), softmax,
    combination (:125-163).  Returns (combined_grad, stats)."""
    T = flat.shape[0]
    lr = learning_rate if learning_rate is not None else (25.0 if T < 50 else 50.0)
    g = flat * torch.clamp(1.0 / (flat.norm(dim=1, keepdim=True) + 1e-8), max=1.0)          # :181-187

    def normalised():
        GG = g @ g.T
        scale = torch.sqrt(torch.diag(GG) + 1e-4).mean()
        GG = GG / scale ** 2
        Gg = GG.mean(dim=1, keepdim=True)
        return GG, Gg, torch.sqrt(Gg.mean() + 1e-4) * c

    GG, Gg, cn = normalised()

    def objective(w):
        ww = w / (w.sum() + 1e-8)
        return (ww.T @ Gg + cn * torch.sqrt(ww.T @ GG @ ww + 1e-4)).squeeze()

    w = torch.zeros(T, 1, dtype=flat.dtype)
    vel = torch.zeros_like(w)
    w_best, obj_best = w.clone(), torch.tensor(float("inf"), dtype=flat.dtype)
    for _ in range(num_iterations - 1):
        wv = w.clone().requires_grad_(True)
        o = objective(wv)
        o.backward()
        if o < obj_best:
            w_best, obj_best = w.clone(), o.detach()
        vel = momentum * vel + wv.grad
        w = w - lr * vel
    o = objective(w)
    if o < obj_best:
        w_best, obj_best = w.clone(), o
    tw = torch.softmax(w_best.squeeze(), dim=0)
    gw_norm = torch.sqrt(tw.reshape(1, -1) @ GG @ tw.reshape(-1, 1) + 1e-4)
    lmbda = cn / (gw_norm + 1e-4)
    comb = 1.0 / T + tw * lmbda.squeeze()
    out = (comb.reshape(-1, 1) * g).sum(dim=0) / (1 + c ** 2)
    return out, {"task_weights": tw, "avg_grad_magnitude": out.norm(), "avg_grad_magnitude_before_surgery": g.norm(dim=1).mean(),
                 "cagrad_objective": obj_best}


def gradnorm(flat: torch.Tensor, task_losses: torch.Tensor, max_grad_norm: float | None = None, asymmetry: float = 0.12,
             lr: float = 3e-4):
    """mtrl/optim/gradnorm.py:86-161 for the first update from init_fn's state (:68-81), literally: per-task clip when
    max_grad_norm is set (:106-107), relative inverse rates (:109-116), the gradnorm loss differentiated with respect to
    the task weights by autograd (:134-144; the loss does not use them, so the gradient is zero), Adam on the weights,
    renormalisation (:150-153) and the weighted sum (:155-157)."""
    T = flat.shape[0]
    g = flat * torch.clamp(1.0 / (flat.norm(dim=1, keepdim=True) + 1e-8), max=1.0) if max_grad_norm else flat
    weights = torch.ones(T, dtype=flat.dtype, requires_grad=True)
    original = task_losses.clone()                       # first step: inf -> the current losses
    improvement = task_losses / (original + 1e-12)
    rate = improvement / (improvement.mean() + 1e-12)
    norms = g.norm(dim=1)
    loss = (norms - norms.mean() * rate ** asymmetry).abs().sum() + 0.0 * weights.sum()   # (the reference's loss ignores `weights`)
    loss.backward()
    grad_w = weights.grad
    m, v = 0.1 * grad_w, 0.001 * grad_w ** 2
    new_w = weights.detach() - lr * (m / 0.1) / (torch.sqrt(v / 0.001) + 1e-5)
    new_w = new_w / new_w.sum() * T
    return (new_w.reshape(-1, 1) * g).sum(dim=0), {"task_weights": new_w, "grad_magnitude": (new_w.reshape(-1, 1) * g).sum(dim=0).norm()}


def _unflatten(flat: torch.Tensor, like) -> dict:
    leaves, off = [], 0
    for x in O.tree_leaves(like):
        leaves.append(flat[off:off + x.numel()].reshape(x.shape))
        off += x.numel()
    it = iter(leaves)
    return O.tree_map(lambda x: next(it), like)


def _surgery(kind: str, flat: torch.Tensor, perm, gradnorm_clip: bool):
    if kind == "pcgrad":
        return pcgrad(flat, perm)
    if kind == "cagrad":
        return cagrad(flat)
    if kind == "dummy":   # mtrl/optim/dummy.py:18: jax.tree.map(lambda x: x.mean(axis=0), updates)
        return flat.mean(dim=0), {"grad_magnitude": flat.mean(dim=0).norm()}
    # gradnorm: the per-task losses only enter through bookkeeping that cannot move the weights (see gradnorm())
    return gradnorm(flat, torch.ones(flat.shape[0], dtype=flat.dtype), max_grad_norm=1.0 if gradnorm_clip else None)


def mtsac_update_pcgrad(state: O.OracleState, batch, eps_c, eps_a, cfg: O.OracleConfig, perm_c=None, perm_a=None,
                        critic: bool = True, actor: bool = True, surgery: str = "pcgrad", gradnorm_clip: bool = False):
    """MTSAC.update (mtsac.py:1173-1251) with split losses and optax.chain(pcgrad, clip_by_global_norm, adam) on the
    chosen networks (surgery="cagrad": cagrad instead, CAGradConfig optim.py:104-124); the other network keeps the plain
    chain.  The split flags follow the reference: split_critic_losses / split_actor_losses are each network's optimiser's
    `requires_split_task_losses` (mtsac.py:272-273), and the split branches differ from the plain ones (see `_targets` and
    `actor_task_grads`).  Returns (new_state, stats); stats["logs"] holds actor_loss / explore_loss as update_actor logs
    them (:704-709; explore_loss as the mean of the reference's per-task vector)."""
    obs = batch[0]
    T = cfg.num_tasks
    opt = dict(state.opt)
    alpha_vals, target = _targets(state, batch, eps_c, cfg, split=critic)
    stats = {}
    ctg = critic_task_grads(state.critic, batch, target, cfg)
    if critic:
        flat, stats["critic"] = _surgery(surgery, flatten(ctg), perm_c, gradnorm_clip)
        cgrads = _unflatten(flat, state.critic)
    else:
        cgrads = O.tree_map(lambda *xs: sum(xs) / len(xs), *ctg)
    stats.setdefault("critic", {})["grad_tree"] = cgrads
    new_critic, opt["critic"] = O.adam_step(state.critic, cgrads, opt["critic"], cfg.lr, cfg.adam_eps, cfg.b1, cfg.b2,
                                            cfg.max_grad_norm)
    new_target = O.tree_map(lambda n, t: cfg.tau * n + (1 - cfg.tau) * t, new_critic, state.critic_target)
    atg, logp, a_losses, a_explore = actor_task_grads(state.actor, new_critic, batch, alpha_vals, eps_a, cfg, explore=actor,
                                                      return_losses=True)
    stats["logs"] = {"losses/actor_loss": a_losses.mean(), "metrics/explore_loss": a_explore.mean()}
    if actor:
        flat, stats["actor"] = _surgery(surgery, flatten(atg), perm_a, gradnorm_clip)
        agrads = _unflatten(flat, state.actor)
    else:
        agrads = O.tree_map(lambda *xs: sum(xs) / len(xs), *atg)
    stats.setdefault("actor", {})["grad_tree"] = agrads
    new_actor, opt["actor"] = O.adam_step(state.actor, agrads, opt["actor"], cfg.lr, cfg.adam_eps, cfg.b1, cfg.b2,
                                          cfg.max_grad_norm)
    la = state.log_alpha.detach().clone().requires_grad_(True)
    (-(obs[..., -T:] @ la.reshape(-1, 1)) * (logp.reshape(-1, 1) + cfg.target_entropy)).mean().backward()
    new_la, opt["alpha"] = O.adam_step(state.log_alpha, la.grad, opt["alpha"], cfg.alpha_lr, cfg.adam_eps, cfg.b1, cfg.b2,
                                       cfg.alpha_max_grad_norm)
    return O.OracleState(new_actor, new_critic, new_target, new_la, opt), stats
