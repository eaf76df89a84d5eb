"""Oracle for the MT-SAC gradient update: a PyTorch-CPU restatement (autograd, fp64 or fp32) of
`MTSAC._update_inner` (/root/reference/mtrl/rl/algorithms/mtsac.py:1173-1247, non-split MSE branch)
and everything it calls.

TEST INFRASTRUCTURE: imported only by tests/, __graft_entry__.smoke() and bench.py's CPU baseline
(`cpu_baseline` / `--impl reference`).  The product path never touches this module.

Parity status: PARITY UNPINNED.  The reference has no test, golden vector or known-answer value for
the update, and it cannot run here or on the GPU box (jax, flax, optax, distrax absent; no network),
so this file is a restatement of the published semantics of the pinned third-party pieces
(uv.lock: jax 0.5.3, flax 0.10.4, optax 0.2.4, distrax 0.1.5) anchored on the reference's call
sites, cited per function.  The only reference-derived known answer is the parameter count
372 880 for the MT10 width-400 actor (plots/get_data.py:51-53), checked in tests/test_mtsac_oracle.py.

Randomness: `jax.random` streams cannot be reproduced without JAX, so the two Gaussian draws
(critic_loss_key at mtsac.py:355,526-528 and actor_loss_key at :629,640-642) are explicit inputs
`eps_c`, `eps_a` of shape (B, A).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import torch

LOG_KEYS = (
    "losses/qf_values", "losses/qf_loss", "metrics/critic_grad_magnitude", "metrics/critic_params_norm",
    "losses/actor_loss", "metrics/actor_grad_magnitude", "metrics/actor_params_norm", "metrics/explore_loss",
    "losses/alpha_loss", "alpha",
)  # mtsac.py:616-621, 704-709, 728-731


@dataclass
class OracleConfig:
    num_tasks: int
    obs_dim: int            # env observation incl. the one-hot task id (39 + T for Meta-World)
    action_dim: int = 4
    width: int = 400
    depth: int = 3
    num_critics: int = 2    # MTSACConfig.num_critics, mtsac.py:122
    gamma: float = 0.99     # AlgorithmConfig.gamma, config/rl.py:17
    tau: float = 0.005      # mtsac.py:123
    lr: float = 3e-4        # OptimizerConfig.lr, config/optim.py:16
    adam_eps: float = 1e-5  # config/optim.py:31-32
    b1: float = 0.9         # optax.adam defaults
    b2: float = 0.999
    max_grad_norm: float | None = 1.0       # actor / critic (experiments/*: OptimizerConfig(max_grad_norm=1.0))
    alpha_lr: float = 3e-4
    alpha_max_grad_norm: float | None = None  # mtsac.py:120
    log_std_min: float = -20.0  # config/networks.py:13-17
    log_std_max: float = 2.0
    clip: bool = False          # AlgorithmConfig.clip, config/rl.py:21; mtsac.py:558-560
    use_task_weights: bool = False  # mtsac.py:124
    initial_temperature: float = 1.0
    # "exact": plain fp32/fp64 matmuls (the reference's CPU numerics).
    # "tf32": round trunk-matmul operands (inputs, kernels, stored activations) to tf32 at the points the
    #         sm_100a kernels do, which is also what XLA's default f32 dot precision does on NVIDIA GPUs.
    #         Used by tests to separate kernel correctness from the tf32 ReLU-gate effect.
    matmul_operands: str = "exact"
    # VanillaNetworkConfig.use_layer_norm / use_skip_connections (mtrl/config/nn.py:33-39): only read by the MLP of the
    # single-task SAC / MT-PPO oracles (mtrl/nn/base.py:32-63); MultiHeadNetwork has neither.
    use_layer_norm: bool = False
    use_skip_connections: bool = False

    @property
    def target_entropy(self) -> float:  # mtsac.py:258
        return -float(self.action_dim)


# --------------------------------------------------------------------------------------------
# Parameters (Flax layout: Dense kernel is (in, out), y = x @ kernel + bias)
# --------------------------------------------------------------------------------------------
def init_multihead(gen: torch.Generator, in_dim: int, cfg: OracleConfig, head_dim: int, head_bound: float,
                   ensemble: int | None = None, dtype=torch.float32) -> dict:
    """he_uniform trunk / zero bias (config/nn.py:15-19), uniform(+-bound) head kernel and bias
    (networks.py:33-34 actor 1e-3, :65-66 critic 3e-3).  Same distributions as the reference, not
    the same draws (jax PRNG)."""
    lead = () if ensemble is None else (ensemble,)

    def u(shape, bound):
        return ((torch.rand(*shape, generator=gen, dtype=torch.float64) * 2 - 1) * bound).to(dtype)

    p = {}
    d = in_dim
    for i in range(cfg.depth):
        p[f"layer_{i}"] = {"kernel": u(lead + (d, cfg.width), math.sqrt(6.0 / d)),
                           "bias": torch.zeros(lead + (cfg.width,), dtype=dtype)}
        d = cfg.width
    p["heads"] = {"kernel": u(lead + (cfg.num_tasks, cfg.width, head_dim), head_bound),
                  "bias": u(lead + (cfg.num_tasks, head_dim), head_bound)}
    return p


def tree_map(fn, *trees):
    t0 = trees[0]
    if isinstance(t0, dict):
        return {k: tree_map(fn, *(t[k] for t in trees)) for k in t0}
    return fn(*trees)


def tree_leaves(tree) -> list:
    if isinstance(tree, dict):
        out = []
        for k in tree:
            out += tree_leaves(tree[k])
        return out
    return [tree]


def num_params(tree) -> int:
    return sum(x.numel() for x in tree_leaves(tree))


@dataclass
class OracleState:
    actor: dict
    critic: dict
    critic_target: dict
    log_alpha: torch.Tensor
    opt: dict = field(default_factory=dict)   # name -> {"m": tree, "v": tree, "count": int}

    def clone(self) -> "OracleState":
        c = lambda t: tree_map(lambda x: x.clone(), t)  # noqa: E731
        return OracleState(c(self.actor), c(self.critic), c(self.critic_target), self.log_alpha.clone(),
                           {k: {"m": c(v["m"]), "v": c(v["v"]), "count": v["count"]} for k, v in self.opt.items()})

    def to(self, dtype) -> "OracleState":
        c = lambda t: tree_map(lambda x: x.to(dtype), t)  # noqa: E731
        return OracleState(c(self.actor), c(self.critic), c(self.critic_target), self.log_alpha.to(dtype),
                           {k: {"m": c(v["m"]), "v": c(v["v"]), "count": v["count"]} for k, v in self.opt.items()})


def init_state(cfg: OracleConfig, seed: int = 1, dtype=torch.float32) -> OracleState:
    """MTSAC.initialize (mtsac.py:152-284): nets, zero Adam moments, target = copy(params),
    log_alpha = log(initial_temperature) per task (mtsac.py:48-58)."""
    gen = torch.Generator().manual_seed(seed)
    actor = init_multihead(gen, cfg.obs_dim, cfg, 2 * cfg.action_dim, 1e-3, None, dtype)
    critic = init_multihead(gen, cfg.action_dim + cfg.obs_dim, cfg, 1, 3e-3, cfg.num_critics, dtype)
    log_alpha = torch.full((cfg.num_tasks,), math.log(cfg.initial_temperature), dtype=dtype)
    zeros = lambda t: tree_map(torch.zeros_like, t)  # noqa: E731
    opt = {
        "actor": {"m": zeros(actor), "v": zeros(actor), "count": 0},
        "critic": {"m": zeros(critic), "v": zeros(critic), "count": 0},
        "alpha": {"m": torch.zeros_like(log_alpha), "v": torch.zeros_like(log_alpha), "count": 0},
    }
    return OracleState(actor, critic, tree_map(lambda x: x.clone(), critic), log_alpha, opt)


# --------------------------------------------------------------------------------------------
# Networks
# --------------------------------------------------------------------------------------------
def tf32_round(x: torch.Tensor) -> torch.Tensor:
    """cvt.rna.tf32.f32: round to 10 mantissa bits, ties away from zero (value semantics only)."""
    i = x.detach().to(torch.float32).contiguous().view(torch.int32)
    return ((i + 0x1000) & ~0x1FFF).view(torch.float32).to(x.dtype)


def _rt(x: torch.Tensor, mode: str) -> torch.Tensor:
    """Operand rounding with a straight-through gradient."""
    if mode == "exact":
        return x
    return x + (tf32_round(x) - x).detach()


def multihead_forward(p: dict, x: torch.Tensor, num_tasks: int, depth: int, all_heads: bool = False,
                      operands: str = "exact") -> torch.Tensor:
    """MultiHeadNetwork.__call__ (mtrl/nn/multi_head.py:21-68): `depth` Dense+ReLU layers on the full
    input (one-hot included), T heads, then pick the head of argmax(one-hot).  `all_heads=True`
    evaluates every head for every row and gathers, exactly as the reference does (:50-66); the
    default computes only the selected head (same numbers, less work)."""
    task = x[..., -num_tasks:].argmax(dim=-1)  # :26, :65
    h = _rt(x, operands)
    for i in range(depth):
        h = _rt(torch.relu(h @ _rt(p[f"layer_{i}"]["kernel"], operands) + p[f"layer_{i}"]["bias"]), operands)  # :34-44
    hk, hb = p["heads"]["kernel"], p["heads"]["bias"]
    if all_heads:
        out = torch.einsum("bw,twh->bth", h, hk) + hb[None]       # :50-62
        return out[torch.arange(x.shape[0]), task]                # :66
    return torch.einsum("bw,bwh->bh", h, hk[task]) + hb[task]


def ensemble_forward(p: dict, x: torch.Tensor, num_tasks: int, depth: int, all_heads: bool = False,
                     operands: str = "exact") -> torch.Tensor:
    """Ensemble (mtrl/rl/networks.py:208-222): params stacked on axis 0, shared input -> (E, B, 1)."""
    E = p["layer_0"]["kernel"].shape[0]
    outs = []
    for e in range(E):
        pe = tree_map(lambda t: t[e], p)
        outs.append(multihead_forward(pe, x, num_tasks, depth, all_heads, operands))
    return torch.stack(outs, 0)


def critic_forward(p: dict, obs: torch.Tensor, act: torch.Tensor, cfg: OracleConfig, all_heads=False) -> torch.Tensor:
    """QValueFunction (networks.py:55-67): input is concatenate((action, state)) (:61)."""
    return ensemble_forward(p, torch.cat((act, obs), dim=-1), cfg.num_tasks, cfg.depth, all_heads, cfg.matmul_operands)


def actor_sample_and_log_prob(p: dict, obs: torch.Tensor, eps: torch.Tensor, cfg: OracleConfig, all_heads=False):
    """ContinuousActionPolicy (networks.py:29-45) + TanhMultivariateNormalDiag.sample_and_log_prob
    (mtrl/nn/distributions.py:6-16; distrax Transformed: log_prob(y) = base.log_prob(x) - fldj(x),
    Tanh.forward_log_det_jacobian(x) = 2 (log 2 - x - softplus(-2x)), summed by Block(.,1))."""
    out = multihead_forward(p, obs, cfg.num_tasks, cfg.depth, all_heads, cfg.matmul_operands)
    mean, log_std = out[..., : cfg.action_dim], out[..., cfg.action_dim:]          # :37
    log_std = torch.clamp(log_std, cfg.log_std_min, cfg.log_std_max)               # :38-40
    std = torch.exp(log_std)                                                       # :41
    x = mean + std * eps
    base_lp = (-0.5 * eps**2 - 0.5 * math.log(2 * math.pi) - log_std).sum(-1)
    fldj = (2.0 * (math.log(2.0) - x - torch.nn.functional.softplus(-2.0 * x))).sum(-1)
    return torch.tanh(x), base_lp - fldj


def actor_action(p: dict, obs: torch.Tensor, cfg: OracleConfig, eps: torch.Tensor | None = None) -> torch.Tensor:
    """`_sample_action` / `_eval_action` (mtsac.py:70-84): dist.sample() = tanh(mu + sigma eps) for the given draw, or
    with eps=None TanhMultivariateNormalDiag.mode() = tanh(mu) (mtrl/nn/distributions.py:15-16)."""
    out = multihead_forward(p, obs, cfg.num_tasks, cfg.depth, False, cfg.matmul_operands)
    mean, log_std = out[..., : cfg.action_dim], out[..., cfg.action_dim:]
    if eps is None:
        return torch.tanh(mean)
    return torch.tanh(mean + torch.exp(torch.clamp(log_std, cfg.log_std_min, cfg.log_std_max)) * eps)


# --------------------------------------------------------------------------------------------
# Optimiser: optax.chain(clip_by_global_norm(max), adam(lr, eps)) (mtrl/config/optim.py:26-43),
# applied by TrainState.apply_gradients (mtrl/rl/algorithms/utils.py:11-46).
# --------------------------------------------------------------------------------------------
def global_norm(tree) -> torch.Tensor:
    return torch.sqrt(sum((g.double() ** 2).sum() for g in tree_leaves(tree))).to(tree_leaves(tree)[0].dtype)


def adam_step(params, grads, opt: dict, lr: float, eps: float, b1: float, b2: float, max_norm: float | None):
    if max_norm is not None:
        gn = global_norm(grads)
        # optax.clip_by_global_norm: g if g_norm < max_norm else g / g_norm * max_norm
        if not bool(gn < max_norm):
            grads = tree_map(lambda g: g / gn * max_norm, grads)
    count = opt["count"] + 1
    m = tree_map(lambda m_, g: b1 * m_ + (1 - b1) * g, opt["m"], grads)
    v = tree_map(lambda v_, g: b2 * v_ + (1 - b2) * g * g, opt["v"], grads)
    bc1, bc2 = 1 - b1**count, 1 - b2**count
    new_params = tree_map(lambda p, m_, v_: p - lr * (m_ / bc1) / (torch.sqrt(v_ / bc2) + eps), params, m, v)
    return new_params, {"m": m, "v": v, "count": count}


# --------------------------------------------------------------------------------------------
# The update
# --------------------------------------------------------------------------------------------
def _with_grad(tree):
    return tree_map(lambda t: t.detach().clone().requires_grad_(True), tree)


def _grads_of(tree):
    return tree_map(lambda t: t.grad if t.grad is not None else torch.zeros_like(t), tree)


def mtsac_update(state: OracleState, batch, eps_c: torch.Tensor, eps_a: torch.Tensor, cfg: OracleConfig,
                 all_heads: bool = False, return_grads: bool = False):
    """One `MTSAC.update` (mtsac.py:1173-1251).  `batch` = (observations, actions, next_observations,
    dones, rewards) in the field order of ReplayBufferSamples (mtrl/types.py:30-35)."""
    obs, actions, next_obs, dones, rewards = batch
    T = cfg.num_tasks
    B = obs.shape[0]
    task_ids = obs[..., -T:]                                                         # :1175
    alpha_vals = torch.exp(task_ids @ state.log_alpha.reshape(-1, 1))                # :1177, :60-63
    if cfg.use_task_weights:                                                         # :103-113
        tw = torch.softmax(-state.log_alpha, dim=0)
        task_weights = (task_ids @ tw.reshape(-1, 1)) * T
    else:
        task_weights = None
    logs = {}
    grads_out = {}
    opt = dict(state.opt)  # the input state is left untouched

    # ---- critic (mtsac.py:513-621) ----
    with torch.no_grad():
        next_actions, next_logp = actor_sample_and_log_prob(state.actor, next_obs, eps_c, cfg, all_heads)   # :526-528
        q_t = critic_forward(state.critic_target, next_obs, next_actions, cfg, all_heads)                   # :534-536
        min_q_next = q_t.min(dim=0).values - alpha_vals * next_logp.reshape(-1, 1)                          # :547-549
        target = rewards + (1 - dones) * cfg.gamma * min_q_next                                             # :551-553
        if cfg.clip:
            target = torch.clamp(target, -5000, 5000)                                                       # :559
    cp = _with_grad(state.critic)
    q_pred = critic_forward(cp, obs, actions, cfg, all_heads)                                               # :555
    if cfg.clip:
        q_pred = torch.clamp(q_pred, -5000, 5000)                                                           # :560
    if task_weights is not None:
        critic_loss = (task_weights * (q_pred - target) ** 2).mean()                                        # :563
    else:
        critic_loss = ((q_pred - target) ** 2).mean()                                                       # :565
    critic_loss.backward()
    cgrads = _grads_of(cp)
    logs["losses/qf_values"] = q_pred.mean().detach()
    logs["losses/qf_loss"] = critic_loss.detach()
    logs["metrics/critic_grad_magnitude"] = global_norm(cgrads)                                             # :619 (pre-clip)
    new_critic, opt["critic"] = adam_step(state.critic, cgrads, opt["critic"], cfg.lr, cfg.adam_eps,
                                                cfg.b1, cfg.b2, cfg.max_grad_norm)                          # :600-606
    new_target = tree_map(lambda n, t: cfg.tau * n + (1 - cfg.tau) * t, new_critic, state.critic_target)    # :607-613
    logs["metrics/critic_params_norm"] = global_norm(new_critic)                                            # :620
    grads_out["critic"] = cgrads

    # ---- actor (mtsac.py:623-711): new critic, old alpha ----
    ap = _with_grad(state.actor)
    a_samples, logp = actor_sample_and_log_prob(ap, obs, eps_a, cfg, all_heads)                             # :640-642
    logp_col = logp.reshape(-1, 1)                                                                          # :649
    q_pi = critic_forward(new_critic, obs, a_samples, cfg, all_heads)                                       # :659-661
    min_q = q_pi.min(dim=0).values                                                                          # :662
    if task_weights is not None:
        actor_loss = (task_weights * (alpha_vals * logp_col - min_q)).mean()                                # :664
    else:
        actor_loss = (alpha_vals * logp_col - min_q).mean()                                                 # :666
    actor_loss.backward()   # explore=False (:277) -> exp_loss = 0.0 (:671-673)
    agrads = _grads_of(ap)
    logs["losses/actor_loss"] = actor_loss.detach()
    logs["metrics/actor_grad_magnitude"] = global_norm(agrads)                                              # :706
    new_actor, opt["actor"] = adam_step(state.actor, agrads, opt["actor"], cfg.lr, cfg.adam_eps,
                                              cfg.b1, cfg.b2, cfg.max_grad_norm)                            # :695-701
    logs["metrics/actor_params_norm"] = global_norm(new_actor)                                              # :707
    logs["metrics/explore_loss"] = torch.zeros((), dtype=obs.dtype)                                         # :708
    grads_out["actor"] = agrads

    # ---- alpha (mtsac.py:713-731): log-probs of the pre-update actor ----
    la = state.log_alpha.detach().clone().requires_grad_(True)
    alpha_loss = (-(task_ids @ la.reshape(-1, 1)) * (logp_col.detach() + cfg.target_entropy)).mean()        # :720-721
    alpha_loss.backward()
    new_la, opt["alpha"] = adam_step(state.log_alpha, la.grad, opt["alpha"], cfg.alpha_lr, cfg.adam_eps,
                                           cfg.b1, cfg.b2, cfg.alpha_max_grad_norm)                         # :726
    logs["losses/alpha_loss"] = alpha_loss.detach()
    logs["alpha"] = torch.exp(new_la).sum()                                                                 # :730
    grads_out["alpha"] = la.grad

    new_state = OracleState(new_actor, new_critic, new_target, new_la, opt)
    aux = {"logp": logp.detach(), "next_logp": next_logp, "actions": a_samples.detach(),
           "next_actions": next_actions, "q_pred": q_pred.detach(), "target": target}
    if return_grads:
        return new_state, logs, grads_out, aux
    return new_state, logs


# --------------------------------------------------------------------------------------------
# Synthetic inputs: Meta-World shaped transitions, no environment
# --------------------------------------------------------------------------------------------
def synthetic_batch(cfg: OracleConfig, per_task: int, seed: int = 1234, dtype=torch.float32, interleaved=True):
    """B = per_task * T rows; row i <-> (sample i // T, task i % T) like buffers.py:547-548."""
    g = torch.Generator().manual_seed(seed)
    T, B = cfg.num_tasks, per_task * cfg.num_tasks
    feat = cfg.obs_dim - T
    task = (torch.arange(B) % T) if interleaved else (torch.arange(B) // per_task)
    onehot = torch.nn.functional.one_hot(task, T).to(torch.float64)
    o = torch.randn(B, feat, generator=g, dtype=torch.float64)
    obs = torch.cat((o, onehot), 1)
    next_obs = torch.cat((o + 0.01 * torch.randn(B, feat, generator=g, dtype=torch.float64), onehot), 1)
    actions = torch.rand(B, cfg.action_dim, generator=g, dtype=torch.float64) * 2 - 1
    rewards = torch.rand(B, 1, generator=g, dtype=torch.float64) * 10
    dones = (torch.rand(B, 1, generator=g, dtype=torch.float64) < 0.002).to(torch.float64)
    eps_c = torch.randn(B, cfg.action_dim, generator=g, dtype=torch.float64)
    eps_a = torch.randn(B, cfg.action_dim, generator=g, dtype=torch.float64)
    cast = lambda t: t.to(torch.float32).to(dtype)  # noqa: E731  (values are exactly fp32-representable)
    return tuple(cast(t) for t in (obs, actions, next_obs, dones, rewards)), cast(eps_c), cast(eps_a)
