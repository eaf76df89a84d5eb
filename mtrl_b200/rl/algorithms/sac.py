"""Mirror of `mtrl.rl.algorithms.sac.SAC` (/root/reference/mtrl/rl/algorithms/sac.py:96-386): the single-task SAC the
reference uses as its parameter-matched baseline (experiments/baselines/mt10_sac_v2.py:36-50), on the MLP of
`VanillaNetwork` (mtrl/nn/base.py:11-89) including its optional pre-layer LayerNorm and skip connections
(`VanillaNetworkConfig.use_layer_norm / use_skip_connections`; csrc/ln_kernels.cuh).

It runs on the same fused CUDA update as MTSAC with `variant = MTRL_VARIANT_SAC`: one "task", the MLP's last Dense is
the single head, the temperature is a scalar, alpha is updated first, the critic loss is 0.5 * sum_e mean_b and the
parameter-norm logs are those of the pre-update parameters (sac.py:292, 334-364).
"""
from __future__ import annotations

import dataclasses
import math

import numpy as np
import torch

from ... import _lib as L
from ...config.networks import ContinuousActionPolicyConfig, QValueFunctionConfig
from ...config.nn import VanillaNetworkConfig
from ...config.optim import OptimizerConfig
from ...config.rl import AlgorithmConfig
from ...nn.multi_head import _kernel_init, uniform
from ...config.utils import Activation, Initializer
from .mtsac import MTSAC, SacConfigC, TrainState, _tree_copy_, precision_code

SAC_LOG_KEYS = (
    "losses/alpha_loss", "alpha", "losses/qf_values", "losses/qf_loss", "metrics/critic_grad_magnitude",
    "metrics/actor_grad_magnitude", "metrics/actor_params_norm", "metrics/critic_params_norm", "losses/actor_loss",
)  # sac.py:300-304, 326-329, 359-364, 383


@dataclasses.dataclass(frozen=True)
class SACConfig(AlgorithmConfig):  # sac.py:96-103
    actor_config: ContinuousActionPolicyConfig = ContinuousActionPolicyConfig()
    critic_config: QValueFunctionConfig = QValueFunctionConfig()
    temperature_optimizer_config: OptimizerConfig = OptimizerConfig(max_grad_norm=None)
    initial_temperature: float = 1.0
    num_critics: int = 2
    tau: float = 0.005


def _mlp_views(flat: torch.Tensor, lay, in_dim: int, ensemble: bool) -> dict:
    """Flax names of MLP (nn/base.py:38-58): layer_0..layer_{depth-1}, output Dense = layer_{depth}."""
    W, D, E, hd = lay.width, lay.depth, lay.members, lay.head_dim
    tree = {}
    o0 = flat.storage_offset()   # as_strided offsets are absolute in the storage
    d = in_dim
    for i in range(D):
        k = flat.as_strided((E, d, W), (lay.member_trunk_stride, W, 1), o0 + lay.kernel_off[i])
        b = flat.as_strided((E, W), (lay.member_trunk_stride, 1), o0 + lay.bias_off[i])
        tree[f"layer_{i}"] = {"kernel": k if ensemble else k[0], "bias": b if ensemble else b[0]}
        d = W
    hk = flat.as_strided((E, W, hd), (lay.member_head_stride, hd, 1), o0 + lay.heads_base + lay.head_kernel_off)
    hb = flat.as_strided((E, hd), (lay.member_head_stride, 1), o0 + lay.heads_base + lay.head_bias_off)
    tree[f"layer_{D}"] = {"kernel": hk if ensemble else hk[0], "bias": hb if ensemble else hb[0]}
    if lay.use_layer_norm:   # nn.LayerNorm() modules in creation order (base.py:35-37, 52-53): LayerNorm_k feeds layer_{k+1}
        for k in range(D):
            sc = flat.as_strided((E, W), (lay.member_trunk_stride, 1), o0 + lay.ln_scale_off[k])
            bi = flat.as_strided((E, W), (lay.member_trunk_stride, 1), o0 + lay.ln_bias_off[k])
            tree[f"LayerNorm_{k}"] = {"scale": sc if ensemble else sc[0], "bias": bi if ensemble else bi[0]}
    return tree


def _wrap(tree: dict, ensemble: bool) -> dict:
    inner = {"VanillaNetwork_0": {"MLP_0": tree}}
    return {"params": {"VmapQValueFunction_0": inner} if ensemble else inner}


class SAC(MTSAC):
    LOG_KEYS = SAC_LOG_KEYS

    @staticmethod
    def initialize(config: SACConfig, env_config, seed: int = 1, *, max_batch: int = 1280,
                   device: str | torch.device | None = None, precision: str | None = None) -> "SAC":
        """sac.py:118-200.  `max_batch` bounds the rows one update may pass (reference batch: 1280)."""
        if not torch.cuda.is_available():
            raise L.MtrlError("SAC needs a CUDA device; there is no CPU fallback")
        self = object.__new__(SAC)
        dev = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self.device, self.config = dev, config
        self.num_tasks = config.num_tasks            # kept for the caller; the networks are task-agnostic MLPs
        self.rank, self.world_size, self.process_group = 0, 1, None
        self.task_begin, self.task_end = 0, 1
        obs_dim = int(np.prod(env_config.observation_space.shape))
        act_dim = int(np.prod(env_config.action_space.shape))
        anc, cnc = config.actor_config.network_config, config.critic_config.network_config
        for nc in (anc, cnc):
            if type(nc) is not VanillaNetworkConfig:
                raise NotImplementedError(f"{type(nc).__name__}: the accelerated SAC path covers VanillaNetworkConfig")
            if nc.activation != Activation.ReLU or not nc.use_bias:
                raise NotImplementedError("the fused path implements Dense(use_bias=True) + ReLU")
        if (anc.width, anc.depth, anc.use_layer_norm, anc.use_skip_connections) != \
                (cnc.width, cnc.depth, cnc.use_layer_norm, cnc.use_skip_connections):
            raise NotImplementedError("actor and critic must share width, depth, use_layer_norm and use_skip_connections")
        a_opt, c_opt, t_opt = anc.optimizer.spawn(), cnc.optimizer.spawn(), config.temperature_optimizer_config.spawn()
        self.gamma, self.tau, self.num_critics = config.gamma, config.tau, config.num_critics
        self.target_entropy = -float(act_dim)
        self.actor_network_type = "vanilla"
        nm = lambda v: -1.0 if v is None else float(v)  # noqa: E731
        max_rows = -(-max_batch // 128) * 128
        self._cfg = SacConfigC(
            num_tasks=1, task_begin=0, num_local_tasks=1, obs_dim=obs_dim, action_dim=act_dim, width=anc.width,
            depth=anc.depth, num_critics=config.num_critics, max_rows=max_rows, max_batch=max_batch, gamma=config.gamma,
            tau=config.tau, actor_lr=a_opt.lr, critic_lr=c_opt.lr, alpha_lr=t_opt.lr, adam_b1=a_opt.b1, adam_b2=a_opt.b2,
            adam_eps=a_opt.eps, actor_max_grad_norm=nm(a_opt.max_grad_norm), critic_max_grad_norm=nm(c_opt.max_grad_norm),
            alpha_max_grad_norm=nm(t_opt.max_grad_norm), log_std_min=config.actor_config.log_std_min,
            log_std_max=config.actor_config.log_std_max, target_entropy=self.target_entropy, clip_q=0, use_task_weights=0,
            noise_seed=int(seed) & (2**63 - 1), variant=1, precision=precision_code(precision),
            use_layer_norm=int(anc.use_layer_norm), use_skip_connections=int(anc.use_skip_connections))
        self.precision = "fp32x3" if self._cfg.precision else "tf32"
        self._allocate(dev, 1)
        lay = self._lay

        def ts(prefix, l, in_dim, ens, tx, step_idx):
            v = lambda name: _wrap(_mlp_views(self._flat[f"{prefix}_{name}"], l, in_dim, ens), ens)  # noqa: E731
            return TrainState(step=self._steps[step_idx], params=v("params"),
                              opt_state={"count": self._steps[step_idx], "mu": v("m"), "nu": v("v")}, tx=tx,
                              target_params=v("target") if ens else None, grads=v("grads"))
        self.actor = ts("actor", lay.actor, obs_dim, False, a_opt, 0)
        self.critic = ts("critic", lay.critic, act_dim + obs_dim, True, c_opt, 1)
        la = self._flat["log_alpha"][:1]
        self.alpha = TrainState(step=self._steps[2], params={"params": {"log_alpha": la}},
                                opt_state={"count": self._steps[2], "mu": {"params": {"log_alpha": self._flat["alpha_m"][:1]}},
                                           "nu": {"params": {"log_alpha": self._flat["alpha_v"][:1]}}}, tx=t_opt)
        gen = torch.Generator().manual_seed(int(seed))

        def init(nc, in_dim, head_dim, bound, ens):
            lead = () if ens is None else (ens,)
            kinit = _kernel_init(nc.kernel_init)
            binit = (lambda g, shape: torch.zeros(*shape)) if nc.bias_init == Initializer.ZEROS else _kernel_init(nc.bias_init)
            p, d = {}, in_dim
            for i in range(nc.depth):
                p[f"layer_{i}"] = {"kernel": kinit(gen, lead + (d, nc.width)), "bias": binit(gen, lead + (nc.width,))}
                d = nc.width
            p[f"layer_{nc.depth}"] = {"kernel": uniform(bound)(gen, lead + (nc.width, head_dim)),
                                      "bias": uniform(bound)(gen, lead + (head_dim,))}
            if nc.use_layer_norm:   # flax LayerNorm: scale = ones, bias = zeros
                for k in range(nc.depth):
                    p[f"LayerNorm_{k}"] = {"scale": torch.ones(lead + (nc.width,)), "bias": torch.zeros(lead + (nc.width,))}
            return p
        _tree_copy_(self.actor.params["params"]["VanillaNetwork_0"]["MLP_0"], init(anc, obs_dim, 2 * act_dim, 1e-3, None))
        _tree_copy_(self.critic.params["params"]["VmapQValueFunction_0"]["VanillaNetwork_0"]["MLP_0"],
                    init(cnc, act_dim + obs_dim, 1, 3e-3, config.num_critics))
        self._flat["critic_target"].copy_(self._flat["critic_params"])
        la.fill_(math.log(config.initial_temperature))
        self._create_handle()
        return self

    def _wrap_tree(self, flat, lay, in_dim, ens):   # state_dict / load_state_dict use the MLP's Flax names
        return _wrap(_mlp_views(flat, lay, in_dim, ens), ens)

    def get_num_params(self) -> dict[str, int]:
        c = self._cfg

        def count(in_dim, head):
            n, d = 0, in_dim
            for _ in range(c.depth):
                n += d * c.width + c.width
                d = c.width
            if c.use_layer_norm:
                n += c.depth * 2 * c.width
            return n + c.width * head + head
        return {"actor_num_params": count(c.obs_dim, 2 * c.action_dim),
                "critic_num_params": c.num_critics * count(c.action_dim + c.obs_dim, 1)}

    def logs(self) -> dict:
        """The nine log scalars of sac.py:300-304, 326-329, 359-364, 383 (0-dim device tensors)."""
        from .mtsac import LOG_KEYS

        full = {k: self._logs[i] for i, k in enumerate(LOG_KEYS)}
        return {k: full[k] for k in SAC_LOG_KEYS}
