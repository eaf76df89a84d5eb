"""Mirror of `mtrl.rl.algorithms.mtppo.MTPPO` (/root/reference/mtrl/rl/algorithms/mtppo.py:83-330) on the fused
CUDA path (csrc/ppo.cu, C-ABI `mtrl_ppo_*`): `MTPPOConfig`, `MTPPO.initialize(config, env_config, seed)`,
`update(rollout) -> (self, logs)` with the six log keys of mtppo.py:234-239, 274-277.

The policy must be built with `ContinuousActionPolicyConfig(squash_tanh=False)`: the reference asks the action
distribution for its entropy (mtppo.py:232), which distrax only defines for the un-squashed Gaussian.
"""
from __future__ import annotations

import ctypes as C
import dataclasses

import numpy as np
import torch

from ... import _lib as L
from ...config.networks import ContinuousActionPolicyConfig, ValueFunctionConfig
from ...config.nn import MultiHeadConfig, VanillaNetworkConfig
from ...config.rl import AlgorithmConfig
from ...config.utils import Activation, Initializer
from ...nn.multi_head import _kernel_init, uniform
from ...types import Rollout
from .mtsac import NetLayoutC, TrainState, _tree_copy_, _views
from .sac import _mlp_views

PPO_LOG_KEYS = ("losses/entropy_loss", "losses/policy_loss", "losses/approx_kl", "losses/clip_fracs",
                "losses/value_function", "losses/values")


class PpoConfigC(C.Structure):
    _fields_ = [
        ("num_tasks", C.c_int), ("obs_dim", C.c_int), ("action_dim", C.c_int), ("width", C.c_int), ("depth", C.c_int),
        ("steps_per_task", C.c_int), ("clip_eps", C.c_float), ("clip_vf_loss", C.c_int),
        ("entropy_coefficient", C.c_float), ("vf_coefficient", C.c_float), ("normalize_advantages", C.c_int),
        ("policy_lr", C.c_float), ("vf_lr", C.c_float), ("adam_b1", C.c_float), ("adam_b2", C.c_float), ("adam_eps", C.c_float),
        ("policy_max_grad_norm", C.c_float), ("vf_max_grad_norm", C.c_float), ("log_std_min", C.c_float),
        ("log_std_max", C.c_float), ("noise_seed", C.c_ulonglong),
        ("use_layer_norm", C.c_int), ("use_skip_connections", C.c_int),
    ]


class PpoLayoutC(C.Structure):
    _fields_ = [("policy", NetLayoutC), ("vf", NetLayoutC), ("workspace_bytes", C.c_longlong), ("k_in", C.c_int),
                ("max_rows", C.c_int)]


class PpoBuffersC(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in (
        "policy_params", "policy_grads", "policy_m", "policy_v", "policy_shadow",
        "vf_params", "vf_grads", "vf_m", "vf_v", "vf_shadow", "steps", "logs", "workspace")]


_vp = C.c_void_p
L._EXTRA_DECLS.update({
    "mtrl_ppo_query_layout": ([C.POINTER(PpoConfigC), C.POINTER(PpoLayoutC)],),
    "mtrl_ppo_create": ([C.POINTER(_vp), C.POINTER(PpoConfigC), C.POINTER(PpoBuffersC)],),
    "mtrl_ppo_destroy": ([_vp], None),
    "mtrl_ppo_refresh_shadows": ([_vp, _vp],),
    "mtrl_ppo_update": ([_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp],),
    "mtrl_ppo_launches_per_update": ([_vp],),
})


@dataclasses.dataclass(frozen=True)
class MTPPOConfig(AlgorithmConfig):  # mtppo.py:83-91
    policy_config: ContinuousActionPolicyConfig = ContinuousActionPolicyConfig()
    vf_config: ValueFunctionConfig = ValueFunctionConfig()
    clip_eps: float = 0.2
    clip_vf_loss: bool = True
    entropy_coefficient: float = 5e-3
    vf_coefficient: float = 0.001
    normalize_advantages: bool = True


class MTPPO:
    LOG_KEYS = PPO_LOG_KEYS

    def __init__(self):
        raise TypeError("use MTPPO.initialize(config, env_config, seed)")

    @staticmethod
    def initialize(config: MTPPOConfig, env_config, seed: int = 1, *, rollout_steps: int = 10_000,
                   device: str | torch.device | None = None) -> "MTPPO":
        """mtppo.py:106-160.  `rollout_steps` = timesteps per task in one rollout (OnPolicyTrainingConfig.rollout_steps,
        mtrl/config/rl.py:86); the update consumes exactly num_tasks * rollout_steps rows."""
        if not torch.cuda.is_available():
            raise L.MtrlError("MTPPO needs a CUDA device; there is no CPU fallback")
        self = object.__new__(MTPPO)
        dev = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self.device, self.config, self.num_tasks = dev, config, config.num_tasks
        obs_dim = int(np.prod(env_config.observation_space.shape))
        act_dim = int(np.prod(env_config.action_space.shape))
        pnc, vnc = config.policy_config.network_config, config.vf_config.network_config
        if type(pnc) is not type(vnc) or type(pnc) not in (MultiHeadConfig, VanillaNetworkConfig):
            raise NotImplementedError("policy and value function must both be MultiHeadConfig or both VanillaNetworkConfig")
        if (pnc.width, pnc.depth) != (vnc.width, vnc.depth):
            raise NotImplementedError("policy and value function must share width and depth")
        if config.policy_config.squash_tanh:
            raise ValueError("MTPPO needs ContinuousActionPolicyConfig(squash_tanh=False): mtppo.py:232 takes the "
                             "entropy of the action distribution, undefined for the tanh-squashed one")
        for nc in (pnc, vnc):
            if nc.activation != Activation.ReLU or not nc.use_bias:
                raise NotImplementedError("the fused path implements Dense(use_bias=True) + ReLU")
        self._multihead = type(pnc) is MultiHeadConfig
        use_ln = (not self._multihead) and bool(pnc.use_layer_norm)
        use_skip = (not self._multihead) and bool(pnc.use_skip_connections)
        if not self._multihead and (pnc.use_layer_norm, pnc.use_skip_connections) != (vnc.use_layer_norm, vnc.use_skip_connections):
            raise NotImplementedError("policy and value function must share use_layer_norm / use_skip_connections")
        heads = config.num_tasks if self._multihead else 1
        p_opt, v_opt = pnc.optimizer.spawn(), vnc.optimizer.spawn()
        nm = lambda v: -1.0 if v is None else float(v)  # noqa: E731
        self.rollout_steps = rollout_steps
        self._rows_per_head = rollout_steps if self._multihead else rollout_steps * config.num_tasks
        self._cfg = PpoConfigC(
            num_tasks=heads, obs_dim=obs_dim, action_dim=act_dim, width=pnc.width, depth=pnc.depth,
            steps_per_task=self._rows_per_head, clip_eps=config.clip_eps, clip_vf_loss=int(config.clip_vf_loss),
            entropy_coefficient=config.entropy_coefficient, vf_coefficient=config.vf_coefficient,
            normalize_advantages=int(config.normalize_advantages), policy_lr=p_opt.lr, vf_lr=v_opt.lr, adam_b1=p_opt.b1,
            adam_b2=p_opt.b2, adam_eps=p_opt.eps, policy_max_grad_norm=nm(p_opt.max_grad_norm),
            vf_max_grad_norm=nm(v_opt.max_grad_norm), log_std_min=config.policy_config.log_std_min,
            log_std_max=config.policy_config.log_std_max, noise_seed=int(seed) & (2**63 - 1),
            use_layer_norm=int(use_ln), use_skip_connections=int(use_skip))
        lay = PpoLayoutC()
        L.check(L.lib().mtrl_ppo_query_layout(C.byref(self._cfg), C.byref(lay)))
        self._lay = lay
        z = lambda n, dt=torch.float32: torch.zeros(int(n), dtype=dt, device=dev)  # noqa: E731
        self._flat = {f"policy_{k}": z(lay.policy.total) for k in ("params", "grads", "m", "v", "shadow")}
        self._flat.update({f"vf_{k}": z(lay.vf.total) for k in ("params", "grads", "m", "v", "shadow")})
        self._steps, self._logs = z(4, torch.int32), z(16)
        self._workspace = z((lay.workspace_bytes + 3) // 4 + 64)
        views = _views if self._multihead else _mlp_views
        net_name = "MultiHeadNetwork_0" if self._multihead else "VanillaNetwork_0"

        def wrap(tree):
            return {"params": {net_name: tree if self._multihead else {"MLP_0": tree}}}

        def ts(prefix, l, tx, idx):
            v = lambda name: wrap(views(self._flat[f"{prefix}_{name}"], l, obs_dim, False))  # noqa: E731
            return TrainState(step=self._steps[idx], params=v("params"), opt_state={"count": self._steps[idx], "mu": v("m"), "nu": v("v")},
                              tx=tx, grads=v("grads"))
        self.policy = ts("policy", lay.policy, p_opt, 0)
        self.value_function = ts("vf", lay.vf, v_opt, 1)
        gen = torch.Generator().manual_seed(int(seed))

        def init(nc, head_dim, bound):
            kinit = _kernel_init(nc.kernel_init)
            binit = (lambda g, shape: torch.zeros(*shape)) if nc.bias_init == Initializer.ZEROS else _kernel_init(nc.bias_init)
            p, d = {}, obs_dim
            for i in range(nc.depth):
                p[f"layer_{i}"] = {"kernel": kinit(gen, (d, nc.width)), "bias": binit(gen, (nc.width,))}
                d = nc.width
            if self._multihead:
                p["VmapDense_0"] = {"kernel": uniform(bound)(gen, (heads, nc.width, head_dim)), "bias": uniform(bound)(gen, (heads, head_dim))}
            else:
                p[f"layer_{nc.depth}"] = {"kernel": uniform(bound)(gen, (nc.width, head_dim)), "bias": uniform(bound)(gen, (head_dim,))}
                if nc.use_layer_norm:   # flax LayerNorm: scale = ones, bias = zeros (mtrl/nn/base.py:35-37, 52-53)
                    for k in range(nc.depth):
                        p[f"LayerNorm_{k}"] = {"scale": torch.ones(nc.width), "bias": torch.zeros(nc.width)}
            return p
        self._inner = (lambda t: t["params"][net_name]) if self._multihead else (lambda t: t["params"][net_name]["MLP_0"])
        _tree_copy_(self._inner(self.policy.params), init(pnc, 2 * act_dim, 1e-3))      # networks.py:33-34
        _tree_copy_(self._inner(self.value_function.params), init(vnc, 1, 3e-3))        # networks.py:199-200
        bufs = PpoBuffersC(**{k: v.data_ptr() for k, v in self._flat.items()}, steps=self._steps.data_ptr(),
                           logs=self._logs.data_ptr(), workspace=self._workspace.data_ptr())
        h = _vp()
        L.check(L.lib().mtrl_ppo_create(C.byref(h), C.byref(self._cfg), C.byref(bufs)))
        self._h = h
        return self

    def __del__(self):
        if getattr(self, "_h", None) is not None and L._lib is not None:
            L._lib.mtrl_ppo_destroy(self._h)
            self._h = None

    def refresh(self) -> None:
        L.check(L.lib().mtrl_ppo_refresh_shadows(self._h, _vp(L.current_stream_ptr())))

    def _dev(self, x, cols):
        t = x if isinstance(x, torch.Tensor) else torch.as_tensor(np.ascontiguousarray(x, dtype=np.float32))
        return t.to(device=self.device, dtype=torch.float32).reshape(-1, cols).contiguous()

    def update(self, data: Rollout, eps=None):
        """mtppo.py:319-330.  `data` is a Rollout whose arrays are (task, timestep, dim) or already flattened
        task-major; log_probs / advantages / returns / values must be present (OnPolicyAlgorithm fills them)."""
        c = self._cfg
        B = self.num_tasks * self.rollout_steps
        obs = self._dev(data.observations, c.obs_dim)
        assert obs.shape[0] == B, f"rollout has {obs.shape[0]} rows, expected num_tasks * rollout_steps = {B}"
        lp, adv, ret, val = (self._dev(x, 1) for x in (data.log_probs, data.advantages, data.returns, data.values))
        e = self._dev(eps, c.action_dim) if eps is not None else None
        L.check(L.lib().mtrl_ppo_update(self._h, _vp(obs.data_ptr()), _vp(lp.data_ptr()), _vp(adv.data_ptr()), _vp(ret.data_ptr()),
                                        _vp(val.data_ptr()), _vp(e.data_ptr() if e is not None else None),
                                        _vp(L.current_stream_ptr())))
        return self, {k: self._logs[i] for i, k in enumerate(PPO_LOG_KEYS)}

    def launches_per_update(self) -> int:
        return int(L.lib().mtrl_ppo_launches_per_update(self._h))
