"""Mirror of `mtrl.rl.algorithms.mtsac.MTSAC` (/root/reference/mtrl/rl/algorithms/mtsac.py:116-311,
1173-1251) on top of the fused CUDA update (csrc/sac.cu, C-ABI `mtrl_sac_*`).

Same surface: `MTSACConfig`, `MTSAC.initialize(config, env_config, seed)`, `update(data) -> (self,
logs)`, `get_num_params()`, state fields `actor`, `critic` (with `.target_params`), `alpha`, each a
TrainState-like object with `step`, `params`, `opt_state`.  Parameter trees carry the Flax names
(`layer_i/kernel (in, W)`, `VmapDense_0/kernel (T, W, head)`, critic under `VmapQValueFunction_0`
with a leading ensemble axis) as zero-copy views of the flat device buffers the kernels use.

Deliberate differences: the update is in place and returns `self` (the
reference returns a new immutable pytree); noise comes from in-kernel Philox unless `eps_c/eps_a`
are passed (jax.random streams cannot be reproduced without JAX).
"""
from __future__ import annotations

import ctypes as C
import dataclasses
import math
import os
from dataclasses import dataclass

import numpy as np
import torch

from ... import _lib as L
from ...config.networks import ContinuousActionPolicyConfig, QValueFunctionConfig
from ...config.nn import MultiHeadConfig
from ...config.optim import OptimizerConfig
from ...config.rl import AlgorithmConfig
from ...nn import get_nn_arch_for_config
from ...nn.multi_head import uniform
from ...types import ReplayBufferSamples

MAX_DEPTH = 4
_GRAPH_DEFAULT = os.environ.get("MTRL_UPDATE_GRAPH", "1") != "0"
LOG_KEYS = (
    "losses/qf_values", "losses/qf_loss", "metrics/critic_grad_magnitude", "metrics/critic_params_norm",
    "losses/actor_loss", "metrics/actor_grad_magnitude", "metrics/actor_params_norm", "metrics/explore_loss",
    "losses/alpha_loss", "alpha",
)  # mtsac.py:616-621, 704-709, 728-731 -- also the order of MTRL_LOG_* in include/mtrl_b200.h


class SacConfigC(C.Structure):
    _fields_ = [
        ("num_tasks", C.c_int), ("task_begin", C.c_int), ("num_local_tasks", C.c_int), ("obs_dim", C.c_int),
        ("action_dim", C.c_int), ("width", C.c_int), ("depth", C.c_int), ("num_critics", C.c_int),
        ("max_rows", C.c_int), ("max_batch", C.c_int),
        ("gamma", C.c_float), ("tau", C.c_float),
        ("actor_lr", C.c_float), ("critic_lr", C.c_float), ("alpha_lr", C.c_float),
        ("adam_b1", C.c_float), ("adam_b2", C.c_float), ("adam_eps", C.c_float),
        ("actor_max_grad_norm", C.c_float), ("critic_max_grad_norm", C.c_float), ("alpha_max_grad_norm", C.c_float),
        ("log_std_min", C.c_float), ("log_std_max", C.c_float), ("target_entropy", C.c_float),
        ("clip_q", C.c_int), ("use_task_weights", C.c_int), ("noise_seed", C.c_ulonglong), ("variant", C.c_int),
        ("precision", C.c_int), ("use_layer_norm", C.c_int), ("use_skip_connections", C.c_int),
    ]


PRECISIONS = {"tf32": 0, "fp32x3": 1}   # MTRL_PRECISION_* (include/mtrl_b200.h)


def precision_code(precision: str | None) -> int:
    """`precision` argument of the agents' initialize(): "tf32" (default; MTRL_PRECISION overrides the default) rounds the
    trunk-GEMM operands to tf32 as XLA does for f32 dots on NVIDIA GPUs; "fp32x3" splits every operand into two tf32
    values and runs three tensor-core passes, the numerics of the reference's fp32 CPU path at 3x the GEMM time."""
    name = precision if precision is not None else os.environ.get("MTRL_PRECISION", "tf32")
    if name not in PRECISIONS:
        raise ValueError(f"precision must be one of {sorted(PRECISIONS)}, got {name!r}")
    return PRECISIONS[name]


class NetLayoutC(C.Structure):
    _fields_ = [
        ("total", C.c_longlong), ("trunk_total", C.c_longlong), ("slots_off", C.c_longlong), ("heads_base", C.c_longlong),
        ("member_trunk_stride", C.c_longlong), ("member_head_stride", C.c_longlong),
        ("kernel_off", C.c_longlong * MAX_DEPTH), ("bias_off", C.c_longlong * MAX_DEPTH),
        ("head_kernel_off", C.c_longlong), ("head_bias_off", C.c_longlong),
        ("in_dim", C.c_int), ("head_dim", C.c_int), ("members", C.c_int), ("num_local_tasks", C.c_int),
        ("width", C.c_int), ("depth", C.c_int),
        ("ln_scale_off", C.c_longlong * MAX_DEPTH), ("ln_bias_off", C.c_longlong * MAX_DEPTH),
        ("use_layer_norm", C.c_int), ("reserved", C.c_int),
    ]


class SacLayoutC(C.Structure):
    _fields_ = [("actor", NetLayoutC), ("critic", NetLayoutC), ("workspace_bytes", C.c_longlong),
                ("k_actor", C.c_int), ("k_critic", C.c_int)]


class SacBuffersC(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in (
        "actor_params", "actor_grads", "actor_m", "actor_v", "actor_shadow",
        "critic_params", "critic_grads", "critic_m", "critic_v", "critic_shadow", "critic_target", "critic_target_shadow",
        "log_alpha", "alpha_m", "alpha_v", "steps", "logs", "workspace")]


_vp, _i = C.c_void_p, C.c_int
L._EXTRA_DECLS.update({
    "mtrl_sac_query_layout": ([C.POINTER(SacConfigC), C.POINTER(SacLayoutC)],),
    "mtrl_sac_create": ([C.POINTER(_vp), C.POINTER(SacConfigC), C.POINTER(SacBuffersC)],),
    "mtrl_sac_destroy": ([_vp], None),
    "mtrl_sac_refresh_shadows": ([_vp, _vp],),
    "mtrl_sac_update": ([_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _vp, _vp, _vp],),
    "mtrl_sac_phase1_critic_grads": ([_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _vp, _vp, _vp],),
    "mtrl_sac_phase2_critic_step_actor_grads": ([_vp, _vp],),
    "mtrl_sac_phase3_actor_step_alpha": ([_vp, _vp],),
    "mtrl_sac_launches_per_update": ([_vp],),
    "mtrl_sac_read_status_async": ([_vp, _vp, _vp],),
    "mtrl_sac_act": ([_vp, _vp, _i, _vp, _i, _vp, _vp],),
    "mtrl_mlp_forward": ([_vp, _i, _vp, _vp, _i, _vp, _vp],),
    "mtrl_sac_task_grads": ([_vp, _vp, _vp, _vp, _vp, _vp, _i, _vp, _vp, _vp, _vp, _vp],),
    "mtrl_sac_enable_pcgrad": ([_vp, _i, _i, _vp, _vp, _vp, _vp, _vp],),
    "mtrl_sac_enable_cagrad": ([_vp, _i, _i, _vp, _vp, _vp],),
    "mtrl_sac_enable_gradnorm": ([_vp, _i, _i, _i, _vp, _vp, _vp],),
    "mtrl_sac_enable_dummy": ([_vp, _i, _i, _vp, _vp, _vp],),
    "mtrl_task_gram": ([_vp, C.c_longlong, _i, C.c_longlong, _vp, _vp],),
    "mtrl_task_elementwise": ([_vp, C.c_longlong, _i, C.c_longlong, C.c_float, C.c_float, C.c_float, _vp, _vp, _vp],),
    "mtrl_task_abs_order_stats": ([_vp, C.c_longlong, _i, C.c_longlong, C.POINTER(C.c_longlong), _vp, _vp, _vp],),
    "mtrl_task_support_pairs": ([_vp, C.c_longlong, _i, C.c_longlong, _vp, _vp, _vp],),
    "mtrl_sac_trunk_owner_mask": ([_vp, _i, _vp],),
    "mtrl_sac_profile_gemms": ([_vp, _i],),
    "mtrl_sac_profile_read": ([_vp, C.POINTER(C.c_double), C.POINTER(_i)],),
    "mtrl_sac_profile_exchange": ([_vp, C.POINTER(C.c_double), C.POINTER(_i)],),
    "mtrl_sac_profile_classes": ([_vp, C.POINTER(C.c_double), C.POINTER(_i)],),
    "mtrl_comm_create": ([C.POINTER(_vp), _i, _i, C.c_longlong, _vp],),
    "mtrl_comm_arena": ([_vp], _vp),
    "mtrl_comm_open_peers": ([_vp, _vp],),
    "mtrl_comm_error": ([_vp, C.POINTER(_i)],),
    "mtrl_comm_destroy": ([_vp], None),
    "mtrl_comm_phase_times": ([_vp, C.POINTER(C.c_double)],),
    "mtrl_memcpy_h2d_batch": ([_i, C.POINTER(_vp), C.POINTER(_vp), C.POINTER(C.c_longlong), _vp],),
    "mtrl_comm_mc_supported": ([C.POINTER(_i)],),
    "mtrl_comm_mc_create": ([_vp, C.c_longlong, C.POINTER(_i)],),
    "mtrl_comm_mc_import": ([_vp, C.c_longlong, _i],),
    "mtrl_comm_mc_add_device": ([_vp],),
    "mtrl_comm_mc_bind": ([_vp],),
    "mtrl_comm_mc_local": ([_vp], _vp),
    "mtrl_comm_mc_ptr": ([_vp], _vp),
    "mtrl_comm_mc_bytes": ([_vp], C.c_longlong),
    "mtrl_sac_attach_comm": ([_vp, _vp, C.c_longlong, C.c_longlong, C.c_longlong, C.c_longlong],),
})

COMM_HEADER_BYTES = 4096
IPC_HANDLE_BYTES = 64


class _DeviceSpan:
    """A raw device allocation (the exchange arena) exposed to torch through __cuda_array_interface__."""

    def __init__(self, ptr: int, n_floats: int, owner):
        self.__cuda_array_interface__ = {"shape": (int(n_floats),), "typestr": "<f4", "data": (int(ptr), False), "version": 2}
        self._owner = owner


def arena_plan(total_critic: int, total_actor: int) -> tuple[dict, int]:
    """Byte offsets of the four exchanged regions inside every rank's arena (identical on all ranks: sized for the
    rank with the most local tasks) and the arena size."""
    off, plan = COMM_HEADER_BYTES, {}
    for name, n in (("critic_grads", total_critic), ("actor_grads", total_actor), ("critic_params", total_critic),
                    ("actor_params", total_actor)):
        plan[name] = off
        off = -(-(off + 4 * n) // 4096) * 4096
    return plan, off


@dataclasses.dataclass(frozen=True)
class MTSACConfig(AlgorithmConfig):  # mtsac.py:116-127
    actor_config: ContinuousActionPolicyConfig = ContinuousActionPolicyConfig()
    critic_config: QValueFunctionConfig = QValueFunctionConfig()
    temperature_optimizer_config: OptimizerConfig = OptimizerConfig(max_grad_norm=None)
    initial_temperature: float = 1.0
    num_critics: int = 2
    tau: float = 0.005
    use_task_weights: bool = False
    v_min: float = -10.0
    v_max: float = 10.0
    n_atoms: int = 51


@dataclass
class TrainState:
    """Shape of flax.training.train_state.TrainState as the reference uses it (algorithms/utils.py:11-46):
    `step`, `params`, `opt_state` = (clip state, (ScaleByAdamState(count, mu, nu), ...)) flattened to a dict."""
    step: torch.Tensor
    params: dict
    opt_state: dict
    tx: object = None
    apply_fn: object = None
    target_params: dict | None = None
    grads: dict | None = None   # not in the reference: the raw (pre-clip) gradients of the last update


def task_partition(num_tasks: int, world_size: int) -> list[tuple[int, int]]:
    """Contiguous task blocks, sizes ceil/floor(T/G): MT50 over 8 -> 7,7,6,6,6,6,6,6."""
    base, extra = divmod(num_tasks, world_size)
    out, start = [], 0
    for r in range(world_size):
        n = base + (1 if r < extra else 0)
        out.append((start, start + n))
        start += n
    return out


SUMMABLE_LOGS = (0, 1, 4, 8, 9)  # per-rank partial sums: qf_values, qf_loss, actor_loss, alpha_loss, alpha


def combine_rank_logs(local: torch.Tensor, summed: torch.Tensor) -> dict:
    """Turn one rank's 16-float log vector and its sum over ranks into the reference's ten scalars.
    Loss terms are partial sums already divided by the GLOBAL batch (summable); gradient norms are global on every
    rank after the trunk all-reduce (take the local value); parameter norms combine the replicated trunk part
    (slots 10 / 12, local) with the sum of every rank's head part (slots 11 / 13)."""
    out = {}
    for i, k in enumerate(LOG_KEYS):
        out[k] = summed[i] if i in SUMMABLE_LOGS else local[i]
    out["metrics/critic_params_norm"] = torch.sqrt(local[10] + summed[11])
    out["metrics/actor_params_norm"] = torch.sqrt(local[12] + summed[13])
    return out


def _views(flat: torch.Tensor, lay: NetLayoutC, in_dim: int, ensemble: bool) -> dict:
    """Flax-named zero-copy views of one network's flat buffer."""
    W, D, T, E, hd = lay.width, lay.depth, lay.num_local_tasks, lay.members, lay.head_dim
    tree = {}
    o0 = flat.storage_offset()   # as_strided offsets are absolute in the storage: `flat` may be a row of a matrix
    d = in_dim
    for i in range(D):
        k = flat.as_strided((E, d, W), (lay.member_trunk_stride, W, 1), o0 + lay.kernel_off[i])
        b = flat.as_strided((E, W), (lay.member_trunk_stride, 1), o0 + lay.bias_off[i])
        tree[f"layer_{i}"] = {"kernel": k if ensemble else k[0], "bias": b if ensemble else b[0]}
        d = W
    hk = flat.as_strided((E, T, W, hd), (lay.member_head_stride, W * hd, hd, 1), o0 + lay.heads_base + lay.head_kernel_off)
    hb = flat.as_strided((E, T, hd), (lay.member_head_stride, hd, 1), o0 + lay.heads_base + lay.head_bias_off)
    tree["VmapDense_0"] = {"kernel": hk if ensemble else hk[0], "bias": hb if ensemble else hb[0]}
    return tree


def _wrap(tree: dict, ensemble: bool) -> dict:
    inner = {"MultiHeadNetwork_0": tree}
    return {"params": {"VmapQValueFunction_0": inner} if ensemble else inner}


def _tree_copy_(dst: dict, src: dict) -> None:
    for k, v in src.items():
        if isinstance(v, dict):
            _tree_copy_(dst[k], v)
        else:
            dst[k].copy_(torch.as_tensor(np.asarray(v) if not isinstance(v, torch.Tensor) else v).to(dst[k].dtype))


class MTSAC:
    """See module docstring."""

    LOG_KEYS = LOG_KEYS

    def __init__(self):
        raise TypeError("use MTSAC.initialize(config, env_config, seed)")

    # ------------------------------------------------------------------ construction
    @staticmethod
    def initialize(config: MTSACConfig, env_config, seed: int = 1, *, max_batch: int | None = None,
                   max_rows: int | None = None, rank: int = 0, world_size: int = 1, process_group=None,
                   device: str | torch.device | None = None, exchange: str = "p2p", precision: str | None = None) -> "MTSAC":
        """mtsac.py:152-284.  `env_config` needs `.observation_space.shape` and `.action_space.shape`
        (the observation includes the one-hot task id).  `max_batch` bounds the rows one update may
        pass (default 128 per local task, the reference's batch).  rank/world_size shard the tasks;
        `exchange` picks how the ranks sum trunk gradients: "p2p" = the fused peer-memory kernel
        (csrc/comm.cuh, sharded Adam, no library collective), "nccl" = all-reduce between the phases.
        `precision`: "tf32" or "fp32x3" (see `precision_code`)."""
        if exchange not in ("p2p", "nccl", "local"):
            raise ValueError(f"exchange must be 'p2p', 'nccl' or 'local', got {exchange!r}")
        # "local" skips the exchange altogether: one rank's share of the kernels for profiling on a single GPU
        # (scripts/shard_profile.sh); its numbers are not the sharded update's.
        if not torch.cuda.is_available():
            raise L.MtrlError("MTSAC needs a CUDA device; there is no CPU fallback")
        self = object.__new__(MTSAC)
        dev = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self.device = dev
        self.config = config
        self.num_tasks = config.num_tasks
        self.rank, self.world_size, self.process_group = rank, world_size, process_group
        self.exchange = exchange if world_size > 1 else "none"
        self._comm = None
        self.task_begin, self.task_end = task_partition(config.num_tasks, world_size)[rank]
        t_local = self.task_end - self.task_begin
        obs_dim = int(np.prod(env_config.observation_space.shape))
        act_dim = int(np.prod(env_config.action_space.shape))
        for nc in (config.actor_config.network_config, config.critic_config.network_config):
            if type(nc) is not MultiHeadConfig:
                get_nn_arch_for_config(nc)  # raises like the reference for the bare base config
                raise NotImplementedError(f"{type(nc).__name__}: the accelerated MT-SAC path covers MultiHeadConfig")
            assert nc.num_tasks == config.num_tasks
        anc, cnc = config.actor_config.network_config, config.critic_config.network_config
        if (anc.width, anc.depth) != (cnc.width, cnc.depth):
            raise NotImplementedError("actor and critic must share width and depth (true for every reference experiment)")
        if config.critic_config.use_classification or not config.actor_config.squash_tanh:
            raise NotImplementedError("c51 critic / unsquashed policy are outside the accelerated path")
        a_opt, c_opt = anc.optimizer.spawn(), cnc.optimizer.spawn()
        t_opt = config.temperature_optimizer_config.spawn()
        if (a_opt.eps, a_opt.b1, a_opt.b2) != (c_opt.eps, c_opt.b1, c_opt.b2) or a_opt.eps != t_opt.eps:
            raise NotImplementedError("one Adam eps / betas for actor, critic and temperature")
        if max_batch is None:
            max_batch = 128 * t_local
        if max_rows is None:
            max_rows = -(-max_batch // 128) * 128 + 128 * (t_local - 1) if max_batch != 128 * t_local else 128 * t_local
        self.gamma, self.tau = config.gamma, config.tau
        self.target_entropy = -float(act_dim)  # mtsac.py:258
        self.use_task_weights, self.clip, self.num_critics = config.use_task_weights, config.clip, config.num_critics
        self.split_actor_losses = self.split_critic_losses = False
        self.explore = False
        self.actor_network_type, self.critic_network_type = "vanilla", "mse"
        nm = lambda v: -1.0 if v is None else float(v)  # noqa: E731
        self._cfg = SacConfigC(
            num_tasks=config.num_tasks, task_begin=self.task_begin, num_local_tasks=t_local, obs_dim=obs_dim,
            action_dim=act_dim, width=anc.width, depth=anc.depth, num_critics=config.num_critics, max_rows=max_rows,
            max_batch=max_batch, gamma=config.gamma, tau=config.tau, actor_lr=a_opt.lr, critic_lr=c_opt.lr,
            alpha_lr=t_opt.lr, adam_b1=a_opt.b1, adam_b2=a_opt.b2, adam_eps=a_opt.eps,
            actor_max_grad_norm=nm(a_opt.max_grad_norm), critic_max_grad_norm=nm(c_opt.max_grad_norm),
            alpha_max_grad_norm=nm(t_opt.max_grad_norm), log_std_min=config.actor_config.log_std_min,
            log_std_max=config.actor_config.log_std_max, target_entropy=self.target_entropy,
            clip_q=int(config.clip), use_task_weights=int(config.use_task_weights), noise_seed=int(seed) & (2**63 - 1),
            precision=precision_code(precision))
        self.precision = "fp32x3" if self._cfg.precision else "tf32"
        self._allocate(dev, t_local)
        lay = self._lay

        # parameter views with the Flax names (mtsac.py:203-246 prints these trees)
        def ts(prefix, l, in_dim, ens, tx, step_idx):
            v = lambda name: _wrap(_views(self._flat[f"{prefix}_{name}"], l, in_dim, ens), ens)  # noqa: E731
            return TrainState(step=self._steps[step_idx], params=v("params"),
                              opt_state={"count": self._steps[step_idx], "mu": v("m"), "nu": v("v")}, tx=tx,
                              target_params=v("target") if ens else None, grads=v("grads"))
        self.actor = ts("actor", lay.actor, obs_dim, False, a_opt, 0)
        self.critic = ts("critic", lay.critic, act_dim + obs_dim, True, c_opt, 1)
        la = self._flat["log_alpha"][:t_local]
        self.alpha = TrainState(step=self._steps[2], params={"params": {"log_alpha": la}},
                                opt_state={"count": self._steps[2], "mu": {"params": {"log_alpha": self._flat["alpha_m"][:t_local]}},
                                           "nu": {"params": {"log_alpha": self._flat["alpha_v"][:t_local]}}}, tx=t_opt)

        # initial weights: same distributions as the reference (he_uniform trunk, zero bias, heads U(+-1e-3) / U(+-3e-3),
        # networks.py:33-34, 65-66; log_alpha = log(initial_temperature), mtsac.py:52-58).  Every rank draws the full
        # T-head tensors from the same seed and keeps its slice, so all ranks agree on the replicated trunk.
        gen = torch.Generator().manual_seed(int(seed))
        a_net = get_nn_arch_for_config(anc)(config=anc, head_dim=2 * act_dim, head_kernel_init=uniform(1e-3),
                                            head_bias_init=uniform(1e-3))
        c_net = get_nn_arch_for_config(cnc)(config=cnc, head_dim=1, head_kernel_init=uniform(3e-3),
                                            head_bias_init=uniform(3e-3))
        a_init = a_net.init(gen, obs_dim)
        c_init = c_net.init(gen, act_dim + obs_dim, ensemble=config.num_critics)
        sl = slice(self.task_begin, self.task_end)
        a_init["VmapDense_0"] = {k: v[sl] for k, v in a_init["VmapDense_0"].items()}
        c_init["VmapDense_0"] = {k: v[:, sl] for k, v in c_init["VmapDense_0"].items()}
        _tree_copy_(self.actor.params["params"]["MultiHeadNetwork_0"], a_init)
        _tree_copy_(self.critic.params["params"]["VmapQValueFunction_0"]["MultiHeadNetwork_0"], c_init)
        self._flat["critic_target"].copy_(self._flat["critic_params"])  # target_params=critic_init_params, mtsac.py:238-242
        la.fill_(math.log(config.initial_temperature))

        self._create_handle()
        surg = lambda o: bool(o.pcgrad or o.cagrad or o.gradnorm or o.dummy)  # noqa: E731
        self._pcgrad = (surg(c_opt), surg(a_opt))   # (critic, actor) start their chain with a multi-task transformation
        kinds = {k for o in (c_opt, a_opt) for k in ("pcgrad", "cagrad", "gradnorm", "dummy") if getattr(o, k)}
        self._gradnorm_clip = bool(c_opt.gradnorm_clip_per_task or a_opt.gradnorm_clip_per_task)
        if len(kinds) > 1:
            raise NotImplementedError("one multi-task optimiser kind per agent (PCGradConfig, CAGradConfig or GradNormConfig)")
        self._surgery = next(iter(kinds), None)
        if any(self._pcgrad):
            if world_size != 1:
                raise NotImplementedError("multi-task optimiser configs need every task on one device (split losses)")
            if config.num_tasks > 64:
                raise NotImplementedError("multi-task optimiser configs: at most 64 tasks")
            # split_actor_losses / split_critic_losses (mtsac.py:272-273)
            self.split_critic_losses, self.split_actor_losses = self._pcgrad
            self._enable_pcgrad(seed)
        return self

    def _task_matrices(self) -> dict:
        """(T, P) work matrices of the per-task gradient path, allocated on first use."""
        if getattr(self, "_tg", None) is None:
            T = self.num_tasks
            self._tg = {"critic": torch.zeros(T, self._lay.critic.total, dtype=torch.float32, device=self.device),
                        "actor": torch.zeros(T, self._lay.actor.total, dtype=torch.float32, device=self.device)}
        return self._tg

    def _enable_pcgrad(self, seed: int) -> None:
        T = self.num_tasks
        tg = self._task_matrices()
        self._pc_scratch = torch.zeros(2 * T * T + 4 * T + 8, dtype=torch.float32, device=self.device)
        self._pc_perm = torch.arange(T, dtype=torch.int32, device=self.device).repeat(2, 1).contiguous()
        self._pc_gen = torch.Generator().manual_seed(int(seed) + 7919)
        if self._surgery == "gradnorm":
            L.check(L.lib().mtrl_sac_enable_gradnorm(self._h, int(self._pcgrad[0]), int(self._pcgrad[1]), int(self._gradnorm_clip),
                                                     _vp(tg["critic"].data_ptr()), _vp(tg["actor"].data_ptr()),
                                                     _vp(self._pc_scratch.data_ptr())))
            return
        if self._surgery == "dummy":
            L.check(L.lib().mtrl_sac_enable_dummy(self._h, int(self._pcgrad[0]), int(self._pcgrad[1]), _vp(tg["critic"].data_ptr()),
                                                  _vp(tg["actor"].data_ptr()), _vp(self._pc_scratch.data_ptr())))
            return
        if self._surgery == "cagrad":
            L.check(L.lib().mtrl_sac_enable_cagrad(self._h, int(self._pcgrad[0]), int(self._pcgrad[1]), _vp(tg["critic"].data_ptr()),
                                                   _vp(tg["actor"].data_ptr()), _vp(self._pc_scratch.data_ptr())))
            return
        L.check(L.lib().mtrl_sac_enable_pcgrad(self._h, int(self._pcgrad[0]), int(self._pcgrad[1]), _vp(tg["critic"].data_ptr()),
                                               _vp(tg["actor"].data_ptr()), _vp(self._pc_scratch.data_ptr()),
                                               _vp(self._pc_perm[0].data_ptr()), _vp(self._pc_perm[1].data_ptr())))

    def pcgrad_stats(self) -> dict:
        """PCGradState of the last update per network (pcgrad.py:12-17, 83-90) as device scalars, plus the norm of the
        plain mean gradient."""
        T, s = self.num_tasks, self._pc_scratch
        base = 2 * T * T + 2 * T
        if self._surgery == "dummy":    # the reference's dummy state is {} (dummy.py:8-10); the combined-gradient norm is a by-product
            return {net: {"grad_magnitude": s[base + 4 * i]} for i, net in enumerate(("critic", "actor")) if self._pcgrad[i]}
        if self._surgery == "gradnorm":
            names = ("grad_magnitude", "avg_grad_magnitude_per_task")
            return {net: dict(zip(names, s[base + 4 * i: base + 4 * i + 2]), task_weights=torch.ones(T, device=self.device))
                    for i, net in enumerate(("critic", "actor")) if self._pcgrad[i]}
        if self._surgery == "cagrad":   # CAGradState (cagrad.py:13-18, 195-203)
            names = ("avg_grad_magnitude", "avg_grad_magnitude_before_surgery", "cagrad_objective")
            return {net: dict(zip(names, s[base + 4 * i: base + 4 * i + 3]), task_weights=s[base + 8 + T * i: base + 8 + T * (i + 1)])
                    for i, net in enumerate(("critic", "actor")) if self._pcgrad[i]}
        names = ("n_grad_conflicts", "avg_grad_magnitude", "avg_grad_magnitude_before_surgery", "mean_grad_norm")
        return {net: dict(zip(names, s[base + 4 * i: base + 4 * i + 4])) for i, net in enumerate(("critic", "actor")) if self._pcgrad[i]}

    def _allocate(self, dev, t_local: int) -> None:
        """Ask the library for the flat layouts of self._cfg and allocate every device buffer it needs."""
        lay = SacLayoutC()
        L.check(L.lib().mtrl_sac_query_layout(C.byref(self._cfg), C.byref(lay)))
        self._lay = lay
        z = lambda n, dt=torch.float32: torch.zeros(int(n), dtype=dt, device=dev)  # noqa: E731
        self._flat = {f"actor_{k}": z(lay.actor.total) for k in ("params", "grads", "m", "v", "shadow")}
        self._flat.update({f"critic_{k}": z(lay.critic.total)
                           for k in ("params", "grads", "m", "v", "shadow", "target", "target_shadow")})
        if getattr(self, "exchange", "none") == "p2p":
            self._flat.update(self._open_arena(dev))
        self._flat.update({"log_alpha": z(max(t_local, 4)), "alpha_m": z(max(t_local, 4)), "alpha_v": z(max(t_local, 4))})
        self._steps = z(4, torch.int32)
        self._logs = z(16)
        self._workspace = z((lay.workspace_bytes + 3) // 4 + 64)
        self._status_host = torch.zeros(4, dtype=torch.int32).pin_memory()

    def _open_arena(self, dev) -> dict:
        """Exchange arena for the fused peer-memory path: gradients and parameters of both networks live in one
        cudaMalloc block per rank that every other rank maps through CUDA IPC (handles all-gathered over the
        process group, which is the only use of torch.distributed on this path besides the log all-reduce)."""
        import torch.distributed as dist

        # every rank sizes the regions for the largest shard so the offsets agree
        cfg_max = SacConfigC.from_buffer_copy(self._cfg)
        cfg_max.num_local_tasks = -(-self.num_tasks // self.world_size)
        cfg_max.task_begin = 0
        cfg_max.max_batch = min(cfg_max.max_batch, cfg_max.max_rows)
        lay_max = SacLayoutC()
        L.check(L.lib().mtrl_sac_query_layout(C.byref(cfg_max), C.byref(lay_max)))
        plan, nbytes = arena_plan(lay_max.critic.total, lay_max.actor.total)
        comm, handle = _vp(), (C.c_ubyte * IPC_HANDLE_BYTES)()
        L.check(L.lib().mtrl_comm_create(C.byref(comm), self.rank, self.world_size, nbytes, handle))
        self._comm = comm
        mine = torch.tensor(list(handle), dtype=torch.uint8, device=dev)
        allh = [torch.empty_like(mine) for _ in range(self.world_size)]
        dist.all_gather(allh, mine, group=self.process_group)
        blob = bytes(torch.stack(allh).cpu().flatten().tolist())
        L.check(L.lib().mtrl_comm_open_peers(comm, C.cast(C.c_char_p(blob), _vp)))
        base = L.lib().mtrl_comm_arena(comm)
        self._arena_plan = dict(plan)
        lay = self._lay
        out = {}
        for name, n in (("critic_grads", lay.critic.total), ("actor_grads", lay.actor.total),
                        ("critic_params", lay.critic.total), ("actor_params", lay.actor.total)):
            out[name] = torch.as_tensor(_DeviceSpan(base + plan[name], n, self), device=dev)
        # parameter all-gather through the NVSwitch multicast engine where the box has one: the parameters then live in
        # the multicast-bound region instead of the arena
        mc_actor_off = -(-(4 * lay_max.critic.total) // 4096) * 4096
        mc_local = self._open_multicast(dev, mc_actor_off + 4 * lay_max.actor.total)
        self.multicast = mc_local is not None
        if mc_local is not None:
            self._arena_plan["critic_params"], self._arena_plan["actor_params"] = 0, mc_actor_off
            out["critic_params"] = torch.as_tensor(_DeviceSpan(mc_local, lay.critic.total, self), device=dev)
            out["actor_params"] = torch.as_tensor(_DeviceSpan(mc_local + mc_actor_off, lay.actor.total, self), device=dev)
        dist.barrier(group=self.process_group)
        return out

    def _open_multicast(self, dev, nbytes: int):
        """Set up the multicast region of csrc/comm.cu on every rank, or return None on ALL ranks (every stage ends with
        an agreement, so no rank is left on a different exchange path): the device attribute, the multicast object on
        rank 0, its file descriptor handed to the other ranks over an abstract unix socket (SCM_RIGHTS), add-device,
        bind + map.  On by default from 8 ranks up (MTRL_MULTICAST=1|0 forces it): measured on 2, 4 and 8 B200s it barely
        shortens the update -- the all-gather is bound by what every GPU RECEIVES (7/8 of the trunk = 60 MB per critic
        step at 8 GPUs, ~0.75 TB/s of NVLink ingress), which a multicast store does not reduce; it only cuts the sender's
        egress (Adam + stores 76 -> 29 us, wait for everyone's stores 36 -> 78 us; +0.5 % at 8 GPUs, -1...2 % at 2 and 4
        where the sender's own copy loops through the switch: profiles/r02_multicast.md)."""
        import os
        import socket

        import torch.distributed as dist

        lib, comm, group = L.lib(), self._comm, self.process_group

        def agree(ok: bool) -> bool:
            if not ok and want:
                import sys

                print(f"rank {self.rank}: no multicast region ({(lib.mtrl_last_error() or b'').decode()}); "
                      "parameters are all-gathered with one store per peer", file=sys.stderr, flush=True)
            t = torch.tensor([1 if ok else 0], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MIN, group=group)
            return bool(int(t.item()))

        sup = _i(0)
        env = os.environ.get("MTRL_MULTICAST")
        want = env == "1" or (env is None and self.world_size >= 8)
        ok = want and lib.mtrl_comm_mc_supported(C.byref(sup)) == 0 and sup.value != 0
        if not agree(ok):
            return None
        fd, srv, name = _i(-1), None, [None]
        if self.rank == 0:
            ok = lib.mtrl_comm_mc_create(comm, nbytes, C.byref(fd)) == 0
            if ok:
                try:
                    srv = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
                    name[0] = "\0mtrl_b200_mc_%d_%d" % (os.getpid(), id(self) & 0xFFFFFF)
                    srv.bind(name[0])
                    srv.listen(self.world_size)
                    srv.settimeout(60.0)
                except OSError:
                    ok = False
        if not agree(ok):
            return None
        dist.broadcast_object_list(name, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        ok = True
        try:
            if self.rank == 0:
                for _ in range(self.world_size - 1):
                    conn, _ = srv.accept()
                    socket.send_fds(conn, [b"m"], [fd.value])
                    conn.close()
                srv.close()
            else:
                s_ = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
                s_.settimeout(60.0)
                s_.connect(name[0])
                _, fds, _, _ = socket.recv_fds(s_, 16, 1)
                s_.close()
                ok = len(fds) == 1 and lib.mtrl_comm_mc_import(comm, nbytes, fds[0]) == 0
                for f in fds:
                    os.close(f)
        except OSError:
            ok = False
        if not agree(ok):
            return None
        if self.rank == 0:
            os.close(fd.value)
        if not agree(lib.mtrl_comm_mc_add_device(comm) == 0):   # every device is in before anyone binds
            return None
        if not agree(lib.mtrl_comm_mc_bind(comm) == 0):         # every rank is bound before anyone stores
            return None
        return lib.mtrl_comm_mc_local(comm)

    def _create_handle(self) -> None:
        bufs = SacBuffersC(**{k: v.data_ptr() for k, v in self._flat.items()}, steps=self._steps.data_ptr(),
                           logs=self._logs.data_ptr(), workspace=self._workspace.data_ptr())
        h = _vp()
        L.check(L.lib().mtrl_sac_create(C.byref(h), C.byref(self._cfg), C.byref(bufs)))
        self._h = h
        if getattr(self, "_comm", None) is not None:
            p = self._arena_plan
            L.check(L.lib().mtrl_sac_attach_comm(h, self._comm, p["critic_grads"], p["actor_grads"], p["critic_params"],
                                                 p["actor_params"]))
            # attaching re-times the backward GEMM plans, whose epilogues add into the peers' gradient buffers: nobody
            # may start an update before every rank is done with that
            import torch.distributed as dist

            torch.cuda.synchronize()
            dist.barrier(group=self.process_group)
        self._status_event = torch.cuda.Event()
        self._pending_status = False
        self._graphs, self._graph_seen, self._prof_enabled = {}, set(), False

    def __del__(self):
        if getattr(self, "_h", None) is not None and L._lib is not None:
            L._lib.mtrl_sac_destroy(self._h)
            self._h = None
        # the arena itself is left to process exit: peers may still have it mapped

    def exchange_error(self) -> int:
        """0, or the code of the in-kernel wait that timed out waiting for a peer (synchronises)."""
        if getattr(self, "_comm", None) is None:
            return 0
        code = _i()
        L.check(L.lib().mtrl_comm_error(self._comm, C.byref(code)))
        return code.value

    # ------------------------------------------------------------------ reference surface
    def get_num_params(self) -> dict[str, int]:
        """mtsac.py:289-296 (counts the logical parameters, not the alignment padding)."""
        c, T = self._cfg, self.num_tasks

        def count(in_dim, head):
            n, d = 0, in_dim
            for _ in range(c.depth):
                n += d * c.width + c.width
                d = c.width
            return n + T * (c.width * head + head)
        return {"actor_num_params": count(c.obs_dim, 2 * c.action_dim),
                "critic_num_params": c.num_critics * count(c.action_dim + c.obs_dim, 1)}

    def refresh(self) -> None:
        """Call after writing into `.params` / `.target_params` views directly (e.g. loading a checkpoint)."""
        L.check(L.lib().mtrl_sac_refresh_shadows(self._h, _vp(L.current_stream_ptr())))

    def _dev(self, x) -> torch.Tensor:
        if isinstance(x, torch.Tensor):
            t = x.to(device=self.device, dtype=torch.float32, non_blocking=True)
        else:
            t = torch.as_tensor(np.ascontiguousarray(x, dtype=np.float32)).to(self.device, non_blocking=True)
        return t.contiguous()

    def _check_status(self) -> None:
        if self._pending_status:
            self._status_event.synchronize()
            self._pending_status = False
            code = int(self._status_host[0])
            if int(self._status_host[1]) != 0:
                raise L.MtrlError(f"rank {self.rank}: the peer-memory exchange timed out waiting for another rank (code "
                                  f"{int(self._status_host[1])}); no rank applied the step in flight and the exchange is disabled -- "
                                  "restart the job (MTRL_COMM_TIMEOUT_S sets how long a straggler is waited for)")
            if code == 1:
                raise ValueError("update: a batch row belongs to a task outside this rank's range "
                                 f"[{self.task_begin}, {self.task_end})")
            if code == 2:
                raise ValueError("update: the batch does not fit max_rows (rows per task are padded to 128)")

    def update(self, data: ReplayBufferSamples, eps_c=None, eps_a=None, *, global_batch: int | None = None,
               check: bool = False, graph: bool | None = None, pcgrad_perm=None):
        """`MTSAC.update` (mtsac.py:1249-1251).  Returns (self, logs) with logs as 0-dim device tensors in
        the reference's keys; nothing here synchronises the host unless `check=True`.

        Launch path: the ~65 kernels of one update are captured into a CUDA graph the second time a batch shape is
        seen and replayed afterwards (inputs are first copied into fixed staging buffers -- for host batches that copy
        IS the H2D transfer), which removes the per-kernel launch gaps.  `graph=False` (or MTRL_UPDATE_GRAPH=0)
        launches kernel by kernel; that is also what happens while the caller is capturing a graph of its own."""
        capturing = torch.cuda.is_current_stream_capturing()
        if not capturing:
            self._check_status()
        if self.world_size > 1 and global_batch is None:
            raise ValueError("multi-GPU update needs global_batch (the B every loss mean divides by)")
        B = int(data[0].shape[0])
        assert data[0].shape[1] == self._cfg.obs_dim and tuple(data[1].shape) == (B, self._cfg.action_dim)
        if graph is None:
            graph = _GRAPH_DEFAULT
        if any(getattr(self, "_pcgrad", (False, False))) and self._surgery == "pcgrad":
            # the row permutation pcgrad draws every step (pcgrad.py:79; jax key there, a seeded host generator here),
            # one per network, or the caller's `pcgrad_perm = (critic_perm, actor_perm)`
            T = self.num_tasks
            perms = pcgrad_perm if pcgrad_perm is not None else (torch.randperm(T, generator=self._pc_gen),
                                                                 torch.randperm(T, generator=self._pc_gen))
            self._pc_perm.copy_(torch.stack([torch.as_tensor(x).to(torch.int32) for x in perms]), non_blocking=True)
        key = (B, eps_c is not None, global_batch)
        entry = None
        if self.world_size > 1 and self.exchange == "nccl":
            graph = False   # a graph holding NCCL collectives blocks the communicator's teardown
        if graph and not capturing and self._prof_enabled is False:
            entry = self._graphs.get(key)
            if entry is None and key in self._graph_seen and len(self._graphs) < 4:
                entry = self._capture(key, B, eps_c is not None, global_batch)
            self._graph_seen.add(key)
        if entry is not None:
            src = list(data) + ([eps_c, eps_a] if eps_c is not None else [])
            self._stage_inputs(entry["inputs"], src)
            entry["graph"].replay()
            logs = entry["logs"] if entry["logs"] is not None else self.logs()
        else:
            obs, act, nxt, done, rew = (self._dev(x) for x in data)
            ec = self._dev(eps_c) if eps_c is not None else None
            ea = self._dev(eps_a) if eps_a is not None else None
            self._launch_update((obs, act, nxt, done, rew), ec, ea, B, global_batch)
            logs = self.logs()
        # asynchronous status read-back (checked at the next call, or now if check=True); not while a CUDA graph
        # is being captured (the captured update is replayed without host involvement)
        if not capturing:
            stream = _vp(L.current_stream_ptr())
            L.check(L.lib().mtrl_sac_read_status_async(self._h, _vp(self._status_host.data_ptr()), stream))
            self._status_event.record()
            self._pending_status = True
            if check:
                self._check_status()
        return self, logs

    def _stage_inputs(self, staging, src) -> None:
        """Copy one batch into the captured graph's input buffers.  Device tensors: device-to-device copies.  Host arrays (CPU
        tensors, NumPy arrays -- what the reference's loop passes, base.py:220-221): ONE library call that enqueues all the
        host-to-device copies (mtrl_memcpy_h2d_batch); five framework-level copies cost more CPU time than the transfer takes."""
        dsts, srcs, sizes, keep = [], [], [], []
        for dst, x in zip(staging, src):
            if isinstance(x, torch.Tensor) and x.is_cuda:
                dst.copy_(x.reshape(dst.shape), non_blocking=True)
                continue
            if isinstance(x, torch.Tensor):
                t = x if (x.dtype == torch.float32 and x.is_contiguous()) else x.to(torch.float32).contiguous()
                ptr, n = t.data_ptr(), t.numel()
            else:
                t = np.ascontiguousarray(x, dtype=np.float32)
                ptr, n = t.ctypes.data, t.size
            if n != dst.numel():
                raise ValueError(f"batch array of {n} elements where {dst.numel()} are expected")
            keep.append(t)
            dsts.append(dst.data_ptr())
            srcs.append(ptr)
            sizes.append(4 * n)
        if dsts:
            k = len(dsts)
            L.check(L.lib().mtrl_memcpy_h2d_batch(k, (_vp * k)(*dsts), (_vp * k)(*srcs), (C.c_longlong * k)(*sizes),
                                                  _vp(L.current_stream_ptr())))
            self._staged_src = keep   # pinned sources are read asynchronously: keep them alive until the next batch replaces them

    def _launch_update(self, tensors, ec, ea, B: int, global_batch: int | None) -> None:
        obs, act, nxt, done, rew = tensors
        stream = _vp(L.current_stream_ptr())
        args = (self._h, _vp(obs.data_ptr()), _vp(act.data_ptr()), _vp(nxt.data_ptr()), _vp(done.data_ptr()),
                _vp(rew.data_ptr()), B)
        p = lambda t: _vp(t.data_ptr() if t is not None else None)  # noqa: E731
        if self.world_size == 1 or self.exchange in ("p2p", "local"):
            # one call: with several ranks the trunk-gradient exchange happens inside the fused kernels (comm.cuh)
            L.check(L.lib().mtrl_sac_update(*args, global_batch or B, p(ec), p(ea), stream))
        else:
            import torch.distributed as dist

            gb = global_batch
            L.check(L.lib().mtrl_sac_phase1_critic_grads(*args, gb, p(ec), p(ea), stream))
            lc, la = self._lay.critic, self._lay.actor
            dist.all_reduce(self._flat["critic_grads"][: lc.trunk_total + 32], group=self.process_group)
            L.check(L.lib().mtrl_sac_phase2_critic_step_actor_grads(self._h, stream))
            dist.all_reduce(self._flat["actor_grads"][: la.trunk_total + 32], group=self.process_group)
            L.check(L.lib().mtrl_sac_phase3_actor_step_alpha(self._h, stream))

    def _capture(self, key, B: int, has_eps: bool, global_batch: int | None) -> dict:
        """Capture one update on fixed staging buffers (the caller has already run this shape once kernel by kernel,
        so every kernel is loaded and nothing initialises lazily under capture)."""
        c = self._cfg
        dims = [c.obs_dim, c.action_dim, c.obs_dim, 1, 1] + ([c.action_dim, c.action_dim] if has_eps else [])
        inputs = [torch.zeros(B, d, dtype=torch.float32, device=self.device) for d in dims]
        torch.cuda.current_stream().synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self._launch_update(tuple(inputs[:5]), inputs[5] if has_eps else None, inputs[6] if has_eps else None, B, global_batch)
            # with several ranks the 16-float log all-reduce (NCCL) stays outside the graph
            logs = self.logs() if self.world_size == 1 else None
        entry = {"graph": g, "inputs": inputs, "logs": logs}
        self._graphs[key] = entry
        return entry

    def logs(self) -> dict:
        """The ten log scalars of the last update as 0-dim device tensors (keys of mtsac.py:616-621, 704-709, 728-731).
        With world_size > 1 the per-rank partial sums are combined here (one 16-float all-reduce)."""
        v = self._logs
        if self.world_size > 1 and self.exchange != "local":
            import torch.distributed as dist

            s = v.clone()
            dist.all_reduce(s, group=self.process_group)
            return combine_rank_logs(v, s)
        return {k: v[i] for i, k in enumerate(LOG_KEYS)}

    def launches_per_update(self) -> int:
        return int(L.lib().mtrl_sac_launches_per_update(self._h))

    def profile_gemms(self, enable: bool) -> None:
        self._prof_enabled = bool(enable)   # event bracketing needs kernel-by-kernel launches
        L.check(L.lib().mtrl_sac_profile_gemms(self._h, int(enable)))

    def profile_read(self) -> tuple[float, int]:
        """(sum of GEMM-launch durations in ms, number of GEMM launches) since profile_gemms(True)."""
        ms, n = C.c_double(), _i()
        L.check(L.lib().mtrl_sac_profile_read(self._h, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def exchange_phase_times(self) -> dict | None:
        """Microseconds spent in each phase of the last critic / actor trunk step on this rank (synchronises)."""
        if getattr(self, "_comm", None) is None:
            return None
        buf = (C.c_double * 14)()
        L.check(L.lib().mtrl_comm_phase_times(self._comm, buf))
        names = ("wait_grads", "norms", "norm_exchange", "adam_allgather", "wait_stores", "derived", "total")
        return {"critic": dict(zip(names, list(buf)[:7])), "actor": dict(zip(names, list(buf)[7:]))}

    PROFILE_CLASSES = ("gemm", "exchange", "adam_polyak", "head_vjp", "critic_loss", "actor_head", "actor_loss", "pack",
                       "grad_norms", "bias_colsum", "layernorm_junction", "unused")

    def profile_classes(self) -> dict:
        """{kernel class: (summed ms, launches)} of the event-bracketed pass, as of the last profile_read()."""
        ms, n = (C.c_double * 12)(), (_i * 12)()
        L.check(L.lib().mtrl_sac_profile_classes(self._h, ms, n))
        return {k: (ms[i], n[i]) for i, k in enumerate(self.PROFILE_CLASSES) if n[i]}

    def profile_exchange(self) -> tuple[float, int]:
        """(sum of exchange-kernel durations in ms, their count) as of the last profile_read()."""
        ms, n = C.c_double(), _i()
        L.check(L.lib().mtrl_sac_profile_exchange(self._h, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    # ------------------------------------------------------------------ per-task gradients
    def per_task_gradients(self, data: ReplayBufferSamples, eps_c=None, eps_a=None) -> dict:
        """The (num_tasks, num_params) per-task gradient matrices of `compute_weights` (mtsac.py:870-1170: batch split
        by task, `jax.vmap(jax.value_and_grad(loss))`), as CUDA tensors in the flat network layout:
        {"critic": (T, P_critic), "actor": (T, P_actor)}.  Row t is the gradient of task t's mean loss.  Parameters are
        not updated.  Needs all tasks on one GPU and the same number of rows per task (as the reference's reshape)."""
        if self.world_size != 1:
            raise NotImplementedError("per-task gradients need every task on one device")
        self._check_status()
        obs, act, nxt, done, rew = (self._dev(x) for x in data)
        B, T = obs.shape[0], self.num_tasks
        if B % T:
            raise ValueError(f"batch of {B} rows cannot be split evenly over {T} tasks (mtsac.py:325 reshapes the same way)")
        ec = self._dev(eps_c) if eps_c is not None else None
        ea = self._dev(eps_a) if eps_a is not None else None
        lc, la = self._lay.critic, self._lay.actor
        self._task_matrices()
        p = lambda t: _vp(t.data_ptr() if t is not None else None)  # noqa: E731
        stream = _vp(L.current_stream_ptr())
        L.check(L.lib().mtrl_sac_task_grads(self._h, p(obs), p(act), p(nxt), p(done), p(rew), B, p(ec), p(ea),
                                            p(self._tg["critic"]), p(self._tg["actor"]), stream))
        L.check(L.lib().mtrl_sac_read_status_async(self._h, _vp(self._status_host.data_ptr()), stream))
        self._status_event.record()
        self._status_event.synchronize()
        code = int(self._status_host[0])
        if code:
            raise ValueError({1: "a batch row belongs to a task this agent does not own", 2: "the batch does not fit max_rows",
                              3: "tasks have different numbers of rows (mtsac.py:325 needs an even split)"}.get(code, str(code)))
        # the kernels differentiate the full-batch mean; the reference's per-task loss is the mean over the task's
        # own B / T rows: scale by T
        return {k: v * float(T) for k, v in self._tg.items()}

    def task_gradient_view(self, flat_row: torch.Tensor, critic: bool) -> dict:
        """Flax-named views of one row of a `per_task_gradients` matrix."""
        c = self._cfg
        lay = self._lay.critic if critic else self._lay.actor
        return self._wrap_tree(flat_row, lay, (c.action_dim if critic else 0) + c.obs_dim, critic)

    def compute_weights(self, data: ReplayBufferSamples, eps_c=None, eps_a=None):
        """`MTSAC.compute_weights` (mtsac.py:870-1170), the metrics that derive from the per-task gradient matrix
        through its Gram matrix: `{critic,actor}_avg_cos_sim`, `_avg_grad_magnitude`, `_conflict_rate`,
        `_mean_conflict_magnitude`, `_mean_conflict_angle`, `_per_task_conflict_rate`, `_per_task_grad_magnitude`,
        `_pairwise_conflict`, `_pairwise_cos_sim`, `_pairwise_angle`, `_effective_rank` (utils.py:49-72, 104-174) and the
        `compute_gram_metrics` set (`_avg_cosine_gram`, `_gram_diag`, `_gram_off_diag_mean/_std`, `_pairwise_gram`,
        `_pairwise_cosine_gram`, mtsac.py:733-771).  Returns (self, logs) with device tensors.  The Gram matrix comes from
        `mtrl_task_gram`; the element-wise set (`_avg_interference_rate`, `_interference_asymmetry`,
        `_per_task_interference_in/_out`, `_pairwise_interference_rate`, `_avg/_per_task_participation_ratio`,
        utils.py:75-101, 146-156) from `mtrl_task_elementwise`; the `compute_support_metrics` set (`_avg_jaccard`,
        `_pairwise_jaccard`, `_avg_genuine/_ghost_conflict_rate`, `_ghost_to_genuine_ratio`, `_per_task/_avg_support_size`,
        `_pairwise_genuine/_ghost_conflict_rate`, mtsac.py:774-860) from `mtrl_task_abs_order_stats` (radix select of
        the 0.8-quantile) and `mtrl_task_support_pairs`; what follows is T x T bookkeeping.  All 66 keys of the
        reference's dictionary (mtsac.py:1085-1170) are present (tests/golden/compute_weights_keys.json)."""
        grads = self.per_task_gradients(data, eps_c, eps_a)
        logs = {}
        T = self.num_tasks
        for name, g in grads.items():
            gram = torch.empty(T, T, dtype=torch.float32, device=self.device)
            L.check(L.lib().mtrl_task_gram(_vp(g.data_ptr()), g.stride(0), T, g.shape[1], _vp(gram.data_ptr()),
                                           _vp(L.current_stream_ptr())))
            mag = torch.sqrt(torch.diagonal(gram).clamp_min(0))
            cos = gram / (mag[:, None] * mag[None, :] + 1e-8)
            off = 1 - torch.eye(T, device=g.device)
            upper = torch.triu(torch.ones(T, T, device=g.device), diagonal=1)
            conflict = (cos < 0).float()
            n_off = T * (T - 1)
            cm = torch.where((conflict * off).bool(), cos.abs() * (mag[:, None] * mag[None, :]), torch.zeros_like(cos))
            angles = torch.rad2deg(torch.arccos(cos.clamp(-1.0, 1.0)))
            off_mean = (gram * off).sum() / n_off
            sv = torch.linalg.svdvals(gram.double())                     # compute_effective_rank (utils.py:104-115) on the Gram
            sv_dist = sv / sv.sum().clamp_min(1e-10)
            logs.update({
                f"{name}_avg_cos_sim": (upper * cos).sum() / (upper.sum() + 1e-8),                   # utils.py:49-72
                f"{name}_avg_grad_magnitude": mag.mean(),                                            # mtsac.py:1043
                f"{name}_conflict_rate": (conflict * off).sum() / n_off,                             # utils.py:118-145
                f"{name}_mean_conflict_magnitude": (cm * off).sum() / n_off,
                f"{name}_mean_conflict_angle": (angles * off).sum() / n_off,
                f"{name}_per_task_conflict_rate": (conflict * off).sum(dim=1) / (T - 1),
                f"{name}_per_task_grad_magnitude": mag,
                f"{name}_pairwise_conflict": conflict,
                f"{name}_pairwise_cos_sim": cos,
                f"{name}_pairwise_angle": angles,
                f"{name}_effective_rank": torch.exp(-(sv_dist * torch.log(sv_dist + 1e-10)).sum()).float(),
                f"{name}_avg_cosine_gram": (cos * off).sum() / n_off,                                # mtsac.py:733-771
                f"{name}_gram_diag": torch.diagonal(gram),
                f"{name}_gram_off_diag_mean": off_mean,
                f"{name}_gram_off_diag_std": torch.sqrt((((gram - off_mean) ** 2) * off).sum() / n_off),
                f"{name}_pairwise_gram": gram,
                f"{name}_pairwise_cosine_gram": cos,
            })
            # element-wise part (utils.py:75-101, 146-156; eps = 1e-3, tau = 1.0 as compute_conflict_metrics' defaults)
            mism = torch.empty(T, T, dtype=torch.float32, device=self.device)
            rstat = torch.empty(T, 2, dtype=torch.float64, device=self.device)
            L.check(L.lib().mtrl_task_elementwise(_vp(g.data_ptr()), g.stride(0), T, g.shape[1], 1.0, 1e-3, 1.0,
                                                  _vp(mism.data_ptr()), _vp(rstat.data_ptr()), _vp(L.current_stream_ptr())))
            d = self.get_num_params()[f"{name}_num_params"]          # the reference's num_params (no layout padding)
            nz_counts = (rstat[:, 1] - (g.shape[1] - d)).clamp_min(1.0)
            rate = (mism.double() / nz_counts[:, None]) * off.double()
            pr = rstat[:, 0] ** 2 / (d * torch.diagonal(gram).double().clamp_min(1e-10))
            logs.update({
                f"{name}_avg_interference_rate": (rate.sum() / n_off).float(),
                f"{name}_interference_asymmetry": ((rate - rate.T).abs().sum() / n_off).float(),
                f"{name}_per_task_interference_in": (rate.sum(dim=0) / (T - 1)).float(),
                f"{name}_per_task_interference_out": (rate.sum(dim=1) / (T - 1)).float(),
                f"{name}_pairwise_interference_rate": rate.float(),
                f"{name}_avg_participation_ratio": pr.mean().float(),
                f"{name}_per_task_participation_ratio": pr.float(),
            })
            # compute_support_metrics (mtsac.py:774-860): support = top 20 % of |g| per task (0.8-quantile, linear
            # interpolation between two order statistics found by radix select), then pairwise counts
            P, pad = g.shape[1], g.shape[1] - d
            pos = 0.8 * (d - 1)
            k, frac = int(math.floor(pos)), pos - math.floor(pos)
            ranks = (C.c_longlong * T)(*([pad + k] * T))      # the layout's padding zeros sort first
            ostat = torch.empty(T, 2, dtype=torch.float32, device=self.device)
            scratch = torch.empty(T * (32 + 2048) // 4, dtype=torch.int32, device=self.device)
            L.check(L.lib().mtrl_task_abs_order_stats(_vp(g.data_ptr()), g.stride(0), T, P, ranks, _vp(ostat.data_ptr()),
                                                      _vp(scratch.data_ptr()), _vp(L.current_stream_ptr())))
            thr = (ostat[:, 0] + frac * (ostat[:, 1] - ostat[:, 0])).contiguous()
            pairs = torch.empty(3, T, T, dtype=torch.float32, device=self.device)
            L.check(L.lib().mtrl_task_support_pairs(_vp(g.data_ptr()), g.stride(0), T, P, _vp(thr.data_ptr()), _vp(pairs.data_ptr()),
                                                    _vp(L.current_stream_ptr())))
            inter, conf_cnt, genuine = pairs[0].double(), pairs[1].double(), pairs[2].double()
            zero_thr = (thr <= 0).double()                      # a zero threshold puts the padding zeros in the support
            inter = inter - pad * zero_thr[:, None] * zero_thr[None, :]
            size = torch.diagonal(inter)
            union = size[:, None] + size[None, :] - inter
            jacc = inter / (union + 1e-8)
            ghost = conf_cnt - genuine
            tot = genuine + ghost + 1e-8
            offd = off.double()
            logs.update({
                f"{name}_avg_jaccard": ((jacc * offd).sum() / n_off).float(),
                f"{name}_pairwise_jaccard": jacc.float(),
                f"{name}_avg_genuine_conflict_rate": (((genuine / tot) * offd).sum() / n_off).float(),
                f"{name}_avg_ghost_conflict_rate": (((ghost / tot) * offd).sum() / n_off).float(),
                f"{name}_ghost_to_genuine_ratio": (ghost.sum() / (genuine.sum() + 1e-8)).float(),
                f"{name}_per_task_support_size": size.float(),
                f"{name}_avg_support_size": size.mean().float(),
                f"{name}_pairwise_genuine_conflict_rate": (genuine / tot).float(),
                f"{name}_pairwise_ghost_conflict_rate": (ghost / tot).float(),
            })
        return self, logs

    # ------------------------------------------------------------------ checkpoints
    def _full_moments(self, prefix: str, critic: bool) -> dict[str, torch.Tensor]:
        """Adam moments of one network with the trunk part complete on every rank: under the sharded exchange a rank
        only keeps the moments of the trunk segments it owns, so they are summed over ranks (mask * moments)."""
        out = {k: self._flat[f"{prefix}_{k}"] for k in ("m", "v")}
        if self.exchange != "p2p":
            return out
        import torch.distributed as dist

        lay = self._lay.critic if critic else self._lay.actor
        mask = torch.empty(lay.trunk_total, dtype=torch.float32, device=self.device)
        L.check(L.lib().mtrl_sac_trunk_owner_mask(self._h, int(critic), _vp(mask.data_ptr())))
        full = {}
        for k, flat in out.items():
            f = flat.clone()
            f[: lay.trunk_total] *= mask
            dist.all_reduce(f[: lay.trunk_total], group=self.process_group)
            full[k] = f
        return full

    def state_dict(self) -> dict:
        """The agent as a host pytree in the reference's own names, the structure `ocp.args.PyTreeSave(agent)` writes
        (mtrl/checkpoint.py:39-44, 66): `actor` / `critic` / `alpha` TrainStates (mtrl/rl/algorithms/utils.py:11-46) with
        `step`, Flax-named `params` (`MultiHeadNetwork_0/layer_i/{kernel,bias}`, `VmapDense_0`; critic under
        `VmapQValueFunction_0` with a leading ensemble axis, plus `target_params`) and the optax state of
        `chain(clip_by_global_norm, adam)` as `opt_state = {"count", "mu", "nu"}`; `key` carries the Philox
        counter.  Leaves are NumPy arrays.  With several ranks every rank calls this (collective); heads / log_alpha
        are the rank's own tasks `[task_begin, task_end)`."""
        def host(tree):
            return {k: host(v) if isinstance(v, dict) else v.detach().cpu().numpy().copy() for k, v in tree.items()}

        def views(flat, lay, in_dim, ens):
            return self._wrap_tree(flat, lay, in_dim, ens)

        c = self._cfg
        out = {}
        for name, prefix, lay, in_dim, ens, idx in (("actor", "actor", self._lay.actor, c.obs_dim, False, 0),
                                                    ("critic", "critic", self._lay.critic, c.action_dim + c.obs_dim, True, 1)):
            mom = self._full_moments(prefix, ens)
            ts = {"step": int(self._steps[idx]), "params": host(views(self._flat[f"{prefix}_params"], lay, in_dim, ens)),
                  "opt_state": {"count": int(self._steps[idx]), "mu": host(views(mom["m"], lay, in_dim, ens)),
                                "nu": host(views(mom["v"], lay, in_dim, ens))}}
            if ens:
                ts["target_params"] = host(views(self._flat["critic_target"], lay, in_dim, ens))
            out[name] = ts
        out["alpha"] = {"step": int(self._steps[2]), "params": host(self.alpha.params),
                        "opt_state": {"count": int(self._steps[2]), "mu": host(self.alpha.opt_state["mu"]),
                                      "nu": host(self.alpha.opt_state["nu"])}}
        out["key"] = {"noise_counter": int(self._steps[3]), "seed": int(c.noise_seed)}
        out["task_range"] = (self.task_begin, self.task_end)
        return out

    def _wrap_tree(self, flat, lay, in_dim, ens):
        return _wrap(_views(flat, lay, in_dim, ens), ens)

    def load_state_dict(self, state: dict) -> None:
        """Inverse of `state_dict` (also accepts a tree exported by the reference through `jax.device_get`, whose
        leaves have the same names and shapes; full-T head tensors are sliced to this rank's tasks)."""
        c = self._cfg
        T_local = self.task_end - self.task_begin

        def put(dst, src, ens):
            for k, v in src.items():
                if isinstance(v, dict):
                    put(dst[k], v, ens)
                    continue
                t = torch.as_tensor(np.asarray(v), dtype=torch.float32)
                if t.shape != dst[k].shape:   # a full-T head tensor: keep this rank's tasks
                    axis = 1 if ens else 0
                    t = t.narrow(axis, self.task_begin, T_local)
                dst[k].copy_(t)

        for name, prefix, lay, in_dim, ens, idx in (("actor", "actor", self._lay.actor, c.obs_dim, False, 0),
                                                    ("critic", "critic", self._lay.critic, c.action_dim + c.obs_dim, True, 1)):
            ts = state[name]
            put(self._wrap_tree(self._flat[f"{prefix}_params"], lay, in_dim, ens), ts["params"], ens)
            put(self._wrap_tree(self._flat[f"{prefix}_m"], lay, in_dim, ens), ts["opt_state"]["mu"], ens)
            put(self._wrap_tree(self._flat[f"{prefix}_v"], lay, in_dim, ens), ts["opt_state"]["nu"], ens)
            if ens:
                put(self._wrap_tree(self._flat["critic_target"], lay, in_dim, ens), ts["target_params"], ens)
            self._steps[idx] = int(ts["opt_state"]["count"])
        a = state["alpha"]
        for dst, src in ((self.alpha.params, a["params"]), (self.alpha.opt_state["mu"], a["opt_state"]["mu"]),
                         (self.alpha.opt_state["nu"], a["opt_state"]["nu"])):
            t = torch.as_tensor(np.asarray(src["params"]["log_alpha"]), dtype=torch.float32)
            if t.shape[0] != T_local:
                t = t[self.task_begin:self.task_end]
            dst["params"]["log_alpha"].copy_(t)
        self._steps[2] = int(a["opt_state"]["count"])
        if "key" in state and isinstance(state["key"], dict):
            self._steps[3] = int(state["key"].get("noise_counter", 0))
        self.refresh()

    def act_device(self, observation, eps=None, deterministic: bool = False) -> torch.Tensor:
        """Actions for `observation` (n, obs_dim) as a CUDA tensor (n, action_dim), without synchronising the host:
        tanh(mu + sigma eps) with `eps` (n, action_dim) or in-kernel Philox noise, or the mode tanh(mu)."""
        obs = self._dev(observation)
        if obs.dim() == 1:
            obs = obs[None]
        n = obs.shape[0]
        assert obs.shape[1] == self._cfg.obs_dim
        e = self._dev(eps) if eps is not None else None
        out = torch.empty(n, self._cfg.action_dim, dtype=torch.float32, device=self.device)
        L.check(L.lib().mtrl_sac_act(self._h, _vp(obs.data_ptr()), n, _vp(e.data_ptr() if e is not None else None),
                                     int(deterministic), _vp(out.data_ptr()), _vp(L.current_stream_ptr())))
        L.check(L.lib().mtrl_sac_read_status_async(self._h, _vp(self._status_host.data_ptr()), _vp(L.current_stream_ptr())))
        self._status_event.record()
        self._pending_status = True
        return out

    def network_forward(self, net: str, observations, actions=None) -> torch.Tensor:
        """`MultiHeadNetwork.__call__` (mtrl/nn/multi_head.py:21-68) of one of the agent's networks on arbitrary rows, as a
        CUDA tensor: net="actor" -> (n, 2 * action_dim) head outputs (mean | log_std before the clip, networks.py:36-40);
        net="critic" / "target" -> (num_critics, n, 1) Q-values of `QValueFunction` on (actions, observations)
        (networks.py:55-67, 208-222).  C entry `mtrl_mlp_forward`."""
        code = {"actor": 0, "critic": 1, "target": 2}[net]
        obs = self._dev(observations)
        if obs.dim() == 1:
            obs = obs[None]
        n, c = obs.shape[0], self._cfg
        assert obs.shape[1] == c.obs_dim
        act = None
        if code:
            if actions is None:
                raise ValueError("the critic takes (observations, actions)")
            act = self._dev(actions).reshape(n, c.action_dim)
        E, hd = (1, 2 * c.action_dim) if code == 0 else (c.num_critics, 1)
        out = torch.empty(E, n, hd, dtype=torch.float32, device=self.device)
        stream = _vp(L.current_stream_ptr())
        L.check(L.lib().mtrl_mlp_forward(self._h, code, _vp(obs.data_ptr()), _vp(act.data_ptr() if act is not None else None), n,
                                         _vp(out.data_ptr()), stream))
        L.check(L.lib().mtrl_sac_read_status_async(self._h, _vp(self._status_host.data_ptr()), stream))
        self._status_event.record()
        self._pending_status = True
        return out[0] if code == 0 else out

    def sample_action(self, observation, task_ids=None, eps=None):
        """mtsac.py:299-304: `(self, action)` with the action on the host as a NumPy array (the reference returns
        `jax.device_get(action)`).  Noise: in-kernel Philox, or `eps` for a reproducible draw."""
        a = self.act_device(observation, eps=eps).cpu().numpy()
        self._check_status()
        return self, a

    def eval_action(self, observations, task_ids=None):
        """mtsac.py:306-311: the mode of the tanh-Gaussian, tanh(mu) (nn/distributions.py:15-16), as a NumPy array."""
        a = self.act_device(observations, deterministic=True).cpu().numpy()
        self._check_status()
        return a
