"""Device-resident mirror of `mtrl.rl.buffers.MultiTaskReplayBuffer`
(/root/reference/mtrl/rl/buffers.py:221-549): same constructor, attributes and methods; storage lives
in HBM and `sample` runs the CUDA sampler (csrc/sampler.cu) through the C-ABI.  Index draws are
bit-identical to the reference's `np.random.default_rng(seed).integers` stream.

Differences a caller can see, all deliberate:
  * returned samples are CUDA torch tensors (fp32) instead of NumPy arrays -- the reference hands its
    NumPy batch straight to a jitted function that copies it to the device (base.py:220-221);
  * `sample` with reward normalisation returns fp32 (the reference returns float64 which JAX then
    casts to fp32 at the jit boundary);
  * sampling more rows per task than the capacity raises IndexError up front (NumPy raises it from
    the fancy index, buffers.py:529).
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from .. import _lib as L
from ..types import ReplayBufferCheckpoint, ReplayBufferSamples, Rollout

_vp, _i, _u32 = C.c_void_p, C.c_int, C.c_uint32
L._EXTRA_DECLS.update({
    "mtrl_sampler_create": ([C.POINTER(_vp), _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp],),
    "mtrl_sampler_destroy": ([_vp], None),
    "mtrl_sampler_set_state": ([_vp, C.POINTER(C.c_uint64), _u32, _u32, _vp],),
    "mtrl_sampler_get_state": ([_vp, C.POINTER(C.c_uint64), C.POINTER(_u32), C.POINTER(_u32), _vp],),
    "mtrl_sampler_add": ([_vp, _i, _vp, _vp, _vp, _vp, _vp, _vp],),
    "mtrl_sampler_sample": ([_vp, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _i, _vp, _vp, _vp],),
    "mtrl_sampler_sample_per_task": ([_vp, _i, C.POINTER(_i), _vp, _vp, _vp, _vp, _vp, _vp],),
    "mtrl_sampler_draw": ([_vp, C.c_ulonglong, _i, _vp, _vp],),
})

_M64 = (1 << 64) - 1


class _DeviceGenerator:
    """Stands in for `np.random.Generator` on `buffer._rng`: the PCG64 state lives on the GPU.
    `bit_generator.state` / `__getstate__` / `__setstate__` speak numpy's PCG64 state dict."""

    def __init__(self, owner: "MultiTaskReplayBuffer"):
        self._owner = owner

    @property
    def bit_generator(self):
        return self

    @property
    def state(self) -> dict:
        st = (C.c_uint64 * 4)()
        has, ui = _u32(), _u32()
        L.check(L.lib().mtrl_sampler_get_state(self._owner._h, st, C.byref(has), C.byref(ui), _vp(L.current_stream_ptr())))
        return {
            "bit_generator": "PCG64",
            "state": {"state": (int(st[0]) << 64) | int(st[1]), "inc": (int(st[2]) << 64) | int(st[3])},
            "has_uint32": int(has.value),
            "uinteger": int(ui.value),
        }

    @state.setter
    def state(self, st: dict) -> None:
        if st.get("bit_generator", "PCG64") != "PCG64":
            raise ValueError("only PCG64 states are supported")
        s, inc = int(st["state"]["state"]), int(st["state"]["inc"])
        arr = (C.c_uint64 * 4)(s >> 64, s & _M64, inc >> 64, inc & _M64)
        L.check(L.lib().mtrl_sampler_set_state(self._owner._h, arr, int(st["has_uint32"]), int(st["uinteger"]),
                                               _vp(L.current_stream_ptr())))

    def __getstate__(self):
        return self.state

    def __setstate__(self, st):
        # accepts numpy's bit_generator.state dict, or {"bit_generator": <that dict>} (Generator pickles)
        if "state" not in st and "bit_generator" in st and isinstance(st["bit_generator"], dict):
            st = st["bit_generator"]
        self.state = st


class MultiTaskReplayBuffer:
    """See module docstring.  `env_obs_space` / `env_action_space` only need `.shape`."""

    def __init__(self, total_capacity: int, num_tasks: int, env_obs_space, env_action_space, seed: int | None = None,
                 max_steps: int = 500, normalize_rewards: bool = False, reward_norm_eps: float = 1e-8,
                 reward_filter: str | None = None, sigma: float | None = None, alpha: float | None = None,
                 delta: float | None = None, filter_mode: str | None = None, returns_normalization: bool = False,
                 discount: float = 0.99, v_max: float = 10.0, device: str | torch.device = "cuda") -> None:
        assert total_capacity % num_tasks == 0, "Total capacity must be divisible by the number of tasks."
        if not torch.cuda.is_available():
            raise L.MtrlError("MultiTaskReplayBuffer needs a CUDA device; there is no CPU fallback")
        self.capacity = total_capacity // num_tasks
        self.num_tasks = num_tasks
        self.device = torch.device(device)
        self._obs_shape = int(np.array(env_obs_space.shape).prod())
        self._action_shape = int(np.array(env_action_space.shape).prod())
        self.full = False
        self.normalize_rewards = normalize_rewards
        self._min_rewards = np.full(num_tasks, np.inf, dtype=np.float64)
        self._max_rewards = np.full(num_tasks, -np.inf, dtype=np.float64)
        self.reward_norm_eps = reward_norm_eps
        self.use_return_normalization = returns_normalization
        self.discount = discount
        self.v_max = v_max
        self.effective_horizon = 1.0 / (1.0 - discount)
        self._returns_min = np.full(num_tasks, np.inf, dtype=np.float64)
        self._returns_max = np.full(num_tasks, -np.inf, dtype=np.float64)
        self._episode_rewards: list[list[float]] = [[] for _ in range(num_tasks)]
        self._h = None
        self._norm_dev = None
        self._seed_state = np.random.PCG64(seed).state  # numpy's SeedSequence -> PCG64 seeding, host side
        self.reset()
        self._rng = _DeviceGenerator(self)
        self._rng.state = self._seed_state

    # ---- storage ----
    def reset(self) -> None:
        """buffers.py:293-306 (zero-filled (capacity, T, dim) fp32 arrays, pos = 0)."""
        c, t, dev = self.capacity, self.num_tasks, self.device
        z = lambda d: torch.zeros((c, t, d), dtype=torch.float32, device=dev)  # noqa: E731
        self.obs, self.actions, self.rewards = z(self._obs_shape), z(self._action_shape), z(1)
        self.next_obs, self.dones = z(self._obs_shape), z(1)
        self.pos = 0
        self._bind()

    def _bind(self) -> None:
        state = None
        if self._h is not None:
            state = self._rng.state
            L.lib().mtrl_sampler_destroy(self._h)
            self._h = None
        h = _vp()
        L.check(L.lib().mtrl_sampler_create(C.byref(h), self.capacity, self.num_tasks, self._obs_shape, self._action_shape,
                                            self.obs.data_ptr(), self.actions.data_ptr(), self.next_obs.data_ptr(),
                                            self.dones.data_ptr(), self.rewards.data_ptr()))
        self._h = h
        if state is not None:
            self._rng.state = state

    def __del__(self):
        if getattr(self, "_h", None) is not None and L._lib is not None:
            L._lib.mtrl_sampler_destroy(self._h)
            self._h = None

    def _advance_position(self, steps: int) -> None:  # buffers.py:337-343
        if steps <= 0:
            return
        new_pos = self.pos + steps
        if new_pos >= self.capacity:
            self.full = True
        self.pos = new_pos % self.capacity

    # ---- return normalisation statistics (host side, env-rate work; buffers.py:347-390) ----
    def _update_return_stats(self, rewards, terminal, truncated) -> None:
        for t in range(self.num_tasks):
            self._episode_rewards[t].append(float(rewards[t]))
            if bool(terminal[t]) or bool(truncated[t]):
                ep = np.array(self._episode_rewards[t], dtype=np.float64)
                values = np.zeros(len(ep), dtype=np.float64)
                bootstrap = float(ep.mean()) * self.effective_horizon if bool(truncated[t]) else 0.0
                for i in reversed(range(len(ep))):
                    values[i] = ep[i] + self.discount * bootstrap
                    bootstrap = values[i]
                self._returns_min[t] = min(self._returns_min[t], float(values.min()))
                self._returns_max[t] = max(self._returns_max[t], float(values.max()))
                self._episode_rewards[t] = []

    @staticmethod
    def _host_f32(x, shape):
        if isinstance(x, torch.Tensor):
            return x.detach().to(torch.float32).reshape(shape).contiguous()
        return np.ascontiguousarray(np.asarray(x, dtype=np.float32).reshape(shape))

    @staticmethod
    def _ptr(x) -> int:
        return x.data_ptr() if isinstance(x, torch.Tensor) else x.ctypes.data

    def add(self, obs, next_obs, action, reward, done, terminal=None, truncated=None) -> None:
        """buffers.py:426-474.  Inputs may be NumPy arrays (host) or torch tensors (host or CUDA)."""
        assert obs.ndim == 2 and action.ndim == 2 and reward.ndim <= 2 and done.ndim <= 2
        assert obs.shape[0] == action.shape[0] == reward.shape[0] == done.shape[0] == self.num_tasks
        t = self.num_tasks
        src = [self._host_f32(obs, (t, self._obs_shape)), self._host_f32(action, (t, self._action_shape)),
               self._host_f32(next_obs, (t, self._obs_shape)), self._host_f32(done, (t, 1)),
               self._host_f32(reward, (t, 1))]
        L.check(L.lib().mtrl_sampler_add(self._h, self.pos, *(_vp(self._ptr(s)) for s in src),
                                         _vp(L.current_stream_ptr())))
        if any(isinstance(s, np.ndarray) for s in src):
            torch.cuda.current_stream().synchronize()  # pageable host sources must outlive the copy
        if self.normalize_rewards or self.use_return_normalization:
            r_host = reward.detach().cpu().numpy() if isinstance(reward, torch.Tensor) else np.asarray(reward)
            d_host = done.detach().cpu().numpy() if isinstance(done, torch.Tensor) else np.asarray(done)
            if self.normalize_rewards:
                self._min_rewards = np.minimum(self._min_rewards, r_host.reshape(-1))
                self._max_rewards = np.maximum(self._max_rewards, r_host.reshape(-1))
                self._norm_dev = None
            if self.use_return_normalization:
                _terminal = terminal if terminal is not None else d_host
                _truncated = truncated if truncated is not None else np.zeros_like(d_host)
                self._update_return_stats(r_host.flatten(), np.asarray(_terminal).flatten().astype(bool),
                                          np.asarray(_truncated).flatten().astype(bool))
                self._norm_dev = None
        self._advance_position(1)

    # ---- sampling ----
    def _fill(self) -> int:
        return self.pos if not self.full else self.capacity

    def _norm_arrays(self):
        """(shift, den) float64 device arrays for rewards' = (r - shift) / den, or None."""
        if self.use_return_normalization:  # buffers.py:392-422
            no_data = np.isinf(self._returns_min) | np.isinf(self._returns_max)
            den = np.where(self._returns_max >= np.abs(self._returns_min), self._returns_max, np.abs(self._returns_min))
            den = den / self.v_max
            den = np.where(no_data | (den < self.reward_norm_eps), 1.0, den)
            shift = np.zeros(self.num_tasks, dtype=np.float64)
        elif self.normalize_rewards:       # buffers.py:534-538
            shift = self._min_rewards.astype(np.float64)
            den = self._max_rewards - self._min_rewards + self.reward_norm_eps
        else:
            return None
        if self._norm_dev is None:
            self._norm_dev = (torch.from_numpy(np.ascontiguousarray(shift)).to(self.device),
                              torch.from_numpy(np.ascontiguousarray(den.astype(np.float64))).to(self.device))
        return self._norm_dev

    def _alloc_out(self, rows: int):
        e = lambda d: torch.empty((rows, d), dtype=torch.float32, device=self.device)  # noqa: E731
        return e(self._obs_shape), e(self._action_shape), e(self._obs_shape), e(1), e(1)

    def sample(self, batch_size, return_indices: bool = False) -> ReplayBufferSamples:
        """buffers.py:494-549.  int -> one shared index vector, rows interleaved (sample, task);
        ndarray of per-task counts -> independent draws, rows concatenated by task."""
        stream = _vp(L.current_stream_ptr())
        if isinstance(batch_size, np.ndarray):
            assert len(batch_size) == self.num_tasks
            assert batch_size.sum() == (128 * self.num_tasks)
            counts = (C.c_int * self.num_tasks)(*[int(x) for x in batch_size])
            outs = self._alloc_out(int(batch_size.sum()))
            L.check(L.lib().mtrl_sampler_sample_per_task(self._h, self._fill(), counts, *(_vp(o.data_ptr()) for o in outs),
                                                         stream))
            return ReplayBufferSamples(*outs)
        assert batch_size % self.num_tasks == 0
        single = batch_size // self.num_tasks
        if max(self._fill(), single) > self.capacity:
            raise IndexError(f"index out of bounds: {single} samples per task from capacity {self.capacity}")
        outs = self._alloc_out(single * self.num_tasks)
        norm = self._norm_arrays()
        idx = torch.empty((single,), dtype=torch.int64, device=self.device) if return_indices else None
        L.check(L.lib().mtrl_sampler_sample(
            self._h, self._fill(), single, _vp(idx.data_ptr() if idx is not None else None),
            *(_vp(o.data_ptr()) for o in outs), 1 if norm is not None else 0,
            _vp(norm[0].data_ptr() if norm is not None else None), _vp(norm[1].data_ptr() if norm is not None else None),
            stream))
        s = ReplayBufferSamples(*outs)
        return (s, idx) if return_indices else s

    def single_task_sample(self, task_idx: int, batch_size: int) -> ReplayBufferSamples:
        """buffers.py:478-492, literally: `self.obs[sample_idx][task_idx]` selects row `task_idx` of the
        SAMPLE axis, i.e. all tasks of that one drawn transition, shape (T, dim)."""
        assert task_idx < self.num_tasks, "Task index out of bounds."
        norm_flags = (self.normalize_rewards, self.use_return_normalization)
        self.normalize_rewards, self.use_return_normalization = False, False
        try:
            s = self.sample(batch_size * self.num_tasks)
        finally:
            self.normalize_rewards, self.use_return_normalization = norm_flags
        t = self.num_tasks
        return ReplayBufferSamples(*(x.reshape(batch_size, t, -1)[task_idx] for x in s))

    # ---- checkpointing (buffers.py:308-335; same keys; arrays returned as NumPy) ----
    def checkpoint(self) -> ReplayBufferCheckpoint:
        c = lambda x: x.detach().cpu().numpy()  # noqa: E731
        return {
            "data": {"obs": c(self.obs), "actions": c(self.actions), "rewards": c(self.rewards),
                     "next_obs": c(self.next_obs), "dones": c(self.dones), "pos": self.pos, "full": self.full,
                     "returns_min": self._returns_min, "returns_max": self._returns_max},
            "rng_state": self._rng.__getstate__(),
        }

    def load_checkpoint(self, ckpt: ReplayBufferCheckpoint) -> None:
        for key in ["data", "rng_state"]:
            assert key in ckpt
        for key in ["obs", "actions", "rewards", "next_obs", "dones", "pos", "full"]:
            assert key in ckpt["data"]
        d = ckpt["data"]
        for key in ["obs", "actions", "rewards", "next_obs", "dones"]:
            getattr(self, key).copy_(torch.as_tensor(np.asarray(d[key]), dtype=torch.float32))
        self.pos = int(d["pos"])
        self.full = bool(d["full"])
        self._returns_min = d.get("returns_min", self._returns_min)
        self._returns_max = d.get("returns_max", self._returns_max)
        self._norm_dev = None
        self._rng.__setstate__(ckpt["rng_state"])


class ReplayBuffer:
    """Mirror of the single-task `mtrl.rl.buffers.ReplayBuffer` (buffers.py:21-218), device resident.

    The reference's `sample` reads `self.num_tasks`, which its `__init__` never sets (buffers.py:35-48 vs :197), so
    calling it raises AttributeError in that snapshot; it is implemented here with the evident intent num_tasks = 1:
    `batch_size` indices from `integers(0, max(fill, batch_size))`, rows gathered from all five arrays.  The index
    stream is the same bit-exact PCG64 stream as the multi-task buffer."""

    def __init__(self, capacity: int, env_obs_space, env_action_space, seed: int | None = None,
                 device: str | torch.device = "cuda") -> None:
        self._mt = MultiTaskReplayBuffer(capacity, 1, env_obs_space, env_action_space, seed=seed, device=device)
        self.capacity = capacity
        self.num_tasks = 1
        self._rng = self._mt._rng
        self._obs_shape, self._action_shape = self._mt._obs_shape, self._mt._action_shape

    # storage views (capacity, dim), names of buffers.py:52-57
    obs = property(lambda self: self._mt.obs[:, 0])
    actions = property(lambda self: self._mt.actions[:, 0])
    rewards = property(lambda self: self._mt.rewards[:, 0])
    next_obs = property(lambda self: self._mt.next_obs[:, 0])
    dones = property(lambda self: self._mt.dones[:, 0])
    pos = property(lambda self: self._mt.pos, lambda self, v: setattr(self._mt, "pos", v))
    full = property(lambda self: self._mt.full, lambda self, v: setattr(self._mt, "full", v))

    def reset(self) -> None:
        self._mt.reset()

    def _advance_position(self, steps: int) -> None:  # buffers.py:83-93
        self._mt._advance_position(steps)

    def add(self, obs, next_obs, action, reward, done) -> None:
        """buffers.py:95-140: a single transition (1-D inputs) or a batch with arbitrary leading dims."""
        dev = self._mt.device
        t = lambda x: torch.as_tensor(np.asarray(x) if not isinstance(x, torch.Tensor) else x, dtype=torch.float32).to(dev)  # noqa: E731
        obs, next_obs, action, reward, done = t(obs), t(next_obs), t(action), t(reward), t(done)
        if obs.ndim >= 2:
            assert obs.shape[0] == action.shape[0] == reward.shape[0] == done.shape[0], \
                "Batch size must be the same for all transition data."
            fo, fn, fa = obs.reshape(-1, obs.shape[-1]), next_obs.reshape(-1, next_obs.shape[-1]), action.reshape(-1, action.shape[-1])
            fr, fd = reward.reshape(-1, 1), done.reshape(-1, 1)
            n = fo.shape[0]
            idx = (self.pos + torch.arange(n, device=dev)) % self.capacity
            self._mt.obs[idx, 0], self._mt.next_obs[idx, 0], self._mt.actions[idx, 0] = fo, fn, fa
            self._mt.rewards[idx, 0], self._mt.dones[idx, 0] = fr, fd
            self._advance_position(n)
        else:
            self._mt.obs[self.pos, 0], self._mt.actions[self.pos, 0], self._mt.next_obs[self.pos, 0] = obs, action, next_obs
            self._mt.dones[self.pos, 0], self._mt.rewards[self.pos, 0] = done.reshape(-1), reward.reshape(-1)
            self._advance_position(1)

    def sample(self, batch_size: int) -> ReplayBufferSamples:
        return self._mt.sample(int(batch_size))

    def checkpoint(self) -> ReplayBufferCheckpoint:  # buffers.py:59-71
        ck = self._mt.checkpoint()
        d = ck["data"]
        return {"data": {k: (d[k][:, 0] if k in ("obs", "actions", "rewards", "next_obs", "dones") else d[k])
                         for k in ("obs", "actions", "rewards", "next_obs", "dones", "pos", "full")},
                "rng_state": ck["rng_state"]}

    def load_checkpoint(self, ckpt: ReplayBufferCheckpoint) -> None:  # buffers.py:73-81
        for key in ["data", "rng_state"]:
            assert key in ckpt
        d = dict(ckpt["data"])
        for k in ("obs", "actions", "rewards", "next_obs", "dones"):
            assert k in d
            d[k] = np.asarray(d[k])[:, None]
        self._mt.load_checkpoint({"data": d, "rng_state": ckpt["rng_state"]})


L._EXTRA_DECLS.update({
    "mtrl_gae": ([_vp, _vp, _vp, _vp, _vp, _i, _i, C.c_float, C.c_float, _vp, _vp, _vp],),
})


class MultiTaskRolloutBuffer:
    """Device-resident mirror of `mtrl.rl.buffers.MultiTaskRolloutBuffer` (/root/reference/mtrl/rl/buffers.py:552-707):
    same constructor, attributes (`observations`, `actions`, `rewards`, `dones`, `values`, `log_probs`, `means`, `stds`
    as (timestep, task, dim) fp32 arrays -- CUDA tensors here -- `pos`, `ready`) and methods.  `get` returns a
    `Rollout` of (task, timestep, dim) CUDA views; advantages come from the CUDA scan `mtrl_gae` (csrc/rollout.cu),
    bit-identical to NumPy's float32 evaluation of the reference loop.

    The reference's `get` cannot run (see oracle/rollout_oracle.py); this implements what its annotations and the
    upstream loop it cites say: fields transposed to (task, timestep, dim), last step bootstrapped from `dones`."""

    def __init__(self, num_rollout_steps: int, num_tasks: int, env_obs_space, env_action_space, seed: int | None = None,
                 device: str | torch.device | None = None) -> None:
        if not torch.cuda.is_available():
            raise L.MtrlError("MultiTaskRolloutBuffer needs a CUDA device; there is no CPU fallback")
        L.lib()
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self.num_rollout_steps, self.num_tasks = num_rollout_steps, num_tasks
        self._rng = np.random.default_rng(seed)
        self._obs_shape = int(np.array(env_obs_space.shape).prod())
        self._action_shape = int(np.array(env_action_space.shape).prod())
        self.reset()

    def reset(self) -> None:   # buffers.py:582-605
        S, T = self.num_rollout_steps, self.num_tasks
        f = lambda d: torch.zeros(S, T, d, dtype=torch.float32, device=self.device)  # noqa: E731
        self.observations, self.actions = f(self._obs_shape), f(self._action_shape)
        self.rewards, self.dones, self.log_probs, self.values = f(1), f(1), f(1), f(1)
        self.means, self.stds = f(self._action_shape), f(self._action_shape)
        self.pos = 0
        self._values_pushed = False

    @property
    def ready(self) -> bool:
        return self.pos == self.num_rollout_steps

    def _row(self, x, shape) -> torch.Tensor:
        if isinstance(x, torch.Tensor):
            return x.detach().to(device=self.device, dtype=torch.float32, non_blocking=True).reshape(shape)
        return torch.as_tensor(np.ascontiguousarray(np.asarray(x, dtype=np.float32))).to(self.device).reshape(shape)

    def add(self, obs, action, reward, done, value=None, log_prob=None, mean=None, std=None) -> None:
        """buffers.py:611-648.  Inputs may be NumPy arrays or torch tensors (host or CUDA)."""
        assert obs.ndim == 2 and action.ndim == 2 and reward.ndim <= 2 and done.ndim <= 2
        assert obs.shape[0] == action.shape[0] == reward.shape[0] == done.shape[0] == self.num_tasks
        if self.pos >= self.num_rollout_steps:
            raise IndexError(f"index {self.pos} is out of bounds for axis 0 with size {self.num_rollout_steps}")
        T, p = self.num_tasks, self.pos
        self.observations[p] = self._row(obs, (T, self._obs_shape))
        self.actions[p] = self._row(action, (T, self._action_shape))
        self.rewards[p] = self._row(reward, (T, 1))
        self.dones[p] = self._row(done, (T, 1))
        if value is not None:
            self.values[p] = self._row(value, (T, 1))
            self._values_pushed = True
        if log_prob is not None:
            self.log_probs[p] = self._row(log_prob, (T, 1))
        if mean is not None:
            self.means[p] = self._row(mean, (T, self._action_shape))
        if std is not None:
            self.stds[p] = self._row(std, (T, self._action_shape))
        self.pos += 1

    def get(self, compute_advantages: bool, last_values=None, dones=None, gamma: float = 0.99,
            gae_lambda: float = 0.97) -> Rollout:
        """buffers.py:650-707."""
        returns = advantages = None
        if compute_advantages:
            assert last_values is not None, "Must provide final value estimates if compute_advantages=True."
            assert dones is not None, "Must provide final value estimates if compute_advantages=True."
            # the reference asserts `not np.all(values == 0)`; checked without a device round trip
            assert self._values_pushed, "Values must have been pushed to the buffer if compute_advantages=True."
            S, T = self.num_rollout_steps, self.num_tasks
            lv = self._row(last_values, (T,)).contiguous()
            ld = self._row(dones, (T,)).contiguous()
            adv = torch.empty_like(self.rewards)
            ret = torch.empty_like(self.rewards)
            L.check(L.lib().mtrl_gae(_vp(self.rewards.data_ptr()), _vp(self.values.data_ptr()), _vp(self.dones.data_ptr()),
                                     _vp(lv.data_ptr()), _vp(ld.data_ptr()), S, T, float(gamma), float(gae_lambda),
                                     _vp(adv.data_ptr()), _vp(ret.data_ptr()), _vp(L.current_stream_ptr())))
            returns, advantages = ret.transpose(0, 1), adv.transpose(0, 1)
        tr = lambda x: x.transpose(0, 1)  # noqa: E731
        return Rollout(tr(self.observations), tr(self.actions), tr(self.rewards), tr(self.dones), tr(self.log_probs),
                       tr(self.means), tr(self.stds), tr(self.values), returns, advantages)
