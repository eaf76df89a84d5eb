"""Oracle for the replay sampler: a NumPy/pure-Python restatement of
`MultiTaskReplayBuffer` (/root/reference/mtrl/rl/buffers.py:221-549) including the exact integer
stream of `np.random.default_rng(seed).integers` (PCG64 XSL-RR + 32-bit Lemire rejection).

TEST INFRASTRUCTURE: imported only by tests/, __graft_entry__.smoke() and bench.py's CPU baseline.

Parity status: PINNED.  tests/golden/sampler_*.npz were produced by importing the reference's own
buffers.py in the build container (tests/golden/make_sampler_golden.py, gymnasium/jax stubbed for
type-annotation imports only) and this restatement reproduces them bit for bit
(tests/test_sampler_oracle.py).  The RNG core is additionally checked against the live
numpy.random.Generator on every run.

The RNG algorithm lives in NumPy (pinned numpy==2.2.4 in the reference's uv.lock:1078-1079), not
in the reference tree; its call site is buffers.py:260 (`np.random.default_rng(seed)`) and
buffers.py:523-527 (`self._rng.integers(low=0, high=..., size=(n,))`).
"""
from __future__ import annotations

from typing import NamedTuple

import numpy as np

_MASK128 = (1 << 128) - 1
_MASK64 = (1 << 64) - 1
_PCG_MULT = 0x2360ED051FC65DA44385DF649FCCF645  # PCG_DEFAULT_MULTIPLIER_128


class ReplayBufferSamples(NamedTuple):
    """Field order of mtrl/types.py:30-35 (differs from the buffer attribute order)."""

    observations: np.ndarray
    actions: np.ndarray
    next_observations: np.ndarray
    dones: np.ndarray
    rewards: np.ndarray


class PCG64:
    """numpy.random.PCG64 restated with Python integers (128-bit LCG, XSL-RR output)."""

    def __init__(self, seed=None, *, state: dict | None = None):
        if state is not None:
            self.set_state(state)
            return
        # numpy/random/_pcg64.pyx: _seed_seq.generate_state(4, uint64) -> pcg64_set_seed
        s = np.random.SeedSequence(seed).generate_state(4, np.uint64)
        initstate = (int(s[0]) << 64) | int(s[1])
        initseq = (int(s[2]) << 64) | int(s[3])
        self.inc = ((initseq << 1) | 1) & _MASK128
        self.state = 0
        self._step()
        self.state = (self.state + initstate) & _MASK128
        self._step()
        self.has_uint32 = 0
        self.uinteger = 0

    def _step(self) -> None:
        self.state = (self.state * _PCG_MULT + self.inc) & _MASK128

    def next64(self) -> int:
        self._step()
        hi, lo = self.state >> 64, self.state & _MASK64
        x = hi ^ lo
        rot = hi >> 58
        return ((x >> rot) | (x << ((-rot) & 63))) & _MASK64

    def next32(self) -> int:
        # pcg64_next32: low half first, high half buffered and kept across calls
        if self.has_uint32:
            self.has_uint32 = 0
            return self.uinteger
        n = self.next64()
        self.has_uint32 = 1
        self.uinteger = n >> 32
        return n & 0xFFFFFFFF

    def get_state(self) -> dict:
        return {
            "bit_generator": "PCG64",
            "state": {"state": self.state, "inc": self.inc},
            "has_uint32": self.has_uint32,
            "uinteger": self.uinteger,
        }

    def set_state(self, st: dict) -> None:
        self.state = int(st["state"]["state"])
        self.inc = int(st["state"]["inc"])
        self.has_uint32 = int(st["has_uint32"])
        self.uinteger = int(st["uinteger"])

    def integers(self, high: int, n: int) -> np.ndarray:
        """`Generator.integers(low=0, high=high, size=(n,))`, int64 result, for high <= 2**32.

        numpy/random/src/distributions/distributions.c: random_bounded_uint64_fill ->
        buffered_bounded_lemire_uint32 (rng = high - 1; rng == 0 consumes no randomness).
        """
        assert 1 <= high <= 0xFFFFFFFF
        out = np.zeros(n, dtype=np.int64)
        if high == 1:
            return out
        thr = ((1 << 32) - high) % high
        for i in range(n):
            m = self.next32() * high
            if (m & 0xFFFFFFFF) < high:
                while (m & 0xFFFFFFFF) < thr:
                    m = self.next32() * high
            out[i] = m >> 32
        return out


class MultiTaskReplayBufferOracle:
    """buffers.py:221-549 restated (storage, ring pointer, sample paths, reward normalisation)."""

    def __init__(self, total_capacity: int, num_tasks: int, obs_dim: int, action_dim: int, seed=None,
                 normalize_rewards: bool = False, reward_norm_eps: float = 1e-8,
                 returns_normalization: bool = False, discount: float = 0.99, v_max: float = 10.0):
        assert total_capacity % num_tasks == 0  # buffers.py:255-257
        self.capacity = total_capacity // num_tasks
        self.num_tasks = num_tasks
        self._rng = PCG64(seed)
        self._obs_shape = obs_dim
        self._action_shape = action_dim
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
        self._episode_rewards = [[] for _ in range(num_tasks)]
        self.reset()

    def reset(self) -> None:  # buffers.py:293-306
        c, t = self.capacity, self.num_tasks
        self.obs = np.zeros((c, t, self._obs_shape), dtype=np.float32)
        self.actions = np.zeros((c, t, self._action_shape), dtype=np.float32)
        self.rewards = np.zeros((c, t, 1), dtype=np.float32)
        self.next_obs = np.zeros((c, t, self._obs_shape), dtype=np.float32)
        self.dones = np.zeros((c, t, 1), dtype=np.float32)
        self.pos = 0

    def _advance_position(self, steps: int) -> None:  # buffers.py:337-343
        if steps <= 0:
            return
        new_pos = self.pos + steps
        if new_pos >= self.capacity:
            self.full = True
        self.pos = new_pos % self.capacity

    def _update_return_stats(self, rewards, terminal, truncated) -> None:  # buffers.py:347-390
        for t in range(self.num_tasks):
            self._episode_rewards[t].append(float(rewards[t]))
            if bool(terminal[t]) or bool(truncated[t]):
                ep = np.array(self._episode_rewards[t], dtype=np.float64)
                n = len(ep)
                values = np.zeros(n, dtype=np.float64)
                bootstrap = float(ep.mean()) * self.effective_horizon if bool(truncated[t]) else 0.0
                for i in reversed(range(n)):
                    values[i] = ep[i] + self.discount * bootstrap
                    bootstrap = values[i]
                self._returns_min[t] = min(self._returns_min[t], float(values.min()))
                self._returns_max[t] = max(self._returns_max[t], float(values.max()))
                self._episode_rewards[t] = []

    def add(self, obs, next_obs, action, reward, done, terminal=None, truncated=None) -> None:
        # buffers.py:426-474
        assert obs.ndim == 2 and action.ndim == 2 and reward.ndim <= 2 and done.ndim <= 2
        assert obs.shape[0] == action.shape[0] == reward.shape[0] == done.shape[0] == self.num_tasks
        self.obs[self.pos] = obs
        self.actions[self.pos] = action
        self.next_obs[self.pos] = next_obs
        self.dones[self.pos] = done.reshape(-1, 1)
        self.rewards[self.pos] = reward.reshape(-1, 1)
        if self.normalize_rewards:
            self._min_rewards = np.minimum(self._min_rewards, reward.reshape(-1))
            self._max_rewards = np.maximum(self._max_rewards, reward.reshape(-1))
        if self.use_return_normalization:
            _terminal = terminal if terminal is not None else done
            _truncated = truncated if truncated is not None else np.zeros_like(done)
            self._update_return_stats(reward.flatten(), np.asarray(_terminal).flatten().astype(bool),
                                      np.asarray(_truncated).flatten().astype(bool))
        self._advance_position(1)

    def _fill(self) -> int:
        return self.pos if not self.full else self.capacity

    def _reward_scale_shift(self):
        """Per-task (shift, scale) so that reward' = (reward - shift) * scale (buffers.py:392-422, 531-538)."""
        if self.use_return_normalization:
            no_data = np.isinf(self._returns_min) | np.isinf(self._returns_max)
            den = np.where(self._returns_max >= np.abs(self._returns_min), self._returns_max,
                           np.abs(self._returns_min))
            den = den / self.v_max
            den = np.where(no_data | (den < self.reward_norm_eps), 1.0, den)
            return np.zeros(self.num_tasks), 1.0 / den, den
        if self.normalize_rewards:
            den = self._max_rewards - self._min_rewards + self.reward_norm_eps
            return self._min_rewards.copy(), 1.0 / den, den
        return None

    def draw_indices(self, batch_size) -> list[np.ndarray]:
        """The index vectors `sample` draws, in draw order (one shared vector, or one per task)."""
        if isinstance(batch_size, np.ndarray):
            out = []
            for t in range(self.num_tasks):
                n = int(batch_size[t])
                if n > 0:
                    out.append(self._rng.integers(max(self._fill(), n), n))
                else:
                    out.append(np.zeros(0, dtype=np.int64))
            return out
        single = batch_size // self.num_tasks
        return [self._rng.integers(max(self._fill(), single), single)]

    def sample(self, batch_size) -> ReplayBufferSamples:
        if isinstance(batch_size, np.ndarray):  # buffers.py:496-519
            assert len(batch_size) == self.num_tasks
            assert batch_size.sum() == 128 * self.num_tasks
            idx = self.draw_indices(batch_size)
            parts = [[], [], [], [], []]
            for t in range(self.num_tasks):
                if len(idx[t]) > 0:
                    parts[0].append(self.obs[idx[t], t])
                    parts[1].append(self.actions[idx[t], t])
                    parts[2].append(self.next_obs[idx[t], t])
                    parts[3].append(self.dones[idx[t], t])
                    parts[4].append(self.rewards[idx[t], t])
            return ReplayBufferSamples(*(np.concatenate(p, axis=0) for p in parts))
        assert batch_size % self.num_tasks == 0  # buffers.py:521
        single = batch_size // self.num_tasks
        (sample_idx,) = self.draw_indices(batch_size)
        rewards = self.rewards[sample_idx]
        if self.use_return_normalization:
            _, _, den = self._reward_scale_shift()
            rewards = rewards / den[np.newaxis, :, np.newaxis]
        elif self.normalize_rewards:
            mn = self._min_rewards[np.newaxis, :, np.newaxis]
            mx = self._max_rewards[np.newaxis, :, np.newaxis]
            rewards = (rewards - mn) / (mx - mn + self.reward_norm_eps)
        batch = (self.obs[sample_idx], self.actions[sample_idx], self.next_obs[sample_idx],
                 self.dones[sample_idx], rewards)
        mt = single * self.num_tasks
        return ReplayBufferSamples(*(x.reshape(mt, *x.shape[2:]) for x in batch))  # buffers.py:547-549

    def single_task_sample(self, task_idx: int, batch_size: int) -> ReplayBufferSamples:
        # buffers.py:478-492.  NOTE the reference indexes `self.obs[sample_idx][task_idx]`, i.e. it
        # takes row `task_idx` of the *sample* axis and returns all tasks of that one sample
        # (shape (num_tasks, dim)); restated literally.
        assert task_idx < self.num_tasks
        idx = self._rng.integers(max(self._fill(), batch_size), batch_size)
        return ReplayBufferSamples(self.obs[idx][task_idx], self.actions[idx][task_idx],
                                   self.next_obs[idx][task_idx], self.dones[idx][task_idx],
                                   self.rewards[idx][task_idx])

    def checkpoint(self) -> dict:  # buffers.py:308-324
        return {
            "data": {"obs": self.obs, "actions": self.actions, "rewards": self.rewards,
                     "next_obs": self.next_obs, "dones": self.dones, "pos": self.pos, "full": self.full,
                     "returns_min": self._returns_min, "returns_max": self._returns_max},
            "rng_state": self._rng.get_state(),
        }

    def load_checkpoint(self, ckpt: dict) -> None:  # buffers.py:326-335
        for key in ["data", "rng_state"]:
            assert key in ckpt
        for key in ["obs", "actions", "rewards", "next_obs", "dones", "pos", "full"]:
            assert key in ckpt["data"]
            setattr(self, key, ckpt["data"][key])
        self._returns_min = ckpt["data"].get("returns_min", self._returns_min)
        self._returns_max = ckpt["data"].get("returns_max", self._returns_max)
        self._rng.set_state(ckpt["rng_state"])
