"""Generates tests/golden/sampler_*.npz by running the REFERENCE's own MultiTaskReplayBuffer
(/root/reference/mtrl/rl/buffers.py) in the build container.  buffers.py is NumPy-only but imports
gymnasium and (through mtrl/types.py -> jaxtyping.Array) jax for type annotations; both are absent
here, so they are stubbed with empty modules.  No reference code is copied: the module is loaded
from where it lies.  Run once: `python tests/golden/make_sampler_golden.py`.
"""
import importlib.util
import os
import sys
import types

import numpy as np

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))


def load_reference_buffers():
    g = types.ModuleType("gymnasium")
    g.Space = object
    sys.modules.setdefault("gymnasium", g)
    j = types.ModuleType("jax")
    j.Array = type("Array", (), {})
    sys.modules.setdefault("jax", j)
    pkg = types.ModuleType("mtrl")
    pkg.__path__ = []
    sys.modules["mtrl"] = pkg

    def load(name, path):
        spec = importlib.util.spec_from_file_location(name, path)
        m = importlib.util.module_from_spec(spec)
        sys.modules[name] = m
        spec.loader.exec_module(m)
        return m

    load("mtrl.types", f"{REF}/mtrl/types.py")
    return load("mtrl.rl.buffers", f"{REF}/mtrl/rl/buffers.py")


class Space:
    def __init__(self, shape):
        self.shape = shape


def transitions(rng, T, obs_dim, act_dim, step):
    obs = rng.standard_normal((T, obs_dim)).astype(np.float32)
    nxt = rng.standard_normal((T, obs_dim)).astype(np.float32)
    act = rng.uniform(-1, 1, (T, act_dim)).astype(np.float32)
    rew = rng.uniform(0, 10, (T,)).astype(np.float32)
    done = (rng.uniform(size=(T,)) < 0.05).astype(np.float32)
    return obs, nxt, act, rew, done


def scenario(buffers, name, T, obs_dim, act_dim, cap_per_task, seed, n_add, batch_sizes, **kw):
    """Adds n_add transitions, then samples with each entry of batch_sizes; records everything."""
    buf = buffers.MultiTaskReplayBuffer(cap_per_task * T, T, Space((obs_dim,)), Space((act_dim,)), seed=seed, **kw)
    data_rng = np.random.default_rng(1000 + seed)
    rec = {"T": T, "obs_dim": obs_dim, "act_dim": act_dim, "cap": cap_per_task, "seed": seed, "n_add": n_add}
    adds = [transitions(data_rng, T, obs_dim, act_dim, i) for i in range(n_add)]
    for k, arrs in enumerate(zip(*adds)):
        rec[f"add_{k}"] = np.stack(arrs)
    for (o, n, a, r, d) in adds:
        buf.add(o, n, a, r, d)
    rec["pos"] = buf.pos
    rec["full"] = buf.full
    for i, bs in enumerate(batch_sizes):
        s = buf.sample(bs)
        rec[f"bs_{i}"] = np.asarray(bs)
        for f, v in zip(s._fields, s):
            rec[f"s{i}_{f}"] = np.asarray(v)
    st = buf._rng.bit_generator.state
    rec["final_state_hi"] = np.uint64(st["state"]["state"] >> 64)
    rec["final_state_lo"] = np.uint64(st["state"]["state"] & ((1 << 64) - 1))
    rec["final_has_uint32"] = st["has_uint32"]
    rec["final_uinteger"] = np.uint32(st["uinteger"])
    np.savez_compressed(os.path.join(OUT, f"sampler_{name}.npz"), **rec)
    print("wrote", name)


def index_stream(buffers, name, T, cap_per_task, seed, fill, calls, single):
    """Golden index streams: obs column 0 stores the ring position so samples reveal the indices."""
    buf = buffers.MultiTaskReplayBuffer(cap_per_task * T, T, Space((1,)), Space((1,)), seed=seed)
    buf.obs[:, :, 0] = np.arange(cap_per_task, dtype=np.float32)[:, None]
    buf.pos = fill % cap_per_task
    buf.full = fill >= cap_per_task
    out = np.zeros((calls, single), dtype=np.int64)
    for c in range(calls):
        s = buf.sample(single * T)
        out[c] = s.observations.reshape(single, T)[:, 0].astype(np.int64)
    st = buf._rng.bit_generator.state
    np.savez_compressed(
        os.path.join(OUT, f"sampler_{name}.npz"), idx=out, T=T, cap=cap_per_task, seed=seed, fill=fill,
        single=single, final_state_hi=np.uint64(st["state"]["state"] >> 64),
        final_state_lo=np.uint64(st["state"]["state"] & ((1 << 64) - 1)),
        final_has_uint32=st["has_uint32"], final_uinteger=np.uint32(st["uinteger"]))
    print("wrote", name)


if __name__ == "__main__":
    b = load_reference_buffers()
    # ragged / not-full / wrap-around / odd counts
    scenario(b, "small_notfull", T=3, obs_dim=5, act_dim=2, cap_per_task=16, seed=0, n_add=7,
             batch_sizes=[3, 9, 15, 3 * 11])                      # high = max(pos, single) incl. single > pos
    scenario(b, "small_wrap", T=4, obs_dim=6, act_dim=3, cap_per_task=8, seed=1, n_add=19,
             batch_sizes=[4, 8, 4 * 5, 4 * 8])                    # buffer wrapped (full=True)
    scenario(b, "mt10_shape", T=10, obs_dim=49, act_dim=4, cap_per_task=256, seed=42, n_add=40,
             batch_sizes=[1280, 1280, 10])                        # MT10 row shapes, B=1280
    scenario(b, "normalized", T=3, obs_dim=4, act_dim=2, cap_per_task=16, seed=3, n_add=12,
             batch_sizes=[6, 12], normalize_rewards=True)
    scenario(b, "per_task_counts", T=2, obs_dim=3, act_dim=2, cap_per_task=300, seed=5, n_add=200,
             batch_sizes=[np.array([100, 156]), np.array([256, 0]), 2 * 128])
    # long index streams incl. Lemire rejections (fill=100000: ~2e-3 rejects per 128-draw call)
    index_stream(b, "idx_mt50", T=2, cap_per_task=100_000, seed=1, fill=100_000, calls=1000, single=128)
    index_stream(b, "idx_partial", T=2, cap_per_task=100_000, seed=7, fill=4001, calls=300, single=128)
    index_stream(b, "idx_big", T=1, cap_per_task=3_000_000, seed=11, fill=3_000_000, calls=50, single=127)
