"""CUDA replay sampler vs (a) the reference's own outputs (golden files), (b) the oracle on seeded
inputs, (c) size-independent properties at the MT50 shapes.  Bit-exact everywhere."""
import numpy as np
import pytest
import torch

import golden_util as GU

pytestmark = pytest.mark.gpu


class Space:
    def __init__(self, shape):
        self.shape = shape


def make(T, obs_dim, act_dim, cap, seed, **kw):
    from mtrl_b200.rl.buffers import MultiTaskReplayBuffer

    return MultiTaskReplayBuffer(cap * T, T, Space((obs_dim,)), Space((act_dim,)), seed=seed, **kw)


@pytest.mark.parametrize("path", GU.scenario_files(), ids=lambda p: p.split("sampler_")[-1][:-4])
def test_matches_reference_golden(cuda, path):
    g = np.load(path)
    kw = {"normalize_rewards": True} if "normalized" in path else {}
    buf = make(int(g["T"]), int(g["obs_dim"]), int(g["act_dim"]), int(g["cap"]), int(g["seed"]), **kw)
    GU.replay_adds(g, buf)
    assert buf.pos == int(g["pos"]) and buf.full == bool(g["full"])
    for i in range(GU.n_samples(g)):
        s = buf.sample(GU.batch_size_of(g, i))
        for f, v in zip(s._fields, s):
            ref = g[f"s{i}_{f}"]
            out = v.cpu().numpy()
            assert out.shape == ref.shape, (i, f)
            # normalised rewards: the reference returns float64, JAX casts to fp32 at the jit boundary
            assert np.array_equal(out, ref.astype(np.float32)), (i, f)
    state, has, ui = GU.final_state(g)
    st = buf._rng.bit_generator.state
    assert st["state"]["state"] == state and st["has_uint32"] == has and st["uinteger"] == ui


@pytest.mark.parametrize("path", GU.index_files(), ids=lambda p: p.split("sampler_")[-1][:-4])
def test_index_streams_match_reference(cuda, path):
    g = np.load(path)
    T, cap, fill, single = int(g["T"]), int(g["cap"]), int(g["fill"]), int(g["single"])
    buf = make(T, 1, 1, cap, int(g["seed"]))
    buf.pos = fill % cap
    buf.full = fill >= cap
    idx = g["idx"]
    got = []
    for c in range(idx.shape[0]):
        _, ix = buf.sample(single * T, return_indices=True)
        got.append(ix)
    got = torch.stack(got).cpu().numpy()
    assert np.array_equal(got, idx)
    state, has, ui = GU.final_state(g)
    st = buf._rng.bit_generator.state
    assert st["state"]["state"] == state and st["has_uint32"] == has and st["uinteger"] == ui


def test_vs_oracle_mt50_rows(cuda):
    """MT50 row shapes (obs 89, act 4, B = 6400) on a smaller ring: oracle and kernel see identical
    storage and seeds; outputs and RNG state must be identical after several calls."""
    from oracle.sampler_oracle import MultiTaskReplayBufferOracle

    T, od, ad, cap = 50, 89, 4, 2048
    buf = make(T, od, ad, cap, seed=1)
    orc = MultiTaskReplayBufferOracle(cap * T, T, od, ad, seed=1)
    g = torch.Generator(device="cpu").manual_seed(7)
    for name, d in (("obs", od), ("actions", ad), ("next_obs", od), ("dones", 1), ("rewards", 1)):
        x = torch.randn(cap, T, d, generator=g)
        getattr(buf, name).copy_(x)
        setattr(orc, name, x.numpy().copy())
    for fill in (cap, 1000, 128, 5):
        buf.pos = orc.pos = fill % cap
        buf.full = orc.full = fill >= cap
        for _ in range(3):
            s, o = buf.sample(128 * T), orc.sample(128 * T)
            for a, b in zip(s, o):
                assert np.array_equal(a.cpu().numpy(), b)
    assert buf._rng.bit_generator.state["state"] == orc._rng.get_state()["state"]


def test_full_size_properties(cuda):
    """MT50 at the reference capacity (100 000 per task, 3.7 GB): every output row is the storage row the
    returned index names; the gather is deterministic given the RNG state; rows are (sample, task) interleaved."""
    T, od, ad, cap = 50, 89, 4, 100_000
    buf = make(T, od, ad, cap, seed=1)
    gen = torch.Generator(device="cuda").manual_seed(3)
    for name in ("obs", "actions", "next_obs", "dones", "rewards"):
        t = getattr(buf, name)
        t.copy_(torch.randn(t.shape, generator=gen, device="cuda"))
    buf.full = True
    state = buf._rng.bit_generator.state
    s, idx = buf.sample(128 * T, return_indices=True)
    assert idx.min() >= 0 and idx.max() < cap
    for out, name in zip(s, ("obs", "actions", "next_obs", "dones", "rewards")):
        store = getattr(buf, name)
        assert torch.equal(out, store[idx].reshape(128 * T, -1)), name
    task = s.observations.reshape(128, T, od)  # row i*T+t is task t
    assert torch.equal(task[:, 7], buf.obs[idx, 7])
    buf._rng.bit_generator.state = state
    s2 = buf.sample(128 * T)
    assert all(torch.equal(a, b) for a, b in zip(s, s2))


def test_checkpoint_roundtrip_and_numpy_state(cuda):
    buf = make(3, 5, 2, 32, seed=9)
    rng = np.random.default_rng(0)
    for _ in range(40):
        buf.add(rng.standard_normal((3, 5)).astype(np.float32), rng.standard_normal((3, 5)).astype(np.float32),
                rng.standard_normal((3, 2)).astype(np.float32), rng.standard_normal(3).astype(np.float32),
                np.zeros(3, np.float32))
    buf.sample(3 * 7)  # odd count leaves a buffered 32-bit half
    ck = buf.checkpoint()
    assert set(ck["data"]) >= {"obs", "actions", "rewards", "next_obs", "dones", "pos", "full", "returns_min", "returns_max"}
    a = buf.sample(3 * 9)
    buf2 = make(3, 5, 2, 32, seed=123)
    buf2.load_checkpoint(ck)
    b = buf2.sample(3 * 9)
    assert all(torch.equal(x, y) for x, y in zip(a, b))
    # the state dict is numpy's: a numpy Generator restored from it continues the same stream
    g = np.random.default_rng(0)
    g.bit_generator.state = ck["rng_state"]
    expect = g.integers(0, 32, size=(9,))
    assert np.array_equal(a.observations.reshape(9, 3, 5)[:, 0].cpu().numpy(), ck["data"]["obs"][expect, 0])


def test_edge_cases(cuda):
    buf = make(2, 3, 1, 8, seed=0)
    with pytest.raises(AssertionError):
        buf.sample(3)  # not divisible by the number of tasks (buffers.py:521)
    with pytest.raises(IndexError):
        buf.sample(2 * 9)  # more rows per task than the ring holds
    st = buf._rng.bit_generator.state
    s = buf.sample(2)  # empty buffer, single = 1 -> high = 1: zeros, no randomness consumed
    assert torch.count_nonzero(s.observations) == 0
    assert buf._rng.bit_generator.state == st
    with pytest.raises(AssertionError):
        buf.add(np.zeros((3, 3), np.float32), np.zeros((3, 3), np.float32), np.zeros((3, 1), np.float32),
                np.zeros(3, np.float32), np.zeros(3, np.float32))


@pytest.mark.parametrize("high", [1, 2, 3, 128, 99_999, 100_000, 1_500_000_000, 3_000_000_000, 4_294_967_295])
def test_raw_draws_match_live_numpy_including_rejections(cuda, high):
    """`rng.integers(0, high, n)` through the jump-ahead draw kernel vs numpy itself.  high = 1.5e9 / 3e9 reject ~30 % of
    the draws (Lemire threshold 2^32 mod high), which exercises the sequential redo; odd sizes exercise the buffered
    32-bit half across calls; the final generator state must equal numpy's."""
    import ctypes as C

    from mtrl_b200 import _lib as L

    seed = 12345 + (high % 97)
    buf = make(2, 1, 1, 64, seed)
    rng = np.random.default_rng(seed)
    out = torch.empty(4096, dtype=torch.int64, device="cuda")
    for n in (1, 128, 127, 1, 2, 33, 1280, 4095, 64):
        ref = rng.integers(low=0, high=high, size=(n,))
        L.check(L.lib().mtrl_sampler_draw(buf._h, C.c_ulonglong(high), n, C.c_void_p(out.data_ptr()),
                                          C.c_void_p(L.current_stream_ptr())))
        got = out[:n].cpu().numpy()
        assert np.array_equal(got, ref), f"high={high} n={n}"
    st, ref_st = buf._rng.bit_generator.state, rng.bit_generator.state
    assert st["state"] == ref_st["state"] and st["has_uint32"] == ref_st["has_uint32"]
    if ref_st["has_uint32"]:
        assert st["uinteger"] == ref_st["uinteger"]


def test_per_task_counts_beyond_4096_rows(cuda):
    """sample(ndarray) only requires sum == 128 T (buffers.py:498), so with MT50 a single task may receive thousands of
    rows; the index scratch is sized from the ring capacity (it used to be a fixed 4096)."""
    from oracle.sampler_oracle import MultiTaskReplayBufferOracle

    T, od, ad, cap = 50, 6, 2, 6000
    buf = make(T, od, ad, cap, seed=4)
    orc = MultiTaskReplayBufferOracle(cap * T, T, od, ad, seed=4)
    rng = np.random.default_rng(1)
    obs = rng.standard_normal((cap, T, od)).astype(np.float32)
    for name in ("obs", "next_obs"):
        getattr(buf, name).copy_(torch.from_numpy(obs))
        getattr(orc, name)[:] = obs
    buf.full = orc.full = True
    counts = np.zeros(T, dtype=np.int64)
    counts[3], counts[17] = 5000, 128 * T - 5000
    a, b = buf.sample(counts), orc.sample(counts)
    for x, y in zip(a, b):
        assert np.array_equal(x.cpu().numpy(), y)
    big = buf.sample(5000 * T)          # sample(int) with more than 4096 rows per task
    ref = orc.sample(5000 * T)
    assert np.array_equal(big.observations.cpu().numpy(), ref.observations)
