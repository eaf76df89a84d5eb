"""GPU: device MultiTaskRolloutBuffer + the CUDA GAE scan (csrc/rollout.cu) vs the reference's storage golden and the
NumPy restatement -- bit-exact (float32, NumPy's evaluation order)."""
import os

import numpy as np
import pytest
import torch

import sac_util as SU
from oracle.rollout_oracle import MultiTaskRolloutBufferOracle

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "rollout_storage.npz")
STORED = ("observations", "actions", "rewards", "dones", "log_probs", "means", "stds", "values")


def test_storage_matches_reference_buffer(cuda):
    from mtrl_b200.rl.buffers import MultiTaskRolloutBuffer

    g = np.load(GOLDEN)
    buf = MultiTaskRolloutBuffer(int(g["S"]), int(g["T"]), SU._Space((int(g["obs_dim"]),)), SU._Space((int(g["act_dim"]),)), seed=0)
    for i in range(int(g["S"])):
        buf.add(g["add_obs"][i], g["add_action"][i], g["add_reward"][i], g["add_done"][i], value=g["add_value"][i],
                log_prob=g["add_log_prob"][i], mean=g["add_mean"][i], std=g["add_std"][i])
    assert buf.ready
    for name in STORED:
        assert np.array_equal(getattr(buf, name).cpu().numpy(), g[f"stored_{name}"]), name
    with pytest.raises(IndexError):
        buf.add(g["add_obs"][0], g["add_action"][0], g["add_reward"][0], g["add_done"][0])


@pytest.mark.parametrize("S,T,p_done", [(1, 7, 0.3), (6, 5, 0.2), (37, 50, 0.05), (10000, 50, 0.002), (300, 200, 0.01)])
def test_gae_bit_exact_vs_numpy(cuda, S, T, p_done):
    from mtrl_b200.rl.buffers import MultiTaskRolloutBuffer

    rng = np.random.default_rng(S * 1000 + T)
    od, ad = 8, 4
    dev = MultiTaskRolloutBuffer(S, T, SU._Space((od,)), SU._Space((ad,)), seed=0)
    orc = MultiTaskRolloutBufferOracle(S, T, od, ad, seed=0)
    # bulk fill (the per-step add path is covered above)
    for name, shape in (("rewards", (S, T, 1)), ("values", (S, T, 1))):
        x = (rng.standard_normal(shape) * (5.0 if name == "rewards" else 20.0)).astype(np.float32)
        setattr(orc, name, x)
        getattr(dev, name).copy_(torch.from_numpy(x))
    d = (rng.uniform(size=(S, T, 1)) < p_done).astype(np.float32)
    orc.dones = d
    dev.dones.copy_(torch.from_numpy(d))
    orc.pos = dev.pos = S
    dev._values_pushed = True
    lv = (rng.standard_normal(T) * 20).astype(np.float32)
    ld = (rng.uniform(size=T) < 0.3).astype(np.float32)
    ref = orc.get(True, lv, ld, gamma=0.99, gae_lambda=0.97)
    got = dev.get(True, lv, ld, gamma=0.99, gae_lambda=0.97)
    assert got.advantages.shape == (T, S, 1) and got.observations.shape == (T, S, od)
    assert np.array_equal(got.advantages.cpu().numpy(), ref[9]), "advantages differ from NumPy float32 evaluation"
    assert np.array_equal(got.returns.cpu().numpy(), ref[8])
    assert np.array_equal(got.values.cpu().numpy(), ref[7])


def test_get_without_values_asserts(cuda):
    from mtrl_b200.rl.buffers import MultiTaskRolloutBuffer

    buf = MultiTaskRolloutBuffer(2, 3, SU._Space((5,)), SU._Space((2,)))
    z = np.zeros((3, 5), np.float32)
    for _ in range(2):
        buf.add(z, np.zeros((3, 2), np.float32), np.zeros(3, np.float32), np.zeros(3, np.float32))
    with pytest.raises(AssertionError):
        buf.get(True, np.zeros(3, np.float32), np.zeros(3, np.float32))
    out = buf.get(False)
    assert out.returns is None and out.advantages is None and out.actions.shape == (3, 2, 2)
