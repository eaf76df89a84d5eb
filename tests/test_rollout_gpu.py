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


def test_rollout_feeds_mtppo_update(cuda):
    """On-policy side end to end: device rollout buffer -> GAE kernel -> `Rollout` of (task, timestep,
    dim) views -> MTPPO.update, equal to the update on the same arrays flattened task-major by hand."""
    from mtrl_b200.config.networks import ContinuousActionPolicyConfig, ValueFunctionConfig
    from mtrl_b200.config.nn import MultiHeadConfig
    from mtrl_b200.config.optim import OptimizerConfig
    from mtrl_b200.rl.algorithms import MTPPO, MTPPOConfig
    from mtrl_b200.rl.buffers import MultiTaskRolloutBuffer
    from mtrl_b200.types import Rollout

    T, S, W, od, ad = 3, 128, 64, 16 + 3, 4
    rng = np.random.default_rng(5)
    buf = MultiTaskRolloutBuffer(S, T, SU._Space((od,)), SU._Space((ad,)), seed=0)
    for _ in range(S):
        obs = rng.standard_normal((T, od)).astype(np.float32)
        obs[:, -T:] = np.eye(T, dtype=np.float32)
        buf.add(obs, rng.uniform(-1, 1, (T, ad)).astype(np.float32), rng.uniform(0, 1, T).astype(np.float32),
                (rng.uniform(size=T) < 0.02).astype(np.float32), value=rng.standard_normal((T, 1)).astype(np.float32),
                log_prob=(rng.standard_normal(T) - 4).astype(np.float32))
    roll = buf.get(True, rng.standard_normal(T).astype(np.float32), np.zeros(T, np.float32))
    assert roll.advantages.shape == (T, S, 1) and not roll.advantages.is_contiguous()

    def agent():
        opt = OptimizerConfig(max_grad_norm=1.0)
        net = MultiHeadConfig(width=W, depth=2, num_tasks=T, optimizer=opt)
        cfg = MTPPOConfig(num_tasks=T, policy_config=ContinuousActionPolicyConfig(network_config=net, squash_tanh=False),
                          vf_config=ValueFunctionConfig(network_config=net))
        return MTPPO.initialize(cfg, SU.EnvSpec(od, ad), seed=3, rollout_steps=S)

    eps = torch.randn(T * S, ad, generator=torch.Generator().manual_seed(1))
    a, b = agent(), agent()
    _, logs_a = a.update(roll, eps=eps)
    flat = Rollout(*(None if x is None else x.reshape(T * S, -1).contiguous() for x in roll))
    _, logs_b = b.update(flat, eps=eps)
    # split-K weight gradients accumulate with atomics, so two runs agree to rounding, not bit for bit
    for k in logs_a:
        assert torch.isfinite(logs_a[k]) and abs(float(logs_a[k]) - float(logs_b[k])) <= 1e-5 * abs(float(logs_b[k])) + 1e-7, k
    assert torch.allclose(a._flat["policy_params"], b._flat["policy_params"], rtol=1e-4, atol=1e-6)
    assert torch.allclose(a._flat["vf_params"], b._flat["vf_params"], rtol=1e-4, atol=1e-6)
