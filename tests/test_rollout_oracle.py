"""CPU: the rollout-buffer restatement (oracle/rollout_oracle.py) against the reference's own storage
(tests/golden/rollout_storage.npz, generated from mtrl/rl/buffers.py:552-648) and GAE known answers."""
import os

import numpy as np

from oracle.rollout_oracle import MultiTaskRolloutBufferOracle, gae

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "rollout_storage.npz")
STORED = ("observations", "actions", "rewards", "dones", "log_probs", "means", "stds", "values")


def replay(g, buf):
    for i in range(int(g["S"])):
        buf.add(g["add_obs"][i], g["add_action"][i], g["add_reward"][i], g["add_done"][i], value=g["add_value"][i],
                log_prob=g["add_log_prob"][i], mean=g["add_mean"][i], std=g["add_std"][i])


def test_storage_matches_reference_buffer():
    g = np.load(GOLDEN)
    buf = MultiTaskRolloutBufferOracle(int(g["S"]), int(g["T"]), int(g["obs_dim"]), int(g["act_dim"]), seed=0)
    replay(g, buf)
    assert buf.ready and buf.pos == int(g["pos"])
    for name in STORED:
        assert np.array_equal(getattr(buf, name), g[f"stored_{name}"]), name


def test_reference_get_is_broken_and_why():
    """The golden file records what the reference's get() does on its own storage: it raises for both modes.  The
    restatement's deviations (transpose to (task, timestep, dim); last step from the `dones` argument) exist because
    of exactly these two errors."""
    g = np.load(GOLDEN)
    assert "axes don't match array" in str(g["reference_get_plain_error"])
    assert "could not broadcast" in str(g["reference_get_gae_error"])


def test_get_layout_and_gae_known_answers():
    g = np.load(GOLDEN)
    S, T = int(g["S"]), int(g["T"])
    buf = MultiTaskRolloutBufferOracle(S, T, int(g["obs_dim"]), int(g["act_dim"]), seed=0)
    replay(g, buf)
    out = buf.get(True, g["last_values"], g["last_dones"], gamma=0.99, gae_lambda=0.97)
    assert out[0].shape == (T, S, int(g["obs_dim"])) and out[9].shape == (T, S, 1)
    assert np.array_equal(out[2], g["stored_rewards"].transpose(1, 0, 2))
    adv, ret = out[9], out[8]
    assert np.array_equal(ret, adv + out[7])
    # closed form in float64: A_t = sum_k (prod_{j<k} c_{t+j}) delta_{t+k}
    r, v, d = (g[f"stored_{n}"].astype(np.float64)[..., 0] for n in ("rewards", "values", "dones"))
    nv = np.concatenate([v[1:], g["last_values"][None].astype(np.float64)])
    nd = np.concatenate([d[1:], g["last_dones"][None].astype(np.float64)])
    delta = r + (1 - nd) * 0.99 * nv - v
    c = (1 - nd) * 0.99 * 0.97
    ref = np.zeros_like(delta)
    for t in range(S):
        w = np.ones(T)
        for k in range(t, S):
            ref[t] += w * delta[k]
            w = w * c[k]
    assert np.allclose(adv[..., 0].T, ref, rtol=1e-5, atol=1e-5)
    # gamma = lambda = 1, no terminations, zero values: advantage = reward-to-go
    rew = np.arange(12, dtype=np.float32).reshape(4, 3, 1)
    z = np.zeros_like(rew)
    a = gae(rew, z, z, np.zeros((3, 1), np.float32), np.zeros((3, 1), np.float32), 1.0, 1.0)
    assert np.array_equal(a[:, :, 0], np.cumsum(rew[::-1, :, 0], axis=0)[::-1])
    # a done flag at step t+1 cuts the bootstrap of step t
    d2 = z.copy()
    d2[2] = 1.0
    a2 = gae(rew, z, d2, np.zeros((3, 1), np.float32), np.zeros((3, 1), np.float32), 1.0, 1.0)
    assert np.array_equal(a2[1], rew[1]) and np.array_equal(a2[0], rew[0] + rew[1])
