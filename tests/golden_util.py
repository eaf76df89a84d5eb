"""Helpers shared by the sampler tests: replay a golden scenario through any buffer implementation."""
import glob
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
FIELDS = ("observations", "actions", "next_observations", "dones", "rewards")


def scenario_files():
    return sorted(f for f in glob.glob(os.path.join(GOLDEN, "sampler_*.npz")) if "sampler_idx_" not in f)


def index_files():
    return sorted(glob.glob(os.path.join(GOLDEN, "sampler_idx_*.npz")))


def n_samples(g):
    return len([k for k in g.files if k.startswith("bs_")])


def replay_adds(g, buf):
    obs, nxt, act, rew, done = (g[f"add_{k}"] for k in range(5))
    for i in range(int(g["n_add"])):
        buf.add(obs[i], nxt[i], act[i], rew[i], done[i])


def batch_size_of(g, i):
    bs = g[f"bs_{i}"]
    return int(bs) if bs.ndim == 0 else np.asarray(bs)


def final_state(g):
    return (int(g["final_state_hi"]) << 64) | int(g["final_state_lo"]), int(g["final_has_uint32"]), int(g["final_uinteger"])
