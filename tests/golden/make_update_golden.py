#!/usr/bin/env python
"""Freezes known answers of the MT-SAC update oracle (SURVEY.md 8(c) "Golden vectors to generate and commit").

The reference's own `MTSAC.update` cannot run in this image (no jax / flax / optax / distrax), so these vectors do NOT pin
the oracle to the reference; they pin the oracle to ITSELF as of the commit that generated them, so that neither the
oracle nor the CUDA path can drift unnoticed (tests/test_update_golden.py checks both against these files).

  python tests/golden/make_update_golden.py          # rewrites tests/golden/update_*.npz

Two configurations x seeds {0, 1}, fp64 oracle, two consecutive updates each (the second one has non-zero Adam moments):
  small  (T=10, W=64,  B=80):   inputs stored in full (fp32), outputs in full (fp64)
  mt10   (T=10, W=400, B=1280): inputs are regenerated from the seed (their checksums are stored and verified), outputs as
                                the 10 log scalars per step + per-leaf {sum, sum of squares, 64 strided samples}
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import mtsac_oracle as O  # noqa: E402

CONFIGS = {
    "small": dict(num_tasks=10, obs_dim=49, action_dim=4, width=64, per_task=8, clip=False),
    "mt10": dict(num_tasks=10, obs_dim=49, action_dim=4, width=400, per_task=128, clip=True),
}
SEEDS = (0, 1)
STEPS = 2
N_SAMPLES = 64


def oracle_config(name: str) -> tuple[O.OracleConfig, int]:
    c = dict(CONFIGS[name])
    per_task = c.pop("per_task")
    return O.OracleConfig(**c), per_task


def flat_leaves(tree, prefix=""):
    out = {}
    for k, v in tree.items():
        if isinstance(v, dict):
            out.update(flat_leaves(v, f"{prefix}{k}/"))
        else:
            out[f"{prefix}{k}"] = v
    return out


def state_leaves(st: O.OracleState) -> dict:
    """Every tensor of an oracle state under a flat name."""
    out = {}
    for name, tree in (("actor", st.actor), ("critic", st.critic), ("target", st.critic_target)):
        for k, v in flat_leaves(tree).items():
            out[f"{name}/{k}"] = v
    for net in ("actor", "critic"):
        for mom in ("m", "v"):
            for k, v in flat_leaves(st.opt[net][mom]).items():
                out[f"{net}_{mom}/{k}"] = v
    out["log_alpha"] = st.log_alpha
    out["alpha_m"] = st.opt["alpha"]["m"]
    out["alpha_v"] = st.opt["alpha"]["v"]
    return out


def sample_idx(n: int) -> np.ndarray:
    return np.unique(np.linspace(0, n - 1, N_SAMPLES).astype(np.int64))


def summarize(x: torch.Tensor) -> np.ndarray:
    f = x.detach().double().flatten().numpy()
    return np.concatenate(([f.sum(), (f * f).sum()], f[sample_idx(f.size)]))


def run(name: str, seed: int):
    cfg, per_task = oracle_config(name)
    st32 = O.init_state(cfg, seed=seed, dtype=torch.float32)
    st = st32.to(torch.float64)
    logs_all, batches = [], []
    for step in range(STEPS):
        batch, ec, ea = O.synthetic_batch(cfg, per_task, seed=1000 * seed + 17 + step, dtype=torch.float32)
        batches.append((batch, ec, ea))
        st, logs = O.mtsac_update(st, tuple(b.double() for b in batch), ec.double(), ea.double(), cfg)
        logs_all.append(np.array([float(logs[k]) for k in O.LOG_KEYS], dtype=np.float64))
    return cfg, st32, st, np.stack(logs_all), batches


def main() -> None:
    torch.set_num_threads(1)   # fixed summation order inside the BLAS calls
    for name in CONFIGS:
        for seed in SEEDS:
            cfg, st32, st, logs, batches = run(name, seed)
            out = {"logs": logs, "log_keys": np.array(O.LOG_KEYS), "steps": np.array(STEPS)}
            inputs = {f"in/{k}": v.numpy() for k, v in state_leaves(st32).items() if "_m/" not in k and "_v/" not in k and not k.startswith("alpha_")}
            for i, (batch, ec, ea) in enumerate(batches):
                for fname, x in zip(("observations", "actions", "next_observations", "dones", "rewards"), batch):
                    inputs[f"in/step{i}/{fname}"] = x.numpy()
                inputs[f"in/step{i}/eps_c"], inputs[f"in/step{i}/eps_a"] = ec.numpy(), ea.numpy()
            if name == "small":
                out.update(inputs)
                out.update({f"out/{k}": v.numpy() for k, v in state_leaves(st).items()})
            else:
                out.update({f"insum/{k[3:]}": np.array([v.astype(np.float64).sum(), (v.astype(np.float64) ** 2).sum()]) for k, v in inputs.items()})
                out.update({f"outsum/{k}": summarize(v) for k, v in state_leaves(st).items()})
            out["counts"] = np.array([st.opt[n]["count"] for n in ("actor", "critic", "alpha")])
            path = os.path.join(HERE, f"update_{name}_seed{seed}.npz")
            np.savez_compressed(path, **out)
            print(path, os.path.getsize(path) // 1024, "KiB", {k: f"{v:.6g}" for k, v in zip(O.LOG_KEYS, logs[-1])})


if __name__ == "__main__":
    main()
