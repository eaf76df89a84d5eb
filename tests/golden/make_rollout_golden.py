"""Generates tests/golden/rollout_*.npz by running the REFERENCE's own MultiTaskRolloutBuffer
(/root/reference/mtrl/rl/buffers.py:552-707) in the build container (gymnasium / jax stubbed as in
make_sampler_golden.py; no reference code is copied).  Run once: `python tests/golden/make_rollout_golden.py`.

What the reference can produce: the storage arrays after `add` (buffers.py:611-648).  Its `get` (:650-707) cannot run:
`.transpose(1, 0)` on the 3-D storage arrays raises "axes don't match array" (:696-705), and with
compute_advantages=True the last timestep multiplies by the whole `self.dones` array (:677) so the (S, T, 1) result
does not fit the (T, 1) slot.  The golden file records both error messages; the restatement implements what the
`Rollout` annotations (mtrl/types.py:48-63, "task timestep") and the cited upstream loop say.
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from make_sampler_golden import OUT, Space, load_reference_buffers  # noqa: E402

STORED = ("observations", "actions", "rewards", "dones", "log_probs", "means", "stds", "values")


def fill(buf, rng, S, T, od, ad):
    adds = []
    for _ in range(S):
        a = dict(obs=rng.standard_normal((T, od)).astype(np.float32), action=rng.uniform(-1, 1, (T, ad)).astype(np.float32),
                 reward=rng.uniform(0, 10, (T,)).astype(np.float32), done=(rng.uniform(size=(T,)) < 0.2).astype(np.float32),
                 value=rng.standard_normal((T, 1)).astype(np.float32), log_prob=rng.standard_normal((T,)).astype(np.float32),
                 mean=rng.standard_normal((T, ad)).astype(np.float32), std=rng.uniform(0.1, 1, (T, ad)).astype(np.float32))
        buf.add(**a)
        adds.append(a)
    return adds


def main():
    B = load_reference_buffers()
    rng = np.random.default_rng(7)
    S, T, od, ad = 6, 5, 12, 4
    buf = B.MultiTaskRolloutBuffer(S, T, Space((od,)), Space((ad,)), seed=0)
    adds = fill(buf, rng, S, T, od, ad)
    assert buf.ready
    rec = {"S": S, "T": T, "obs_dim": od, "act_dim": ad, "pos": np.array(buf.pos)}
    for k in adds[0]:
        rec[f"add_{k}"] = np.stack([a[k] for a in adds])
    for name in STORED:
        rec[f"stored_{name}"] = getattr(buf, name).copy()
    last_values = rng.standard_normal((T,)).astype(np.float32)
    last_dones = (rng.uniform(size=(T,)) < 0.5).astype(np.float32)
    rec["last_values"], rec["last_dones"] = last_values, last_dones
    # what the reference's get() does with its own storage
    for flag, kw in (("plain", dict(compute_advantages=False)),
                     ("gae", dict(compute_advantages=True, last_values=last_values, dones=last_dones))):
        try:
            buf.get(**kw)
            msg = ""
        except ValueError as e:
            msg = str(e)
        rec[f"reference_get_{flag}_error"] = np.array(msg)
        print(f"reference get({flag}):", repr(msg) if msg else "ok")
    np.savez_compressed(os.path.join(OUT, "rollout_storage.npz"), **rec)


if __name__ == "__main__":
    main()
