"""Known-answer tests of the MT-SAC update: tests/golden/update_*.npz hold fp64 outputs of
oracle/mtsac_oracle.py frozen by tests/golden/make_update_golden.py (two configs x seeds {0, 1}, two consecutive updates).

  * CPU: the oracle, re-run now, reproduces the frozen answers (the restatement cannot drift silently);
  * GPU: the CUDA path, fed the STORED inputs, matches the frozen answers -- every leaf to 1e-3 in the fp32x3 precision,
    the per-network / log tolerances in the tf32 precision.

The reference's own update cannot run in this image, so these files pin the oracle to itself, not to the reference
(DESIGN.md "Oracle and parity status")."""
import os
import sys

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))

import make_update_golden as MG  # noqa: E402
from oracle import mtsac_oracle as O  # noqa: E402

CASES = [(name, seed) for name in MG.CONFIGS for seed in MG.SEEDS]


def load(name, seed):
    return np.load(os.path.join(HERE, "golden", f"update_{name}_seed{seed}.npz"))


def regenerated_inputs(name, seed):
    """Initial state and the per-step (batch, eps_c, eps_a), from the file (small) or from the seed (mt10, verified against
    the stored checksums so that a change of torch's generators is reported as such)."""
    z = load(name, seed)
    cfg, per_task = MG.oracle_config(name)
    st = O.init_state(cfg, seed=seed, dtype=torch.float32)
    batches = [O.synthetic_batch(cfg, per_task, seed=1000 * seed + 17 + step, dtype=torch.float32) for step in range(MG.STEPS)]
    have = {f"in/{k}": v.numpy() for k, v in MG.state_leaves(st).items()
            if "_m/" not in k and "_v/" not in k and not k.startswith("alpha_")}
    for i, (batch, ec, ea) in enumerate(batches):
        for fname, x in zip(("observations", "actions", "next_observations", "dones", "rewards"), batch):
            have[f"in/step{i}/{fname}"] = x.numpy()
        have[f"in/step{i}/eps_c"], have[f"in/step{i}/eps_a"] = ec.numpy(), ea.numpy()
    for k, v in have.items():
        if name == "small":
            assert np.array_equal(z[k], v), f"{k}: regenerated input differs from the stored one"
        else:
            s = np.array([v.astype(np.float64).sum(), (v.astype(np.float64) ** 2).sum()])
            assert np.allclose(z[f"insum/{k[3:]}"], s, rtol=1e-12, atol=0), f"{k}: input checksum differs (torch RNG changed?)"
    return z, cfg, st, batches


@pytest.mark.parametrize("name,seed", CASES)
def test_oracle_reproduces_frozen_answers(name, seed):
    torch.set_num_threads(1)
    z, cfg, st32, batches = regenerated_inputs(name, seed)
    st = st32.to(torch.float64)
    for step, (batch, ec, ea) in enumerate(batches):
        st, logs = O.mtsac_update(st, tuple(b.double() for b in batch), ec.double(), ea.double(), cfg)
        got = np.array([float(logs[k]) for k in O.LOG_KEYS])
        assert np.allclose(got, z["logs"][step], rtol=1e-9, atol=1e-12), (step, dict(zip(O.LOG_KEYS, got - z["logs"][step])))
    for k, v in MG.state_leaves(st).items():
        if name == "small":
            assert np.allclose(v.numpy(), z[f"out/{k}"], rtol=1e-8, atol=1e-13), k
        else:
            assert np.allclose(MG.summarize(v), z[f"outsum/{k}"], rtol=1e-8, atol=1e-13), k
    assert list(z["counts"]) == [MG.STEPS] * 3


def _agent_leaves(agent) -> dict:
    """The agent's device state under the names of MG.state_leaves."""
    import sac_util as SU

    out = {}
    trees = (("actor", agent.actor.params, False), ("critic", agent.critic.params, True), ("target", agent.critic.target_params, True),
             ("actor_m", agent.actor.opt_state["mu"], False), ("actor_v", agent.actor.opt_state["nu"], False),
             ("critic_m", agent.critic.opt_state["mu"], True), ("critic_v", agent.critic.opt_state["nu"], True))
    for name, tree, ens in trees:
        net = SU._net(tree, ens)
        for lk, lv in net.items():
            ok = "heads" if lk == "VmapDense_0" else lk
            for leaf in ("kernel", "bias"):
                out[f"{name}/{ok}/{leaf}"] = lv[leaf]
    out["log_alpha"] = agent.alpha.params["params"]["log_alpha"]
    out["alpha_m"] = agent.alpha.opt_state["mu"]["params"]["log_alpha"]
    out["alpha_v"] = agent.alpha.opt_state["nu"]["params"]["log_alpha"]
    return out


@pytest.mark.gpu
@pytest.mark.parametrize("precision", ["fp32x3", "tf32"])
@pytest.mark.parametrize("name,seed", CASES)
def test_cuda_update_matches_frozen_answers(cuda, name, seed, precision):
    import sac_util as SU

    z, cfg, st32, batches = regenerated_inputs(name, seed)
    agent = SU.make_agent(cfg, MG.CONFIGS[name]["per_task"], seed=seed, precision=precision)
    SU.load_oracle_state(agent, st32)
    for step, (batch, ec, ea) in enumerate(batches):
        _, logs = agent.update(tuple(b.cuda() for b in batch), eps_c=ec.cuda(), eps_a=ea.cuda(), check=True)
        got = np.array([float(logs[k]) for k in O.LOG_KEYS])
        ref = z["logs"][step]
        tol = 1e-3
        assert np.all(np.abs(got - ref) <= tol * np.abs(ref) + 1e-12), (precision, step, dict(zip(O.LOG_KEYS, (got - ref) / np.maximum(np.abs(ref), 1e-12))))
    leaves = _agent_leaves(agent)
    worst = {}
    for k, v in leaves.items():
        if name == "small":
            ref = torch.from_numpy(z[f"out/{k}"])
            worst[k] = SU.rel(v, ref)
        else:   # sum of squares + 64 strided samples of the leaf
            ref = z[f"outsum/{k}"]
            f = v.detach().double().flatten().cpu().numpy()
            samp = f[MG.sample_idx(f.size)]
            e_s = np.linalg.norm(samp - ref[2:]) / max(np.linalg.norm(ref[2:]), 1e-300)
            e_n = abs(np.sqrt((f * f).sum()) - np.sqrt(ref[1])) / max(np.sqrt(ref[1]), 1e-300)
            worst[k] = max(e_s, e_n)
    if precision == "fp32x3":
        # EVERY leaf of the parameters, the target and the Adam moments after two updates: the north-star 1e-3
        bad = {k: e for k, e in worst.items() if e > 1e-3}
        assert not bad, bad
    else:
        # tf32: kernels and the target to 1e-3; step-dominated leaves (zero-initialised biases, 1e-3-scale heads) and the
        # raw moments carry the ReLU-gate effect described in DESIGN.md "Precision" and are bounded loosely
        for k, e in worst.items():
            strict = k.split("/")[0] in ("actor", "critic", "target") and k.endswith("kernel") and "heads" not in k
            assert e <= (1e-3 if strict else 0.3), (k, e)
    print(precision, name, seed, {k: f"{e:.1e}" for k, e in sorted(worst.items(), key=lambda kv: -kv[1])[:6]})
