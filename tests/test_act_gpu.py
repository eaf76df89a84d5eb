"""MTSAC.sample_action / eval_action (mtsac.py:70-84, 299-311) on the CUDA actor vs the fp64 oracle."""
import numpy as np
import pytest
import torch

import sac_util as SU
from oracle import mtsac_oracle as O

pytestmark = pytest.mark.gpu
TOL_EXACT = 5e-3   # vs exact arithmetic: tf32 trunk operands (~1e-3 relative on the pre-tanh mean of magnitude ~1)
TOL_TF32 = 5e-4    # vs the oracle run with tf32-rounded matmul operands: isolates the kernels' arithmetic


def _setup(T=10, W=256, seed=3):
    cfg = O.OracleConfig(num_tasks=T, obs_dim=39 + T, action_dim=4, width=W)
    st = O.init_state(cfg, seed=seed, dtype=torch.float32)
    # the reference initialises heads at U(+-1e-3): scale them up so actions are not all ~0
    for k in ("kernel", "bias"):
        st.actor["heads"][k] = st.actor["heads"][k] * 100.0
    agent = SU.make_agent(cfg, 16, seed=seed)
    SU.load_oracle_state(agent, st)
    return cfg, st, agent


@pytest.mark.parametrize("n_per_task", [1, 3])
def test_sample_and_eval_action_match_oracle(cuda, n_per_task):
    cfg, st, agent = _setup()
    g = torch.Generator().manual_seed(11)
    T = cfg.num_tasks
    task = torch.arange(T).repeat(n_per_task)[torch.randperm(T * n_per_task, generator=g)]
    obs = torch.zeros(T * n_per_task, cfg.obs_dim)
    obs[:, :39] = torch.randn(T * n_per_task, 39, generator=g)
    obs[torch.arange(obs.shape[0]), 39 + task] = 1.0
    eps = torch.randn(obs.shape[0], 4, generator=g)
    import dataclasses

    p64 = O.tree_map(lambda x: x.double(), st.actor)
    ref_mode = O.actor_action(p64, obs.double(), cfg)
    ref_samp = O.actor_action(p64, obs.double(), cfg, eps.double())
    cfg_t = dataclasses.replace(cfg, matmul_operands="tf32")
    tf_mode = O.actor_action(p64, obs.double(), cfg_t)
    tf_samp = O.actor_action(p64, obs.double(), cfg_t, eps.double())
    got_mode = agent.eval_action(obs.numpy())
    _, got_samp = agent.sample_action(obs.numpy(), eps=eps)
    assert isinstance(got_mode, np.ndarray) and got_mode.shape == (obs.shape[0], 4)
    assert float(ref_mode.abs().max()) > 0.05, "degenerate test: actions are all ~0"
    assert np.abs(got_mode - ref_mode.numpy()).max() <= TOL_EXACT
    assert np.abs(got_samp - ref_samp.numpy()).max() <= TOL_EXACT
    assert np.abs(got_mode - tf_mode.numpy()).max() <= TOL_TF32
    assert np.abs(got_samp - tf_samp.numpy()).max() <= TOL_TF32
    # Philox draws: bounded, different between calls, identical for the rows of one call with the same inputs
    _, a1 = agent.sample_action(obs.numpy())
    _, a2 = agent.sample_action(obs.numpy())
    assert np.all(np.abs(a1) <= 1.0) and not np.allclose(a1, a2)
    # the update still works afterwards (the action path borrows its actor buffers)
    batch, ec, ea = O.synthetic_batch(cfg, 16, seed=5)
    _, logs = agent.update(tuple(b.cuda() for b in batch), eps_c=ec.cuda(), eps_a=ea.cuda(), check=True)
    assert torch.isfinite(torch.stack([v for v in logs.values()])).all()


def test_action_of_unowned_task_is_rejected(cuda):
    cfg, st, agent = _setup()
    obs = torch.zeros(2, cfg.obs_dim)
    obs[:, -1] = 1.0
    bad = torch.zeros(2, cfg.obs_dim + 1)
    with pytest.raises(AssertionError):
        agent.eval_action(bad.numpy())
    a = agent.eval_action(obs.numpy())   # last task is owned on one GPU: fine
    assert a.shape == (2, 4)
