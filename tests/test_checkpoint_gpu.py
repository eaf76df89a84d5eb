"""state_dict / load_state_dict: the exported tree carries the reference's names and shapes
(mtrl/rl/algorithms/mtsac.py:203-246 param trees, utils.py:11-46 TrainState, optax adam mu/nu), and restoring it into a
fresh agent continues the run bit for bit."""
import pytest
import torch

import sac_util as SU
from oracle import mtsac_oracle as O

pytestmark = pytest.mark.gpu


def test_state_dict_names_shapes_and_resume(cuda):
    T, W = 6, 128
    cfg = O.OracleConfig(num_tasks=T, obs_dim=39 + T, action_dim=4, width=W)
    a = SU.make_agent(cfg, 16, seed=1)
    batches = [O.synthetic_batch(cfg, 16, seed=70 + i) for i in range(4)]

    def step(agent, i):
        b, ec, ea = batches[i]
        agent.update(tuple(x.cuda() for x in b), eps_c=ec.cuda(), eps_a=ea.cuda(), check=True)

    step(a, 0)
    step(a, 1)
    sd = a.state_dict()
    net = sd["actor"]["params"]["params"]["MultiHeadNetwork_0"]
    assert net["layer_0"]["kernel"].shape == (39 + T, W) and net["layer_2"]["bias"].shape == (W,)
    assert net["VmapDense_0"]["kernel"].shape == (T, W, 8) and net["VmapDense_0"]["bias"].shape == (T, 8)
    cnet = sd["critic"]["params"]["params"]["VmapQValueFunction_0"]["MultiHeadNetwork_0"]
    assert cnet["layer_0"]["kernel"].shape == (2, 4 + 39 + T, W) and cnet["VmapDense_0"]["kernel"].shape == (2, T, W, 1)
    assert set(sd["critic"]) == {"step", "params", "opt_state", "target_params"}
    assert sd["actor"]["opt_state"]["count"] == 2 and sd["alpha"]["params"]["params"]["log_alpha"].shape == (T,)
    mu = sd["critic"]["opt_state"]["mu"]["params"]["VmapQValueFunction_0"]["MultiHeadNetwork_0"]["layer_1"]["kernel"]
    assert mu.shape == (2, W, W) and abs(mu).max() > 0

    b = SU.make_agent(cfg, 16, seed=99)       # different initial weights
    b.load_state_dict(sd)
    step(a, 2)
    step(b, 2)
    step(a, 3)
    _, logs_b = b.update(tuple(x.cuda() for x in batches[3][0]), eps_c=batches[3][1].cuda(), eps_a=batches[3][2].cuda(), check=True)
    for k in ("actor_params", "critic_params", "critic_target", "actor_m", "critic_v", "log_alpha"):
        # dW accumulates with atomics (split-K), so the two runs agree to rounding, not bit for bit
        assert torch.allclose(a._flat[k], b._flat[k], rtol=1e-4, atol=1e-6), k
    assert int(b._steps[0]) == 4
