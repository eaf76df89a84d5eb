"""CPU checks of the single-task SAC and MT-PPO oracles (ordering facts and closed forms the kernels rely on)."""
import math

import torch

from oracle import mtsac_oracle as O
from oracle import ppo_oracle as P
from oracle import sac_oracle as S


def test_sac_alpha_first_then_critic_uses_new_alpha():
    """sac.py:334-351: alpha is stepped first and the Bellman target uses the NEW alpha."""
    cfg = O.OracleConfig(num_tasks=1, obs_dim=10, action_dim=2, width=16, initial_temperature=0.5)
    st = S.init_state(cfg, dtype=torch.float64)
    b, ec, ea = S.synthetic_batch(cfg, 24, dtype=torch.float64)
    new, logs, grads = S.sac_update(st, b, ec, ea, cfg, return_grads=True)
    _, logp = S.actor_sample_and_log_prob(st.actor, b[0], ea, cfg)
    g_alpha = -(logp + cfg.target_entropy).mean()
    assert torch.allclose(grads["alpha"], g_alpha.reshape(1), rtol=1e-10)
    assert torch.allclose(logs["alpha"], torch.exp(new.log_alpha).sum())
    # param-norm logs are those of the PRE-update parameters (sac.py:360-364)
    assert torch.allclose(logs["metrics/critic_params_norm"], O.global_norm(st.critic))
    assert not torch.allclose(O.global_norm(new.critic), O.global_norm(st.critic), rtol=1e-12)
    # critic loss = 0.5 * sum_e mean_b (sac.py:292)
    with torch.no_grad():
        na, nlp = S.actor_sample_and_log_prob(st.actor, b[2], ec, cfg)
        qt = S.critic_forward(st.critic_target, b[2], na, cfg)
        y = b[4] + (1 - b[3]) * cfg.gamma * (qt.min(0).values - torch.exp(new.log_alpha) * nlp.reshape(-1, 1))
        q = S.critic_forward(st.critic, b[0], b[1], cfg)
    assert torch.allclose(logs["losses/qf_loss"], 0.5 * ((q - y) ** 2).mean(1).sum(), rtol=1e-10)
    assert st.opt["critic"]["count"] == 0 and new.opt["critic"]["count"] == 1


def test_ppo_surrogate_trains_only_log_std_and_closed_form_gradient():
    """mtppo.py:205-207 takes the log-prob of a fresh sample: no gradient reaches the mean head; d loss / d log_std has
    the closed form the CUDA kernel implements."""
    cfg = P.PPOConfig(net=O.OracleConfig(num_tasks=3, obs_dim=12, action_dim=2, width=16))
    st = P.init_state(cfg, dtype=torch.float64)
    r, eps = P.synthetic_rollout(cfg, 30, dtype=torch.float64)
    new, logs, g = P.ppo_update(st, r, eps, cfg, return_grads=True)
    A = cfg.net.action_dim
    assert float(g["policy"]["heads"]["kernel"][..., :A].abs().max()) == 0.0
    assert float(g["policy"]["heads"]["bias"][..., :A].abs().max()) == 0.0
    obs, old_logp, adv, ret, val = r
    out = O.multihead_forward(st.policy, obs, 3, 3)
    ls = out[:, A:]
    nl = (-0.5 * eps**2 - 0.5 * math.log(2 * math.pi) - ls).sum(-1, keepdim=True)
    ratio = torch.exp(nl - old_logp)
    ah = (adv - adv.mean()) / (adv.std(unbiased=False) + 1e-8)
    inside = (ratio > 1 - cfg.clip_eps) & (ratio < 1 + cfg.clip_eps)
    l1, l2 = -ah * ratio, -ah * torch.clamp(ratio, 1 - cfg.clip_eps, 1 + cfg.clip_eps)
    B = obs.shape[0]
    g_nl = torch.where((l1 >= l2) | inside, -ah * ratio, torch.zeros_like(ratio)) / B
    d_ls = (-g_nl - cfg.entropy_coefficient / B).expand(-1, A)
    task = obs[:, -3:].argmax(1)
    gb = torch.zeros(3, A, dtype=torch.float64).index_add_(0, task, d_ls)
    assert torch.allclose(g["policy"]["heads"]["bias"][:, A:], gb, rtol=1e-9, atol=1e-14)
    assert set(logs) == set(P.PPO_LOG_KEYS)


def test_ppo_value_loss_branches():
    cfg = P.PPOConfig(net=O.OracleConfig(num_tasks=2, obs_dim=10, action_dim=2, width=16), vf_coefficient=1.0)
    st = P.init_state(cfg, dtype=torch.float64)
    r, eps = P.synthetic_rollout(cfg, 16, dtype=torch.float64)
    _, logs_c, g_c = P.ppo_update(st, r, eps, cfg, return_grads=True)
    import dataclasses

    _, logs_u, g_u = P.ppo_update(st, r, eps, dataclasses.replace(cfg, clip_vf_loss=False), return_grads=True)
    assert float(logs_c["losses/value_function"]) >= float(logs_u["losses/value_function"]) - 1e-12  # max(.,.) >= unclipped
    assert torch.allclose(logs_c["losses/values"], logs_u["losses/values"])
