#!/usr/bin/env python
"""Device-timed update rates of the two other algorithm variants BASELINE.json's `configs` name (parity-test cases,
not bench.py lines): the parameter-matched single-task SAC baseline (configs[3]) and the MT50 width-4096 MT-PPO
policy/value update on a 50 x 10 000-row rollout (configs[4]).  CUDA events, synthetic inputs resident in HBM.
usage: python scripts/bench_variants.py [--ppo-steps 10000] [--out gpurun_out/variants.json]"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from mtrl_b200.config.networks import ContinuousActionPolicyConfig, QValueFunctionConfig, ValueFunctionConfig  # noqa: E402
from mtrl_b200.config.nn import MultiHeadConfig, VanillaNetworkConfig  # noqa: E402
from mtrl_b200.config.optim import OptimizerConfig  # noqa: E402
from mtrl_b200.presets import EnvSpec  # noqa: E402
from mtrl_b200.types import Rollout  # noqa: E402


def timed(fn, warmup=3, steps=10):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def sac_baseline(out):
    from mtrl_b200.rl.algorithms.sac import SAC, SACConfig

    # MT50 multi-head actor at W=2048 has 9 396 624 parameters; a 3-layer MLP on the same 89-dim input with
    # an 8-wide output matches that at W = 2144 (9 405 736).
    T, W, B, od, ad = 50, 2144, 6400, 89, 4
    opt = OptimizerConfig(max_grad_norm=1.0)
    net = VanillaNetworkConfig(width=W, depth=3, optimizer=opt)
    cfg = SACConfig(num_tasks=T, gamma=0.99, actor_config=ContinuousActionPolicyConfig(network_config=net),
                    critic_config=QValueFunctionConfig(network_config=net), num_critics=2)
    agent = SAC.initialize(cfg, EnvSpec(od, ad), seed=1, max_batch=B)
    g = torch.Generator(device="cuda").manual_seed(0)
    r = lambda *s: torch.randn(*s, generator=g, device="cuda")  # noqa: E731
    batch = (r(B, od), torch.rand(B, ad, generator=g, device="cuda") * 2 - 1, r(B, od), torch.zeros(B, 1, device="cuda"),
             torch.rand(B, 1, generator=g, device="cuda") * 10)
    ms = timed(lambda: agent.update(batch), warmup=5, steps=30)
    n = agent.get_num_params()
    out.append({"variant": "single-task SAC, parameter-matched MLP (configs[3])", "width": W, "batch": B,
                "actor_params": n["actor_num_params"], "ms_per_update": ms, "updates_per_s": 1e3 / ms})


def ppo(out, steps):
    from mtrl_b200.rl.algorithms import MTPPO, MTPPOConfig

    T, W, od, ad = 50, 4096, 89, 4
    opt = OptimizerConfig(max_grad_norm=1.0)
    net = MultiHeadConfig(width=W, depth=3, num_tasks=T, optimizer=opt)
    cfg = MTPPOConfig(num_tasks=T, policy_config=ContinuousActionPolicyConfig(network_config=net, squash_tanh=False),
                      vf_config=ValueFunctionConfig(network_config=net))
    agent = MTPPO.initialize(cfg, EnvSpec(od, ad), seed=1, rollout_steps=steps)
    B = T * steps
    g = torch.Generator(device="cuda").manual_seed(0)
    obs = torch.zeros(B, od, device="cuda")
    obs[:, :39] = torch.randn(B, 39, generator=g, device="cuda")
    obs[torch.arange(B, device="cuda"), 39 + torch.arange(B, device="cuda") // steps] = 1.0
    r = lambda: torch.randn(B, 1, generator=g, device="cuda")  # noqa: E731
    data = Rollout(obs, None, None, None, log_probs=r() - 5.0, advantages=r(), returns=r(), values=r())
    ms = timed(lambda: agent.update(data), warmup=2, steps=5)
    flops = 2 * 3 * 2.0 * B * (od * W + 2.0 * W * W)   # two networks x (fwd + dX + dW) trunk contractions
    out.append({"variant": "MT50 MT-PPO policy + value update (configs[4])", "width": W, "rows": B, "ms_per_update": ms,
                "updates_per_s": 1e3 / ms, "trunk_tflops": flops / ms / 1e9, "workspace_gb": agent._lay.workspace_bytes / 1e9})


def task_grads(out):
    from mtrl_b200.presets import metaworld_mtmhsac
    from mtrl_b200.rl.algorithms.mtsac import MTSAC

    T, W, B = 50, 2048, 6400
    mcfg, env = metaworld_mtmhsac(T, W)
    agent = MTSAC.initialize(mcfg, env, seed=1, max_batch=B)
    g = torch.Generator(device="cuda").manual_seed(0)
    obs = torch.zeros(B, 39 + T, device="cuda")
    obs[:, :39] = torch.randn(B, 39, generator=g, device="cuda")
    obs[torch.arange(B, device="cuda"), 39 + torch.arange(B, device="cuda") % T] = 1.0
    batch = (obs, torch.rand(B, 4, generator=g, device="cuda") * 2 - 1, obs + 0.01, torch.zeros(B, 1, device="cuda"),
             torch.rand(B, 1, generator=g, device="cuda") * 10)
    ms_g = timed(lambda: agent.per_task_gradients(batch), warmup=2, steps=5)
    ms_w = timed(lambda: agent.compute_weights(batch), warmup=1, steps=5)
    matrix_gb = (agent._lay.critic.total + agent._lay.actor.total) * T * 4 / 1e9
    # the same update with PCGradConfig on both networks
    import dataclasses

    from mtrl_b200.config.optim import PCGradConfig
    pc = PCGradConfig(max_grad_norm=1.0, num_tasks=T)
    net = dataclasses.replace(mcfg.actor_config.network_config, optimizer=pc)
    pcfg = dataclasses.replace(mcfg, actor_config=dataclasses.replace(mcfg.actor_config, network_config=net),
                               critic_config=dataclasses.replace(mcfg.critic_config, network_config=net))
    del agent
    torch.cuda.empty_cache()
    pagent = MTSAC.initialize(pcfg, env, seed=1, max_batch=B)
    ms_pc = timed(lambda: pagent.update(batch), warmup=3, steps=10)
    out.append({"variant": "MT50/W2048 MT-SAC update with PCGradConfig on actor and critic", "rows": B, "ms_per_update": ms_pc,
                "updates_per_s": 1e3 / ms_pc, "n_grad_conflicts_critic": float(pagent.pcgrad_stats()["critic"]["n_grad_conflicts"])})
    out.append({"variant": "MT50/W2048 per-task gradient matrices (T x P) + Gram metrics (compute_weights, SURVEY 8f row 1)",
                "rows": B, "ms_per_task_gradients": ms_g, "ms_compute_weights": ms_w,
                "matrix_gb": matrix_gb})


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--ppo-steps", type=int, default=10_000)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "variants.json"))
    a = ap.parse_args()
    res = []
    sac_baseline(res)
    torch.cuda.empty_cache()
    ppo(res, a.ppo_steps)
    torch.cuda.empty_cache()
    task_grads(res)
    for r in res:
        print(json.dumps(r))
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    with open(a.out, "w") as f:
        json.dump(res, f, indent=1)
