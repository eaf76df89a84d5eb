#!/usr/bin/env python
"""Precision table SURVEY.md 7 ("Hard parts") asks for: tf32 / 3xTF32 (fp32x3) / bf16 operands against fp64, (1) for one
trunk-shaped contraction per K and (2) for a whole MT-SAC update per width (worst leaf of gradients and of updated
parameters, ten log scalars), with the fp32 oracle (= the reference's CPU arithmetic) beside them as the yardstick.

  gpurun -- 'python scripts/precision_probe.py > gpurun_out/precision_probe.json'

Not on the product path; imports oracle/ as the checker (like tests/)."""
from __future__ import annotations

import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import gemm_cases as G  # noqa: E402
import sac_util as SU  # noqa: E402
from mtrl_b200 import _lib as L  # noqa: E402
from oracle import mtsac_oracle as O  # noqa: E402


def rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-300))


def gemm_table():
    rows = []
    M = N = 512
    for K in (64, 256, 1024, 2048, 4096):
        g = torch.Generator().manual_seed(K)
        A = torch.randn(M, K, generator=g).cuda()
        B = torch.randn(N, K, generator=g).cuda()
        ref = A.double() @ B.double().t()
        out = {}
        # tf32 operands (rounded to nearest by the producer, as the update path does)
        Ah, Bh = G.tf32_round(A), G.tf32_round(B)
        Bt = Bh.t().contiguous()
        D = torch.zeros(M, N, device="cuda")
        p = L.GemmProblem(A=Ah.data_ptr(), lda=K, a_major=0, B=Bt.data_ptr(), ldb=N, b_major=1, D=D.data_ptr(), ldd=N, M=M, N=N, K=K,
                          block_n=256, k_splits=1, epilogue=L.EPI_STORE)
        L.GemmPlan([p]).run()
        torch.cuda.synchronize()
        out["tf32"] = rel(D, ref)
        # fp32x3
        (Ah, Al), (Bh, Bl) = G.split_tf32(A), G.split_tf32(B)
        Bth, Btl = Bh.t().contiguous(), Bl.t().contiguous()
        D3 = torch.zeros(M, N, device="cuda")
        p = L.GemmProblem(A=Ah.data_ptr(), lda=K, a_major=0, B=Bth.data_ptr(), ldb=N, b_major=1, D=D3.data_ptr(), ldd=N, M=M, N=N,
                          K=K, block_n=256, k_splits=1, epilogue=L.EPI_STORE, A_lo=Al.data_ptr(), B_lo=Btl.data_ptr())
        L.GemmPlan([p]).run()
        torch.cuda.synchronize()
        out["fp32x3"] = rel(D3, ref)
        # yardsticks (library arithmetic, not the product): fp32 SIMT-class product and bf16 operands
        torch.backends.cuda.matmul.allow_tf32 = False
        out["fp32_cublas"] = rel(A @ B.t(), ref)
        out["bf16_operands"] = rel((A.bfloat16().float().double() @ B.bfloat16().float().double().t()), ref)
        rows.append({"K": K, **out})
    return rows


def leaves_err(oracle_tree, agent_tree, ens):
    return SU.compare_trees(oracle_tree, agent_tree, ens)


def update_table():
    rows = []
    dev = torch.device("cuda")
    for T, W, per in ((10, 64, 8), (10, 400, 128), (10, 1024, 128), (10, 2048, 128), (50, 2048, 128)):
        cfg = O.OracleConfig(num_tasks=T, obs_dim=39 + T, action_dim=4, width=W)
        st = O.init_state(cfg, seed=1, dtype=torch.float32)
        batch, ec, ea = O.synthetic_batch(cfg, per, seed=100, dtype=torch.float32)
        mv = lambda t, dt: O.tree_map(lambda x: x.to(dev).to(dt), t)  # noqa: E731

        def to_dev(s, dt):
            return O.OracleState(mv(s.actor, dt), mv(s.critic, dt), mv(s.critic_target, dt), s.log_alpha.to(dev).to(dt),
                                 {k: {"m": mv(v["m"], dt) if isinstance(v["m"], dict) else v["m"].to(dev).to(dt),
                                      "v": mv(v["v"], dt) if isinstance(v["v"], dict) else v["v"].to(dev).to(dt),
                                      "count": v["count"]} for k, v in s.opt.items()})
        b64 = tuple(b.to(dev).double() for b in batch)
        new64, logs64, g64, _ = O.mtsac_update(to_dev(st, torch.float64), b64, ec.to(dev).double(), ea.to(dev).double(), cfg,
                                               return_grads=True)
        row = {"T": T, "W": W, "B": per * T}
        # fp32 oracle on the CPU: the reference's own arithmetic (skipped at the largest sizes: minutes of CPU time)
        if W <= 1024:
            new32, logs32, g32, _ = O.mtsac_update(st, batch, ec, ea, cfg, return_grads=True)

            def worst(ref_tree, got_tree):
                return max(rel(a.cpu(), o.cpu()) for o, a in zip(O.tree_leaves(ref_tree), O.tree_leaves(got_tree)))
            worst_p = max(worst(getattr(new64, n), getattr(new32, n)) for n in ("actor", "critic"))
            worst_g = max(worst(g64[n], g32[n]) for n in ("actor", "critic"))
            row["fp32_cpu_oracle"] = {"worst_param_leaf": worst_p, "worst_grad_leaf": worst_g,
                                      "worst_log": max(abs(float(logs32[k]) - float(logs64[k])) / max(abs(float(logs64[k])), 1e-12)
                                                       for k in O.LOG_KEYS if float(logs64[k]) != 0)}
        for precision in ("tf32", "fp32x3"):
            agent = SU.make_agent(cfg, per, seed=1, precision=precision)
            SU.load_oracle_state(agent, st)
            _, logs = agent.update(tuple(b.cuda() for b in batch), eps_c=ec.cuda(), eps_a=ea.cuda(), check=True)
            pe, ge = {}, {}
            for name, tree, ens in (("actor", agent.actor.params, False), ("critic", agent.critic.params, True)):
                for leaf, e in leaves_err(getattr(new64, name), tree, ens).items():
                    pe[f"{name}/{leaf}"] = e
            for name, tree, ens in (("actor", agent.actor.grads, False), ("critic", agent.critic.grads, True)):
                for leaf, e in leaves_err(g64[name], tree, ens).items():
                    ge[f"{name}/{leaf}"] = e
            le = {k: abs(float(logs[k]) - float(logs64[k])) / max(abs(float(logs64[k])), 1e-12) for k in O.LOG_KEYS if float(logs64[k]) != 0}
            wp, wg, wl = max(pe, key=pe.get), max(ge, key=ge.get), max(le, key=le.get)
            row[precision] = {"worst_param_leaf": pe[wp], "worst_param_leaf_name": wp, "worst_grad_leaf": ge[wg],
                              "worst_grad_leaf_name": wg, "worst_log": le[wl], "worst_log_name": wl,
                              "param_leaves": pe, "grad_leaves": ge}
            del agent
            torch.cuda.empty_cache()
        rows.append(row)
        print(json.dumps({k: (v if not isinstance(v, dict) else {kk: vv for kk, vv in v.items() if not kk.endswith("leaves")})
                          for k, v in row.items()}), file=sys.stderr, flush=True)
    return rows


if __name__ == "__main__":
    out = {"x3_chunk_kb": int(os.environ.get("MTRL_X3_CHUNK", "4")), "gemm": gemm_table()}
    print(json.dumps(out["gemm"]), file=sys.stderr, flush=True)
    if "--gemm-only" not in sys.argv:
        out["update"] = update_table()
    print(json.dumps(out))
