"""Per-role cycle accounting of the GEMM kernel (debug aid): where do the TMA / MMA / epilogue warps wait?"""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import torch

import gemm_cases as G
from mtrl_b200 import _lib as L

NAMES = ["prod wait-empty", "prod tma-issue", "mma wait-full", "mma wait-tmem-empty", "mma issue", "mma total",
         "epi wait-tmem-full", "epi work"]
for (M, N, K, am, bm, epi, ks, tag) in [
    (6400, 2048, 2048, 0, 0, L.EPI_STORE, 1, "dX-plain(KK)"),
    (6400, 2048, 2048, 0, 0, L.EPI_RELU_MASK, 1, "dX-mask(KK)"),
    (6400, 2048, 2048, 0, 1, L.EPI_BIAS_RELU, 1, "fwd(K,MN)"),
    (2048, 2048, 6400, 1, 1, L.EPI_STORE, 1, "dW(MN,MN)"),
    (6400, 2048, 96, 0, 1, L.EPI_BIAS_RELU, 1, "fwd-L0(K=96)"),
    (6400, 2048, 96, 0, 1, L.EPI_STORE, 1, "plain-L0(K=96)"),
]:
    p, D, ref, keep = G.make_problem(M, N, K, am, bm, epi, 256, ks)
    plan = L.GemmPlan([p])
    for _ in range(3):
        plan.run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        plan.run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    dbg = torch.zeros(8, dtype=torch.int64, device="cuda")
    plan.set_debug(dbg.data_ptr())
    plan.run()
    torch.cuda.synchronize()
    plan.set_debug(None)
    d = dbg.cpu().tolist()
    nworkers = 148 // plan.ctas
    nprod = 148
    kblocks = (M // (128 * plan.ctas)) * (N // 256) * (K // 32)  # total k-block iterations over all workers
    tiles = (M // (128 * plan.ctas)) * (N // 256)
    print(f"== {tag} ctas={plan.ctas} {ms*1e3:.1f} us {2*M*N*K/ms/1e9:.0f} TFLOP/s; k-block iters/worker={kblocks/nworkers:.0f}")
    per = [nprod, nprod, nworkers, nworkers, nworkers, nworkers, 148, 148]
    print(f"   tiles/worker={tiles/nworkers:.1f}  epilogue work per tile = {d[7]/148/(tiles/nworkers):.0f} cycles, out bytes/tile/CTA = {128*256*4}")
    for n, v, c in zip(NAMES, d, per):
        print(f"   {n:22s} {v/c:12.0f} cycles per CTA   ({v/c/(kblocks/nworkers):8.1f} per k-block)")
