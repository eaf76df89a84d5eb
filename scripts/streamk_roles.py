"""Per-role cycle counters of one launch class as whole tiles vs stream-K (debug aid)."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch

from mtrl_b200 import _lib as L

rows, W, nprob = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
NAMES = ["prod wait-empty", "prod tma-issue", "mma wait-full", "mma wait-tmem-empty", "mma issue", "mma total",
         "epi wait-tmem-full", "epi work"]
keep = []


def buf(*shape):
    t = torch.randn(*shape, device="cuda")
    keep.append(t)
    return t


def fwd(M, N, K):
    A, B, D, bias = buf(M, K), buf(K, N), buf(M, N), buf(N)
    return L.GemmProblem(A=A.data_ptr(), lda=K, a_major=0, B=B.data_ptr(), ldb=N, b_major=1, D=D.data_ptr(), ldd=N, M=M, N=N, K=K,
                         block_n=256, k_splits=1, epilogue=L.EPI_BIAS_RELU, bias=bias.data_ptr())


probs = [fwd(rows, W, W) for _ in range(nprob)]
for ctas in (2, 1):
    for sk in (0, L.GEMM_STREAMK):
        plan = L.GemmPlan(probs, ctas=ctas | sk)
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
        nw = 148 // plan.ctas
        per = [148, 148, nw, nw, nw, nw, 148, 148]
        print(f"== {'stream-K' if sk else 'tiles'} ctas={plan.ctas}: {ms * 1e3:.1f} us, units {plan.units}")
        for n, v, c in zip(NAMES, d, per):
            print(f"   {n:22s} {v / c:12.0f} cycles per CTA")
