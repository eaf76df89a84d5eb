"""Forward chain of `depth` Dense layers for `nchain` networks (the update's fwd x4 / fwd x2 passes): one launch per layer vs one
phased launch (grid barriers) vs one launch with per-row-tile dependencies (MTRL_GEMM_ROWDEPS).
usage: python scripts/chain_probe.py ROWS WIDTH [K0]"""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch

from mtrl_b200 import _lib as L

rows, W = int(sys.argv[1]), int(sys.argv[2])
K0 = int(sys.argv[3]) if len(sys.argv) > 3 else 96
D = 3
keep = []


def buf(*shape):
    t = torch.randn(*shape, device="cuda") * 0.05
    t = ((t.view(torch.int32) + 0x1000) & ~0x1FFF).view(torch.float32)
    keep.append(t)
    return t


def block_n_for(n):
    tiles = (n + 255) // 256
    bn = ((n + tiles - 1) // tiles + 31) // 32 * 32
    return min(bn, 256)


def chain(phased):
    x, ps = buf(rows, K0), []
    for l in range(D):
        K = K0 if l == 0 else W
        w, b, out = buf(K, W), buf(W), buf(rows, W)
        ps.append(L.GemmProblem(A=x.data_ptr(), lda=K, a_major=0, B=w.data_ptr(), ldb=W, b_major=1, D=out.data_ptr(), ldd=W, M=rows,
                                N=W, K=K, block_n=block_n_for(W), k_splits=1, epilogue=L.EPI_BIAS_RELU, bias=b.data_ptr(),
                                phase=l if phased else 0))
        x = out
    return ps


def timeit(run, reps=20):
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        run()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


for nchain in (4, 2):
    flops = nchain * 2 * rows * W * (K0 + (D - 1) * W)
    print(f"== {nchain} chains x {D} layers: rows {rows}, width {W}   (ideal at 617 TFLOP/s: {flops / 617e6:.1f} us)")
    for ctas in (2, 1):
        keep.clear()
        chains = [chain(False) for _ in range(nchain)]
        layer_plans = [L.GemmPlan([c[l] for c in chains], ctas=ctas) for l in range(D)]
        t0 = timeit(lambda: [p.run() for p in layer_plans])
        res = [f"per layer {t0:6.1f} us"]
        for name, flags in (("phased", 0), ("row deps", L.GEMM_ROWDEPS)):
            keep.clear()
            chains = [chain(True) for _ in range(nchain)]
            plan = L.GemmPlan([p for c in chains for p in c], ctas=ctas | flags)
            res.append(f"{name} {timeit(plan.run):6.1f} us")
        print(f"   ctas {ctas}: " + "   ".join(res))
