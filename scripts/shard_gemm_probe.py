"""Tile-shape probe for the grouped GEMM launches of ONE rank of a task-sharded update (debug / tuning aid).

For every launch class of a rank that holds `rows` batch rows (MT50 over 8 ranks: 896) it times the same group of
problems as (CTA pairs | single CTAs) x tile widths and prints the time, the TFLOP/s and the unit count, next to the
ideal time at the measured TF32 peak.  `python scripts/shard_gemm_probe.py [rows] [width]`."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch

from mtrl_b200 import _lib as L

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 896
W = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
PEAK = 617.0  # TFLOP/s, cuBLAS TF32 sustained (profiles/tf32_peak.json)
keep = []


def buf(*shape):
    t = torch.randn(*shape, device="cuda")
    t = ((t.view(torch.int32) + 0x1000) & ~0x1FFF).view(torch.float32)
    keep.append(t)
    return t


def fwd(M, N, K, bn):
    A, B, D, bias = buf(M, K), buf(K, N), buf(M, N), buf(N)
    bits = torch.zeros(M, (N + 31) // 32, dtype=torch.int32, device="cuda")
    keep.append(bits)
    return L.GemmProblem(A=A.data_ptr(), lda=K, a_major=0, B=B.data_ptr(), ldb=N, b_major=1, D=D.data_ptr(), ldd=N, M=M, N=N, K=K,
                         block_n=bn, k_splits=1, epilogue=L.EPI_BIAS_RELU, bias=bias.data_ptr(), relu_bits_out=bits.data_ptr(),
                         ldbits=(N + 31) // 32)


def dx(M, N, K, bn):
    A, B, D = buf(M, K), buf(N, K), buf(M, N)
    bits = torch.full((M, (N + 31) // 32), 0x55555555, dtype=torch.int32, device="cuda")
    cs = buf((M + 31) // 32, N)
    keep.append(bits)
    return L.GemmProblem(A=A.data_ptr(), lda=K, a_major=0, B=B.data_ptr(), ldb=K, b_major=0, D=D.data_ptr(), ldd=N, M=M, N=N, K=K,
                         block_n=bn, k_splits=1, epilogue=L.EPI_RELU_MASK, mask_bits=bits.data_ptr(), ldbits=(N + 31) // 32,
                         colsum_partial=cs.data_ptr())


def dw(n_in, N, rows_, bn, splits):
    X, dZ, D = buf(rows_, n_in), buf(rows_, N), buf(n_in, N)
    return L.GemmProblem(A=X.data_ptr(), lda=n_in, a_major=1, B=dZ.data_ptr(), ldb=N, b_major=1, D=D.data_ptr(), ldd=N, M=n_in, N=N,
                         K=rows_, block_n=bn, k_splits=splits, epilogue=L.EPI_ATOMIC_ADD if splits > 1 else L.EPI_STORE)


def timeit(plan, reps=20):
    for _ in range(3):
        plan.run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        plan.run()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


GROUPS = {
    "fwd x4 (AN, AO, C0, C1)": lambda bn: ([fwd(rows, W, W, bn) for _ in range(4)], 4 * 2 * rows * W * W),
    "fwd x2 (target / pi)": lambda bn: ([fwd(rows, W, W, bn) for _ in range(2)], 2 * 2 * rows * W * W),
    "bwd critic: 2 dW + 2 dX": lambda bn: ([dw(W, W, rows, bn, 1) for _ in range(2)] + [dx(rows, W, W, bn) for _ in range(2)],
                                           4 * 2 * rows * W * W),
    "bwd pi: 2 dX": lambda bn: ([dx(rows, W, W, bn) for _ in range(2)], 2 * 2 * rows * W * W),
    "bwd actor: dW + dX": lambda bn: ([dw(W, W, rows, bn, 1), dx(rows, W, W, bn)], 2 * 2 * rows * W * W),
    # first layer (K = 96 after padding): output-store bound, not tensor bound
    "fwd L0 x4 (K = 96)": lambda bn: ([fwd(rows, W, 96, bn) for _ in range(4)], 4 * 2 * rows * W * 96),
    "fwd L0 x2 (K = 96)": lambda bn: ([fwd(rows, W, 96, bn) for _ in range(2)], 2 * 2 * rows * W * 96),
}
only = os.environ.get("PROBE_ONLY")
for name, make in GROUPS.items():
    if only and only not in name:
        continue
    print(f"== {name}: rows {rows}, width {W}")
    widths = [int(x) for x in os.environ.get("PROBE_WIDTHS", "256,224,192,160,128,96,64").split(",")]
    for sk in (0, L.GEMM_STREAMK):
        for ctas in (2, 1):
            for bn in widths:
                if bn > W:
                    continue
                keep.clear()
                probs, flops = make(bn)
                try:
                    plan = L.GemmPlan(probs, ctas=ctas | sk)
                except Exception as e:  # noqa: BLE001
                    print(f"   ctas {ctas} block_n {bn}: {e}")
                    continue
                us = timeit(plan)
                workers = 148 // plan.ctas
                print(f"   {'stream-K' if sk else 'tiles   '} ctas {plan.ctas} block_n {bn:3d}: {us:7.1f} us  {flops / us / 1e6:6.0f} TFLOP/s  units "
                      f"{plan.units:4d} ({plan.units / workers:5.2f} per worker)  ideal {flops / PEAK / 1e6:5.1f} us")
                del plan
