"""Runs every GEMM case and prints the error of each (debug aid; tests/test_gemm_gpu.py asserts)."""
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import torch

import gemm_cases as G
from mtrl_b200 import _lib as L

for case in G.CASES:
    try:
        r, a = G.run_case(case)
        print(f"{case[0]:24s} rel={r:.3e} maxabs={a:.3e}", flush=True)
    except Exception as e:  # noqa: BLE001
        print(f"{case[0]:24s} EXC {e}", flush=True)
for name, r, a in G.run_grouped():
    print(f"grouped/{name:16s} rel={r:.3e} maxabs={a:.3e}", flush=True)

# timing: the MT50 / width-2048 trunk shapes
for (M, N, K, am, bm, epi, ks, tag) in [
    (6400, 2048, 2048, 0, 1, L.EPI_BIAS_RELU, 1, "fwd"),
    (6400, 2048, 2048, 0, 0, L.EPI_RELU_MASK, 1, "dX"),
    (2048, 2048, 6400, 1, 1, L.EPI_STORE, 1, "dW"),
    (2048, 2048, 6400, 1, 1, L.EPI_ATOMIC_ADD, 3, "dW_split3"),
]:
    p, D, ref, keep = G.make_problem(M, N, K, am, bm, epi, 256, ks)
    plan = L.GemmPlan([p])
    for _ in range(3):
        plan.run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        plan.run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print(f"time {tag:10s} {M}x{N}x{K}: {ms*1e3:.1f} us  {2*M*N*K/ms/1e9:.1f} TFLOP/s", flush=True)
torch.backends.cuda.matmul.allow_tf32 = True
a = torch.randn(6400, 2048, device="cuda"); b = torch.randn(2048, 2048, device="cuda")
for _ in range(3): a @ b
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): a @ b
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20
print(f"cublas tf32 6400x2048x2048: {ms*1e3:.1f} us {2*6400*2048*2048/ms/1e9:.1f} TFLOP/s")
