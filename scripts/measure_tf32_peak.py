#!/usr/bin/env python
"""cuBLAS TF32 throughput on this pool's B200s, measured with the recipe MEASURED_PEAKS.json documents for bf16
(torch.matmul 8192^3, 2 N^3 FLOPs: best of 10 = burst, back to back for 4 s = sustained) -- the roofline denominator of
the tf32 grouped GEMM (bench.py reads profiles/tf32_peak.json when it exists).

  gpurun -- 'python scripts/measure_tf32_peak.py > gpurun_out/tf32_peak.json'   then copy to profiles/tf32_peak.json"""
import json
import time

import torch

torch.backends.cuda.matmul.allow_tf32 = True
N = 8192
a = torch.randn(N, N, device="cuda")
b = torch.randn(N, N, device="cuda")
c = torch.empty(N, N, device="cuda")
flops = 2.0 * N ** 3
for _ in range(3):
    torch.matmul(a, b, out=c)
torch.cuda.synchronize()
best = 0.0
for _ in range(10):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    torch.matmul(a, b, out=c)
    e1.record()
    e1.synchronize()
    best = max(best, flops / (e0.elapsed_time(e1) * 1e-3) / 1e12)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 0
t0 = time.perf_counter()
e0.record()
while time.perf_counter() - t0 < 4.0:
    for _ in range(20):
        torch.matmul(a, b, out=c)
    n += 20
    torch.cuda.synchronize()
e1.record()
e1.synchronize()
sustained = n * flops / (e0.elapsed_time(e1) * 1e-3) / 1e12
print(json.dumps({"tf32_tflops": best, "tf32_tflops_sustained": sustained, "gpu_name": torch.cuda.get_device_name(0),
                  "torch": torch.__version__, "how": "torch.matmul fp32 with allow_tf32 (cuBLAS TF32), 8192^3 (2*N^3): best of 10 (burst) "
                  "and back to back for 4 s (sustained), CUDA events", "when": time.strftime("%Y-%m-%dT%H:%M:%SZ", time.gmtime())}))
