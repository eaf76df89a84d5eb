#!/bin/bash
# bench every workload on one GPU; prints value / e2e / gemm frac
mkdir -p gpurun_out
for w in "$@"; do
  python bench.py --workload $w --steps 50 --warmup 5 > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/bench_$w.json").read().strip().splitlines()[-1])
print("$w", round(d["value"],1), round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1), "gemm TF/s", round(d["roofline"]["achieved"],1), "frac", round(d["roofline"]["frac"],3), "cpu", round(d.get("cpu_baseline",{}).get("value",0),3), d.get("cpu_baseline",{}).get("cores"))
PY
done
