#!/bin/bash
# Same-box A/B of an environment switch on the headline bench, with the per-class kernel times:
#   scripts/ab_env.sh VAR [workload] [extra bench args]   (runs VAR=1 / VAR=0 alternately, twice each)
VAR=$1; WL=${2:-mt50_w2048}; shift; shift
for rep in 1 2; do for v in 1 0; do
  env $VAR=$v python bench.py --workload $WL --steps 60 --warmup 5 --no-cpu-baseline "$@" 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$VAR=$v', '$WL', 'ms/step', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['ms_per_step'],4), 'gemm avg launch', round(d['roofline']['avg_launch_ms'],5), 'x', d['roofline']['launches_per_step'], [(k['kernel'][:12], round(k['ms_per_step'],4)) for k in d['hbm_kernels']['kernels']])
"
done; done
