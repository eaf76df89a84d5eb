#!/bin/bash
# A/B of an environment switch on the same box: scripts/ab.sh VAR  (runs VAR=1 / VAR=0 alternately, twice each)
VAR=$1
for rep in 1 2; do for v in 1 0; do
  a=$(env $VAR=$v python bench.py --steps 60 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['ms_per_step'],4), round(d['e2e']['ms_per_step'],4))")
  b=$(env $VAR=$v MTRL_EMULATE_WORLD=8 python bench.py --steps 60 --warmup 5 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['ms_per_step'],4))")
  c=$(env $VAR=$v python bench.py --workload mt10_w400 --steps 60 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['ms_per_step'],4))")
  echo "$VAR=$v  mt50_w2048 (value e2e): $a   shard8: $b   mt10_w400: $c"
done; done
