#!/bin/bash
# Per-kernel timeline of ONE rank's share of an N-rank job, on a single GPU (no exchange): bench + ncu launch list.
N=${1:-8}
mkdir -p gpurun_out
MTRL_EMULATE_WORLD=$N python bench.py --steps 30 --warmup 5 > gpurun_out/shard${N}_bench.json 2> gpurun_out/shard${N}_bench.err
echo "exit=$?"; cut -c1-300 gpurun_out/shard${N}_bench.json
MTRL_EMULATE_WORLD=$N ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/shard${N}_launches.csv \
  python bench.py --steps 2 --warmup 3 --no-graph --capacity 2000 > gpurun_out/shard${N}_ncu.log 2>&1
python scripts/launch_summary.py gpurun_out/shard${N}_launches.csv > gpurun_out/shard${N}_launches_summary.txt; cat gpurun_out/shard${N}_launches_summary.txt
