B="python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline --capacity 4000"
for v in 1 0; do
MTRL_FUSED_HEADS=$v ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_heads$v.csv $B > /dev/null 2>&1
python scripts/gemm_launch_times.py gpurun_out/r02_launches_heads$v.csv
done
