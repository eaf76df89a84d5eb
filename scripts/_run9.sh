cd /root/repo
N=8
MG_T=50 MG_W=2048 bash scripts/mg_check.sh $N p2p 1; cp gpurun_out/mgcheck_${N}_p2p_g1.log gpurun_out/r02d_mgcheck_${N}_p2p_T50_W2048.log
for dp in 1 0; do
MTRL_DEFER_POLYAK=$dp bash scripts/mg_bench.sh $N p2p | head -1; cp gpurun_out/bench_g${N}_p2p.json gpurun_out/r02d_bench_g${N}_dp$dp.json
python - <<PY
import json
d=json.loads(open('gpurun_out/r02d_bench_g${N}_dp$dp.json').read().strip().splitlines()[-1])
print($N, 'defer_polyak=$dp', round(d['value'],1), round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value'],1), 'exch', round(d['exchange']['ms_per_step'],4), d['exchange']['phase_us']['critic'], d['parity_check']['ok'], 'gemm', round(d['roofline']['avg_launch_ms'],5), d['config']['parallelism'][-45:])
PY
done
