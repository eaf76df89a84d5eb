#!/bin/bash
# usage: scripts/mg_bench.sh N [exchange ...]   -- torchrun bench at N GPUs for each exchange mode
N=$1; shift
mkdir -p gpurun_out
for ex in "$@"; do
  MTRL_EXCHANGE=$ex timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
    --master-port 29544 bench.py --gpus $N --steps 30 --warmup 5 > gpurun_out/bench_g${N}_${ex}.json 2> gpurun_out/bench_g${N}_${ex}.err
  echo "exit=$? N=$N ex=$ex"; cat gpurun_out/bench_g${N}_${ex}.json | cut -c1-400
done
