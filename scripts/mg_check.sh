#!/bin/bash
# usage: scripts/mg_check.sh N exchange [graph 0|1]  -- tests/multigpu_check.py under torchrun with a hard timeout
N=$1; EX=$2; G=${3:-1}
mkdir -p gpurun_out
MG_HANG_DUMP_S=60 MG_EXCHANGE=$EX MTRL_UPDATE_GRAPH=$G timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N \
  --master-addr 127.0.0.1 --master-port 29577 tests/multigpu_check.py > gpurun_out/mgcheck_${N}_${EX}_g${G}.log 2>&1
echo "exit=$? N=$N ex=$EX graph=$G"; grep -E "multigpu_check ok|Error|error|File \"/|assert" gpurun_out/mgcheck_${N}_${EX}_g${G}.log | head -20
