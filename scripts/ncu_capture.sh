#!/bin/bash
# ncu --set full captures of the dominant kernels (one GPU; run only after the plain bench exited 0).
# usage: scripts/ncu_capture.sh <round-tag>   (MTRL_GEMM_AUTOTUNE=0: no timing launches at plan creation, so --launch-skip counts the update's own launches; the launches captured here, >= 6 rounds of tiles, are never re-tuned anyway)
TAG=${1:-r01}
mkdir -p gpurun_out
B="env MTRL_GEMM_AUTOTUNE=0 python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline --capacity 4000"
NCU="ncu --set full --clock-control none --import-source on -f"
# GEMM launches of an update, in order: fwd L0 L1 L2 | fwd_target L0 L1 L2 | bwd_critic L2 L1 L0 | fwd_pi ... (skip counts launches)
$NCU -k regex:gemm_tf32_grouped --launch-skip 1 --launch-count 2 -o gpurun_out/${TAG}_gemm_fwd $B > gpurun_out/${TAG}_ncu_gemm.log 2>&1; echo gemm=$?
$NCU -k regex:gemm_tf32_grouped --launch-skip 6 --launch-count 1 -o gpurun_out/${TAG}_gemm_bwd $B > gpurun_out/${TAG}_ncu_gemm_bwd.log 2>&1; echo gemm_bwd=$?
$NCU -k regex:"adam_kernel|sumsq_kernel" --launch-skip 0 --launch-count 6 -o gpurun_out/${TAG}_adam $B > gpurun_out/${TAG}_ncu_adam.log 2>&1; echo adam=$?
$NCU -k regex:"gather_slabs|draw_indices" --launch-skip 0 --launch-count 2 -o gpurun_out/${TAG}_sampler $B > gpurun_out/${TAG}_ncu_sampler.log 2>&1; echo sampler=$?
$NCU -k regex:"head_bwd|critic_loss|actor_loss|actor_head" --launch-skip 0 --launch-count 7 -o gpurun_out/${TAG}_heads $B > gpurun_out/${TAG}_ncu_heads.log 2>&1; echo heads=$?
ls -la gpurun_out/${TAG}_*.ncu-rep
