#!/bin/bash
# One GPU trip: validate every GEMM configuration in its own process (a trap in one must not hide the others).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
for cfg in 1 3 4 2; do
  timeout 300 python tools/gpu_check_gemm.py $cfg > gpurun_out/gemm_cfg$cfg.log 2>&1
  echo "cfg $cfg exit $?" >> gpurun_out/gemm_summary.txt
  tail -3 gpurun_out/gemm_cfg$cfg.log
done
cat gpurun_out/gemm_summary.txt
