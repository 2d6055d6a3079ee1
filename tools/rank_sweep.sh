#!/bin/bash
# BASELINE.json configs[4]: FLUX.1-dev LoRA rank sweep r in {4, 8, 16, 32, 64} at bs 4, 1024^2 -- one bench line per rank and an
# ncu --set full capture of the fused LoRA-Linear at M = 16384 (N = 12288, K = 3072) per rank (+ the other two shapes at r 16)
mkdir -p gpurun_out
for r in ${RANKS:-4 8 16 32 64}; do
  timeout 600 python bench.py --batch 4 --rank $r --steps ${STEPS:-5} --warmup 3 --skip-cpu-baseline --skip-gpu-reference \
      > gpurun_out/sweep_bs4_r$r.log 2>&1; echo "bench r=$r exit $?"; tail -c 1800 gpurun_out/sweep_bs4_r$r.log | grep -o '"ms_per_step": [0-9.]*' | head -1
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16 -s 4 -c 1 -o gpurun_out/sweep_r${r}_mlp_up -f \
      python tools/profile_lora_linear.py $r mlp_up > gpurun_out/sweep_prof_r$r.log 2>&1; echo "ncu r=$r exit $?"
  timeout 120 python tools/profile_lora_linear.py $r mlp_up | tail -1
done
for s in attn_out single_out; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16 -s 4 -c 1 -o gpurun_out/sweep_r16_$s -f \
      python tools/profile_lora_linear.py 16 $s > gpurun_out/sweep_prof_r16_$s.log 2>&1; echo "ncu r=16 $s exit $?"
  timeout 120 python tools/profile_lora_linear.py 16 $s | tail -1
done
