#!/bin/bash
# ncu --set full on the kernels added / changed late in round 2 (one GPU; never a bench number).
mkdir -p gpurun_out
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"gemm_bf16_kernel|ln_modulate_bwd|col_reduce|attn_fwd_pp2|attn_bwd_r2|small_attn" \
  -o gpurun_out/r2_new_kernels -f python tools/profile_kernels.py r2new > gpurun_out/r2_new_kernels_ncu.log 2>&1
echo "ncu exit $?"; tail -5 gpurun_out/r2_new_kernels_ncu.log; ls -la gpurun_out/r2_new_kernels.ncu-rep
