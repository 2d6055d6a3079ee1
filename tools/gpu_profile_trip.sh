#!/bin/bash
# Round-1 profile set: step launch list + ncu --set full of the dominant kernels (read with tools/ncu_summary.py).
mkdir -p gpurun_out
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r1_step_launches.csv \
    python tools/profile_step.py > gpurun_out/profile_step.log 2>&1; echo "launch list exit $?"
ncu --set full --clock-control none --import-source on -k regex:gemm_bf16 -s 2 -c 1 -o gpurun_out/r1_gemm_fwd -f \
    python tools/profile_kernels.py gemm > gpurun_out/prof_gemm.log 2>&1; echo "gemm fwd full exit $?"
ncu --set full --clock-control none --import-source on -k regex:gemm_bf16 -s 5 -c 1 -o gpurun_out/r1_gemm_dgrad -f \
    python tools/profile_kernels.py gemm >> gpurun_out/prof_gemm.log 2>&1; echo "gemm dgrad full exit $?"
ncu --set full --clock-control none --import-source on -k regex:attn_fwd -s 1 -c 1 -o gpurun_out/r1_attn_fwd -f \
    python tools/profile_kernels.py attn > gpurun_out/prof_attn.log 2>&1; echo "attn fwd full exit $?"
ncu --set full --clock-control none --import-source on -k regex:attn_bwd -s 2 -c 2 -o gpurun_out/r1_attn_bwd -f \
    python tools/profile_kernels.py attn >> gpurun_out/prof_attn.log 2>&1; echo "attn bwd full exit $?"
