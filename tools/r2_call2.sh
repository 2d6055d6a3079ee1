#!/bin/bash
# round-2 GPU call: full GPU test tier, bench line (with gpu_reference), launch lists, ncu --set full of the hot kernels
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -p no:cacheprovider -rs -s > gpurun_out/r2_gputests.log 2>&1; echo "gpu tests exit $?"
grep -E "passed|failed|error|FLUX dims|100-step|SKIP" gpurun_out/r2_gputests.log | tail -30
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench.log 2>&1; echo "bench exit $?"; tail -c 3000 gpurun_out/r2_bench.log
bash tools/r2_profile_trip.sh
ncu --set full --clock-control none --import-source on -k regex:gemm_bf16 -s 2 -c 1 -o gpurun_out/r2_gemm_fwd -f \
    python tools/profile_kernels.py gemm > gpurun_out/prof_gemm.log 2>&1; echo "gemm fwd full exit $?"
ncu --set full --clock-control none --import-source on -k regex:gemm_bf16 -s 5 -c 1 -o gpurun_out/r2_gemm_dgrad -f \
    python tools/profile_kernels.py gemm >> gpurun_out/prof_gemm.log 2>&1; echo "gemm dgrad full exit $?"
ncu --set full --clock-control none --import-source on -k regex:attn_fwd -s 1 -c 1 -o gpurun_out/r2_attn_fwd -f \
    python tools/profile_kernels.py attn > gpurun_out/prof_attn.log 2>&1; echo "attn fwd full exit $?"
ncu --set full --clock-control none --import-source on -k regex:attn_bwd -s 2 -c 2 -o gpurun_out/r2_attn_bwd -f \
    python tools/profile_kernels.py attn >> gpurun_out/prof_attn.log 2>&1; echo "attn bwd full exit $?"
ls -la gpurun_out | tail -20
