#!/bin/bash
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:attn_ -s 4 -c 4 -o gpurun_out/r1b_attn -f \
    python tools/profile_kernels.py attn > gpurun_out/prof_attn.log 2>&1; echo "attn full exit $?"
timeout 300 python -m pytest tests/test_gpu_gemm.py -x -q 2>&1 | tail -2
timeout 900 python bench.py --steps 5 --warmup 3 --skip-cpu-baseline > gpurun_out/bench_full3.log 2>&1; echo "full exit $?"
tail -1 gpurun_out/bench_full3.log | cut -c1-400
