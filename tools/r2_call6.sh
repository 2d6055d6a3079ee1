#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -p no:cacheprovider -rs > gpurun_out/r2_gputests3.log 2>&1; echo "gpu tests exit $?"
tail -6 gpurun_out/r2_gputests3.log; grep -E "^FAILED|^E  " gpurun_out/r2_gputests3.log | head -20
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_v6.log 2>&1; echo "bench exit $?"
grep -o '"ms_per_step": [0-9.]*' gpurun_out/r2_bench_v6.log | head -2
ncu --set full --clock-control none --import-source on -k regex:attn_fwd -s 1 -c 1 -o gpurun_out/r2_attn_fwd_pp2 -f \
    python tools/profile_kernels.py attn > gpurun_out/prof_attn2.log 2>&1; echo "attn fwd full exit $?"
