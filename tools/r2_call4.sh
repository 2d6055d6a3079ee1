#!/bin/bash
# Wan2.1 path on the GPU (tests + bench line for BASELINE.json configs[3]) and the full GPU tier with the promoted kernels
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_wan.py -m gpu -q -p no:cacheprovider -s -x > gpurun_out/r2_wan_tests.log 2>&1; echo "wan tests exit $?"
grep -E "passed|failed|\[wan\]|wan loss|Error|error" gpurun_out/r2_wan_tests.log | tail -20
timeout 1800 python -m pytest tests -m gpu -q -p no:cacheprovider -rs > gpurun_out/r2_gputests2.log 2>&1; echo "gpu tests exit $?"
tail -5 gpurun_out/r2_gputests2.log
timeout 900 python bench.py --model wan --steps 5 --warmup 3 > gpurun_out/r2_bench_wan.log 2>&1; echo "wan bench exit $?"; tail -c 2500 gpurun_out/r2_bench_wan.log
timeout 600 python bench.py --steps 10 --warmup 3 --skip-cpu-baseline --skip-gpu-reference > gpurun_out/r2_bench_pp.log 2>&1; echo "bench exit $?"
grep -o '"ms_per_step": [0-9.]*' gpurun_out/r2_bench_pp.log | head -1
