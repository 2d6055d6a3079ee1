#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -X faulthandler -m pytest tests/test_gpu_flux_engine.py tests/test_gpu_gemm.py -x -q -s --timeout 300 --timeout-method=thread 2>&1 | grep -E "rel err|passed|failed|rror" | tail -5
timeout 900 python bench.py --steps 10 --warmup 3 --skip-cpu-baseline > gpurun_out/bench_full10.log 2>&1; echo "full exit $?"
grep -c watchdog gpurun_out/bench_full10.log
tail -1 gpurun_out/bench_full10.log | cut -c1-300
