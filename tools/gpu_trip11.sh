#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -X faulthandler -m pytest tests/test_gpu_flux_engine.py -x -q -s --timeout 200 --timeout-method=thread 2>&1 | grep -E "rel err|loss curve|passed|failed|rror" | tail -6
timeout 900 python bench.py --steps 5 --warmup 3 --skip-cpu-baseline > gpurun_out/bench_full8.log 2>&1; echo "full exit $?"
grep -c watchdog gpurun_out/bench_full8.log
tail -1 gpurun_out/bench_full8.log | cut -c1-330
