#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -X faulthandler -m pytest tests/test_gpu_attention.py tests/test_gpu_flux_engine.py -x -q --timeout 200 --timeout-method=thread -k "not fwd_bwd[1-4-4608" 2>&1 | tail -3
timeout 300 python -m pytest tests/test_gpu_attention.py -q -s -k timing 2>&1 | grep -E "attention fwd|passed|failed"
timeout 900 python bench.py --steps 5 --warmup 3 --skip-cpu-baseline > gpurun_out/bench_full5.log 2>&1; echo "full exit $?"
grep -c watchdog gpurun_out/bench_full5.log
tail -1 gpurun_out/bench_full5.log | cut -c1-330
