#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -X faulthandler -m pytest tests/test_gpu_attention.py tests/test_gpu_flux_engine.py -x -q -s --timeout 200 --timeout-method=thread 2>&1 | grep -E "attention fwd|passed|failed|rror" | tail -5
timeout 200 python tools/stress_attn.py 512 30 2>&1 | tail -1
timeout 900 python bench.py --steps 5 --warmup 3 --skip-cpu-baseline > gpurun_out/bench_full7.log 2>&1; echo "full exit $?"
grep -c watchdog gpurun_out/bench_full7.log
tail -1 gpurun_out/bench_full7.log | cut -c1-330
