#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -X faulthandler -m pytest tests -m gpu -x -q --timeout 300 --timeout-method=thread -k "not fwd_bwd[1-4-4608" 2>&1 | tail -2
timeout 120 python tools/time_attn.py 2>&1 | tail -3
timeout 300 python -m pytest tests/test_gpu_attention.py -q -s -k timing 2>&1 | grep -E "attention fwd"
timeout 200 python tools/stress_attn.py 512 30 2>&1 | tail -1
timeout 900 python bench.py --steps 10 --warmup 3 --skip-cpu-baseline > gpurun_out/bench_full9.log 2>&1; echo "full exit $?"
grep -c watchdog gpurun_out/bench_full9.log
tail -1 gpurun_out/bench_full9.log | cut -c1-300
