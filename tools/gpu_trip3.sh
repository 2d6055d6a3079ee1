#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -X faulthandler -m pytest tests -m gpu -x -q --timeout 200 --timeout-method=thread -k "not fwd_bwd[1-4-4608" 2>&1 | tail -15
timeout 300 python -m pytest tests/test_gpu_attention.py -q -s -k timing 2>&1 | grep -E "attention fwd|passed|failed"
timeout 900 python bench.py --steps 5 --warmup 3 --skip-cpu-baseline > gpurun_out/bench_full2.log 2>&1; echo "full exit $?"
tail -2 gpurun_out/bench_full2.log | cut -c1-900
