#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -X faulthandler -m pytest tests -m gpu -x -q --timeout 200 --timeout-method=thread -k "not fwd_bwd[1-4-4608" 2>&1 | tail -4
timeout 300 python -m pytest tests/test_gpu_attention.py -q -s -k timing 2>&1 | grep -E "attention fwd|passed|failed"
timeout 300 python tools/gpu_check_gemm.py 4 2>&1 | tail -4
timeout 900 python bench.py --steps 5 --warmup 3 --skip-cpu-baseline > gpurun_out/bench_full4.log 2>&1; echo "full exit $?"
tail -1 gpurun_out/bench_full4.log | cut -c1-330
