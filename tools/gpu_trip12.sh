#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -X faulthandler -m pytest tests -m gpu -x -q --timeout 300 --timeout-method=thread 2>&1 | tail -3
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r1_final.log 2>&1; echo "bench exit $?"
tail -1 gpurun_out/bench_r1_final.log | cut -c1-300
timeout 600 python tools/ref_gpu_step.py ckpt 3 2>&1 | tail -2 | tee gpurun_out/ref_gpu_ckpt.log
timeout 600 python tools/ref_gpu_step.py nockpt 3 2>&1 | tail -2 | tee gpurun_out/ref_gpu_nockpt.log
