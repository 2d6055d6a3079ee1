#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -X faulthandler -m pytest tests/test_gpu_gemm.py tests/test_gpu_flux_engine.py -x -q --timeout 300 --timeout-method=thread 2>&1 | tail -3
timeout 200 python tools/epi_probe.py 2>&1 | tail -7
timeout 900 python bench.py --steps 10 --warmup 3 --skip-cpu-baseline > gpurun_out/bench_full11.log 2>&1; echo "full exit $?"
grep -c watchdog gpurun_out/bench_full11.log
tail -1 gpurun_out/bench_full11.log | cut -c1-300
