#!/bin/bash
# UNet host: parity test of the tiny SDXL-form UNet step, then the SDXL bench line (configs[1], hybrid arm).
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_unet_blocks.py -m gpu -x -q -s > gpurun_out/r2_unet_host_tests.log 2>&1
echo "unet host tests exit $?"; tail -5 gpurun_out/r2_unet_host_tests.log
timeout 900 python bench.py --model sdxl --steps 5 --warmup 3 > gpurun_out/r2_bench_sdxl.json.log 2> gpurun_out/r2_bench_sdxl.err
echo "sdxl bench exit $?"; tail -c 3000 gpurun_out/r2_bench_sdxl.json.log; tail -5 gpurun_out/r2_bench_sdxl.err
