#!/bin/bash
# UNet host parity + whole-step CUDA graph; 16-byte col_reduce (tests + A/B timing); SDXL bench line with the graph.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_unet_blocks.py tests/test_gpu_ops.py -m gpu -q -s -k "unet_host or col_reduce or ln_modulate" > gpurun_out/r2_unet_host_tests.log 2>&1
echo "tests exit $?"; grep -E "passed|failed|\[unet|^E  " gpurun_out/r2_unet_host_tests.log | head -20
timeout 300 python tools/time_rows.py > gpurun_out/r2_time_rows.log 2>&1
B200_COL_REDUCE_NARROW=1 timeout 300 python tools/time_rows.py >> gpurun_out/r2_time_rows.log 2>&1
cat gpurun_out/r2_time_rows.log
timeout 900 python bench.py --model sdxl --steps 10 --warmup 3 > gpurun_out/r2_bench_sdxl.json.log 2> gpurun_out/r2_bench_sdxl.err
echo "sdxl bench exit $?"; tail -c 2500 gpurun_out/r2_bench_sdxl.json.log; tail -5 gpurun_out/r2_bench_sdxl.err
