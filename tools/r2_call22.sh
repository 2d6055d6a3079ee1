#!/bin/bash
# CUDA-core attention for head dim 160 (SD1.5), SD1.5-form UNet parity, SD1.5 bench line; refreshed SDXL per-kernel profile.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_attention.py tests/test_unet_blocks.py -m gpu -q -x -s -k "small_attention or unet or t2d or transformer2d" > gpurun_out/r2_call22_tests.log 2>&1; echo "tests exit $?"; grep -E "passed|failed|\[unet|\[t2d|^E  " gpurun_out/r2_call22_tests.log | head -20
timeout 600 python bench.py --model sd15 --steps 10 --warmup 3 > gpurun_out/r2_bench_sd15.json.log 2> gpurun_out/r2_bench_sd15.err; echo "sd15 bench exit $?"; tail -c 1500 gpurun_out/r2_bench_sd15.json.log; tail -3 gpurun_out/r2_bench_sd15.err
timeout 600 python tools/profile_sdxl.py > gpurun_out/r2_sdxl_profile2.md 2> gpurun_out/r2_sdxl_profile2.err; head -36 gpurun_out/r2_sdxl_profile2.md
