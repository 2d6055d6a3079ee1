#!/bin/bash
# smem-staged ln_modulate kernels: tests + timing; UNet host tests; SDXL per-kernel profile (torch.profiler, eager launches).
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_unet_blocks.py tests/test_gpu_ops.py -m gpu -q -s > gpurun_out/r2_rows_unet_tests.log 2>&1
echo "tests exit $?"; grep -E "passed|failed|\[unet|^E  " gpurun_out/r2_rows_unet_tests.log | head -20
timeout 300 python tools/time_rows.py > gpurun_out/r2_time_rows2.log 2>&1; cat gpurun_out/r2_time_rows2.log
timeout 600 python tools/profile_sdxl.py > gpurun_out/r2_sdxl_profile.md 2> gpurun_out/r2_sdxl_profile.err; head -50 gpurun_out/r2_sdxl_profile.md; tail -3 gpurun_out/r2_sdxl_profile.err
