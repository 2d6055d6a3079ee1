#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_unet_blocks.py -m gpu -q -p no:cacheprovider -s > gpurun_out/r2_unet_tests.log 2>&1; echo "unet tests exit $?"
grep -E "passed|failed|\[t2d|^E  |Error" gpurun_out/r2_unet_tests.log | head -30
