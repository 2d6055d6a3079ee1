#!/bin/bash
# after the final validation: the plugin's UNet route on the GPU (adopted eager UNet == host UNet), UNet tests again; small attention timing
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_unet_blocks.py tests/test_gpu_attention.py -m gpu -q -x -s -k "unet or t2d or transformer2d or adopted or small_attention or row_kernels" > gpurun_out/r2_call26_tests.log 2>&1; echo "tests exit $?"; grep -E "passed|failed|\[adopted|^E  " gpurun_out/r2_call26_tests.log | head
timeout 600 python bench.py --model sd15 --steps 10 --warmup 3 --skip-gpu-reference 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('sd15 ms_per_step', round(d['ms_per_step'], 2), 'launches', d['launches_per_step'], 'loss', d['loss_last'])
" | tee gpurun_out/r2_call26_sd15.log
