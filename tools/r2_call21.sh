#!/bin/bash
# live head dim 64 in the attention kernels + grouped q/k/v in the UNet engine: tests, timing, SDXL bench; FLUX sanity.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_attention.py tests/test_unet_blocks.py -m gpu -q -x -s > gpurun_out/r2_call21_tests.log 2>&1; echo "tests exit $?"; grep -E "passed|failed|\[unet|\[t2d|^E  " gpurun_out/r2_call21_tests.log | head -20
timeout 300 python tools/time_attn64.py > gpurun_out/r2_time_attn64.md 2>&1; cat gpurun_out/r2_time_attn64.md
for i in 1 2; do
timeout 600 python bench.py --model sdxl --steps 10 --warmup 3 --skip-gpu-reference 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('sdxl ms_per_step', round(d['ms_per_step'], 2), 'launches', d['launches_per_step'], 'loss', d['loss_last'])
"
done | tee gpurun_out/r2_call21_sdxl.log
timeout 600 python bench.py --steps 10 --warmup 3 --skip-gpu-reference --skip-cpu-baseline 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('flux ms_per_step', round(d['ms_per_step'], 2), 'loss', d['loss_last'])
" | tee -a gpurun_out/r2_call21_sdxl.log
