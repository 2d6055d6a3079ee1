#!/bin/bash
# q/k/v head re-layout in one launch; channels_last body experiment (profiler totals); SDXL bench.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_unet_blocks.py -m gpu -q -x > gpurun_out/r2_call23_tests.log 2>&1; echo "tests exit $?"; tail -3 gpurun_out/r2_call23_tests.log
timeout 600 python tools/profile_sdxl.py 2>/dev/null | head -3 | tee gpurun_out/r2_call23_profile_heads.log
CHANNELS_LAST=1 timeout 600 python tools/profile_sdxl.py 2>/dev/null | head -30 | tee gpurun_out/r2_call23_profile_cl.log
timeout 600 python bench.py --model sdxl --steps 10 --warmup 3 --skip-gpu-reference 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('sdxl ms_per_step', round(d['ms_per_step'], 2), 'launches', d['launches_per_step'], 'loss', d['loss_last'])
" | tee gpurun_out/r2_call23_sdxl.log
