#!/bin/bash
# SDXL: programmatic-dependent-launch build vs default, same box (5,500 small launches per step).
mkdir -p gpurun_out
for lib in libb200lora.so libb200lora_pdl.so libb200lora.so libb200lora_pdl.so; do
  B200_LIB=$PWD/ai_toolkit_b200/lib/$lib timeout 600 python bench.py --model sdxl --steps 10 --warmup 3 --skip-gpu-reference 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('$lib', 'ms_per_step', round(d['ms_per_step'], 2), 'e2e', round(d['e2e']['ms_per_step'], 2), 'loss', d['loss_last'])
"
done | tee gpurun_out/r2_sdxl_pdl_ab.log
B200_LIB=$PWD/ai_toolkit_b200/lib/libb200lora_pdl.so timeout 600 python -m pytest tests/test_unet_blocks.py -m gpu -q 2>&1 | tail -2 | tee -a gpurun_out/r2_sdxl_pdl_ab.log
