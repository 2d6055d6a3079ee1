#!/bin/bash
# CUDA-core rank-side kernel: unit tests, timing vs the tensor-core skinny GEMM, engine parity tests, FLUX + SDXL bench A/B.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_rank_simt.py -q -x > gpurun_out/r2_rank_simt_tests.log 2>&1; echo "unit exit $?"; tail -5 gpurun_out/r2_rank_simt_tests.log
timeout 900 python -m pytest tests/test_gpu_flux_engine.py tests/test_unet_blocks.py tests/test_wan.py -m gpu -q -x > gpurun_out/r2_rank_simt_engine_tests.log 2>&1; echo "engine exit $?"; tail -4 gpurun_out/r2_rank_simt_engine_tests.log
for simt in 1 0; do
  B200_RANK_SIMT=$simt timeout 600 python bench.py --steps 10 --warmup 3 --skip-gpu-reference --skip-cpu-baseline 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('flux simt=$simt ms_per_step', round(d['ms_per_step'], 2), 'loss', d['loss_last'])
"
  B200_RANK_SIMT=$simt timeout 600 python bench.py --model sdxl --steps 10 --warmup 3 --skip-gpu-reference 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('sdxl simt=$simt ms_per_step', round(d['ms_per_step'], 2), 'loss', d['loss_last'])
"
done | tee gpurun_out/r2_rank_simt_ab.log
