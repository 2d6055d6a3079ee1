#!/bin/bash
# rank_simt v2 (3 chunks of X in flight) + 128x160/128x192 GEMM tiles: tests, timing, bench A/B.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_rank_simt.py tests/test_gpu_gemm.py -q -x > gpurun_out/r2_call20_tests.log 2>&1; echo "unit exit $?"; tail -5 gpurun_out/r2_call20_tests.log
run() {  # name, model args, env...
  local name=$1; shift; local margs=$1; shift
  env "$@" timeout 600 python bench.py $margs --steps 10 --warmup 3 --skip-gpu-reference --skip-cpu-baseline 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('$name', 'ms_per_step', round(d['ms_per_step'], 2), 'loss', d['loss_last'])
"
}
{
run "sdxl simt=0 wave=0" "--model sdxl" B200_RANK_SIMT=0 B200_GEMM_WAVE_TUNE=0
run "sdxl simt=0 wave=1" "--model sdxl" B200_RANK_SIMT=0 B200_GEMM_WAVE_TUNE=1
run "sdxl simt=1 wave=1" "--model sdxl" B200_RANK_SIMT=1 B200_GEMM_WAVE_TUNE=1
run "flux simt=0" "" B200_RANK_SIMT=0
run "flux simt=1" "" B200_RANK_SIMT=1
} | tee gpurun_out/r2_call20_ab.log
