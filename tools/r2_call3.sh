#!/bin/bash
# attention forward variant 5 (ping-pong): parity + same-box timing against variant 3, unit tests, and the fixed tests
mkdir -p gpurun_out
LOG=gpurun_out/r2_pp.log
: > $LOG
run() { echo "== $*" >> $LOG; timeout 300 env "$@" >> $LOG 2>&1; echo "exit $?" >> $LOG; }
run B200_ATTN_FWD=3 python tools/time_attn_variants.py 88 320 1000 4608
for p in 1 0 2 3; do
  run B200_ATTN_FWD=5 B200_ATTN_PP_POLY=$p python tools/time_attn_variants.py 88 320 1000 4608
done
run B200_ATTN_FWD=5 python -m pytest tests/test_gpu_attention.py -x -q -p no:cacheprovider
run B200_ATTN_FWD=5 python -m pytest tests/test_gpu_flux_engine.py -x -q -p no:cacheprovider -k "oracle or golden"
run python -m pytest tests/test_gpu_batch_ops.py -q -p no:cacheprovider
run B200_ATTN_FWD=5 python bench.py --steps 10 --warmup 3 --skip-cpu-baseline --skip-gpu-reference
grep -E "^\[|^== |exit|passed|failed|\"value\"" $LOG | cut -c1-250
# programmatic dependent launch build (griddepcontrol in every kernel, PDL attribute on every launch): never measured in round 1
PDL=ai_toolkit_b200/lib/libb200lora_pdl.so
run B200_LIB=$PDL python -m pytest tests/test_gpu_gemm.py tests/test_gpu_ops.py tests/test_gpu_attention.py -x -q -p no:cacheprovider
run B200_LIB=$PDL python -m pytest tests/test_gpu_flux_engine.py -x -q -p no:cacheprovider -k "oracle or golden or 20"
run B200_LIB=$PDL python bench.py --steps 10 --warmup 3 --skip-cpu-baseline --skip-gpu-reference
run python bench.py --steps 10 --warmup 3 --skip-cpu-baseline --skip-gpu-reference
for n in 74 40 20; do
  run B200_GEMM_PAIR_MIN=$n python bench.py --steps 10 --warmup 3 --skip-cpu-baseline --skip-gpu-reference
done
grep -E "^\[|^== |exit|passed|failed" $LOG | cut -c1-250 | tail -40
grep -o '"ms_per_step": [0-9.]*' $LOG
