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
