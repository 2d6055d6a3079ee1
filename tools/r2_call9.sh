#!/bin/bash
mkdir -p gpurun_out
LOG=gpurun_out/r2_pp3.log
: > $LOG
run() { echo "== $*" >> $LOG; timeout 300 env "$@" >> $LOG 2>&1; echo "exit $?" >> $LOG; }
run B200_ATTN_FWD=6 python tools/time_attn_variants.py 88 320 1000 4608
for p in 0 1 2; do
  run B200_ATTN_FWD=7 B200_ATTN_PP_POLY=$p python tools/time_attn_variants.py 88 320 1000 4608
done
run B200_ATTN_FWD=7 python -m pytest tests/test_gpu_attention.py -x -q -p no:cacheprovider
run B200_ATTN_FWD=7 python -m pytest tests/test_gpu_flux_engine.py tests/test_wan.py -m gpu -x -q -p no:cacheprovider -k "oracle or golden or wan"
run python tools/stress_attn.py
grep -E "^\[|^== |exit|passed|failed" $LOG | cut -c1-220
