#!/bin/bash
mkdir -p gpurun_out
LOG=gpurun_out/r2_pp2.log
: > $LOG
run() { echo "== $*" >> $LOG; timeout 300 env "$@" >> $LOG 2>&1; echo "exit $?" >> $LOG; }
run B200_ATTN_FWD=5 python tools/time_attn_variants.py 88 320 1000 4608
for p in 1 2 0; do
  run B200_ATTN_FWD=6 B200_ATTN_PP_POLY=$p python tools/time_attn_variants.py 88 320 1000 4608
done
run B200_ATTN_FWD=6 python -m pytest tests/test_gpu_attention.py -x -q -p no:cacheprovider
run B200_ATTN_FWD=6 python -m pytest tests/test_gpu_flux_engine.py tests/test_wan.py -m gpu -x -q -p no:cacheprovider -k "oracle or golden or wan"
grep -E "^\[|^== |exit|passed|failed" $LOG | cut -c1-220
bash tools/rank_sweep.sh 2>&1 | tee gpurun_out/r2_rank_sweep.log
