#!/bin/bash
mkdir -p gpurun_out
LOG=gpurun_out/r2_diag.log
: > $LOG
run() { echo "== $*" >> $LOG; timeout 300 env "$@" >> $LOG 2>&1; echo "exit $?" >> $LOG; }
for d in 0 1 2; do
  run B200_ATTN_FWD=6 B200_ATTN_PP_DBG=$d python tools/time_attn_variants.py 4608
done
run B200_ATTN_FWD=5 python bench.py --steps 10 --warmup 3 --skip-cpu-baseline --skip-gpu-reference
run B200_ATTN_FWD=6 python bench.py --steps 10 --warmup 3 --skip-cpu-baseline --skip-gpu-reference
run B200_ATTN_FWD=5 python bench.py --steps 10 --warmup 3 --skip-cpu-baseline --skip-gpu-reference
run B200_ATTN_FWD=6 python bench.py --steps 10 --warmup 3 --skip-cpu-baseline --skip-gpu-reference
grep -E "^\[|^== |exit" $LOG | cut -c1-220; grep -o '"ms_per_step": [0-9.]*' $LOG
