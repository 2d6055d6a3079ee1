#!/bin/bash
mkdir -p gpurun_out
LOG=gpurun_out/r2_batched_issue.log
: > $LOG
run() { echo "== $*" >> $LOG; timeout 600 env "$@" >> $LOG 2>&1; echo "exit $?" >> $LOG; }
run B200_ATTN_BWD_DBG=3 python tools/time_attn_variants.py 88 320 1000 4608
run B200_ATTN_BWD_DBG=0 python tools/time_attn_variants.py 88 320 1000 4608
run B200_ATTN_BWD_DBG=3 python tools/time_attn_variants.py 4608
run B200_ATTN_BWD_DBG=0 python tools/time_attn_variants.py 4608
run python -m pytest tests/test_gpu_attention.py -q -p no:cacheprovider
run python -m pytest tests/test_gpu_flux_engine.py tests/test_wan.py -m gpu -x -q -p no:cacheprovider -k "oracle or golden or wan"
run python tools/stress_attn.py 512 20
run B200_ATTN_BWD_DBG=3 python bench.py --steps 10 --warmup 3 --skip-cpu-baseline --skip-gpu-reference
run python bench.py --steps 10 --warmup 3 --skip-cpu-baseline --skip-gpu-reference
grep -E "^\[|^== |exit|passed|failed" $LOG | cut -c1-220; grep -o '"ms_per_step": [0-9.]*' $LOG
