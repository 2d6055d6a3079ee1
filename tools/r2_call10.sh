#!/bin/bash
mkdir -p gpurun_out
LOG=gpurun_out/r2_bwd_diag.log
: > $LOG
run() { echo "== $*" >> $LOG; timeout 600 env "$@" >> $LOG 2>&1; echo "exit $?" >> $LOG; }
run B200_ATTN_BWD_DBG=0 python tools/time_attn_variants.py 4608
run B200_ATTN_BWD_DBG=2 python tools/time_attn_variants.py 4608
run python -m pytest tests/test_gpu_attention.py -q -p no:cacheprovider
# memcheck of the kernels written this round (small shapes; tcgen05 / TMA kernels run under the sanitizer too, slowly)
run compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_batch_ops.py -q -p no:cacheprovider -x -k "ddpm or train_loss or nchw or im2col or flux_packed"
run compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_wan.py -q -p no:cacheprovider -x -m gpu -k "rms_rope"
run compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_attention.py -q -p no:cacheprovider -x -k "cross_attention_fwd_bwd and 300 or growing"
grep -E "^\[|^== |exit|passed|failed|ERROR SUMMARY|Invalid" $LOG | cut -c1-220
