#!/bin/bash
# Round-2, first GPU trip: parity + same-box timing of the opt-in attention candidates (csrc/attention_r2.cu).
#   gpurun --timeout 1500 -- 'bash tools/r2_attn_trip.sh'
# Every process runs under its own `timeout` (a barrier-protocol mistake ends in the 2.5 s mbarrier watchdog trap, not
# in a hung box).  Output: gpurun_out/r2_attn_trip.log.  Promote a candidate to the default only if parity=OK on every
# shape, tests/test_gpu_attention.py passes with it, and it is faster than variant 1 in THIS log.
mkdir -p gpurun_out
LOG=gpurun_out/r2_attn_trip.log
: > $LOG
run() { echo "== $*" >> $LOG; timeout 300 env "$@" >> $LOG 2>&1; echo "exit $?" >> $LOG; }
# 1. baseline + each candidate alone: parity on ragged/short/long shapes, timing at the FLUX shape
for v in "B200_ATTN_FWD=1 B200_ATTN_BWD=1" "B200_ATTN_FWD=3 B200_ATTN_BWD=1" "B200_ATTN_FWD=4 B200_ATTN_BWD=1" \
         "B200_ATTN_FWD=1 B200_ATTN_BWD=3" "B200_ATTN_FWD=1 B200_ATTN_BWD=2"; do
  run $v python tools/time_attn_variants.py 88 320 1000 4608
done
# 1b. share of the exponentials computed by the FMA-pipe polynomial in the r2 forward (default 1 of 4 pairs): 0 and 2
for n in 0 2; do
  make variant NAME=poly$n DEFS="-DB200_ATTN_POLY_R2=$n" > /dev/null 2>&1
  run B200_LIB=ai_toolkit_b200/lib/libb200lora_poly$n.so B200_ATTN_FWD=3 B200_ATTN_BWD=1 python tools/time_attn_variants.py 4608
  run B200_LIB=ai_toolkit_b200/lib/libb200lora_poly$n.so B200_ATTN_FWD=4 B200_ATTN_BWD=1 python tools/time_attn_variants.py 4608
done
# 2. the unit tests (ragged tails, split = 0, large scores that force the lazy rescale) with each candidate
for v in "B200_ATTN_FWD=3 B200_ATTN_BWD=3" "B200_ATTN_FWD=4 B200_ATTN_BWD=2"; do
  run $v python -m pytest tests/test_gpu_attention.py -x -q -p no:cacheprovider
done
# 3. best-looking pairs through the whole step (parity of the engine + bench line)
for v in "B200_ATTN_FWD=3 B200_ATTN_BWD=3" "B200_ATTN_FWD=4 B200_ATTN_BWD=2"; do
  run $v python -m pytest tests/test_gpu_flux_engine.py -x -q -p no:cacheprovider -k "oracle or golden"
  run $v python bench.py --steps 10 --warmup 3
done
run B200_ATTN_FWD=1 B200_ATTN_BWD=1 python bench.py --steps 10 --warmup 3
grep -E "^\[|exit|passed|failed|\"value\"" $LOG | cut -c1-260
# 4. programmatic dependent launch build (same kernels + griddepcontrol, every launch with the PDL attribute)
PDL=ai_toolkit_b200/lib/libb200lora_pdl.so
run B200_LIB=$PDL python -m pytest tests/test_gpu_gemm.py tests/test_gpu_ops.py tests/test_gpu_attention.py -x -q -p no:cacheprovider
run B200_LIB=$PDL python -m pytest tests/test_gpu_flux_engine.py -x -q -p no:cacheprovider
run B200_LIB=$PDL python bench.py --steps 10 --warmup 3
grep -E "exit|passed|failed|\"value\"" $LOG | tail -8 | cut -c1-260
# 5. shared-address-space variant of the VALIDATED kernels (LDS/STS instead of generic LD/ST in the GEMM epilogue
#    transposes and the attention statistics): parity + sustained GEMM rate + bench
make variant NAME=lds DEFS="-DB200_SMEM_SHARED_ADDR=1" > /dev/null 2>&1
LDS=ai_toolkit_b200/lib/libb200lora_lds.so
run B200_LIB=$LDS python -m pytest tests/test_gpu_gemm.py tests/test_gpu_attention.py -x -q -p no:cacheprovider
run B200_LIB=$LDS python tools/sustained.py
run python tools/sustained.py
run B200_LIB=$LDS python bench.py --steps 10 --warmup 3
grep -E "exit|passed|failed|\"value\"|TF" $LOG | tail -12 | cut -c1-260
