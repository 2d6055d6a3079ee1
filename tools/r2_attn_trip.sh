#!/bin/bash
# Round-2 GPU trips for the opt-in candidates.  Sections (run any subset, default = all, ~25 min of box time):
#   gpurun --timeout 1800 -- 'bash tools/r2_attn_trip.sh attn poly unit'      (first call: ~8 min)
#   gpurun --timeout 1800 -- 'bash tools/r2_attn_trip.sh step pdl lds'        (second call)
#  attn  parity + same-box timing of every attention candidate alone (csrc/attention_r2.cu)
#  poly  share of the exponentials on the FMA-pipe polynomial in the r2 forward (0 / 2 of 4 pairs; default 1)
#  unit  tests/test_gpu_attention.py with the candidates (ragged tails, split = 0, big scores -> lazy rescale)
#  step  the candidate pairs through the whole step: engine parity tests + bench line, baseline bench beside it
#  pdl   programmatic-dependent-launch build: GPU test suite + bench
#  skinny  cluster split-K rank-side GEMM with the PUSH reduction (B200_SKINNY_PUSH=1): tests, hot-L2 timing vs the pull
#          reduction and the persistent kernel, bench with it forced for every rank-side GEMM (B200_SKINNY_ALWAYS=1)
#  tiles   threshold between 128x128 and 256x256 (2-CTA) tiles for the GEMMs of the short text stream (B200_GEMM_PAIR_MIN)
#  lds   shared-address-space variant of the VALIDATED kernels: tests + sustained GEMM rate + bench
# Every process runs under its own `timeout` (a barrier-protocol mistake ends in the 2.5 s mbarrier watchdog trap, not in
# a hung box).  Output: gpurun_out/r2_trip.log.  Promote a candidate to the default only if parity = OK on every shape,
# the unit tests pass with it, and it is faster than the baseline IN THE SAME LOG (box-to-box spread is +-2 %).
mkdir -p gpurun_out
LOG=gpurun_out/r2_trip.log
SECTIONS="${*:-attn poly unit step pdl lds skinny tiles}"
echo "### $(date -u +%H:%M:%S) sections: $SECTIONS" >> $LOG
run() { echo "== $*" >> $LOG; timeout 400 env "$@" >> $LOG 2>&1; echo "exit $?" >> $LOG; }
want() { case " $SECTIONS " in *" $1 "*) return 0;; *) return 1;; esac; }

if want attn; then
  for v in "B200_ATTN_FWD=1 B200_ATTN_BWD=1" "B200_ATTN_FWD=3 B200_ATTN_BWD=1" "B200_ATTN_FWD=4 B200_ATTN_BWD=1" \
           "B200_ATTN_FWD=1 B200_ATTN_BWD=3" "B200_ATTN_FWD=1 B200_ATTN_BWD=2" "B200_ATTN_FWD=1 B200_ATTN_BWD=5" \
           "B200_ATTN_FWD=1 B200_ATTN_BWD=4"; do
    run $v python tools/time_attn_variants.py 88 320 1000 4608
  done
fi
if want poly; then
  for n in 0 2; do
    make variant NAME=poly$n DEFS="-DB200_ATTN_POLY_R2=$n" > /dev/null 2>&1
    run B200_LIB=ai_toolkit_b200/lib/libb200lora_poly$n.so B200_ATTN_FWD=3 B200_ATTN_BWD=1 python tools/time_attn_variants.py 4608
    run B200_LIB=ai_toolkit_b200/lib/libb200lora_poly$n.so B200_ATTN_FWD=4 B200_ATTN_BWD=1 python tools/time_attn_variants.py 4608
  done
fi
if want unit; then
  for v in "B200_ATTN_FWD=3 B200_ATTN_BWD=3" "B200_ATTN_FWD=4 B200_ATTN_BWD=2" "B200_ATTN_FWD=3 B200_ATTN_BWD=5" \
           "B200_ATTN_FWD=4 B200_ATTN_BWD=4"; do
    run $v python -m pytest tests/test_gpu_attention.py -x -q -p no:cacheprovider
  done
fi
if want step; then
  for v in "B200_ATTN_FWD=3 B200_ATTN_BWD=3" "B200_ATTN_FWD=4 B200_ATTN_BWD=2"; do
    run $v python -m pytest tests/test_gpu_flux_engine.py -x -q -p no:cacheprovider -k "oracle or golden"
    run $v python bench.py --steps 10 --warmup 3
  done
  run B200_ATTN_FWD=1 B200_ATTN_BWD=1 python bench.py --steps 10 --warmup 3
fi
if want pdl; then
  PDL=ai_toolkit_b200/lib/libb200lora_pdl.so
  run B200_LIB=$PDL python -m pytest tests/test_gpu_gemm.py tests/test_gpu_ops.py tests/test_gpu_attention.py -x -q -p no:cacheprovider
  run B200_LIB=$PDL python -m pytest tests/test_gpu_flux_engine.py -x -q -p no:cacheprovider
  run B200_LIB=$PDL python bench.py --steps 10 --warmup 3
fi
if want lds; then
  make variant NAME=lds DEFS="-DB200_SMEM_SHARED_ADDR=1" > /dev/null 2>&1
  LDS=ai_toolkit_b200/lib/libb200lora_lds.so
  run B200_LIB=$LDS python -m pytest tests/test_gpu_gemm.py tests/test_gpu_attention.py -x -q -p no:cacheprovider
  run B200_LIB=$LDS python tools/sustained.py
  run python tools/sustained.py
  run B200_LIB=$LDS python bench.py --steps 10 --warmup 3
fi
if want tiles; then  # 512-token text stream: 128x128 tiles (L2-bound, measured 490 TF/s cold on qkv) vs 256x256 pair tiles
  for n in 74 64 40 20; do
    run B200_GEMM_PAIR_MIN=$n python bench.py --steps 10 --warmup 3
  done
fi
if want skinny; then
  run B200_SKINNY_PUSH=1 python -m pytest tests/test_gpu_gemm.py -x -q -p no:cacheprovider
  run B200_SKINNY_PUSH=1 B200_SKINNY_ALWAYS=1 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_flux_engine.py -x -q -p no:cacheprovider -k "not curve"
  run python tools/time_skinny.py
  run B200_SKINNY_PUSH=1 python tools/time_skinny.py
  run B200_SKINNY_PUSH=1 python bench.py --steps 10 --warmup 3
  run B200_SKINNY_PUSH=1 B200_SKINNY_ALWAYS=1 python bench.py --steps 10 --warmup 3
fi
grep -E "^\[|^== |exit|passed|failed|\"value\"|TF|persistent" $LOG | cut -c1-240 | tail -80
