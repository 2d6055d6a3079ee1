#!/bin/bash
# Round-2 profile set.  Same as tools/gpu_profile_trip.sh plus a WARM-cache launch list: ncu flushes the caches before
# every kernel by default, which makes each memory-bound kernel look as if its input came from HBM although the
# producer has just written it into the 126 MB L2.  Shares in the two lists bracket the real in-graph share.
mkdir -p gpurun_out
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --cache-control none --csv \
    --log-file gpurun_out/r2_step_launches_warm.csv python tools/profile_step.py > gpurun_out/profile_step_warm.log 2>&1
echo "warm launch list exit $?"
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/r2_step_launches.csv python tools/profile_step.py > gpurun_out/profile_step.log 2>&1
echo "cold launch list exit $?"
python tools/summarize_launches.py gpurun_out/r2_step_launches_warm.csv > gpurun_out/r2_step_launches_warm.md
python tools/summarize_launches.py gpurun_out/r2_step_launches.csv > gpurun_out/r2_step_launches.md
head -20 gpurun_out/r2_step_launches_warm.md
# decision input for folding dQ into the dK/dV kernel: L2 reduction throughput (see the header of the .cu file)
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/bulk_reduce tools/microbench/bulk_reduce.cu \
  && timeout 120 gpurun_out/bulk_reduce > gpurun_out/r2_bulk_reduce.log 2>&1; echo "bulk_reduce exit $?"; cat gpurun_out/r2_bulk_reduce.log
# rank-side (N = 64) GEMMs with a hot L2, persistent vs cluster split-K: the cold-cache launch list overstates them
timeout 300 python tools/time_skinny.py > gpurun_out/r2_time_skinny.log 2>&1; echo "time_skinny exit $?"; cat gpurun_out/r2_time_skinny.log
