#!/bin/bash
# final validation of the tree as committed: smoke, full GPU tier, bench line (the same three things the driver runs)
mkdir -p gpurun_out
timeout 600 python __graft_entry__.py smoke > gpurun_out/r2_final_smoke.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/r2_final_smoke.log
timeout 1800 python -m pytest tests -m gpu -q -p no:cacheprovider -rs -s > gpurun_out/r2_final_gputests.log 2>&1; echo "gpu tests exit $?"
grep -E "passed|failed|FLUX dims|100-step|\[wan\]|wan loss|SKIP" gpurun_out/r2_final_gputests.log | tail -12
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_final_bench.log 2>&1; echo "bench exit $?"
grep "^{" gpurun_out/r2_final_bench.log | tail -1 | cut -c1-400
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_final_ref.log 2>&1; echo "reference arm exit $?"
grep "^{" gpurun_out/r2_final_ref.log | tail -1 | cut -c1-300
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --cache-control none --csv \
    --log-file gpurun_out/r2_final_step_launches_warm.csv python tools/profile_step.py > gpurun_out/profile_step_final.log 2>&1
python tools/summarize_launches.py gpurun_out/r2_final_step_launches_warm.csv > gpurun_out/r2_final_step_launches_warm.md; head -12 gpurun_out/r2_final_step_launches_warm.md
