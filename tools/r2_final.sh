#!/bin/bash
# final validation of the tree as committed: smoke, full GPU tier, bench line (the same three things the driver runs)
mkdir -p gpurun_out
timeout 600 python __graft_entry__.py smoke > gpurun_out/r2_final_smoke.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/r2_final_smoke.log
timeout 1800 python -m pytest tests -m gpu -q -p no:cacheprovider -rs -s > gpurun_out/r2_final_gputests.log 2>&1; echo "gpu tests exit $?"
grep -E "passed|failed|FLUX dims|100-step|\[wan\]|wan loss|SKIP" gpurun_out/r2_final_gputests.log | tail -12
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_final_bench.log 2>&1; echo "bench exit $?"
grep "^{" gpurun_out/r2_final_bench.log | tail -1 | cut -c1-400
for m in sdxl sd15 wan; do
  timeout 900 python bench.py --model $m --steps 10 --warmup 3 > gpurun_out/r2_final_bench_$m.log 2>&1; echo "bench $m exit $?"
  grep "^{" gpurun_out/r2_final_bench_$m.log | tail -1 | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['metric'], '| ms/step', round(d['ms_per_step'], 2), '| e2e', round(d['e2e']['ms_per_step'], 2), '| frac', round(d['step_roofline']['frac'], 3), '| ref', {k: (round(v.get('ms_per_step', 0), 1) if isinstance(v, dict) else v) for k, v in d.get('gpu_reference', {}).items() if k in ('ms_per_step', 'checkpointing', 'no_checkpointing')})
"
done
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_final_ref.log 2>&1; echo "reference arm exit $?"
grep "^{" gpurun_out/r2_final_ref.log | tail -1 | cut -c1-300
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --cache-control none --csv \
    --log-file gpurun_out/r2_final_step_launches_warm.csv python tools/profile_step.py > gpurun_out/profile_step_final.log 2>&1
python tools/summarize_launches.py gpurun_out/r2_final_step_launches_warm.csv > gpurun_out/r2_final_step_launches_warm.md; head -12 gpurun_out/r2_final_step_launches_warm.md
