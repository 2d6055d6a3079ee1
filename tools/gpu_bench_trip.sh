#!/bin/bash
# First bench trip: small-depth sanity run, then the full FLUX.1-dev step.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" | tee -a gpurun_out/smoke.log
timeout 600 python bench.py --layers 2 2 --steps 3 --warmup 3 --skip-cpu-baseline > gpurun_out/bench_small.log 2>&1; echo "small exit $?"
tail -3 gpurun_out/bench_small.log
timeout 1500 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_full.log 2>&1; echo "full exit $?"
tail -5 gpurun_out/bench_full.log
nvidia-smi --query-gpu=memory.used,memory.total --format=csv
