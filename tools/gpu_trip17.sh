#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -X faulthandler -m pytest tests -m gpu -x -q --timeout 300 --timeout-method=thread 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r1_final.log 2>&1; echo "bench exit $?"
grep -c watchdog gpurun_out/bench_r1_final.log
tail -1 gpurun_out/bench_r1_final.log | cut -c1-300
