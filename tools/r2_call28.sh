#!/bin/bash
# 8 GPUs on the final library, launched as the driver launches it (FLUX headline config).
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r2_bench_flux_8gpu_final.json.log 2> gpurun_out/r2_bench_flux_8gpu_final.err; echo "8-GPU bench exit $?"
grep "^{" gpurun_out/r2_bench_flux_8gpu_final.json.log | tail -1 | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('n_gpus', d['n_gpus'], 'value', round(d['value'], 3), 'ms_per_step', round(d['ms_per_step'], 2), 'e2e', round(d['e2e']['ms_per_step'], 2))
"
timeout 600 python bench.py --steps 10 --warmup 3 --skip-gpu-reference --skip-cpu-baseline 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('same box n_gpus 1 ms_per_step', round(d['ms_per_step'], 2))
"
