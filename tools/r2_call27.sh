#!/bin/bash
# 2 GPUs on the final library: the NCCL equivalence test and the FLUX bench launched as the driver launches it.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_flux_engine.py -m gpu -q -x -s -k "two_ranks" > gpurun_out/r2_call27_nccl_test.log 2>&1; echo "nccl test exit $?"; tail -3 gpurun_out/r2_call27_nccl_test.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2_bench_flux_2gpu_final.json.log 2> gpurun_out/r2_bench_flux_2gpu_final.err; echo "2-GPU bench exit $?"
grep "^{" gpurun_out/r2_bench_flux_2gpu_final.json.log | tail -1 | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('n_gpus', d['n_gpus'], 'value', round(d['value'], 3), 'ms_per_step', round(d['ms_per_step'], 2), 'e2e', round(d['e2e']['ms_per_step'], 2))
"
