#!/bin/bash
# 2 GPUs: the NCCL equivalence test, and the bench at N = 2 for FLUX and Wan (launched as the driver launches it)
mkdir -p gpurun_out
nvidia-smi -L | head -4
timeout 900 python -m pytest tests/test_gpu_flux_engine.py -m gpu -q -p no:cacheprovider -s -k "nccl" > gpurun_out/r2_nccl_test.log 2>&1; echo "nccl test exit $?"
tail -5 gpurun_out/r2_nccl_test.log
for m in flux wan; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus 2 --model $m --steps 10 --warmup 3 \
      > gpurun_out/r2_bench_${m}_2gpu.log 2>&1; echo "bench $m 2 gpus exit $?"
  grep -o '"value": [0-9.]*, "unit": "steps/s", "n_gpus": [0-9]*' gpurun_out/r2_bench_${m}_2gpu.log | head -1
  grep -o '"ms_per_step": [0-9.]*' gpurun_out/r2_bench_${m}_2gpu.log | head -1
done
timeout 600 python bench.py --model flux --steps 10 --warmup 3 --skip-cpu-baseline --skip-gpu-reference > gpurun_out/r2_bench_flux_1of2.log 2>&1
grep -o '"ms_per_step": [0-9.]*' gpurun_out/r2_bench_flux_1of2.log | head -1
