#!/bin/bash
# 8 GPUs: the bench exactly as the driver launches it, FLUX (headline) and Wan (configs[3] names 8xB200), + rank 16 bs 4
mkdir -p gpurun_out
nvidia-smi -L | wc -l
for spec in "flux 1 16 10" "wan 1 16 5" "flux 4 16 5"; do
  set -- $spec
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29721 bench.py --gpus 8 --model $1 --batch $2 --rank $3 --steps $4 --warmup 3 \
      > gpurun_out/r2_bench_$1_bs$2_8gpu.log 2>&1; echo "bench $1 bs$2 8 gpus exit $?"
  grep "^{" gpurun_out/r2_bench_$1_bs$2_8gpu.log | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['ms_per_step'])"
done
timeout 600 python bench.py --steps 10 --warmup 3 --skip-cpu-baseline --skip-gpu-reference > gpurun_out/r2_bench_flux_1of8.log 2>&1
grep -o '"ms_per_step": [0-9.]*' gpurun_out/r2_bench_flux_1of8.log | head -1
timeout 600 python bench.py --model wan --steps 5 --warmup 3 --skip-cpu-baseline --skip-gpu-reference > gpurun_out/r2_bench_wan_1of8.log 2>&1
grep -o '"ms_per_step": [0-9.]*' gpurun_out/r2_bench_wan_1of8.log | head -1
