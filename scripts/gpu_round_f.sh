#!/bin/bash
# GPU-box call F (gpurun --gpus 8): 8-rank parity against 1 GPU, strong-scaling bench at N = 4, 8.
mkdir -p gpurun_out
for sc in "dambreak2d_72k 3" "dambreak3d_123k 2"; do
  set -- $sc
  timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29702 tests/multi_gpu_worker.py $1 $2 2>&1 | grep -E "MGPU \{\"rank\": 0|rror" | cut -c1-600
done
for g in 4 8; do
  ( timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $g --master-addr 127.0.0.1 --master-port 29700 bench.py --gpus $g --no-cpu-baseline --steps 5 --warmup 3 ) > gpurun_out/bench_n$g.log 2>&1
  grep '^{' gpurun_out/bench_n$g.log | tail -1 | cut -c1-330; grep -i "error\|trap" gpurun_out/bench_n$g.log | head -3
done
