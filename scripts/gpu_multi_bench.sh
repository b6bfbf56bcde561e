#!/bin/bash
# GPU-box call J (gpurun --gpus N): strong-scaling bench lines for the listed rank counts, optionally another workload.
# usage: gpu_round_j.sh "4 8" [workload]
mkdir -p gpurun_out
WL=${2:-dambreak2d_1m}
for g in $1; do
  ( timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $g --master-addr 127.0.0.1 --master-port 29700 bench.py --gpus $g --workload $WL --no-cpu-baseline --steps 5 --warmup 3 ) > gpurun_out/bench_${WL}_n$g.log 2>&1
  grep '^{' gpurun_out/bench_${WL}_n$g.log | tail -1 | cut -c1-200; grep -i "error\|trap" gpurun_out/bench_${WL}_n$g.log | head -3
done
