#!/bin/bash
# GPU-box call (gpurun --gpus 8): the default bench line on 8 GPUs (weak 2-D 8 x 1M + large 3-D 12M strong)
mkdir -p gpurun_out
( timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29701 bench.py --gpus 8 --no-cpu-baseline ) > gpurun_out/bench_final_n8.log 2>&1
grep '^{' gpurun_out/bench_final_n8.log | tail -1 | cut -c1-250; grep -i "error\|trap" gpurun_out/bench_final_n8.log | head -3
