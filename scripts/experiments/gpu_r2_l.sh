#!/bin/bash
# GPU-box call (gpurun --gpus 8): BASELINE configs[4] on 8 and 4 GPUs with the per-stage breakdown
mkdir -p gpurun_out
run() { local g=$1; shift; timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $g --master-addr 127.0.0.1 --master-port 29700 "$@"; }
for g in 8 4; do
run $g scripts/mgpu_large.py dambreak3d_100m 2 2 stages > gpurun_out/large_100m_n$g.log 2>&1; grep -E "LARGE|rror|Traceback" gpurun_out/large_100m_n$g.log | cut -c1-1100 | head -3
done
