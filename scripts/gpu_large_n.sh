#!/bin/bash
# GPU-box call (gpurun --gpus N): BASELINE configs[4] (101M particles, 3-D) on N GPUs with the per-stage breakdown
g=${1:-8}
mkdir -p gpurun_out
# on ONE GPU the 101M block only just fits (160.6 of 180 GB): no head-room on growing buffers, exact blob sizing
if [ "$g" = 1 ]; then export MPS_ALLOC_SLACK=0; fi
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $g --master-addr 127.0.0.1 --master-port 29700 scripts/mgpu_large.py dambreak3d_100m 2 2 stages > gpurun_out/large_100m_n$g.log 2>&1
grep -E "LARGE|rror|Traceback" gpurun_out/large_100m_n$g.log | cut -c1-1100 | head -3
