#!/bin/bash
# GPU-box call G (gpurun --gpus N): N-rank parity on a mid-size 2-D block, both couplings, adaptive split on / off.
NG=${1:-2}
W() { timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29702 tests/multi_gpu_worker.py "$@" 2>&1 | grep -E "MGPU \{\"rank\": 0|rror" | cut -c1-500; }
echo "== peer adaptive"; W dambreak2d_72k 3
echo "== peer frozen"; MPS_CG_ADAPTIVE=0 W dambreak2d_72k 3
echo "== nccl frozen"; MPS_CG_ADAPTIVE=0 MPS_COMM_NCCL_ONLY=1 W dambreak2d_72k 3
