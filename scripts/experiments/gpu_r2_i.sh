#!/bin/bash
# GPU-box call (gpurun --gpus 2): halo-only gathers — 2-GPU parity tests, forced-distributed variant, bench lines at N = 2
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests/test_multi_gpu.py -m gpu -x -q ) > gpurun_out/pytest_mgpu.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_mgpu.log; tail -8 gpurun_out/pytest_mgpu.log | cut -c1-1500
for sc in dambreak2d_72k dambreak3d_123k; do
  MPS_MG_DIST_CELLS=0 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29702 tests/multi_gpu_worker.py $sc 5 2>&1 | grep -E "MGPU|rror" | cut -c1-400
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29701 scripts/mgpu_large.py dambreak3d_10m 3 3 stages 2>&1 | grep LARGE | cut -c1-900
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29700 bench.py --gpus 2 --no-large ) > gpurun_out/bench_n2.log 2>&1
python scripts/show_line.py gpurun_out/bench_n2.log
