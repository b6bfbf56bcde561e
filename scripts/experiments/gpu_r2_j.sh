#!/bin/bash
# GPU-box call (gpurun --gpus 8): 8- and 4-rank parity, BASELINE configs[4] (101M particles, 3-D) on 8 / 4 / 2 GPUs, default bench line on 8 GPUs
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo8.txt 2>&1; free -g | head -2
run() { local g=$1; shift; timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $g --master-addr 127.0.0.1 --master-port 29700 "$@"; }
run 8 tests/multi_gpu_worker.py dambreak3d_123k 3 2>&1 | grep -E "MGPU|rror" | grep -E '"rank": 0|rror' | cut -c1-500
MPS_MG_DIST_CELLS=0 run 4 tests/multi_gpu_worker.py dambreak2d_72k 5 2>&1 | grep -E "MGPU|rror" | grep -E '"rank": 0|rror' | cut -c1-500
for g in 8 4 2; do
  run $g scripts/mgpu_large.py dambreak3d_100m 2 2 > gpurun_out/large_100m_n$g.log 2>&1; grep -E "LARGE|rror|Traceback" gpurun_out/large_100m_n$g.log | cut -c1-700 | head -8
done
( run 8 bench.py --gpus 8 ) > gpurun_out/bench_n8.log 2>&1; python scripts/show_line.py gpurun_out/bench_n8.log; grep -E "rror|Traceback" gpurun_out/bench_n8.log | head -5
