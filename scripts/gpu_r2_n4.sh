#!/bin/bash
# GPU-box call (gpurun --gpus N, N = 4 or 8): N-rank parity tests and the default bench line on N GPUs (weak 2-D N x 1M + large 3-D 12M strong)
g=${1:-4}
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_multi_gpu.py -m gpu -x -q -k "four_and_eight" ) > gpurun_out/pytest_mgpu_n$g.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_mgpu_n$g.log; tail -6 gpurun_out/pytest_mgpu_n$g.log | cut -c1-600
( timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $g --master-addr 127.0.0.1 --master-port 29701 bench.py --gpus $g --no-cpu-baseline ) > gpurun_out/bench_final_n$g.log 2>&1
grep '^{' gpurun_out/bench_final_n$g.log | tail -1 | cut -c1-250; grep -i "error\|trap" gpurun_out/bench_final_n$g.log | head -3
