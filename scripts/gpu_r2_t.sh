#!/bin/bash
# GPU-box call (gpurun --gpus 2): multi-GPU parity tests and the default bench line on 2 GPUs (weak 2-D 2M + large 3-D 12M strong)
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_multi_gpu.py -m gpu -x -q ) > gpurun_out/pytest_mgpu_t.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_mgpu_t.log; tail -6 gpurun_out/pytest_mgpu_t.log | cut -c1-600
( timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29701 bench.py --gpus 2 --no-cpu-baseline ) > gpurun_out/bench_t_n2.log 2>&1
grep '^{' gpurun_out/bench_t_n2.log | tail -1 | cut -c1-250; grep -i "error\|trap" gpurun_out/bench_t_n2.log | head -3
