#!/bin/bash
# GPU-box call (gpurun --gpus 2): multi-rank preconditioned CG — parity tests, then 1- and 2-GPU bench lines of the 1M 2-D and 12M 3-D blocks
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_multi_gpu.py -m gpu -x -q ) > gpurun_out/pytest_mgpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_mgpu.log
tail -40 gpurun_out/pytest_mgpu.log | cut -c1-1500
for sc in dambreak2d_72k dambreak3d_123k; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29702 tests/multi_gpu_worker.py $sc 5 2>&1 | grep -E "MGPU|rror" | cut -c1-600
  MPS_MG_DIST_CELLS=0 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29702 tests/multi_gpu_worker.py $sc 5 2>&1 | grep -E "MGPU|rror" | cut -c1-600
done
run() { # gpus extra-args...
  local g=$1; shift
  if [ "$g" = 1 ]; then timeout 600 python bench.py --no-cpu-baseline "$@"
  else timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $g --master-addr 127.0.0.1 --master-port 29700 bench.py --gpus $g --no-cpu-baseline "$@"; fi
}
for w in dambreak2d_1m dambreak3d_10m; do
for g in 1 2; do
  ( run $g --workload $w --steps 5 --warmup 3 --no-e2e ) > gpurun_out/bench_${w}_n$g.log 2>&1; grep '^{' gpurun_out/bench_${w}_n$g.log | tail -1 | cut -c1-1200
  grep -E "rror|Traceback" gpurun_out/bench_${w}_n$g.log | head -5
done
done
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log | cut -c1-1500
