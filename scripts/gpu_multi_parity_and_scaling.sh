#!/bin/bash
# GPU-box call B (gpurun --gpus N): multi-GPU parity tests, then the strong-scaling bench on 1..N GPUs for both CG couplings.
NG=${1:-2}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
( time timeout 900 python -m pytest tests/test_multi_gpu.py -m gpu -x -q ) > gpurun_out/pytest_mgpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_mgpu.log
tail -15 gpurun_out/pytest_mgpu.log
for sc in dambreak2d static dambreak3d; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29702 tests/multi_gpu_worker.py $sc 5 2>&1 | grep MGPU
  MPS_COMM_NCCL_ONLY=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29703 tests/multi_gpu_worker.py $sc 5 2>&1 | grep MGPU
done
run() { # gpus extra-args...
  local g=$1; shift
  if [ "$g" = 1 ]; then timeout 600 python bench.py --no-cpu-baseline "$@"
  else timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $g --master-addr 127.0.0.1 --master-port 29700 bench.py --gpus $g --no-cpu-baseline "$@"; fi
}
for g in 1 2 4 8; do
  [ $g -le $NG ] || continue
  ( run $g --steps 5 --warmup 3 ) > gpurun_out/bench_n$g.log 2>&1; grep '^{' gpurun_out/bench_n$g.log | tail -1 | cut -c1-400
done
( MPS_COMM_NCCL_ONLY=1 run $NG --steps 2 --warmup 3 --no-e2e ) > gpurun_out/bench_n${NG}_nccl.log 2>&1; grep '^{' gpurun_out/bench_n${NG}_nccl.log | tail -1 | cut -c1-300
tail -5 gpurun_out/bench_n${NG}.log | cut -c1-600
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29701 scripts/mgpu_probe.py dambreak2d_1m 2 > gpurun_out/probe_n$NG.log 2>&1; grep PROBE gpurun_out/probe_n$NG.log
