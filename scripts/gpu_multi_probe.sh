#!/bin/bash
# GPU-box call H (gpurun --gpus N): N-rank parity (peer-memory all-gathers + persistent CG), strong-scaling bench at N.
NG=${1:-4}
mkdir -p gpurun_out
W() { timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29702 tests/multi_gpu_worker.py "$@" 2>&1 | grep -E "MGPU \{\"rank\": 0|rror|Trace" | cut -c1-500; }
W dambreak2d 5
W dambreak3d_123k 2
( timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29700 bench.py --gpus $NG --no-cpu-baseline --steps 5 --warmup 3 ) > gpurun_out/bench_n$NG.log 2>&1
grep '^{' gpurun_out/bench_n$NG.log | tail -1 | cut -c1-330; grep -i "error\|trap" gpurun_out/bench_n$NG.log | head -3
timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29701 scripts/mgpu_probe.py dambreak2d_1m 2 2>&1 | grep PROBE | python scripts/probe_digest.py
