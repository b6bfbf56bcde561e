#!/bin/bash
# GPU-box call (gpurun --gpus 2): new tests (observables, wall motion), per-stage probe of the 2-rank step, small-workload bench lines
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests/test_longrun_observables.py tests/test_gpu_parity.py -m gpu -x -q -k "longrun or long_run or wall_motion or zhou" ) > gpurun_out/pytest_new.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_new.log
tail -30 gpurun_out/pytest_new.log | cut -c1-2000
for w in dambreak2d_1m dambreak3d_10m; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29701 scripts/mgpu_probe.py $w 3 > gpurun_out/probe_${w}_n2.log 2>&1; grep PROBE gpurun_out/probe_${w}_n2.log | cut -c1-1800
done
timeout 300 python scripts/stage_probe.py dambreak2d_default static_pressure dambreak2d_72k 2>&1 | tail -3
