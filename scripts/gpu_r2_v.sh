#!/bin/bash
# GPU-box call (1 GPU): gather kernels with G lanes per particle (MPS_GATHER_LANES) — parity and effect
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_multigrid.py tests/test_upstream_gtests.py -m gpu -x -q ) > gpurun_out/pytest_lanes.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_lanes.log; tail -8 gpurun_out/pytest_lanes.log | cut -c1-2500
for g in 1 2 4 8; do
  echo "== MPS_GATHER_LANES=$g"
  MPS_GATHER_LANES=$g timeout 900 python scripts/stage_probe.py dambreak2d_1m dambreak3d_10m 2>&1 | grep workload | cut -c1-40,250-700
done | tee gpurun_out/stage_probe_v.log
