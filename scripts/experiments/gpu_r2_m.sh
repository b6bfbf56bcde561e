#!/bin/bash
# GPU-box call (1 GPU): odd ring depths (per-(stage, group) full barriers) — correctness and the 3-D / 2-D effect
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_multigrid.py tests/test_gpu_parity.py -m gpu -x -q ) > gpurun_out/pytest_ring.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_ring.log; tail -5 gpurun_out/pytest_ring.log | cut -c1-1500
for st in 0 2; do
  echo "== 3-D 12M MPS_CG_STAGES=$st (0: default = as many as fit)"
  MPS_CG_STAGES=$st timeout 600 python scripts/stage_probe.py dambreak3d_10m 2>&1 | grep workload | cut -c1-420
done
for st in 0 4; do
  echo "== 2-D 1M MPS_CG_STAGES=$st"
  MPS_CG_STAGES=$st timeout 600 python scripts/stage_probe.py dambreak2d_1m 2>&1 | grep workload | cut -c1-420
done
MPS_CG_PRECOND=0 timeout 600 python scripts/stage_probe.py dambreak2d_250k 2>&1 | grep workload | cut -c1-300
timeout 600 python scripts/cg_probe.py dambreak3d_10m 2>&1 | grep -E "per iteration|CTA 0"
