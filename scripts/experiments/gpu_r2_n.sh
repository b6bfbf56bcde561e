#!/bin/bash
# GPU-box call (1 GPU): window ranges by LDGSTS (MPS_CG_LDGSTS=1) vs one bulk copy per range — correctness and effect
mkdir -p gpurun_out
( MPS_CG_LDGSTS=1 timeout 900 python -m pytest tests/test_multigrid.py tests/test_gpu_parity.py -m gpu -x -q ) > gpurun_out/pytest_ldgsts.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_ldgsts.log; tail -5 gpurun_out/pytest_ldgsts.log | cut -c1-1500
for f in 0 1; do
  echo "== MPS_CG_LDGSTS=$f"
  MPS_CG_LDGSTS=$f timeout 600 python scripts/stage_probe.py dambreak3d_10m dambreak2d_1m 2>&1 | grep workload | cut -c1-200
  MPS_CG_LDGSTS=$f timeout 600 python scripts/cg_probe.py dambreak3d_10m 2>&1 | grep -E "per iteration"
done
