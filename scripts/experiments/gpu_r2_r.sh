#!/bin/bash
# GPU-box call (1 GPU): several producer warps per CTA (MPS_CG_PRODUCERS) — parity and effect on the solve
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_multigrid.py -m gpu -x -q ) > gpurun_out/pytest_prod.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_prod.log; tail -4 gpurun_out/pytest_prod.log | cut -c1-1500
for p in 1 2 4; do
  echo "== MPS_CG_PRODUCERS=$p"
  MPS_CG_PRODUCERS=$p timeout 900 python scripts/stage_probe.py dambreak2d_1m dambreak3d_10m 2>&1 | grep workload | cut -c1-200
done | tee gpurun_out/stage_probe_r.log
timeout 600 python scripts/cg_probe.py dambreak3d_10m 2>&1 | grep -E "per iteration|CTA 0" | tee -a gpurun_out/stage_probe_r.log
