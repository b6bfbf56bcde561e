#!/bin/bash
# GPU-box call (1 GPU): prefetching V-cycle A/B (2-D), correctness, 3-D pipeline-geometry sweep
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_multigrid.py tests/test_gpu_parity.py tests/test_longrun_observables.py -m gpu -x -q ) > gpurun_out/pytest_pf.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_pf.log; tail -8 gpurun_out/pytest_pf.log | cut -c1-1500
for pf in 0 1; do
  echo "== MPS_MG_PREFETCH=$pf"
  MPS_MG_PREFETCH=$pf timeout 600 python scripts/cg_probe.py dambreak2d_1m > gpurun_out/probe_2d1m_pf$pf.log 2>&1; grep -E "per iteration|CTA 0|rror" gpurun_out/probe_2d1m_pf$pf.log
  MPS_MG_PREFETCH=$pf timeout 600 python scripts/stage_probe.py dambreak2d_1m dambreak2d_72k dambreak2d_default 2>&1 | grep workload | cut -c1-200
done
for cfg in "16 4 0" "12 4 0" "8 4 0" "16 8 0" "12 8 0" "16 2 0"; do
  set -- $cfg
  echo "== 3-D 12M: MPS_CG_WARPS=$1 MPS_CG_LPR=$2"
  MPS_CG_WARPS=$1 MPS_CG_LPR=$2 timeout 600 python scripts/stage_probe.py dambreak3d_10m 2>&1 | grep workload | cut -c1-260
done
