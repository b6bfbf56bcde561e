#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -5 gpurun_out/pytest_gpu.log
F="per iteration|us_per_iter|chunks\"|rror|^phase1|^wait_data|^barriers|slowest"
for cfg in "12 1 12000" "12 1 30000" "16 1 12000" "16 2 12000" "8 1 12000"; do
  set -- $cfg
  echo "== warps $1 lpr $2 cost $3"; MPS_CG_WARPS=$1 MPS_CG_LPR=$2 MPS_CG_COST_FIXED=$3 timeout 300 python scripts/cg_probe.py dambreak2d_1m 2>&1 | grep -E "$F"
done
echo "== 250k"; timeout 300 python scripts/cg_probe.py dambreak2d_250k 2>&1 | grep -E "$F"
echo "== 3d 123k 16 4"; MPS_CG_WARPS=16 MPS_CG_LPR=4 timeout 300 python scripts/cg_probe.py dambreak3d_123k 2>&1 | grep -E "$F"
echo "== 3d 123k 16 2"; MPS_CG_WARPS=16 MPS_CG_LPR=2 timeout 300 python scripts/cg_probe.py dambreak3d_123k 2>&1 | grep -E "$F"
