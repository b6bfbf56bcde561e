#!/bin/bash
# GPU-box call (1 GPU): 3-D 12M solve — consumer warps / lanes per row / ring depth sweep with 4 producer warps
mkdir -p gpurun_out
for cfg in "16 4 0" "16 8 0" "16 2 0" "12 4 0" "8 4 0" "8 2 0" "16 4 3" ; do
  set -- $cfg
  echo "== WARPS=$1 LPR=$2 STAGES=$3"
  MPS_CG_WARPS=$1 MPS_CG_LPR=$2 MPS_CG_STAGES=$3 timeout 300 python scripts/stage_probe.py dambreak3d_10m 2>&1 | grep workload | cut -c1-180
done | tee gpurun_out/sweep3d_s.log
