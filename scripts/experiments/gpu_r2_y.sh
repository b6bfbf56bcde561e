#!/bin/bash
# GPU-box call (1 GPU): how many cells a level may have to be run by CTA 0 alone (MPS_MG_SMALL_CELLS)
mkdir -p gpurun_out
for c in 256 600 2048 8192; do
  echo "== MPS_MG_SMALL_CELLS=$c"
  MPS_MG_SMALL_CELLS=$c timeout 900 python scripts/stage_probe.py dambreak2d_1m dambreak2d_72k dambreak3d_1m dambreak3d_10m 2>&1 | grep workload | cut -c1-190
done | tee gpurun_out/stage_probe_y.log
