#!/bin/bash
# GPU-box call (1 GPU): size limit of the top level of the V-cycle (MPS_MG_TOP_CELLS) and its extra sweeps
mkdir -p gpurun_out
for cfg in "64 4" "128 4" "256 4" "256 8" "600 8"; do
  set -- $cfg
  echo "== TOP_CELLS=$1 TOP_SWEEPS=$2"
  MPS_MG_TOP_CELLS=$1 MPS_MG_TOP_SWEEPS=$2 timeout 600 python scripts/stage_probe.py dambreak2d_1m dambreak3d_1m 2>&1 | grep workload | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['workload'], round(d['ms_per_step'], 3), 'cg', round(d['cg_ms'],3), 'its', round(d['iters_per_step'],1), 'levels', d['mg_levels'])"
done | tee gpurun_out/top_cells.log
