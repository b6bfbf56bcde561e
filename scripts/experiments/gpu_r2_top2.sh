#!/bin/bash
# GPU-box call (1 GPU): top level of the V-cycle at <= 256 cells — extra sweeps, and the other workloads
mkdir -p gpurun_out
for cfg in "256 2" "256 3" "256 6" "256 4"; do
  set -- $cfg
  echo "== TOP_CELLS=$1 TOP_SWEEPS=$2"
  MPS_MG_TOP_CELLS=$1 MPS_MG_TOP_SWEEPS=$2 timeout 600 python scripts/stage_probe.py dambreak2d_1m 2>&1 | grep workload | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['workload'], round(d['ms_per_step'], 3), 'cg', round(d['cg_ms'],3), 'its', round(d['iters_per_step'],1), 'levels', d['mg_levels'])"
done | tee gpurun_out/top_cells2.log
for tc in 64 256; do
  echo "== TOP_CELLS=$tc (4 sweeps), other workloads"
  MPS_MG_TOP_CELLS=$tc timeout 900 python scripts/stage_probe.py dambreak2d_default static_pressure dambreak2d_72k dambreak2d_250k central_gravity_4m dambreak3d_123k dambreak3d_10m 2>&1 | grep workload | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['workload'], round(d['ms_per_step'], 3), 'cg', round(d['cg_ms'],3), 'its', round(d['iters_per_step'],1), 'levels', d['mg_levels'])"
done | tee -a gpurun_out/top_cells2.log
