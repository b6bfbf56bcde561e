#!/bin/bash
# GPU-box call D (1 GPU): parity tests after the live-chunk list, per-CTA profiles, 3-D chunk geometry variants.
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -4 gpurun_out/pytest_gpu.log
F="per iteration|us_per_iter|rror|^phase1|^wait_data|^barriers|slowest|\"stages\"|\"chunks\""
timeout 300 python scripts/cg_probe.py dambreak2d_1m 2>&1 | grep -E "$F"
echo "== 3d 1m default (16 warps, lpr 4)"; timeout 300 python scripts/cg_probe.py dambreak3d_1m 2>&1 | grep -E "$F"
for cfg in "12 4" "8 4" "8 2" "16 8"; do
  set -- $cfg
  echo "== 3d 1m warps $1 lpr $2"; MPS_CG_WARPS=$1 MPS_CG_LPR=$2 timeout 300 python scripts/cg_probe.py dambreak3d_1m 2>&1 | grep -E "$F"
done
( timeout 600 python bench.py --no-cpu-baseline ) > gpurun_out/bench_quick.log 2>&1; grep '^{' gpurun_out/bench_quick.log | cut -c1-250
