#!/bin/bash
# GPU-box call E (1 GPU): parity tests with the adaptive CTA split, per-CTA profiles, default bench.
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -6 gpurun_out/pytest_gpu.log
F="per iteration|us_per_iter|rror|^phase1|^phase2|^wait_data|^barriers|slowest"
timeout 300 python scripts/cg_probe.py dambreak2d_1m 2>&1 | grep -E "$F"
echo "== adaptive off"; MPS_CG_ADAPTIVE=0 timeout 300 python scripts/cg_probe.py dambreak2d_1m 2>&1 | grep -E "$F"
echo "== 3d 1m"; timeout 300 python scripts/cg_probe.py dambreak3d_1m 2>&1 | grep -E "$F"
( timeout 600 python bench.py --no-cpu-baseline ) > gpurun_out/bench_quick.log 2>&1; grep '^{' gpurun_out/bench_quick.log | cut -c1-250
