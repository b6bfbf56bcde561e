#!/bin/bash
# GPU-box call (1 GPU): hierarchy set-up with one launch per table for all levels, levels capped by the dense grid — parity and effect
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_multigrid.py tests/test_upstream_gtests.py -m gpu -x -q -k "not c3_c4" ) > gpurun_out/pytest_mgsetup.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_mgsetup.log; tail -6 gpurun_out/pytest_mgsetup.log | cut -c1-1500
timeout 900 python scripts/stage_probe.py dambreak2d_default static_pressure dambreak2d_1m dambreak3d_10m 2>&1 | grep workload | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['workload'], round(d['ms_per_step'], 3), 'launches', d['launches_per_step'], 'its', round(d['iters_per_step'],1), 'levels', d['mg_levels'], 'ppe_assemble', d['stage_ms_per_step']['ppe_assemble'], 'cg', d['stage_ms_per_step']['cg'])"
