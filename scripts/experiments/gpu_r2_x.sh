#!/bin/bash
# GPU-box call (1 GPU): k_ppe_fill with 3 window ranges in 2-D and a 96-register cap — parity and effect
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_multigrid.py -m gpu -x -q -k "not c3_c4" ) > gpurun_out/pytest_x.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_x.log; tail -4 gpurun_out/pytest_x.log | cut -c1-1500
timeout 900 python scripts/stage_probe.py dambreak2d_1m dambreak3d_10m 2>&1 | grep workload | cut -c1-40,250-700 | tee gpurun_out/stage_probe_x.log
