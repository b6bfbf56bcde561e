#!/bin/bash
# GPU-box call (1 GPU): single-launch scan for short inputs — parity and effect on small and large scenes
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_multigrid.py tests/test_upstream_gtests.py tests/test_driver_gpu.py -m gpu -x -q -k "not c3_c4" ) > gpurun_out/pytest_z.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_z.log; tail -4 gpurun_out/pytest_z.log | cut -c1-1500
timeout 900 python scripts/stage_probe.py dambreak2d_default static_pressure dambreak2d_1m 2>&1 | grep workload | cut -c1-60,100-170,250-700 | tee gpurun_out/stage_probe_z.log
