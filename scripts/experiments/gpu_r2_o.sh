#!/bin/bash
# GPU-box call (1 GPU): gather kernels with batched list walks / one sqrt per pair / fused PPE walk / sqrt-free search — parity and effect
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_multigrid.py tests/test_upstream_gtests.py -m gpu -x -q ) > gpurun_out/pytest_gather.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gather.log; tail -8 gpurun_out/pytest_gather.log | cut -c1-1500
timeout 900 python scripts/stage_probe.py dambreak2d_1m dambreak3d_1m dambreak3d_10m 2>&1 | grep workload | cut -c1-700 | tee gpurun_out/stage_probe_o.log
