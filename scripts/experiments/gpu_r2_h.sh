#!/bin/bash
# GPU-box call (1 GPU): full GPU test suite after the allocation changes; BASELINE configs[4] (101M particles, 3-D) on ONE GPU
mkdir -p gpurun_out
free -g | head -2; nvidia-smi --query-gpu=memory.total,memory.used --format=csv
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu.log; tail -6 gpurun_out/pytest_gpu.log | cut -c1-1500
MPS_ALLOC_SLACK=0 timeout 900 python scripts/mgpu_large.py dambreak3d_100m 2 2 stages > gpurun_out/large_100m_n1.log 2>&1; grep -E "LARGE|rror|Traceback" gpurun_out/large_100m_n1.log | cut -c1-1500
tail -3 gpurun_out/large_100m_n1.log | cut -c1-600
