#!/bin/bash
# GPU-box call (1 GPU): bulk-copy issue rate with several issuing threads (tools/tma_bench), register-only chunk walk (parity + time)
mkdir -p gpurun_out
timeout 300 tools/tma_bench 2>&1 | grep -E "multi|clock" | tee gpurun_out/tma_multi.log
( timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_multigrid.py -m gpu -x -q ) > gpurun_out/pytest_chunk.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_chunk.log; tail -4 gpurun_out/pytest_chunk.log | cut -c1-1500
timeout 900 python scripts/stage_probe.py dambreak2d_1m dambreak3d_10m 2>&1 | grep workload | cut -c1-700 | tee gpurun_out/stage_probe_q.log
