#!/bin/bash
# GPU-box call (1 GPU): last check of the round on the final code — all GPU tests, smoke, default bench, reference arm
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu_final2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_final2.log; tail -5 gpurun_out/pytest_gpu_final2.log
( timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/smoke_final2.log 2>&1; tail -1 gpurun_out/smoke_final2.log
( time timeout 900 python bench.py ) > gpurun_out/bench_final2.log 2>&1; grep '^{' gpurun_out/bench_final2.log | cut -c1-200
( time timeout 600 python bench.py --impl reference --steps 3 --warmup 1 ) > gpurun_out/bench_ref_final2.log 2>&1; grep '^{' gpurun_out/bench_ref_final2.log | cut -c1-200
