#!/bin/bash
# quick GPU-box check: parity tests + default bench (no CPU baseline)
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -25 gpurun_out/pytest_gpu.log
( time timeout 600 python bench.py --no-cpu-baseline "$@" ) > gpurun_out/bench_quick.log 2>&1
tail -4 gpurun_out/bench_quick.log
