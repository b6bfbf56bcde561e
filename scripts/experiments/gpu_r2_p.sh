#!/bin/bash
# GPU-box call (1 GPU): default bench line + reference arm + launch list of the same command
mkdir -p gpurun_out
( time timeout 900 python bench.py ) > gpurun_out/bench_p.log 2>&1; grep '^{' gpurun_out/bench_p.log | cut -c1-300
( time timeout 600 python bench.py --impl reference --steps 3 --warmup 1 ) > gpurun_out/bench_ref_p.log 2>&1; grep '^{' gpurun_out/bench_ref_p.log | cut -c1-300
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_p.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-large > gpurun_out/bench_under_ncu_p.log 2>&1
tail -3 gpurun_out/bench_p.log
