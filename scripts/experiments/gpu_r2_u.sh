#!/bin/bash
# GPU-box call (1 GPU): the small configurations (C1 StaticPressure 6 040, C2a Sample 1 323) as us/step, C3 (CentralGravity 4 M) with its CPU baseline
mkdir -p gpurun_out
timeout 600 python scripts/stage_probe.py dambreak2d_default static_pressure dambreak2d_72k 2>&1 | grep workload | cut -c1-700 | tee gpurun_out/stage_probe_small.log
( timeout 900 python bench.py --workload central_gravity_4m --steps 5 --warmup 3 --no-large ) > gpurun_out/bench_c3.log 2>&1; grep '^{' gpurun_out/bench_c3.log | cut -c1-300
( timeout 900 python bench.py --workload static_pressure --steps 50 --warmup 10 --no-large ) > gpurun_out/bench_c1.log 2>&1; grep '^{' gpurun_out/bench_c1.log | cut -c1-300
( timeout 900 python bench.py --workload dambreak2d_default --steps 50 --warmup 10 --no-large ) > gpurun_out/bench_c2a.log 2>&1; grep '^{' gpurun_out/bench_c2a.log | cut -c1-300
