#!/bin/bash
# GPU-box call (1 GPU): ncu launch lists (per-kernel durations) of one step of the 3-D 12M block and of the 2-D 1M block
mkdir -p gpurun_out
for w in dambreak3d_10m dambreak2d_1m; do
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_$w.csv \
    python bench.py --workload $w --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_under_ncu_$w.log 2>&1
python scripts/summarize_launches.py gpurun_out/launches_$w.csv 2>&1 | head -60
done
