#!/bin/bash
# GPU-box call C (1 GPU): per-CTA profile of the CG kernel, the other BASELINE configs that fit one GPU.
mkdir -p gpurun_out
timeout 300 python scripts/cg_probe.py dambreak2d_1m > gpurun_out/probe_2d_1m.log 2>&1; tail -12 gpurun_out/probe_2d_1m.log
timeout 300 python scripts/cg_probe.py dambreak3d_1m > gpurun_out/probe_3d_1m.log 2>&1; tail -12 gpurun_out/probe_3d_1m.log
( timeout 900 python bench.py --workload central_gravity_4m --steps 3 --warmup 3 --no-cpu-baseline ) > gpurun_out/bench_cg4m.log 2>&1; grep '^{' gpurun_out/bench_cg4m.log | cut -c1-300; tail -3 gpurun_out/bench_cg4m.log | cut -c1-300
( timeout 900 python bench.py --workload dambreak3d_10m --steps 3 --warmup 3 --no-cpu-baseline ) > gpurun_out/bench_3d10m.log 2>&1; grep '^{' gpurun_out/bench_3d10m.log | cut -c1-300; tail -3 gpurun_out/bench_3d10m.log | cut -c1-300
nvidia-smi --query-gpu=memory.used,memory.total --format=csv
