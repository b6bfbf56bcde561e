#!/bin/bash
# GPU-box call I (1 GPU): the round's reference run — parity tests, smoke, default bench (+ CPU reference), reference arm,
# ncu launch list and one full capture of the CG kernel, the 3-D 10M configuration.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -5 gpurun_out/pytest_gpu.log
( timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
( time timeout 900 python bench.py ) > gpurun_out/bench.log 2>&1; grep '^{' gpurun_out/bench.log | cut -c1-200
( time timeout 600 python bench.py --impl reference --steps 5 --warmup 1 ) > gpurun_out/bench_ref.log 2>&1; grep '^{' gpurun_out/bench_ref.log | cut -c1-300
( timeout 900 python bench.py --workload dambreak3d_10m --steps 3 --warmup 3 --no-cpu-baseline ) > gpurun_out/bench_3d10m.log 2>&1; grep '^{' gpurun_out/bench_3d10m.log | cut -c1-200
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_under_ncu.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_cg_stream -s 3 -c 1 -o gpurun_out/prof_cg \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/prof_cg.log 2>&1
# FP64 pipe peak (the roofline denominator of the gather kernels)
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/fp64_peak tools/fp64_peak.cu && tools/fp64_peak > gpurun_out/fp64_peak.json 2>&1; cat gpurun_out/fp64_peak.json
ls -la gpurun_out | head -30
