#!/bin/bash
# GPU-box call A: parity tests, smoke, default bench, full ncu capture of every gather/grid kernel of one 1M-particle step.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
nproc > gpurun_out/nproc.txt; lscpu | grep -E 'Model name|^CPU\(s\)' >> gpurun_out/nproc.txt
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
( time timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/smoke.log 2>&1
( time timeout 900 python bench.py ) > gpurun_out/bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on \
    -k regex:'k_density|k_search|k_ppe_fill|k_gradient|k_ecs|k_explicit_accel|k_ds|k_chunk_build|k_reorder' -s 36 -c 12 -o gpurun_out/prof_gather \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/prof_gather.log 2>&1
ls -la gpurun_out
tail -3 gpurun_out/pytest_gpu.log; tail -2 gpurun_out/smoke.log; tail -3 gpurun_out/bench.log
