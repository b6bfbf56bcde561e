#!/bin/bash
# GPU-box call (1 GPU): where the step goes after the preconditioner — per-stage times (2-D 1M, 3-D 1M, 3-D 12M), the kernel's own
# cycle breakdown on the 12M block, the FP64 pipe peaks, and the 12M bench line
mkdir -p gpurun_out
timeout 900 python scripts/stage_probe.py dambreak2d_1m dambreak3d_1m dambreak3d_10m > gpurun_out/stage_probe.log 2>&1; cat gpurun_out/stage_probe.log
timeout 600 python scripts/cg_probe.py dambreak3d_10m > gpurun_out/probe_dambreak3d_10m.log 2>&1; grep -E "per iteration|us_per_iter|ms_per_step|iters_last|precond|CTA 0|rror" gpurun_out/probe_dambreak3d_10m.log
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/fp64_peak tools/fp64_peak.cu && tools/fp64_peak > gpurun_out/fp64_peak.json 2>&1; cat gpurun_out/fp64_peak.json
( timeout 900 python bench.py --workload dambreak3d_10m --steps 3 --warmup 3 --no-cpu-baseline ) > gpurun_out/bench_3d10m_pcg.log 2>&1; grep '^{' gpurun_out/bench_3d10m_pcg.log | cut -c1-400
