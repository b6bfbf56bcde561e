#!/bin/bash
# GPU-box call (round 2, first contact of the preconditioned solve): parity tests, smoke, probe of the new kernel, bench lines
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -30 gpurun_out/pytest_gpu.log
( timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
for w in dambreak2d_72k dambreak2d_1m; do
  timeout 300 python scripts/cg_probe.py $w > gpurun_out/probe_$w.log 2>&1; grep -E "per iteration|us_per_iter|ms_per_step|iters_last|precond|rror" gpurun_out/probe_$w.log
done
( timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline ) > gpurun_out/bench_pcg.log 2>&1; grep '^{' gpurun_out/bench_pcg.log | cut -c1-1500; tail -3 gpurun_out/bench_pcg.log | cut -c1-300
( MPS_CG_PRECOND=0 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e ) > gpurun_out/bench_plain.log 2>&1; grep '^{' gpurun_out/bench_plain.log | cut -c1-400
