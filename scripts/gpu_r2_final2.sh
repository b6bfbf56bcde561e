#!/bin/bash
# GPU-box call (1 GPU): last check of the round on the final code — all GPU tests, smoke, default bench, reference arm, ncu launch list
# of the bench command, per-stage probe of the small and large configurations
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu_final2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_final2.log; tail -5 gpurun_out/pytest_gpu_final2.log
( timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/smoke_final2.log 2>&1; tail -1 gpurun_out/smoke_final2.log
( time timeout 900 python bench.py ) > gpurun_out/bench_final2.log 2>&1; grep '^{' gpurun_out/bench_final2.log | cut -c1-200
( time timeout 600 python bench.py --impl reference --steps 3 --warmup 1 ) > gpurun_out/bench_ref_final2.log 2>&1; grep '^{' gpurun_out/bench_ref_final2.log | cut -c1-200
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_final2.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-large > gpurun_out/bench_under_ncu_final2.log 2>&1
timeout 900 python scripts/stage_probe.py dambreak2d_default static_pressure dambreak2d_72k dambreak2d_1m dambreak3d_1m dambreak3d_10m 2>&1 | grep workload > gpurun_out/stage_probe_final2.log
( timeout 600 python bench.py --workload static_pressure --steps 50 --warmup 10 --no-large ) > gpurun_out/bench_c1_final2.log 2>&1
( timeout 600 python bench.py --workload dambreak2d_default --steps 50 --warmup 10 --no-large ) > gpurun_out/bench_c2a_final2.log 2>&1
( timeout 900 python bench.py --workload central_gravity_4m --steps 5 --warmup 3 --no-large --no-cpu-baseline ) > gpurun_out/bench_c3_final2.log 2>&1
