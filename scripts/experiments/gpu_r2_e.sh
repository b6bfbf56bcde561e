#!/bin/bash
# GPU-box call (gpurun --gpus 2): new tests, default bench lines at N = 1 and N = 2 (weak scaling + the `large` record), reference arm,
# full ncu captures of the preconditioned CG kernel (2-D 1M and 3-D 12M) and of the gather kernels in 3-D
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests/test_longrun_observables.py tests/test_gpu_parity.py -m gpu -x -q -k "longrun or long_run or wall_motion or zhou" ) > gpurun_out/pytest_new.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_new.log
tail -30 gpurun_out/pytest_new.log | cut -c1-2000
( time timeout 900 python bench.py ) > gpurun_out/bench_n1.log 2>&1; grep '^{' gpurun_out/bench_n1.log | cut -c1-3000
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29700 bench.py --gpus 2 ) > gpurun_out/bench_n2.log 2>&1; grep '^{' gpurun_out/bench_n2.log | cut -c1-3000
grep -E "rror|Traceback" gpurun_out/bench_n2.log | head -5
( time timeout 600 python bench.py --impl reference --steps 3 --warmup 1 ) > gpurun_out/bench_ref.log 2>&1; grep '^{' gpurun_out/bench_ref.log | cut -c1-1500
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29701 scripts/mgpu_probe.py dambreak3d_10m 3 > gpurun_out/probe_dambreak3d_10m_n2.log 2>&1; grep PROBE gpurun_out/probe_dambreak3d_10m_n2.log | cut -c1-300
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_pcg_stream -s 3 -c 1 -o gpurun_out/prof_pcg_2d1m \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-large > gpurun_out/prof_pcg_2d1m.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_pcg_stream -s 3 -c 1 -o gpurun_out/prof_pcg_3d10m \
    python bench.py --workload dambreak3d_10m --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/prof_pcg_3d10m.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_search|k_density|k_ecs|k_explicit_accel|k_ppe_fill|k_gradient|k_ds|k_chunk_build|k_reorder' -s 36 -c 12 -o gpurun_out/prof_gather_3d1m \
    python bench.py --workload dambreak3d_1m --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/prof_gather_3d1m.log 2>&1
ls -la gpurun_out/*.ncu-rep
