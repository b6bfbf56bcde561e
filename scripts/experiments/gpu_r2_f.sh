#!/bin/bash
# GPU-box call (gpurun --gpus 2): tests, N = 2 probes and bench lines after the work-weighted slab split, ncu captures exported as text
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests/test_longrun_observables.py tests/test_gpu_parity.py tests/test_multi_gpu.py -m gpu -x -q -k "longrun or long_run or wall_motion or zhou or two_gpus" ) > gpurun_out/pytest_new.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_new.log
tail -30 gpurun_out/pytest_new.log | cut -c1-2000
for w in dambreak2d_2m dambreak3d_10m; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29701 scripts/mgpu_probe.py $w 3 > gpurun_out/probe_${w}_n2.log 2>&1; grep PROBE gpurun_out/probe_${w}_n2.log | cut -c1-2000
done
timeout 300 python scripts/stage_probe.py dambreak2d_2m 2>&1 | tail -1
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29700 bench.py --gpus 2 ) > gpurun_out/bench_n2.log 2>&1
( time timeout 900 python bench.py ) > gpurun_out/bench_n1.log 2>&1
python scripts/show_line.py gpurun_out/bench_n1.log gpurun_out/bench_n2.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_pcg_stream -s 3 -c 1 -o gpurun_out/prof_pcg_2d1m \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-large > gpurun_out/prof_pcg_2d1m.log 2>&1
bash scripts/ncu_export.sh gpurun_out/prof_pcg_2d1m.ncu-rep
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_pcg_stream -s 3 -c 1 -o gpurun_out/prof_pcg_3d10m \
    python bench.py --workload dambreak3d_10m --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/prof_pcg_3d10m.log 2>&1
bash scripts/ncu_export.sh gpurun_out/prof_pcg_3d10m.ncu-rep
timeout 900 ncu --set full --clock-control none -k regex:'k_search|k_density|k_ecs|k_explicit_accel|k_ppe_fill|k_gradient|k_ds|k_chunk_build|k_reorder' -s 36 -c 12 -o gpurun_out/prof_gather_3d1m \
    python bench.py --workload dambreak3d_1m --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/prof_gather_3d1m.log 2>&1
bash scripts/ncu_export.sh gpurun_out/prof_gather_3d1m.ncu-rep
du -sh gpurun_out
