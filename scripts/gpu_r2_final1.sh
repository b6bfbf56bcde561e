#!/bin/bash
# GPU-box call (1 GPU): the round's reference run on the final code — all GPU tests, smoke, default bench (+ large), reference arm,
# ncu launch list of the bench command, full ncu captures: the preconditioned CG kernel (2-D 1M) and every gather /
# grid kernel of one step (2-D 1M and 3-D 1M).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu_final.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_final.log; tail -5 gpurun_out/pytest_gpu_final.log
( timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/smoke_final.log 2>&1; tail -1 gpurun_out/smoke_final.log
( time timeout 900 python bench.py ) > gpurun_out/bench_final.log 2>&1; grep '^{' gpurun_out/bench_final.log | cut -c1-200
( time timeout 600 python bench.py --impl reference --steps 3 --warmup 1 ) > gpurun_out/bench_ref_final.log 2>&1; grep '^{' gpurun_out/bench_ref_final.log | cut -c1-200
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_final.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-large > gpurun_out/bench_under_ncu_final.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_pcg_stream -s 3 -c 1 -o gpurun_out/prof_pcg_2d1m_final \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-large > gpurun_out/prof_pcg_2d1m_final.log 2>&1
bash scripts/ncu_export.sh gpurun_out/prof_pcg_2d1m_final.ncu-rep > /dev/null 2>&1
for wl in dambreak2d_1m dambreak3d_1m; do
  timeout 1200 ncu --set full --clock-control none \
    -k regex:'k_density|k_search|k_ppe_fill|k_gradient|k_ecs|k_explicit_accel|k_ds|k_chunk_build|k_reorder' -s 39 -c 13 -o gpurun_out/prof_gather_${wl}_final \
    python bench.py --workload $wl --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-large > gpurun_out/prof_gather_${wl}_final.log 2>&1
  python tools/ncu_summary.py gpurun_out/prof_gather_${wl}_final.ncu-rep > gpurun_out/prof_gather_${wl}_final_summary.txt 2>&1; rm -f gpurun_out/prof_gather_${wl}_final.ncu-rep
done
ls -la gpurun_out | grep final
