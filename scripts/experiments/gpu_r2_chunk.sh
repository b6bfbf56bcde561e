#!/bin/bash
# GPU-box call (1 GPU): lane-parallel greedy chunking — parity and kernel time
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_multigrid.py tests/test_upstream_gtests.py -m gpu -x -q -k "not c3_c4" ) > gpurun_out/pytest_chunk2.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_chunk2.log; tail -6 gpurun_out/pytest_chunk2.log | cut -c1-1500
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_chunk_build --csv --log-file gpurun_out/chunk_launches.csv python scripts/stage_probe.py dambreak2d_default dambreak2d_1m dambreak3d_1m > gpurun_out/chunk_probe.log 2>&1
python - <<'PY'
import csv
rows = [r for r in csv.reader(open('gpurun_out/chunk_launches.csv')) if len(r) > 10 and r[0].isdigit()]
import collections
d = collections.defaultdict(list)
for r in rows:
    d[(r[4][:60], r[7])].append(float(r[-1].replace(',', '')))
for k, v in d.items():
    print(k, len(v), 'median', sorted(v)[len(v)//2], r[-2] if rows else '')
PY
timeout 900 python scripts/stage_probe.py dambreak2d_1m dambreak3d_10m 2>&1 | grep workload | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['workload'], round(d['ms_per_step'], 3), 'ppe_assemble', d['stage_ms_per_step']['ppe_assemble'])"
