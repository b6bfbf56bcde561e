#!/usr/bin/env python
"""Summarise an `ncu --set full` report (one line per profiled launch): duration, DRAM bytes and rate, FP64 pipe,
occupancy, hit rates.  Usage: tools/ncu_summary.py report.ncu-rep [algorithmic bytes per launch, by kernel-name substring: name=bytes ...]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
alg = dict(a.split("=") for a in sys.argv[2:])
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}


def val(r, name, scale=None):
    v = float(r[ix[name]].replace(",", ""))
    u = units[ix[name]]
    if scale == "us":
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(u, 1.0) if u in ("ns", "us", "ms", "s") else {"nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3, "second": 1e6}[u]
    if scale == "B":
        v *= {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[u]
    return v


print(f"{'kernel':34s} {'us':>9s} {'dram MB':>9s} {'dram GB/s':>10s} {'alg MB':>8s} {'alg GB/s':>9s} {'fp64 pipe %':>11s} {'SM thr %':>8s} {'warps %':>8s} {'L1 hit %':>8s} {'L2 hit %':>8s} {'regs':>5s}")
for r in data:
    name = r[ix["Kernel Name"]]
    import re
    m = re.search(r"(k_\w+(<[^>]*>)?)", name)
    short = m.group(1) if m else name[:34]
    us = val(r, "gpu__time_duration.sum", "us")
    dram = val(r, "dram__bytes_read.sum", "B") + val(r, "dram__bytes_write.sum", "B")
    a = next((float(v) for k, v in alg.items() if k in short), None)
    print(f"{short:34s} {us:9.1f} {dram / 1e6:9.1f} {dram / us / 1e3:10.0f} {(a / 1e6 if a else float('nan')):8.1f} {(a / us / 1e3 if a else float('nan')):9.0f} "
          f"{val(r, 'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active'):11.1f} {val(r, 'sm__throughput.avg.pct_of_peak_sustained_elapsed'):8.1f} "
          f"{val(r, 'sm__warps_active.avg.pct_of_peak_sustained_active'):8.1f} {val(r, 'l1tex__t_sector_hit_rate.pct'):8.1f} {val(r, 'lts__t_sector_hit_rate.pct'):8.1f} "
          f"{int(val(r, 'launch__registers_per_thread')):5d}")
