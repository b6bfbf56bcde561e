#!/usr/bin/env python
"""prints the essentials of bench.py JSON lines found in the given log files (development aid)"""
import json
import sys

for path in sys.argv[1:]:
    for l in open(path):
        if not l.startswith("{"):
            continue
        d = json.loads(l)
        r = d.get("roofline") or {}
        print(path, {k: d.get(k) for k in ("impl", "n_gpus", "value", "ms_per_step", "scaling")}, "workload", d["config"]["workload"],
              "e2e", (d.get("e2e") or {}).get("value"), "frac", r.get("frac"), "cg_ms", r.get("kernel_ms_per_launch"),
              "iters/step", (d.get("detail") or {}).get("cg_iterations_per_step"))
        print("   stages", (d.get("detail") or {}).get("stage_ms_per_step"))
        print("   large", d.get("large"))
        print("   cpu", d.get("cpu_baseline"))
