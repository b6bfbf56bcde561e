#!/usr/bin/env python
"""Development probe: per-stage device time of one step (CUDA events per stage) on the listed workloads."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from openmps_b200 import capi  # noqa: E402
import bench  # noqa: E402

for name in sys.argv[1:] or ["dambreak2d_1m"]:
    sc = bench.WORKLOADS[name][0]()
    gpu = capi.GpuComputer.from_scene(sc, device=0)
    gpu.forward(3)
    gpu.set_stage_timing(True)
    gpu.reset_stats()
    steps = 3
    gpu.forward(steps)
    st = gpu.stats_dict()
    gpu.set_stage_timing(False)
    gpu.reset_stats()
    ms = gpu.run_steps(steps)
    st2 = gpu.stats_dict()
    print(json.dumps({"workload": name, "n": sc.count, "ms_per_step": ms / steps, "cg_ms": st2["cg_ms"] / steps,
                      "iters_per_step": st2["cg_iterations"] / steps, "launches_per_step": st2["kernel_launches"] / steps,
                      "nnz": st2["nnz"], "mg_levels": st2["mg_levels"], "mg_cells": st2["mg_cells"],
                      "stage_ms_per_step": {k: round(v / steps, 4) for k, v in st["stage_ms"].items() if v}}))
    gpu.close()
