#!/usr/bin/env python
"""Development probe: in-kernel cycle breakdown of the streaming CG kernel on one workload."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from openmps_b200 import capi, scenes  # noqa: E402
import bench  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "dambreak2d_1m"
sc = bench.WORKLOADS[name][0]()
gpu = capi.GpuComputer.from_scene(sc, device=0)
gpu.forward(3)
gpu.set_cg_profile(True)
gpu.reset_stats()
ms = gpu.run_steps(2)
st = gpu.stats_dict()
pr = gpu.cg_profile()
it = st["last_cg_iterations"]
print(json.dumps({"workload": name, "n": sc.count, "ms_per_step": ms / 2, "cg_ms": st["cg_ms"] / 2, "iters_last": it,
                  "us_per_iter": 1e3 * st["cg_ms"] / max(st["cg_iterations"], 1), "nnz": st["nnz"], "profile": pr}, indent=1))
m = pr["mean"]
if it:
    print("per iteration (cycles, mean over CTAs): phase1 %.0f  wait_data %.0f  phase2 %.0f  barriers %.0f  producer_wait %.0f  chunks/CTA %.1f  vcycle %.0f  total %.0f" % (
        m["phase1"] / it, m["wait_data"] / (it + 1), m["phase2"] / it, m["barriers"] / it, m["wait_stage"] / (it + 1), m["chunks_per_cta"] / (it + 1), m["vcycle"] / it,
        m["iteration_cycles_total"] / it))
    print("preconditioner: levels %d, cells %d, matrix sweeps %d" % (st["mg_levels"], st["mg_cells"], st["matrix_sweeps"]))

stg = gpu.cg_profile_stages().astype(float)
if it and stg.sum() > 0:
    us = stg / it / 1.965e3   # cycles per iteration -> microseconds at 1965 MHz
    print("CTA 0, us per iteration: down " + " ".join(f"{v:.1f}" for v in us[0:st["mg_levels"] - 1]) + f" | top {us[16]:.1f} | hand-back wait {us[17]:.1f} | up "
          + " ".join(f"{v:.1f}" for v in us[20:20 + st["mg_levels"] - 1]) + f" | rows 2a {us[40]:.1f} 2b {us[41]:.1f} | reductions pAp {us[43]:.1f} rr {us[42]:.1f} rz {us[44]:.1f}")
raw = gpu.cg_profile_raw().astype(float)
if len(raw) and it:
    import numpy as np
    os.makedirs("gpurun_out", exist_ok=True)
    np.save(f"gpurun_out/cg_profile_{name}.npy", raw / it)
    for k, nm in enumerate(["phase1", "wait_data", "phase2", "barriers", "producer_wait", "chunks"]):
        v = raw[:, k] / it
        print(f"{nm:14s} min {v.min():10.0f}  p10 {np.percentile(v,10):10.0f}  median {np.median(v):10.0f}  p90 {np.percentile(v,90):10.0f}  max {v.max():10.0f}")
    order = np.argsort(-raw[:, 0])[:6]
    print("slowest CTAs (id, phase1, chunks):", [(int(i), int(raw[i,0]/it), round(raw[i,5]/(it+1),1)) for i in order])
