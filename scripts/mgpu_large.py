#!/usr/bin/env python
"""Strong-scaling run of one LARGE block (BASELINE.json configs[3] / configs[4]) on WORLD_SIZE GPUs (torchrun) or on one GPU
(plain python): a few steps, device time (max over ranks), per-rank memory, no state download.
usage: [torchrun ...] scripts/mgpu_large.py WORKLOAD [steps] [warmup]"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from openmps_b200 import capi  # noqa: E402
import bench  # noqa: E402

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist = None
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
name = sys.argv[1] if len(sys.argv) > 1 else "dambreak3d_10m"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
warm = int(sys.argv[3]) if len(sys.argv) > 3 else 2
t0 = time.perf_counter()
sc = bench.WORKLOADS[name][0]()
t_scene = time.perf_counter() - t0
g = capi.GpuComputer.from_scene(sc, device=local)
n, fluid = sc.count, int((sc.type == 0).sum())
del sc
if world > 1:
    bench.attach(g, dist, torch, capi, rank, world, local)
t0 = time.perf_counter()
g.forward(warm)
t_warm = time.perf_counter() - t0
g.reset_stats()
if dist is not None:
    dist.barrier()
torch.cuda.synchronize()
ms = g.run_steps(steps)
torch.cuda.synchronize()
st = g.stats_dict()
t = torch.tensor([ms], dtype=torch.float64, device=f"cuda:{local}")
if dist is not None:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
free, total = torch.cuda.mem_get_info()
out = {"rank": rank, "workload": name, "n_gpus": world, "particles": n, "fluid": fluid, "steps": steps, "warmup": warm,
       "ms_per_step": float(t.item()) / steps, "value": n * steps / (float(t.item()) * 1e-3), "cg_ms_per_step": st["cg_ms"] / steps,
       "cg_iterations_per_step": st["cg_iterations"] / steps, "nnz_this_rank": st["nnz"], "mg_levels": st["mg_levels"], "mg_cells": st["mg_cells"],
       "own": g.comm_info()["own"] if world > 1 else [0, n], "mode": g.comm_info()["mode"] if world > 1 else "1 GPU",
       "hbm_used_gb": round((total - free) / 1e9, 1), "hbm_total_gb": round(total / 1e9, 1), "scene_s": round(t_scene, 1), "warmup_s": round(t_warm, 1)}
if len(sys.argv) > 4 and sys.argv[4] == "stages":
    out["stage_ms_per_step"] = bench.stage_breakdown(g, 1)
for r in range(world):
    if r == rank:
        print("LARGE " + json.dumps(out), flush=True)
    if dist is not None:
        dist.barrier()
g.close()
if dist is not None:
    dist.destroy_process_group()
