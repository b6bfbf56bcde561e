#!/usr/bin/env python
"""Development probe (torchrun, one process per GPU): in-kernel cycle breakdown of the multi-GPU persistent CG kernel.
usage: torchrun ... scripts/mgpu_probe.py WORKLOAD [steps]"""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from openmps_b200 import capi  # noqa: E402
import bench  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
name = sys.argv[1] if len(sys.argv) > 1 else "dambreak2d_1m"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
sc = bench.WORKLOADS[name][0]()
uid = torch.zeros(128, dtype=torch.uint8, device=f"cuda:{local}")
if rank == 0:
    uid.copy_(torch.tensor(list(capi.GpuComputer.comm_unique_id()), dtype=torch.uint8))
dist.broadcast(uid, 0)
g = capi.GpuComputer.from_scene(sc, device=local)
g.attach_comm(rank, world, bytes(uid.cpu().numpy().tobytes()))
g.forward(3)
g.set_cg_profile(True)
g.reset_stats()
dist.barrier()
ms = g.run_steps(steps)
st = g.stats_dict()
it = st["last_cg_iterations"]
out = {"rank": rank, "mode": g.comm_info()["mode"], "own": g.comm_info()["own"], "ms_per_step": ms / steps, "cg_ms": st["cg_ms"] / steps,
       "iters": st["cg_iterations"], "us_per_iter": 1e3 * st["cg_ms"] / max(st["cg_iterations"], 1), "nnz": st["nnz"]}
if g.comm_info()["mode"] == "peer-memory" and it:
    raw = g.cg_profile_raw().astype(float)
    for k, nm in enumerate(["phase1", "wait_data", "phase2", "reductions", "producer_wait", "chunks", "iteration"]):
        v = raw[:, k] / (it + (1 if nm in ("wait_data", "producer_wait", "chunks") else 0))
        out[nm] = {"min": float(v.min()), "median": float(np.median(v)), "max": float(v.max()), "cta0": float(v[0])}
# per-stage device time (adds a sync per stage): where the non-CG part of a multi-GPU step goes, rank by rank
g.set_cg_profile(False)
g.set_stage_timing(True)
g.reset_stats()
dist.barrier()
g.forward(2)
st2 = g.stats_dict()
out["stage_ms_per_step"] = {k: round(v / 2, 3) for k, v in st2["stage_ms"].items()}
for r in range(world):
    if r == rank:
        print("PROBE " + json.dumps(out), flush=True)
    dist.barrier()
dist.destroy_process_group()
