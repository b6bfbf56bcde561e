"""Worker of the multi-GPU parity test (launched with torchrun, one process per GPU): the same scene is advanced by the
slab-decomposed solver on WORLD_SIZE GPUs and, on rank 0, by a single-GPU solver; the states must agree to 1e-9
(only the order of the dot-product sums differs)."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from openmps_b200 import capi, scenes  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    which = sys.argv[1] if len(sys.argv) > 1 else "dambreak2d"
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    sc = {"dambreak2d": lambda: scenes.dambreak2d_fast(1.6e-3), "dambreak3d": lambda: scenes.dambreak3d(1.2e-2),
          "static": lambda: scenes.static_pressure(),
          # larger blocks for 4 / 8 ranks (a slab must stay thicker than the neighbour stencil)
          "dambreak2d_72k": lambda: scenes.dambreak2d_fast(8e-4), "dambreak3d_123k": lambda: scenes.dambreak3d(8e-3)}[which]()
    uid = torch.zeros(128, dtype=torch.uint8, device=f"cuda:{local}")
    if rank == 0:
        uid.copy_(torch.tensor(list(capi.GpuComputer.comm_unique_id()), dtype=torch.uint8))
    dist.broadcast(uid, 0)
    g = capi.GpuComputer.from_scene(sc, device=local)
    g.attach_comm(rank, world, bytes(uid.cpu().numpy().tobytes()))
    g.forward(steps)
    info = g.comm_info()
    st = g.state()
    stats = g.stats_dict()
    out = {"rank": rank, "own": info["own"], "n": sc.count, "iters": stats["cg_iterations"], "comm_calls": stats.get("comm_calls"), "mode": info["mode"]}
    if rank == 0:
        ref = capi.GpuComputer.from_scene(sc, device=local)
        ref.forward(steps)
        rs = ref.state()
        rstats = ref.stats_dict()

        def rel(a, b):
            return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))
        out.update({"type_equal": bool(np.array_equal(st["type"], rs["type"])), "err_x": rel(st["x"], rs["x"]), "err_u": rel(st["u"], rs["u"]),
                    "err_p": rel(st["p"], rs["p"]), "err_n": rel(st["n"], rs["n"]), "iters_1gpu": rstats["cg_iterations"]})
    # every rank must hold the same replicated state
    x = torch.from_numpy(np.ascontiguousarray(st["x"])).cuda()
    x0 = x.clone()
    dist.broadcast(x0, 0)
    out["replicas_equal"] = bool(torch.equal(x, x0))
    print("MGPU " + json.dumps(out), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
