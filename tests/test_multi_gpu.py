"""Multi-GPU slab decomposition.  CPU part: the partition arithmetic and the launcher-side plumbing on a
2-rank gloo group.  GPU part (needs >= 2 GPUs): 2-GPU run == 1-GPU run on the same input."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from openmps_b200 import capi  # noqa: E402


@pytest.mark.parametrize("n,nranks", [(0, 1), (1, 4), (10, 3), (1008104, 8), (123147, 2), (7, 8)])
def test_partition_tiles_the_slots(n, nranks):
    ranges = [capi.partition_range(n, nranks, r) for r in range(nranks)]
    assert ranges[0][0] == 0 and ranges[-1][1] == n
    for (a, b), (c, d) in zip(ranges, ranges[1:]):
        assert a <= b == c <= d
    sizes = [b - a for a, b in ranges]
    assert max(sizes) == (n + nranks - 1) // nranks if n else max(sizes) == 0


_GLOO_WORKER = r"""
import os, sys
import torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from openmps_b200 import capi
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
n = 1008104
# launcher plumbing: a 128-byte id made on rank 0 reaches every rank unchanged
uid = torch.zeros(128, dtype=torch.uint8)
if rank == 0:
    uid = torch.arange(128, dtype=torch.uint8)
dist.broadcast(uid, 0)
assert bytes(uid.numpy().tobytes()) == bytes(range(128))
# every rank derives the same decomposition; together the slabs tile the slots exactly once
mine = torch.tensor(capi.partition_range(n, world, rank), dtype=torch.int64)
alls = [torch.zeros(2, dtype=torch.int64) for _ in range(world)]
dist.all_gather(alls, mine)
cover = torch.zeros(n, dtype=torch.int32)
for a, b in (t.tolist() for t in alls):
    cover[a:b] += 1
assert int(cover.min()) == 1 and int(cover.max()) == 1
# the weak-scaling reduction bench.py uses: max over ranks of the per-rank time
t = torch.tensor([1.0 + rank], dtype=torch.float64)
dist.all_reduce(t, op=dist.ReduceOp.MAX)
assert float(t) == float(world)
dist.destroy_process_group()
print("GLOO_OK", rank)
"""


def test_two_rank_gloo_partition_and_plumbing(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(_GLOO_WORKER)
    env = dict(os.environ, OMP_NUM_THREADS="1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", "29611", str(script), ROOT], capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count("GLOO_OK") == 2


def _gpu_count():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.gpu
@pytest.mark.parametrize("coupling", ["peer-memory", "nccl"])
@pytest.mark.parametrize("scene,steps", [("dambreak2d", 5), ("static", 3), ("dambreak3d", 2)])
def test_two_gpus_match_one_gpu(scene, steps, coupling):
    """The slab-decomposed step on 2 GPUs == the 1-GPU step, for both couplings of the CG iteration: the persistent kernel
    over NVLink peer memory (the default where the ranks can map each other's memory) and NCCL between per-phase launches."""
    if _gpu_count() < 2:
        pytest.skip("needs two GPUs (run with gpurun --gpus 2)")
    env = dict(os.environ)
    if coupling == "nccl":
        # the NCCL coupling runs the reference's plain CG between per-phase launches: the one-GPU run it is compared with does too
        env["MPS_COMM_NCCL_ONLY"] = "1"
        env["MPS_CG_PRECOND"] = "0"
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", "29612", os.path.join(ROOT, "tests", "multi_gpu_worker.py"), scene, str(steps)],
                       capture_output=True, text=True, timeout=900, env=env)
    assert r.returncode == 0, r.stdout[-4000:] + r.stderr[-4000:]
    rows = [json.loads(l[5:]) for l in r.stdout.splitlines() if l.startswith("MGPU ")]
    assert len(rows) == 2
    r0 = next(x for x in rows if x["rank"] == 0)
    assert all(x["mode"] == coupling for x in rows), [x["mode"] for x in rows]
    assert all(x["replicas_equal"] for x in rows)
    assert r0["type_equal"]
    # only the summation order of the CG dot products differs between 1 and 2 GPUs
    assert r0["err_x"] <= 1e-9 and r0["err_u"] <= 1e-6 and r0["err_n"] <= 1e-9 and r0["err_p"] <= 1e-5, r0
    assert abs(r0["iters"] - r0["iters_1gpu"]) <= max(3, 0.01 * r0["iters_1gpu"]), r0
    assert r0["comm_calls"] and r0["comm_calls"] > 0


@pytest.mark.gpu
@pytest.mark.parametrize("nranks", [4, 8])
@pytest.mark.parametrize("scene,steps", [("dambreak2d_72k", 5), ("dambreak3d_123k", 3)])
def test_four_and_eight_gpus_match_one_gpu(scene, steps, nranks):
    """The same on 4 and 8 ranks (peer-memory coupling, multigrid-preconditioned CG with the first levels of the cell hierarchy
    distributed over the slabs).  Bounds: x / n as on 2 GPUs; p and u looser on the 2-D block -- the PPE of a dam break at rest is
    ill-conditioned (two 1-GPU runs that differ only in the CG split differ by as much, test_adaptive_split_changes_rounding_only)."""
    if _gpu_count() < nranks:
        pytest.skip(f"needs {nranks} GPUs (run with gpurun --gpus {nranks})")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nranks}", "--master-addr", "127.0.0.1",
                        "--master-port", "29613", os.path.join(ROOT, "tests", "multi_gpu_worker.py"), scene, str(steps)],
                       capture_output=True, text=True, timeout=900, env=dict(os.environ))
    assert r.returncode == 0, r.stdout[-4000:] + r.stderr[-4000:]
    rows = [json.loads(l[5:]) for l in r.stdout.splitlines() if l.startswith("MGPU ")]
    assert len(rows) == nranks
    r0 = next(x for x in rows if x["rank"] == 0)
    assert all(x["mode"] == "peer-memory" for x in rows)
    assert all(x["replicas_equal"] for x in rows)
    owns = sorted(tuple(x["own"]) for x in rows)
    assert owns[0][0] == 0 and owns[-1][1] == r0["n"] and all(a[1] == b[0] for a, b in zip(owns, owns[1:]))   # the slabs tile the slots
    assert r0["type_equal"]
    assert r0["err_x"] <= 1e-9 and r0["err_n"] <= 1e-8 and r0["err_u"] <= 1e-3 and r0["err_p"] <= 1e-3, r0
    assert abs(r0["iters"] - r0["iters_1gpu"]) <= max(3, 0.02 * r0["iters_1gpu"]), r0


def test_bench_leaves_openmp_binding_alone_on_several_gpus():
    """bench.py must not export OMP_PROC_BIND to multi-GPU ranks (libgomp would bind every rank's host thread to the same
    core), must give the CPU reference arm every core even under torchrun, and keeps the one-GPU settings for cpu_baseline."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_under_test", os.path.join(ROOT, "bench.py"))
    env_before = dict(os.environ)
    try:
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        nproc = str(mod.NPROC)
        e = {"WORLD_SIZE": "4", "OMP_NUM_THREADS": "1"}
        assert mod.configure_openmp(["bench.py", "--gpus", "4"], e) == (False, 4)
        assert e == {"WORLD_SIZE": "4", "OMP_NUM_THREADS": "1"}
        e = {"WORLD_SIZE": "4", "OMP_NUM_THREADS": "1"}
        assert mod.configure_openmp(["bench.py", "--impl", "reference", "--gpus", "4"], e) == (True, 4)
        assert e["OMP_NUM_THREADS"] == nproc and e["OMP_PROC_BIND"] == "close"
        e = {}
        assert mod.configure_openmp(["bench.py"], e) == (False, 1)
        assert e == {"OMP_NUM_THREADS": nproc, "OMP_PROC_BIND": "close"}
    finally:
        os.environ.clear(); os.environ.update(env_before)
