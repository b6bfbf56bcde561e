#!/usr/bin/env python
"""bench.py — throughput of the MPS hot path (Computer::ForwardTime) on B200, one JSON line on stdout.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]

Metric (BASELINE.json): particle-steps/s = particles x steps / time.  A "step" is one ForwardTime() (DetermineDt ->
neighbour grid -> densities -> explicit forces -> PPE assembly -> CG -> pressure gradient -> DS) over the whole particle
block.  Workload at N = 1: BASELINE.json configs[1], the 2-D Koshizuka & Oka dam break scaled to 1 008 104 particles
(l0 = 2.08e-4, time-step cap scaled with l0), synthetic, started from rest.

  value     device time: CUDA events on the solver's stream around exactly K steps, state resident in HBM
  e2e       the same K steps through the public C-ABI calls with HOST buffers: every step uploads the particle state
            from pinned host memory (mps_upload), steps (mps_forward_time_auto) and downloads it (mps_download)
  roofline  the dominant kernel (the persistent CG solve): algorithmic bytes = iterations x (12 nnz + 92 active rows)
             / CUDA-event time of that kernel, both summed over the timed steps, vs MEASURED_PEAKS.json
  cpu_baseline (N = 1, rank 0): the reference's own CPU code (oracle/_ref, OpenMP, all host cores) on ONE step of the
            same 1M-particle workload (~10-30 s)

--impl reference times the reference's CPU implementation (oracle/_ref if present, else the CPU restatement) on the same
workload, bounded in wall-clock time (cpu_baseline.sample says how many steps were timed).
N > 1: one process per GPU (torchrun).  The block is cut into N x-slabs of the cell-sorted slots (cell-column aligned, equal
modelled work): every rank computes the neighbour lists, gather stages, PPE rows and CG rows of its slab; the fields neighbours
read are copied over NVLink peer memory after each stage, and the preconditioned CG runs as ONE persistent kernel per rank
coupled over peer memory (rim rows and coarse-level halo cells read from the neighbour ranks' buffers, dot products exchanged
through mailboxes) -- csrc/mps_comm.cu, csrc/mps_cg.cu.
Default workload by N: the 2-D dam break with N x 1M particles (dambreak2d_1m / _2m / _4m / _8m: per-GPU work fixed, "scaling":
"weak"; N = 1 is BASELINE.json configs[1]).  --workload NAME runs that block on N GPUs (strong scaling).  `large` in the line is
BASELINE.json configs[3], the 12M-particle 3-D dam break, on the same N GPUs for a few steps: the strong-scaling curve of a block
that is large enough to shard (--no-large skips it).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

NPROC = os.cpu_count() or 1


def configure_openmp(argv, environ):
    """OpenMP settings of the CPU arms; must run before any library that brings libgomp is loaded.

    * `--impl reference`: the reference's CPU code uses every host core, also under torchrun (which exports
      OMP_NUM_THREADS=1 to its workers; only rank 0 computes in this arm), threads bound close.
    * our arm on one GPU: the same settings, for the `cpu_baseline` leg that runs in this process.
    * our arm on several GPUs: nothing is touched.  OMP_PROC_BIND makes libgomp bind the INITIAL thread of every process
      that loads it to the first core of the (shared) affinity mask: all ranks' host threads would time-share one core, and the
      lockstep multi-GPU step waits for the slowest host (measured: 69 instead of 51 ms/step on 4 GPUs)."""
    is_reference = any(a == "reference" and i > 0 and argv[i - 1] == "--impl" for i, a in enumerate(argv)) or "--impl=reference" in argv
    world = int(environ.get("WORLD_SIZE", "1"))
    if is_reference:
        environ["OMP_NUM_THREADS"] = str(NPROC)
        environ.setdefault("OMP_PROC_BIND", "close")
    elif world == 1:
        environ.setdefault("OMP_NUM_THREADS", str(NPROC))
        environ.setdefault("OMP_PROC_BIND", "close")
    return is_reference, world


configure_openmp(sys.argv, os.environ)

import numpy as np  # noqa: E402

from openmps_b200 import scenes  # noqa: E402

WORKLOADS = {
    # name: (factory, description)
    "dambreak2d_1m": (lambda: scenes.dambreak2d_fast(2.08e-4), "DamBreak 2D (Koshizuka&Oka 1996) l0=2.08e-4, 1008104 particles, from rest"),
    "dambreak2d_2m": (lambda: scenes.dambreak2d_fast(2.08e-4 / 2 ** 0.5), "DamBreak 2D l0=1.47e-4, 2003902 particles (2 x the 1M block), from rest"),
    "dambreak2d_4m": (lambda: scenes.dambreak2d_fast(2.08e-4 / 2), "DamBreak 2D l0=1.04e-4, 3987392 particles (4 x the 1M block), from rest"),
    "dambreak2d_8m": (lambda: scenes.dambreak2d_fast(2.08e-4 / 8 ** 0.5), "DamBreak 2D l0=7.35e-5, 7949974 particles (8 x the 1M block), from rest"),
    "dambreak2d_250k": (lambda: scenes.dambreak2d_fast(4.2e-4), "DamBreak 2D l0=4.2e-4"),
    "dambreak2d_72k": (lambda: scenes.dambreak2d_fast(8e-4), "DamBreak 2D l0=8e-4, 72667 particles"),
    "dambreak2d_default": (lambda: scenes.dambreak2d(), "DamBreak 2D default (Benchmark/Sample), 1323 particles"),
    "static_pressure": (lambda: scenes.static_pressure(), "StaticPressure 2D default, 6040 particles"),
    "central_gravity_4m": (lambda: scenes.central_gravity(half=1000, l0=5e-5), "CentralGravity 2D 2001^2 = 4004001 particles"),
    "dambreak3d_123k": (lambda: scenes.dambreak3d(8e-3), "DamBreak 3D l0=8e-3, 123147 particles"),
    "dambreak3d_1m": (lambda: scenes.dambreak3d(3.6e-3), "DamBreak 3D l0=3.6e-3"),
    "dambreak3d_10m": (lambda: scenes.dambreak3d(1.36e-3), "DamBreak 3D l0=1.36e-3, ~12.2M particles (~10M fluid)"),
    "dambreak3d_100m": (lambda: scenes.dambreak3d(6.5e-4), "DamBreak 3D l0=6.5e-4, ~101M particles (~91M fluid): BASELINE.json configs[4]"),
}
LARGE_WORKLOAD = "dambreak3d_10m"     # BASELINE.json configs[3]: the `large` sub-record of every line (strong scaling over --gpus)
LARGE_STEPS, LARGE_WARMUP = 3, 3


def default_workload(world):
    """BASELINE.json configs[1] on one GPU; N x that block on N GPUs (weak scaling: per-GPU work fixed)."""
    return {1: "dambreak2d_1m", 2: "dambreak2d_2m", 4: "dambreak2d_4m", 8: "dambreak2d_8m"}.get(world, "dambreak2d_1m")


def workload_config(name, sc, world, forced):
    """`config` of the JSON line: what is run, identical in both arms (--impl ours / reference) for the same command line."""
    D = sc.env.dim
    k = 57 if D == 3 else 21                 # matrix entries per row at r_e = 2.4 l0
    mb = sc.count * (10.0 * k + 40.0) / 1e6
    return {"workload": name, "description": WORKLOADS[name][1], "particles": sc.count, "dim": D, "l0": sc.env.l0,
            "r_e_by_l0": sc.env.r_e_by_l0, "eps": sc.env.eps, "gpus": world,
            "scaling_rule": ("one GPU" if world == 1 else
                             (f"strong: the {name} block on {world} GPUs" if forced else
                              f"weak: {world} x the 1M-particle block of BASELINE.json configs[1] on {world} GPUs (per-GPU work fixed)")),
            "start_state": "from rest, warm-up steps untimed, every timed leg starts from the same warmed state",
            "l2_policy": f"inputs larger than L2: every step rebuilds and then streams ~{mb:.0f} MB of matrix + vectors per matrix sweep vs 126 MB L2 "
                         f"(no explicit flush)" if mb > 126 * world else
                         f"matrix + vectors ~{mb:.0f} MB fit the L2 of {world} GPU(s): the sweeps of one solve re-read the same matrix by construction; "
                         f"every step rebuilds neighbour lists and matrix from moved particles, nothing is cached across steps"}


REFERENCE_SAMPLE = "dambreak2d_72k"   # fallback block when only the scalar CPU restatement is available
REFERENCE_BUDGET_S = 150              # wall-clock bound of the --impl reference loop


def read_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"   # B200_PROFILING.md


class ClockSampler:
    """SM clock and throttle reasons of one GPU DURING the timed region (B200_PROFILING.md clocks line).  In-process NVML
    (pynvml) polled by a thread every 250 ms; falls back to an `nvidia-smi -lms` child process if pynvml is unusable.
    Samples taken before mark() (warm-up) are dropped."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
    MASKS = (("sw_power_cap", 0x4), ("hw_slowdown", 0x8), ("sw_thermal_slowdown", 0x20), ("hw_thermal_slowdown", 0x40))

    def __init__(self, index, uuid=None):
        self.index = index
        self.uuid = uuid
        self.rows = []        # (sm_mhz, max_mhz, [reasons])
        self.proc = None
        self.nvml = None
        self.t_begin = None   # host time at which the timed region starts: earlier samples are dropped
        self.stop_flag = False
        self.how = None

    def mark(self):
        self.t_begin = time.perf_counter()

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            h = None
            if self.uuid:
                try:
                    h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + str(self.uuid)).encode())
                except Exception:
                    h = None
            if h is None:
                h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.nvml, self.handle, self.how = pynvml, h, "pynvml"
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "250"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.how = "nvidia-smi"
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _poll(self):
        n = self.nvml
        while not self.stop_flag:
            try:
                sm = n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)
                mx = n.nvmlDeviceGetMaxClockInfo(self.handle, n.NVML_CLOCK_SM)
                try:
                    bits = n.nvmlDeviceGetCurrentClocksEventReasons(self.handle)
                except Exception:
                    bits = n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
                if self.t_begin is not None and time.perf_counter() >= self.t_begin:
                    self.rows.append((float(sm), float(mx), [nm for nm, m in self.MASKS if bits & m]))
            except Exception:
                pass
            time.sleep(0.25)

    def _read(self):
        for line in self.proc.stdout:
            if self.t_begin is None or time.perf_counter() < self.t_begin:
                continue
            c = [x.strip() for x in line.split(",")]
            try:
                self.rows.append((float(c[0]), float(c[1]), [nm for nm, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[3:7])
                                                             if v.lower().startswith("active")]))
            except Exception:
                continue

    def stop(self):
        time.sleep(0.15)
        self.stop_flag = True
        if self.nvml is not None:
            self.thread.join(timeout=2)
            try:
                self.nvml.nvmlShutdown()
            except Exception:
                pass
        elif self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        if not self.rows:
            return None
        reasons = sorted({r for row in self.rows for r in row[2]})
        return {"sm_mhz": float(np.median([r[0] for r in self.rows])), "sm_max_mhz": float(max(r[1] for r in self.rows)), "reasons": reasons,
                "samples": len(self.rows), "how": self.how}


def dist_info():
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


def reference_arm(args, rank, world):
    """The reference's own CPU implementation of the path on the host cores (rank 0 only), on the SAME workload as our arm.
    Bounded: a 1M-particle ForwardTime() costs ~9 s on 16 cores, so the timed loop stops after K steps or REFERENCE_BUDGET_S
    seconds, whichever comes first; `steps` in the line is what was actually timed."""
    if rank != 0:
        return 0
    from oracle import bind
    name = args.workload
    sc = WORKLOADS[name][0]()
    config = workload_config(name, sc, world, args.workload_forced)   # the config our arm prints for the same command line
    note = ""
    if sc.count > CPU_STEP_BUDGET_PARTICLES:
        # bounded sample: the reference keeps a 13.8 kB/particle neighbour table (Computer.hpp:1763-1765) and needs minutes per step
        # on blocks of several million particles; its particle-steps/s falls with resolution (plain-CG iterations grow), so the
        # smaller block of the same family OVER-estimates what it would do on the full one
        small = "dambreak3d_1m" if sc.env.dim == 3 else "dambreak2d_1m"
        note = f"; bounded sample of {name} ({sc.count} particles): timed on {small}, an upper bound of the CPU's rate on the full block"
        name = small
        sc = WORKLOADS[name][0]()
    use_ref = bind.available(sc.env.dim, sc.env.central_gravity, fast=True)
    if not use_ref and sc.count > 200000:
        name = REFERENCE_SAMPLE          # the scalar restatement cannot do a 1M step in bounded time: coarser block, said in `sample`
        sc = WORKLOADS[name][0]()
    eng = bind.RefComputer.from_scene(sc, fast=True) if use_ref else bind.PortComputer.from_scene(sc)
    kind = "reference" if use_ref else "port"
    cores = min(NPROC, int(os.environ.get("OMP_NUM_THREADS", NPROC))) if use_ref else 1
    t_start = time.perf_counter()
    warm = 0
    while warm < args.warmup and time.perf_counter() - t_start < REFERENCE_BUDGET_S / 3:
        eng.forward(1); warm += 1
    t0 = time.perf_counter()
    steps = 0
    while steps < args.steps and (steps == 0 or time.perf_counter() - t_start < REFERENCE_BUDGET_S):
        eng.forward(1); steps += 1
    sec = time.perf_counter() - t0
    alive = int((eng.state()["type"] != 3).sum())
    value = alive * steps / sec
    sample = (f"{WORKLOADS[name][1]}: N={sc.count}, {steps} ForwardTime() steps after {warm} warm-up steps from rest "
              f"(asked {args.steps}/{args.warmup}; bounded to {REFERENCE_BUDGET_S} s), OMP_NUM_THREADS={os.environ['OMP_NUM_THREADS']}; " + CPU_BUILD_NOTE + note)
    line = {
        "impl": "reference", "metric": "particle-steps/sec", "value": value, "unit": "particle-steps/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": warm, "ms_per_step": 1e3 * sec / steps, "higher_is_better": True,
        "scaling": "weak" if (world == 1 or not args.workload_forced) else "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config,
        "detail": {"host": "CPU only (OpenMP)", "ran": name, "particles_alive": alive,
                   "ppe_solver": "plain CG (the reference's algorithm, Computer.hpp:1359-1429, ViennaCL host backend)"},
        "cpu_baseline": {"value": value, "unit": "particle-steps/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "particle-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


CPU_BUILD_NOTE = ("reference headers compiled here with -O3 -march=x86-64-v3 -fopenmp against the repo's Boost stand-in (oracle/boost_shim: "
                  "uBLAS-compatible sparse rows; Boost is not installed), not upstream's real uBLAS / -march=native: PPE insertion is <= 18 % of a CPU step")
CPU_STEP_BUDGET_PARTICLES = 2500000   # the largest block whose single CPU step stays within ~30 s on the box's cores


def cpu_baseline_one_step(sc, name, snap=None):
    """Reference CPU code, all host cores, on a BOUNDED sample of the same workload: ONE ForwardTime() from the warmed state the
    GPU legs are timed from (`snap`), ~10-30 s.  Blocks too large for that budget (3-D 12M: minutes per CPU step) are sampled on
    the largest smaller block of the same family, and the line says so (`extrapolated`)."""
    from oracle import bind
    extrapolated = None
    if sc.count > CPU_STEP_BUDGET_PARTICLES:
        small_name = "dambreak3d_1m" if sc.env.dim == 3 else "dambreak2d_1m"
        extrapolated = f"sampled on {small_name} instead of {name}: a CPU step of {sc.count} particles does not fit the ~30 s budget; plain-CG " \
                       f"iterations grow with resolution, so the CPU's particle-steps/s on {name} itself would be LOWER than this figure"
        sc, name, snap = WORKLOADS[small_name][0](), small_name, None
    use_ref = bind.available(sc.env.dim, sc.env.central_gravity, fast=True)
    if use_ref:
        eng = bind.RefComputer.from_scene(sc, fast=True); kind = "reference"
        cores = min(NPROC, int(os.environ.get("OMP_NUM_THREADS", NPROC)))
        steps = 1 if sc.count > 300000 else (5 if sc.count > 30000 else 50)
    else:
        # the scalar restatement is too slow for a 1M step: bounded sub-sample instead
        small = WORKLOADS[REFERENCE_SAMPLE][0]() if sc.count > 200000 else sc
        if small is not sc:
            snap = None
        eng = bind.PortComputer.from_scene(small); kind = "port"; cores = 1; steps = 1
        sc = small
    start = "from rest"
    if snap is not None and not (snap["type"] == 3).any():
        eng.set_state(x=snap["x"], u=snap["u"], p=snap["p"], n=snap["n"])
        start = "from the warmed state the GPU legs start from"
    t0 = time.perf_counter()
    eng.forward(steps)
    sec = time.perf_counter() - t0
    alive = int((eng.state()["type"] != 3).sum())
    out = {"value": alive * steps / sec, "unit": "particle-steps/s", "cores": cores, "kind": kind,
           "sample": f"{steps} ForwardTime() step(s) {start} on N={sc.count} ({name if kind == 'reference' else REFERENCE_SAMPLE}), "
                     f"{sec:.1f} s wall, OMP_NUM_THREADS={os.environ['OMP_NUM_THREADS']}; " + CPU_BUILD_NOTE}
    if extrapolated:
        out["extrapolated"] = extrapolated
    return out


def attach(gpu, dist, torch, capi, rank, world, local):
    """slab decomposition: rank 0's NCCL id reaches every rank through the launcher's own group"""
    uid = torch.zeros(128, dtype=torch.uint8, device=f"cuda:{local}")
    if rank == 0:
        uid.copy_(torch.tensor(list(capi.GpuComputer.comm_unique_id()), dtype=torch.uint8))
    dist.broadcast(uid, 0)
    gpu.attach_comm(rank, world, bytes(uid.cpu().numpy().tobytes()))


def stage_breakdown(gpu, steps=2):
    """per-stage device time of a step (CUDA events per stage, a sync per stage: a separate, untimed pass)"""
    gpu.set_stage_timing(True)
    gpu.reset_stats()
    gpu.forward(steps)
    st = gpu.stats_dict()
    gpu.set_stage_timing(False)
    return {k: round(v / steps, 4) for k, v in st["stage_ms"].items() if v}


def solver_text(st):
    return (f"CG preconditioned by Jacobi + one multigrid V(1,1) cycle on the cell hierarchy ({st['mg_levels']} levels, {st['mg_cells']} cells), "
            "reference stopping rule (Computer.hpp:1386,1408)" if st["mg_levels"] else "plain CG (the reference's algorithm, Computer.hpp:1359-1429)")


def large_record(args, torch, dist, capi, rank, world, local, barrier):
    """BASELINE.json configs[3] (3-D dam break, 12.2M particles) on the same GPUs, LARGE_STEPS steps: strong scaling of a block that
    is large enough to shard.  Device time, max over ranks."""
    sc = WORKLOADS[LARGE_WORKLOAD][0]()
    gpu = capi.GpuComputer.from_scene(sc, device=local)
    if world > 1:
        attach(gpu, dist, torch, capi, rank, world, local)
    gpu.forward(LARGE_WARMUP)
    gpu.reset_stats()
    barrier()
    ms = gpu.run_steps(LARGE_STEPS)
    barrier()
    st = gpu.stats_dict()
    alive = int((gpu.state()["type"] != 3).sum())
    t = torch.tensor([ms], dtype=torch.float64, device=f"cuda:{local}")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    stages = stage_breakdown(gpu) if world == 1 else None
    gpu.close()
    return {"workload": LARGE_WORKLOAD, "description": WORKLOADS[LARGE_WORKLOAD][1], "particles": sc.count, "n_gpus": world, "scaling": "strong",
            "steps": LARGE_STEPS, "warmup": LARGE_WARMUP, "ms_per_step": ms / LARGE_STEPS, "value": alive * LARGE_STEPS / (ms * 1e-3),
            "unit": "particle-steps/s", "cg_iterations_per_step": st["cg_iterations"] / LARGE_STEPS, "cg_ms_per_step": st["cg_ms"] / LARGE_STEPS,
            "nnz_this_rank": st["nnz"], "ppe_solver": solver_text(st), "stage_ms_per_step": stages}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=None, choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-large", action="store_true")
    args = ap.parse_args()
    rank, world, local = dist_info()
    args.workload_forced = args.workload is not None
    if args.workload is None:
        args.workload = default_workload(world)
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    if args.impl == "reference":
        return reference_arm(args, rank, world)

    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU path (use --impl reference for the CPU arm)")
    from openmps_b200 import capi

    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    sc = WORKLOADS[args.workload][0]()
    n = sc.count
    D = sc.env.dim
    strong = world > 1 and args.workload_forced
    gpu = capi.GpuComputer.from_scene(sc, device=local)
    if world > 1:
        attach(gpu, dist, torch, capi, rank, world, local)

    # clocks / throttle reasons of rank 0's GPU, sampled DURING the timed region.  The poller is started before the warm-up:
    # initialising NVML touches every GPU of the box and stalls running CUDA work for tens of milliseconds — invisible
    # behind one long kernel, but the multi-GPU step runs in lockstep with host syncs.  One poller (rank 0) only.
    sampler = None
    if rank == 0:
        try:
            uuid = torch.cuda.get_device_properties(local).uuid
        except Exception:
            uuid = None
        sampler = ClockSampler(local, uuid)
        sampler.start()
    # ---- warm-up (untimed), then exactly K steps timed on the device ----
    gpu.forward(args.warmup)
    # both timed legs start from this state (the CG iteration count depends on the state: same state, same work)
    snap = gpu.state()
    snap_t = gpu.time()
    gpu.reset_stats()
    barrier()
    if sampler:
        sampler.mark()
    dev_ms = gpu.run_steps(args.steps)
    barrier()
    clocks = sampler.stop() if sampler else None
    st = gpu.stats_dict()
    alive = int((gpu.state()["type"] != 3).sum())

    t = torch.tensor([dev_ms], dtype=torch.float64, device=f"cuda:{local}")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms_max = float(t.item())
    value = alive * args.steps / (dev_ms_max * 1e-3)   # whole job: the ranks share ONE block (N x 1M particles by default)
    comm = gpu.comm_info() if world > 1 else None

    # ---- end to end through the public C ABI with host buffers (pinned), copies inside the timed region ----
    e2e = None
    if not args.no_e2e:
        hx = torch.empty((n, D), dtype=torch.float64).pin_memory().numpy()
        hu = torch.empty((n, D), dtype=torch.float64).pin_memory().numpy()
        hp = torch.empty(n, dtype=torch.float64).pin_memory().numpy()
        hn = torch.empty(n, dtype=torch.float64).pin_memory().numpy()
        ht = torch.empty(n, dtype=torch.int32).pin_memory().numpy()
        hx[:] = snap["x"]; hu[:] = snap["u"]; hp[:] = snap["p"]; hn[:] = snap["n"]
        gpu.set_time(*snap_t)
        h2d = n * (2 * D + 2) * 8
        d2h = n * (2 * D + 2) * 8 + n * 4
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            gpu.set_state(x=hx, u=hu, p=hp, n=hn)     # host -> device: this step's input state
            gpu.forward(1)                            # ForwardTime()
            gpu.download_into(hx, hu, hp, hn, ht)     # device -> host: the step's result (what Particles() returns)
        barrier()
        sec = time.perf_counter() - t0
        te = torch.tensor([sec], dtype=torch.float64, device=f"cuda:{local}")
        if dist is not None:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e2e = {"value": alive * args.steps / float(te.item()), "unit": "particle-steps/s",
               "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": 1e3 * float(te.item()) / args.steps,
               "timing": "wall clock around the synchronous C-ABI calls (mps_upload, mps_forward_time_auto, mps_download), max over ranks; "
                         "same start state and time as the device-timed leg" + ("; per rank (the state is replicated)" if world > 1 else "")}

    # per-stage device time and the FP64-pipe share of the gather stages (an extra, untimed pass; one GPU)
    stages = stage_breakdown(gpu) if world == 1 else None
    gpu.close()

    large = None
    if not args.no_large and not args.workload_forced:
        try:
            large = large_record(args, torch, dist, capi, rank, world, local, barrier)
        except Exception as ex:   # reported, never allowed to break the bench line
            large = {"workload": LARGE_WORKLOAD, "error": repr(ex)}

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return 0

    peak, peak_src = read_peaks()
    cg_ms, cg_bytes = st["cg_ms"], st["cg_bytes"]
    traffic, traffic_src = None, None
    try:
        # DRAM bytes of the kernel from the committed `ncu --set full` capture, per CG iteration, scaled to this run's
        # iterations per launch (the solve is ONE launch whose length is the iteration count)
        with open(os.path.join(ROOT, "profiles", "cg_traffic.json")) as f:
            tj = json.load(f)
        if tj.get("workload") == args.workload and bool(tj.get("preconditioned")) == bool(st["mg_levels"]) and world == 1:
            traffic = tj["dram_bytes_per_iteration"] * st["cg_iterations"] / max(args.steps, 1)
            traffic_src = tj["source"]
    except Exception:
        pass
    achieved = (cg_bytes / (cg_ms * 1e-3)) / 1e9 if cg_ms > 0 else 0.0
    iters = st["cg_iterations"]
    pcg = bool(st["mg_levels"])
    roofline = {
        "kernel": ("k_pcg_stream (persistent cooperative preconditioned CG, one launch per step: chunk blobs streamed through shared memory by bulk async "
                   "copies, SpMV + dots + vector updates + the multigrid V-cycle on the cell hierarchy fused)" if pcg else
                   "k_cg_stream (persistent cooperative CG, one launch per step: chunk blobs streamed through shared memory by bulk async copies, "
                   "SpMV + dots + vector updates fused)"),
        "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak if peak else None,
        "peak_source": f"MEASURED_PEAKS.json hbm_gbs ({peak_src}); nominal 8000 GB/s -> frac {achieved / 8000.0:.3f}",
        "traffic": traffic, "traffic_source": traffic_src,
        "dram_frac": (traffic / (cg_ms / max(args.steps, 1) * 1e-3) / 1e9 / peak) if (traffic and cg_ms > 0 and peak) else None,
        "algorithmic_bytes_per_launch": cg_bytes / max(args.steps, 1),
        "bytes_model": "matrix sweeps actually made x (12 nnz + 92 active_rows) per launch (SURVEY.md 8d: one sweep per CG iteration); the kernel's own "
                       "layout moves ~10 B/nnz from HBM and keeps the vectors L2-resident" +
                       ("; the preconditioner's own traffic (cell-level stencils, a second pass over the row vectors) is NOT counted as useful bytes, "
                        "so its time lowers this fraction" if pcg else ""),
        "launches": args.steps, "cg_iterations": iters, "cg_iterations_per_s": iters / (cg_ms * 1e-3) if cg_ms > 0 else None,
        "kernel_ms_per_launch": cg_ms / max(args.steps, 1), "kernel_share_of_step": cg_ms / dev_ms if dev_ms > 0 else None,
        "nnz": st["nnz"], "active_rows": st["active_rows"],
    }
    cpu = None
    if not args.no_cpu_baseline and world == 1:
        try:
            cpu = cpu_baseline_one_step(sc, args.workload, snap)
        except Exception as ex:  # the baseline is reported, never allowed to break the bench line
            cpu = {"value": None, "unit": "particle-steps/s", "cores": NPROC, "kind": "unavailable", "sample": repr(ex)}

    detail = {"particles_alive": alive,
              "parallelism": "1 GPU" if world == 1 else f"{world} x-slabs of the cell-sorted slots (cell-column aligned, equal modelled work), one per GPU; "
                             f"CG coupling: {comm['mode']}; {st['comm_calls']} NCCL calls in the timed region",
              "cg_iterations_per_step": iters / max(args.steps, 1),
              "matrix_sweeps_per_step": st["matrix_sweeps"] / max(args.steps, 1),
              "ppe_solver": solver_text(st),
              "stage_ms_per_step": stages,
              "fp64": "gather stages keep true IEEE sqrt and division per in-range pair (parity <= 4 ulp with the reference); measured FP64 peaks of this "
                      "GPU (tools/fp64_peak.cu, profiles/r02_fp64_peak.json): 18.55 T DFMA/s, 0.651 T sqrt+div pairs/s"}
    line = {
        "metric": "particle-steps/sec", "value": value, "unit": "particle-steps/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dev_ms_max / args.steps, "higher_is_better": True,
        "scaling": "strong" if strong else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args.workload, sc, world, args.workload_forced),
        "detail": detail,
        "e2e": e2e, "gpu_launches": int(st["kernel_launches"]),
        "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu, "large": large,
    }
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
