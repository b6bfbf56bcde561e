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
N > 1: one process per GPU (torchrun); the SAME block is cut into N x-slabs of the cell-sorted slots (strong scaling):
every rank computes the neighbour lists, gather stages, PPE rows and CG rows of its slab; NCCL all-gathers the fields
neighbours read after each stage, and the CG iteration runs as ONE persistent kernel per rank coupled over NVLink peer
memory (rim rows pushed by P2P stores, dot products exchanged through mailboxes) -- csrc/mps_comm.cu, csrc/mps_cg.cu.
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
    "dambreak2d_250k": (lambda: scenes.dambreak2d_fast(4.2e-4), "DamBreak 2D l0=4.2e-4"),
    "dambreak2d_72k": (lambda: scenes.dambreak2d_fast(8e-4), "DamBreak 2D l0=8e-4, 72667 particles"),
    "dambreak2d_default": (lambda: scenes.dambreak2d(), "DamBreak 2D default (Benchmark/Sample), 1323 particles"),
    "static_pressure": (lambda: scenes.static_pressure(), "StaticPressure 2D default, 6040 particles"),
    "central_gravity_4m": (lambda: scenes.central_gravity(half=1000, l0=5e-5), "CentralGravity 2D 2001^2 = 4004001 particles"),
    "dambreak3d_123k": (lambda: scenes.dambreak3d(8e-3), "DamBreak 3D l0=8e-3, 123147 particles"),
    "dambreak3d_1m": (lambda: scenes.dambreak3d(3.6e-3), "DamBreak 3D l0=3.6e-3"),
    "dambreak3d_10m": (lambda: scenes.dambreak3d(1.36e-3), "DamBreak 3D l0=1.36e-3, ~12.2M particles (~10M fluid)"),
}
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
              f"(asked {args.steps}/{args.warmup}; bounded to {REFERENCE_BUDGET_S} s), OMP_NUM_THREADS={os.environ['OMP_NUM_THREADS']}")
    line = {
        "impl": "reference", "metric": "particle-steps/sec", "value": value, "unit": "particle-steps/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": warm, "ms_per_step": 1e3 * sec / steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": name, "particles": sc.count, "description": WORKLOADS[name][1], "host": "CPU only (OpenMP)"},
        "cpu_baseline": {"value": value, "unit": "particle-steps/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "particle-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


def cpu_baseline_one_step(sc, name):
    """Reference CPU code, all host cores, ONE step of the same workload from rest (bounded: ~10-30 s at 1M particles)."""
    from oracle import bind
    use_ref = bind.available(sc.env.dim, sc.env.central_gravity, fast=True)
    if use_ref:
        eng = bind.RefComputer.from_scene(sc, fast=True); kind = "reference"
        cores = min(NPROC, int(os.environ.get("OMP_NUM_THREADS", NPROC)))
        steps = 1 if sc.count > 300000 else (5 if sc.count > 30000 else 50)
    else:
        # the scalar restatement is too slow for a 1M step: bounded sub-sample instead
        small = WORKLOADS[REFERENCE_SAMPLE][0]() if sc.count > 200000 else sc
        eng = bind.PortComputer.from_scene(small); kind = "port"; cores = 1; steps = 1
        sc = small
    t0 = time.perf_counter()
    eng.forward(steps)
    sec = time.perf_counter() - t0
    alive = int((eng.state()["type"] != 3).sum())
    return {"value": alive * steps / sec, "unit": "particle-steps/s", "cores": cores, "kind": kind,
            "sample": f"{steps} ForwardTime() step(s) from rest on N={sc.count} ({name if kind == 'reference' else REFERENCE_SAMPLE}), "
                      f"{sec:.1f} s wall, OMP_NUM_THREADS={os.environ['OMP_NUM_THREADS']}"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=None, choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    args.workload_forced = args.workload is not None
    if args.workload is None:
        args.workload = "dambreak2d_1m"
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    rank, world, local = dist_info()

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
    gpu = capi.GpuComputer.from_scene(sc, device=local)
    if world > 1:
        # slab decomposition: rank 0's NCCL id reaches every rank through the launcher's own group
        uid = torch.zeros(128, dtype=torch.uint8, device=f"cuda:{local}")
        if rank == 0:
            uid.copy_(torch.tensor(list(capi.GpuComputer.comm_unique_id()), dtype=torch.uint8))
        dist.broadcast(uid, 0)
        gpu.attach_comm(rank, world, bytes(uid.cpu().numpy().tobytes()))

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
    value = alive * args.steps / (dev_ms_max * 1e-3)   # N > 1 is strong scaling: the ranks share ONE block
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
        if tj.get("workload") == args.workload:
            traffic = tj["dram_bytes_per_iteration"] * st["cg_iterations"] / max(args.steps, 1)
            traffic_src = tj["source"]
    except Exception:
        pass
    achieved = (cg_bytes / (cg_ms * 1e-3)) / 1e9 if cg_ms > 0 else 0.0
    iters = st["cg_iterations"]
    roofline = {
        "kernel": "k_cg_stream (persistent cooperative CG, one launch per step: chunk blobs streamed through shared memory by bulk async copies, SpMV + dots + vector updates fused)",
        "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak if peak else None,
        "peak_source": f"MEASURED_PEAKS.json hbm_gbs ({peak_src}); nominal 8000 GB/s -> frac {achieved / 8000.0:.3f}",
        "traffic": traffic, "traffic_source": traffic_src,
        "algorithmic_bytes_per_launch": cg_bytes / max(args.steps, 1),
        "bytes_model": "iterations x (12 nnz + 92 active_rows) per launch (SURVEY.md 8d); the kernel's own layout moves ~10 B/nnz from HBM and keeps the vectors L2-resident",
        "launches": args.steps, "cg_iterations": iters, "cg_iterations_per_s": iters / (cg_ms * 1e-3) if cg_ms > 0 else None,
        "kernel_ms_per_launch": cg_ms / max(args.steps, 1), "kernel_share_of_step": cg_ms / dev_ms if dev_ms > 0 else None,
        "nnz": st["nnz"], "active_rows": st["active_rows"],
    }
    cpu = None
    if not args.no_cpu_baseline and world == 1:
        try:
            cpu = cpu_baseline_one_step(sc, args.workload)
        except Exception as ex:  # the baseline is reported, never allowed to break the bench line
            cpu = {"value": None, "unit": "particle-steps/s", "cores": NPROC, "kind": "unavailable", "sample": repr(ex)}

    mat_mb = (10.0 * st["nnz"] + 40.0 * n) / 1e6
    l2_policy = (f"inputs larger than L2: matrix blobs + vectors ~{mat_mb:.0f} MB per step vs 126 MB L2 (no explicit flush)" if mat_mb > 126 else
                 f"per-rank matrix slab + vectors ~{mat_mb:.0f} MB fit the 126 MB L2: the CG iterations of one step re-read the same matrix by construction "
                 "(the reuse is the algorithm's); every step rebuilds neighbour lists and matrix from moved particles, nothing is cached across steps")
    line = {
        "metric": "particle-steps/sec", "value": value, "unit": "particle-steps/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dev_ms_max / args.steps, "higher_is_better": True,
        "scaling": "weak" if world == 1 else "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": args.workload, "description": WORKLOADS[args.workload][1], "particles": n, "dim": D,
                   "particles_alive": alive, "l0": sc.env.l0, "r_e_by_l0": sc.env.r_e_by_l0, "eps": sc.env.eps,
                   "parallelism": "1 GPU" if world == 1 else f"{world} x-slabs of the cell-sorted slots, one per GPU (CG coupling: {comm['mode']}; "
                                  f"{st['comm_calls']} NCCL calls in the timed region)",
                   "l2_policy": l2_policy,
                   "cg_iterations_per_step": iters / max(args.steps, 1),
                   "matrix_sweeps_per_step": st["matrix_sweeps"] / max(args.steps, 1),
                   "ppe_solver": (f"CG preconditioned by Jacobi + one multigrid V(1,1) cycle on the cell hierarchy ({st['mg_levels']} levels, "
                                  f"{st['mg_cells']} cells), reference stopping rule" if st["mg_levels"] else
                                  "plain CG (the reference's algorithm, Computer.hpp:1359-1429)")},
        "e2e": e2e, "gpu_launches": int(st["kernel_launches"]),
        "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
    }
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
