"""ctypes binding of libopenmps_b200.so — the product library (hand-written CUDA for sm_100a behind include/mps_capi.h).

``GpuComputer`` mirrors the stage-level surface of the reference's ``Computer`` (Computer.hpp:437-460 friend access +
public part :1659-1790), so the parity tests read like the upstream gtests.  There is NO CPU fallback: if the shared
library is missing or no CUDA device is present, construction raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libopenmps_b200.so")

MPS_OK, MPS_CG_NOT_CONVERGED, MPS_CELL_OVERFLOW, MPS_CUDA_ERROR, MPS_NCCL_ERROR, MPS_BAD_ARG = range(6)

VEC = {"x": 0, "b": 1, "r": 2, "p": 3, "Ap": 4, "ecs": 5, "nWithoutSpp": 6, "du": 7, "originalX": 8}


class MpsEnv(C.Structure):
    _fields_ = [("dim", C.c_int32), ("central_gravity", C.c_int32), ("max_dt", C.c_double), ("courant", C.c_double),
                ("g", C.c_double), ("rho", C.c_double), ("nu", C.c_double), ("r_e_by_l0", C.c_double), ("l0", C.c_double),
                ("min_x", C.c_double * 3), ("max_x", C.c_double * 3)]


class MpsEnvInfo(C.Structure):
    _fields_ = [("t", C.c_double), ("dt", C.c_double), ("n0", C.c_double), ("max_dt", C.c_double), ("max_dx", C.c_double),
                ("r_e", C.c_double), ("neighbor_length", C.c_double), ("l0", C.c_double), ("rho", C.c_double), ("nu", C.c_double),
                ("grid_cells", C.c_int64 * 3), ("cell_capacity", C.c_uint64)]


class MpsStats(C.Structure):
    _fields_ = [("steps", C.c_uint64), ("cg_iterations", C.c_uint64), ("last_cg_iterations", C.c_uint64),
                ("last_rr0", C.c_double), ("last_rr", C.c_double),
                ("particles", C.c_uint64), ("neighbors", C.c_uint64), ("nnz", C.c_uint64), ("active_rows", C.c_uint64),
                ("kernel_launches", C.c_uint64), ("disabled_last", C.c_uint64), ("comm_calls", C.c_uint64), ("cg_ms", C.c_double), ("cg_bytes", C.c_double), ("stage_ms", C.c_double * 16), ("stage_calls", C.c_uint64 * 16),
                ("matrix_sweeps", C.c_uint64), ("mg_levels", C.c_uint64), ("mg_cells", C.c_uint64)]


_OBS_FIELDS = ("edge_x", "top_z", "r_max_surface", "r_min_surface", "center_r2", "center_id", "center_p", "inner", "sum_d", "sum_p", "sum_dd",
               "sum_dp", "max_dev", "h1", "h2", "p2_sum", "p2_count", "fluid", "wall", "dummy", "disabled", "p_max", "u_max2", "surface_count")


class MpsObserveParams(C.Structure):
    _fields_ = [(k, C.c_double) for k in ("surface_n", "rho_g", "surface_z", "x_h1", "x_h2", "min_n", "z_p2", "d")]


class MpsObservables(C.Structure):
    _fields_ = [(k, C.c_double) for k in _OBS_FIELDS]


class MpsWallMotion(C.Structure):
    _fields_ = [("amplitude", C.c_double * 3), ("velocity", C.c_double * 3), ("omega", C.c_double), ("phase", C.c_double),
                ("t_begin", C.c_double), ("t_end", C.c_double)]


class MpsError(RuntimeError):
    """Carries the mps_status; codes 1 / 2 correspond to Computer::Exception / Grid::Exception of the reference."""

    def __init__(self, code, msg):
        super().__init__(f"[mps_status {code}] {msg}")
        self.code = code
        self.message = msg


_lib = None


def load_library():
    """Loads libopenmps_b200.so; raises (never falls back) when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise FileNotFoundError(f"{LIB_PATH} is missing: build it with `make -C openmps_b200/csrc` "
                                f"(or __graft_entry__.build()); there is no CPU fallback")
    lib = C.CDLL(LIB_PATH)
    vp, u64, dbl = C.c_void_p, C.c_uint64, C.c_double
    pd = C.POINTER(C.c_double)

    def sig(name, args, res=C.c_int):
        f = getattr(lib, name)
        f.argtypes = args
        f.restype = res

    sig("mps_create", [C.POINTER(MpsEnv), dbl, C.c_int, C.POINTER(vp)])
    sig("mps_destroy", [vp])
    sig("mps_last_error", [vp], C.c_char_p)
    sig("mps_get_env_info", [vp, C.POINTER(MpsEnvInfo)])
    sig("mps_add_particles", [vp, u64, vp, vp, vp, vp, vp])
    sig("mps_count", [vp], u64)
    sig("mps_download", [vp, vp, vp, vp, vp, vp])
    sig("mps_upload", [vp, vp, vp, vp, vp])
    sig("mps_set_wall_positions", [vp, u64, vp, vp])
    sig("mps_set_wall_motion", [vp, u64, vp, C.POINTER(MpsWallMotion)])
    sig("mps_determine_dt", [vp, pd])
    sig("mps_forward_time", [vp, dbl])
    sig("mps_forward_time_auto", [vp])
    sig("mps_run_until", [vp, dbl, C.POINTER(u64)])
    sig("mps_run_steps", [vp, u64, pd])
    sig("mps_get_time", [vp, pd, pd])
    sig("mps_set_dt", [vp, dbl, C.c_int])
    sig("mps_set_time", [vp, dbl, dbl])
    for st in ("search_neighbor", "compute_density", "error_correction", "explicit_forces", "save_x", "set_ppe", "solve_ppe",
               "assign_pressure", "implicit_forces", "pressure_gradient", "dynamic_stabilize"):
        sig("mps_" + st, [vp])
    sig("mps_dndt", [vp, u64, pd])
    sig("mps_get_cells", [vp, vp])
    sig("mps_get_neighbors", [vp, vp, vp])
    sig("mps_get_csr_nnz", [vp, C.POINTER(u64)])
    sig("mps_get_csr", [vp, vp, vp, vp])
    sig("mps_get_vec", [vp, C.c_int, vp])
    sig("mps_set_system", [vp, u64, vp, vp, vp, vp, vp])
    sig("mps_get_solution", [vp, u64, vp])
    sig("mps_observe", [vp, C.POINTER(MpsObserveParams), C.POINTER(MpsObservables)])
    sig("mps_set_stage_timing", [vp, C.c_int])
    sig("mps_get_stats", [vp, C.POINTER(MpsStats)])
    sig("mps_reset_stats", [vp])
    sig("mps_stage_name", [C.c_int], C.c_char_p)
    sig("mps_time_kernel", [vp, C.c_char_p, C.c_int, pd, pd])
    sig("mps_flush_l2", [vp])
    sig("mps_comm_unique_id", [vp])
    sig("mps_comm_init", [vp, C.c_int, C.c_int, vp])
    sig("mps_comm_info", [vp, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(u64), C.POINTER(u64)])
    sig("mps_comm_mode", [vp, C.POINTER(C.c_int)])
    sig("mps_partition_range", [u64, C.c_int, C.c_int, C.POINTER(u64), C.POINTER(u64)])
    sig("mps_debug_mg", [vp, C.c_int, C.c_int, vp, u64, C.POINTER(u64)])
    sig("mps_set_cg_profile", [vp, C.c_int])
    sig("mps_get_cg_profile", [vp, pd])
    sig("mps_get_cg_profile_raw", [vp, vp, u64, C.POINTER(u64)])
    sig("mps_get_cg_profile_stages", [vp, vp])
    _lib = lib
    return lib


def partition_range(n, nranks, rank):
    """Slots [first, last) that `rank` of `nranks` computes for n particles (pure host arithmetic of the slab decomposition)."""
    lib = load_library()
    a, b = C.c_uint64(), C.c_uint64()
    rc = lib.mps_partition_range(int(n), int(nranks), int(rank), C.byref(a), C.byref(b))
    if rc != MPS_OK:
        raise MpsError(rc, "bad partition arguments")
    return a.value, b.value


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f64(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float64)


_STAGES = {
    "search": "mps_search_neighbor", "density": "mps_compute_density", "ecs": "mps_error_correction",
    "explicit": "mps_explicit_forces", "savex": "mps_save_x", "setppe": "mps_set_ppe", "solveppe": "mps_solve_ppe",
    "pressure": "mps_assign_pressure", "implicit": "mps_implicit_forces", "gradient": "mps_pressure_gradient",
    "ds": "mps_dynamic_stabilize",
}


class GpuComputer:
    """One MPS solver on one B200 (the drop-in for the reference's ``Computer`` on the Python side)."""
    kind = "gpu"

    def __init__(self, env, device=0):
        self.lib = load_library()
        self.env = env
        self.dim = env.dim
        e = MpsEnv()
        e.dim = env.dim; e.central_gravity = int(env.central_gravity)
        e.max_dt = env.max_dt; e.courant = env.courant; e.g = env.g; e.rho = env.rho; e.nu = env.nu
        e.r_e_by_l0 = env.r_e_by_l0; e.l0 = env.l0
        for k in range(env.dim):
            e.min_x[k] = env.min_x[k]; e.max_x[k] = env.max_x[k]
        h = C.c_void_p()
        rc = self.lib.mps_create(C.byref(e), env.eps, device, C.byref(h))
        if rc != MPS_OK:
            raise MpsError(rc, self.lib.mps_last_error(None).decode())
        self.h = h

    @classmethod
    def from_scene(cls, scene, device=0):
        c = cls(scene.env, device)
        c.add_particles(scene.x, scene.u, scene.p, scene.n, scene.type)
        return c

    def close(self):
        if getattr(self, "h", None):
            self.lib.mps_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != MPS_OK:
            raise MpsError(rc, self.lib.mps_last_error(self.h).decode())

    # ---- particles ----
    def add_particles(self, x, u, p, n, type):
        x, u, p, n = _f64(x), _f64(u), _f64(p), _f64(n)
        t = np.ascontiguousarray(type, dtype=np.int32)
        self._check(self.lib.mps_add_particles(self.h, len(t), _ptr(x), _ptr(u), _ptr(p), _ptr(n), _ptr(t)))

    @property
    def count(self):
        return int(self.lib.mps_count(self.h))

    def state(self):
        n, d = self.count, self.dim
        x = np.empty((n, d)); u = np.empty((n, d)); p = np.empty(n); nd = np.empty(n); t = np.empty(n, np.int32)
        self._check(self.lib.mps_download(self.h, _ptr(x), _ptr(u), _ptr(p), _ptr(nd), _ptr(t)))
        return {"x": x, "u": u, "p": p, "n": nd, "type": t}

    def download_into(self, x=None, u=None, p=None, n=None, type=None):
        """Device -> caller-provided (e.g. pinned) host buffers; used by the end-to-end timing."""
        self._check(self.lib.mps_download(self.h, _ptr(x), _ptr(u), _ptr(p), _ptr(n), _ptr(type)))

    def set_state(self, x=None, u=None, p=None, n=None):
        x, u, p, n = _f64(x), _f64(u), _f64(p), _f64(n)
        self._check(self.lib.mps_upload(self.h, _ptr(x), _ptr(u), _ptr(p), _ptr(n)))

    def set_wall_positions(self, ids, x):
        ids = np.ascontiguousarray(ids, np.uint64); x = _f64(x)
        self._check(self.lib.mps_set_wall_positions(self.h, len(ids), _ptr(ids), _ptr(x)))

    def set_wall_motion(self, ids=None, amplitude=(0, 0, 0), velocity=(0, 0, 0), omega=0.0, phase=0.0, t_begin=0.0, t_end=float("inf"), clear=False):
        """mps_set_wall_motion: the listed (default: all non-fluid) particles follow an analytic motion evaluated on the device."""
        if clear:
            self._check(self.lib.mps_set_wall_motion(self.h, 0, None, None))
            return
        pad = lambda v: tuple(float(c) for c in v) + (0.0,) * (3 - len(v))  # noqa: E731
        m = MpsWallMotion((C.c_double * 3)(*pad(amplitude)), (C.c_double * 3)(*pad(velocity)), omega, phase, t_begin, t_end)
        if ids is None:
            self._check(self.lib.mps_set_wall_motion(self.h, 0, None, C.byref(m)))
        else:
            ids = np.ascontiguousarray(ids, np.uint64)
            self._check(self.lib.mps_set_wall_motion(self.h, len(ids), _ptr(ids), C.byref(m)))

    # ---- environment / time ----
    def env_info(self):
        info = MpsEnvInfo()
        self._check(self.lib.mps_get_env_info(self.h, C.byref(info)))
        return info

    def env_values(self):
        i = self.env_info()
        return {"t": i.t, "dt": i.dt, "n0": i.n0, "MaxDt": i.max_dt, "MaxDx": i.max_dx, "R_e": i.r_e,
                "NeighborLength": i.neighbor_length, "L_0": i.l0, "Rho": i.rho, "Nu": i.nu}

    def grid_capacity(self):
        return int(self.env_info().cell_capacity)

    def set_dt(self, dt, advance=True):
        self._check(self.lib.mps_set_dt(self.h, dt, int(advance)))

    def determine_dt(self):
        dt = C.c_double()
        self._check(self.lib.mps_determine_dt(self.h, C.byref(dt)))
        return dt.value

    def set_time(self, t, dt):
        self._check(self.lib.mps_set_time(self.h, float(t), float(dt)))

    def time(self):
        t, dt = C.c_double(), C.c_double()
        self._check(self.lib.mps_get_time(self.h, C.byref(t), C.byref(dt)))
        return t.value, dt.value

    # ---- stepping ----
    def stage(self, name):
        self._check(getattr(self.lib, _STAGES[name])(self.h))

    def forward(self, steps=1, dt=None):
        for _ in range(steps):
            self._check(self.lib.mps_forward_time_auto(self.h) if dt is None else self.lib.mps_forward_time(self.h, dt))
        return steps

    def run_steps(self, steps):
        """``steps`` x ForwardTime(); returns device milliseconds measured with CUDA events on the solver's stream."""
        ms = C.c_double()
        self._check(self.lib.mps_run_steps(self.h, steps, C.byref(ms)))
        return ms.value

    def run_until(self, t_end):
        k = C.c_uint64()
        rc = self.lib.mps_run_until(self.h, t_end, C.byref(k))
        self._check(rc)
        return k.value

    # ---- inspection ----
    def cells(self):
        out = np.empty((self.count, self.dim), np.int64)
        self._check(self.lib.mps_get_cells(self.h, _ptr(out)))
        return out

    def neighbors(self):
        n = self.count
        rowptr = np.empty(n + 1, np.uint64)
        self._check(self.lib.mps_get_neighbors(self.h, _ptr(rowptr), None))
        idx = np.empty(int(rowptr[-1]), np.uint64)
        self._check(self.lib.mps_get_neighbors(self.h, _ptr(rowptr), _ptr(idx)))
        return rowptr, idx

    def neighbor_counts(self):
        """Neighbour-list lengths per particle (original order) without downloading the lists."""
        rowptr = np.empty(self.count + 1, np.uint64)
        self._check(self.lib.mps_get_neighbors(self.h, _ptr(rowptr), None))
        return np.diff(rowptr).astype(np.int64)

    def csr(self):
        nnz = C.c_uint64()
        self._check(self.lib.mps_get_csr_nnz(self.h, C.byref(nnz)))
        n = self.count
        rowptr = np.empty(n + 1, np.uint64); col = np.empty(nnz.value, np.uint32); val = np.empty(nnz.value)
        self._check(self.lib.mps_get_csr(self.h, _ptr(rowptr), _ptr(col), _ptr(val)))
        return rowptr, col, val

    def vec(self, name, n=None):
        which = VEC[name]
        n = self.count if n is None else n
        out = np.empty((n, self.dim)) if which >= 7 else np.empty(n)
        self._check(self.lib.mps_get_vec(self.h, which, _ptr(out)))
        return out

    def dndt(self, i):
        out = C.c_double()
        self._check(self.lib.mps_dndt(self.h, int(i), C.byref(out)))
        return out.value

    def set_system(self, rowptr, col, val, b, x0):
        rowptr = np.ascontiguousarray(rowptr, np.uint64); col = np.ascontiguousarray(col, np.uint32)
        val, b, x0 = _f64(val), _f64(b), _f64(x0)
        self._n_sys = len(b)
        self._check(self.lib.mps_set_system(self.h, len(b), _ptr(rowptr), _ptr(col), _ptr(val), _ptr(b), _ptr(x0)))

    def solution(self):
        x = np.empty(self._n_sys)
        self._check(self.lib.mps_get_solution(self.h, self._n_sys, _ptr(x)))
        return x

    def last_iterations(self):
        return int(self.stats().last_cg_iterations)

    def mg_table(self, level, which):
        """One table of the preconditioner's cell hierarchy (see mps_debug_mg)."""
        n = C.c_uint64(0)
        self._check(self.lib.mps_debug_mg(self.h, level, which, None, 0, C.byref(n)))
        if level < 0:
            dt = {0: np.uint32, 1: np.uint64, 2: np.float64, 3: np.uint32}[which]
        else:
            dt = np.uint32 if which <= 3 else np.float64
        out = np.zeros(n.value, dtype=dt)
        if n.value:
            self._check(self.lib.mps_debug_mg(self.h, level, which, _ptr(out), out.nbytes, C.byref(n)))
        return out

    # ---- measurement ----
    def observe(self, surface_n=0.0, rho_g=0.0, surface_z=0.0, x_h1=0.0, x_h2=0.0, min_n=float("inf"), z_p2=0.0, d=0.0):
        """mps_observe: the benchmark observables of the resident state, reduced on the device (see observables.device_*)."""
        prm = MpsObserveParams(surface_n, rho_g, surface_z, x_h1, x_h2, min_n, z_p2, d)
        out = MpsObservables()
        self._check(self.lib.mps_observe(self.h, C.byref(prm), C.byref(out)))
        return {k: getattr(out, k) for k in _OBS_FIELDS}

    def set_stage_timing(self, on):
        self._check(self.lib.mps_set_stage_timing(self.h, int(on)))

    def stats(self):
        st = MpsStats()
        self._check(self.lib.mps_get_stats(self.h, C.byref(st)))
        return st

    def stats_dict(self):
        st = self.stats()
        names = []
        k = 0
        while True:
            nm = self.lib.mps_stage_name(k)
            if nm is None:
                break
            names.append(nm.decode()); k += 1
        return {"steps": st.steps, "cg_iterations": st.cg_iterations, "last_cg_iterations": st.last_cg_iterations,
                "last_rr0": st.last_rr0, "last_rr": st.last_rr, "particles": st.particles, "neighbors": st.neighbors,
                "nnz": st.nnz, "active_rows": st.active_rows, "kernel_launches": st.kernel_launches, "comm_calls": st.comm_calls, "disabled_last": st.disabled_last, "cg_ms": st.cg_ms, "cg_bytes": st.cg_bytes,
                "matrix_sweeps": st.matrix_sweeps, "mg_levels": st.mg_levels, "mg_cells": st.mg_cells,
                "stage_ms": {nm: st.stage_ms[i] for i, nm in enumerate(names)},
                "stage_calls": {nm: st.stage_calls[i] for i, nm in enumerate(names)}}

    def reset_stats(self):
        self._check(self.lib.mps_reset_stats(self.h))

    def time_kernel(self, name, reps=5):
        ms, by = C.c_double(), C.c_double()
        self._check(self.lib.mps_time_kernel(self.h, name.encode(), reps, C.byref(ms), C.byref(by)))
        return ms.value, by.value

    # ---- multi-GPU ----
    @staticmethod
    def comm_unique_id():
        """128-byte NCCL unique id (call on rank 0, distribute with the launcher's own transport)."""
        lib = load_library()
        buf = (C.c_ubyte * 128)()
        rc = lib.mps_comm_unique_id(buf)
        if rc != MPS_OK:
            raise MpsError(rc, "mps_comm_unique_id failed (is libnccl loadable?)")
        return bytes(buf)

    def attach_comm(self, rank, nranks, unique_id):
        buf = (C.c_ubyte * 128).from_buffer_copy(bytes(unique_id))
        self._check(self.lib.mps_comm_init(self.h, int(rank), int(nranks), buf))

    def comm_info(self):
        r, n, a, b = C.c_int(), C.c_int(), C.c_uint64(), C.c_uint64()
        self._check(self.lib.mps_comm_info(self.h, C.byref(r), C.byref(n), C.byref(a), C.byref(b)))
        m = C.c_int()
        self._check(self.lib.mps_comm_mode(self.h, C.byref(m)))
        return {"rank": r.value, "nranks": n.value, "own": (a.value, b.value), "mode": {0: "single", 1: "peer-memory", 2: "nccl"}[m.value]}

    def set_cg_profile(self, on):
        self._check(self.lib.mps_set_cg_profile(self.h, 1 if on else 0))

    def cg_profile(self):
        """Cycle counters of the last streaming CG solve (see mps_get_cg_profile)."""
        out = (C.c_double * 19)()
        self._check(self.lib.mps_get_cg_profile(self.h, out))
        names = ["phase1", "wait_data", "phase2", "barriers", "wait_stage", "chunks_per_cta", "iteration_cycles_total", "vcycle"]
        d = {"mean": dict(zip(names, out[0:8])), "max": dict(zip(names, out[8:16])), "chunks": out[16], "blob_bytes": out[17], "ctas": out[18]}
        return d

    def cg_profile_stages(self):
        out = np.zeros(64, dtype=np.uint64)
        self._check(self.lib.mps_get_cg_profile_stages(self.h, _ptr(out)))
        return out

    def cg_profile_raw(self):
        out = np.zeros((1024, 8), dtype=np.uint64)
        n = C.c_uint64(0)
        self._check(self.lib.mps_get_cg_profile_raw(self.h, _ptr(out), 1024, C.byref(n)))
        return out[: n.value]

    def flush_l2(self):
        self._check(self.lib.mps_flush_l2(self.h))
