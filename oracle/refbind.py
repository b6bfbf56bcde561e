"""ctypes binding of oracle/_ref/libref_*.so — the UNMODIFIED reference step engine.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
See oracle/ref_capi.cpp for what the library is and how it is built (oracle/Makefile, target ``ref``).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_REF_DIR = os.path.join(_HERE, "_ref")

_dp = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")
_ip = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
_u32p = np.ctypeslib.ndpointer(np.uint32, flags="C_CONTIGUOUS")
_u64p = np.ctypeslib.ndpointer(np.uint64, flags="C_CONTIGUOUS")
_i64p = np.ctypeslib.ndpointer(np.int64, flags="C_CONTIGUOUS")

VEC = {"x": 0, "b": 1, "r": 2, "p": 3, "Ap": 4, "ecs": 5, "nWithoutSpp": 6, "du": 7, "originalX": 8}


def variant_name(dim, central_gravity=False, fast=False):
    name = "3d" if dim == 3 else "2d"
    if central_gravity:
        name += "_cg"
    if fast:
        name += "_fast"
    return name


def available(dim=2, central_gravity=False, fast=False):
    return os.path.exists(os.path.join(_REF_DIR, f"libref_{variant_name(dim, central_gravity, fast)}.so"))


_libs = {}


def _load(dim, central_gravity, fast):
    key = variant_name(dim, central_gravity, fast)
    if key in _libs:
        return _libs[key]
    path = os.path.join(_REF_DIR, f"libref_{key}.so")
    if not os.path.exists(path):
        raise FileNotFoundError(f"{path} missing: build it with `make -C oracle ref` where /root/reference exists")
    lib = C.CDLL(path)
    lib.ref_create.restype = C.c_void_p
    lib.ref_create.argtypes = [C.c_double] * 7 + [_dp, _dp, C.c_double]
    lib.ref_destroy.argtypes = [C.c_void_p]
    lib.ref_last_error.restype = C.c_char_p
    lib.ref_last_error.argtypes = [C.c_void_p]
    lib.ref_add_particles.argtypes = [C.c_void_p, C.c_uint64, _dp, _dp, _dp, _dp, _ip]
    lib.ref_set_wall_positions.argtypes = [C.c_void_p, C.c_uint64, _u64p, _dp]
    lib.ref_count.restype = C.c_uint64
    lib.ref_count.argtypes = [C.c_void_p]
    lib.ref_get_state.argtypes = [C.c_void_p, _dp, _dp, _dp, _dp, _ip]
    lib.ref_set_state.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.ref_get_env.argtypes = [C.c_void_p, _dp]
    lib.ref_set_dt.argtypes = [C.c_void_p, C.c_double, C.c_int]
    lib.ref_determine_dt.restype = C.c_double
    lib.ref_determine_dt.argtypes = [C.c_void_p]
    lib.ref_stage.argtypes = [C.c_void_p, C.c_char_p]
    lib.ref_forward.argtypes = [C.c_void_p, C.c_uint64, C.c_double, C.POINTER(C.c_uint64), C.POINTER(C.c_double)]
    lib.ref_run_until.argtypes = [C.c_void_p, C.c_double, C.POINTER(C.c_uint64)]
    lib.ref_get_cells.argtypes = [C.c_void_p, _i64p]
    lib.ref_grid_capacity.restype = C.c_uint64
    lib.ref_grid_capacity.argtypes = [C.c_void_p]
    lib.ref_get_neighbors.argtypes = [C.c_void_p, _u64p, C.c_void_p]
    lib.ref_csr_nnz.restype = C.c_uint64
    lib.ref_csr_nnz.argtypes = [C.c_void_p]
    lib.ref_get_csr.argtypes = [C.c_void_p, _u32p, _u32p, _dp]
    lib.ref_set_system.argtypes = [C.c_void_p, C.c_uint64, _u32p, _u32p, _dp, _dp, _dp]
    lib.ref_get_vec.argtypes = [C.c_void_p, C.c_int, _dp]
    lib.ref_dndt.restype = C.c_double
    lib.ref_dndt.argtypes = [C.c_void_p, C.c_uint64]
    assert lib.ref_dim() == dim and bool(lib.ref_central_gravity()) == bool(central_gravity)
    _libs[key] = lib
    return lib


class RefError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(msg)
        self.code = code


class RefComputer:
    """The reference ``Computer`` (Computer.hpp) driven stage by stage, as the upstream gtests drive it."""

    def __init__(self, env, fast=False):
        self.env = env
        self.dim = env.dim
        self.lib = _load(env.dim, env.central_gravity, fast)
        lo = np.ascontiguousarray(env.min_x, dtype=np.float64)
        hi = np.ascontiguousarray(env.max_x, dtype=np.float64)
        self.h = self.lib.ref_create(env.max_dt, env.courant, env.g, env.rho, env.nu, env.r_e_by_l0, env.l0, lo, hi, env.eps)

    def close(self):
        if self.h:
            self.lib.ref_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @classmethod
    def from_scene(cls, scene, fast=False):
        c = cls(scene.env, fast)
        c.add_particles(scene.x, scene.u, scene.p, scene.n, scene.type)
        return c

    def _check(self, rc):
        if rc != 0:
            raise RefError(rc, self.lib.ref_last_error(self.h).decode())

    def add_particles(self, x, u, p, n, type):
        cnt = len(type)
        self.lib.ref_add_particles(self.h, cnt, np.ascontiguousarray(x, np.float64), np.ascontiguousarray(u, np.float64),
                                   np.ascontiguousarray(p, np.float64), np.ascontiguousarray(n, np.float64),
                                   np.ascontiguousarray(type, np.int32))

    def set_wall_positions(self, ids, x):
        ids = np.ascontiguousarray(ids, np.uint64)
        self.lib.ref_set_wall_positions(self.h, len(ids), ids, np.ascontiguousarray(x, np.float64))

    @property
    def count(self):
        return int(self.lib.ref_count(self.h))

    def state(self):
        n, d = self.count, self.dim
        x = np.empty((n, d)); u = np.empty((n, d)); p = np.empty(n); nd = np.empty(n); t = np.empty(n, np.int32)
        self.lib.ref_get_state(self.h, x, u, p, nd, t)
        return {"x": x, "u": u, "p": p, "n": nd, "type": t}

    def set_state(self, x=None, u=None, p=None, n=None):
        arrs = [None if a is None else np.ascontiguousarray(a, np.float64) for a in (x, u, p, n)]
        self.lib.ref_set_state(self.h, *[None if a is None else a.ctypes.data for a in arrs])

    def env_values(self):
        out = np.empty(10)
        self.lib.ref_get_env(self.h, out)
        keys = ["t", "dt", "n0", "MaxDt", "MaxDx", "R_e", "NeighborLength", "L_0", "Rho", "Nu"]
        return dict(zip(keys, out.tolist()))

    def set_dt(self, dt, advance=True):
        self.lib.ref_set_dt(self.h, dt, int(advance))

    def determine_dt(self):
        return float(self.lib.ref_determine_dt(self.h))

    def stage(self, name):
        self._check(self.lib.ref_stage(self.h, name.encode()))

    def forward(self, steps=1, dt=None):
        done = C.c_uint64(0); sec = C.c_double(0)
        rc = self.lib.ref_forward(self.h, steps, -1.0 if dt is None else dt, C.byref(done), C.byref(sec))
        self.last_seconds = sec.value
        self.last_steps = done.value
        self._check(rc)
        return done.value

    def run_until(self, t_end):
        done = C.c_uint64(0)
        rc = self.lib.ref_run_until(self.h, t_end, C.byref(done))
        self._check(rc)
        return done.value

    def cells(self):
        out = np.empty((self.count, self.dim), np.int64)
        self.lib.ref_get_cells(self.h, out)
        return out

    def grid_capacity(self):
        return int(self.lib.ref_grid_capacity(self.h))

    def neighbors(self):
        n = self.count
        rowptr = np.empty(n + 1, np.uint64)
        self.lib.ref_get_neighbors(self.h, rowptr, None)
        idx = np.empty(int(rowptr[-1]), np.uint64)
        self.lib.ref_get_neighbors(self.h, rowptr, idx.ctypes.data)
        return rowptr, idx

    def csr(self):
        n = self.count
        nnz = int(self.lib.ref_csr_nnz(self.h))
        rowptr = np.empty(n + 1, np.uint32); col = np.empty(nnz, np.uint32); val = np.empty(nnz)
        self.lib.ref_get_csr(self.h, rowptr, col, val)
        return rowptr, col, val

    def set_system(self, rowptr, col, val, b, x0):
        n = len(b)
        self._n_sys = n
        self.lib.ref_set_system(self.h, n, np.ascontiguousarray(rowptr, np.uint32), np.ascontiguousarray(col, np.uint32),
                                np.ascontiguousarray(val, np.float64), np.ascontiguousarray(b, np.float64),
                                np.ascontiguousarray(x0, np.float64))

    def vec(self, name, n=None):
        which = VEC[name]
        n = self.count if n is None else n
        out = np.empty((n, self.dim)) if which >= 7 else np.empty(n)
        self.lib.ref_get_vec(self.h, which, out)
        return out

    def dndt(self, i):
        return float(self.lib.ref_dndt(self.h, int(i)))
