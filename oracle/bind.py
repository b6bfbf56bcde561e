"""ctypes bindings of the two CPU checkers.  TEST INFRASTRUCTURE ONLY.

* ``RefComputer``  — oracle/_ref/libref_*.so: the UNMODIFIED reference step engine (oracle/ref_capi.cpp; built by
                     ``make -C oracle ref`` where /root/reference exists; the prebuilt .so travels to the GPU box).
* ``PortComputer`` — oracle/_build/libmps_oracle.so: our CPU restatement (oracle/mps_oracle.cpp; ``make -C oracle port``).

Both expose the same stage-level interface.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
``--impl reference`` legs may import this module; the product never does.
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


class _Api:
    """Prefix-stripping view of a library: api.stage -> lib.ref_stage / lib.orc_stage."""

    def __init__(self, lib, prefix):
        self._lib, self._prefix = lib, prefix

    def __getattr__(self, name):
        return getattr(self._lib, self._prefix + name)


def _bind_common(api):
    lib = api
    lib.destroy.argtypes = [C.c_void_p]
    lib.last_error.restype = C.c_char_p
    lib.last_error.argtypes = [C.c_void_p]
    lib.add_particles.argtypes = [C.c_void_p, C.c_uint64, _dp, _dp, _dp, _dp, _ip]
    lib.set_wall_positions.argtypes = [C.c_void_p, C.c_uint64, _u64p, _dp]
    lib.count.restype = C.c_uint64
    lib.count.argtypes = [C.c_void_p]
    lib.get_state.argtypes = [C.c_void_p, _dp, _dp, _dp, _dp, _ip]
    lib.set_state.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.get_env.argtypes = [C.c_void_p, _dp]
    lib.set_dt.argtypes = [C.c_void_p, C.c_double, C.c_int]
    lib.determine_dt.restype = C.c_double
    lib.determine_dt.argtypes = [C.c_void_p]
    lib.stage.argtypes = [C.c_void_p, C.c_char_p]
    lib.forward.argtypes = [C.c_void_p, C.c_uint64, C.c_double, C.POINTER(C.c_uint64), C.POINTER(C.c_double)]
    lib.run_until.argtypes = [C.c_void_p, C.c_double, C.POINTER(C.c_uint64)]
    lib.get_cells.argtypes = [C.c_void_p, _i64p]
    lib.grid_capacity.restype = C.c_uint64
    lib.grid_capacity.argtypes = [C.c_void_p]
    lib.get_neighbors.argtypes = [C.c_void_p, _u64p, C.c_void_p]
    lib.csr_nnz.restype = C.c_uint64
    lib.csr_nnz.argtypes = [C.c_void_p]
    lib.get_csr.argtypes = [C.c_void_p, _u32p, _u32p, _dp]
    lib.set_system.argtypes = [C.c_void_p, C.c_uint64, _u32p, _u32p, _dp, _dp, _dp]
    lib.get_vec.argtypes = [C.c_void_p, C.c_int, _dp]
    lib.dndt.restype = C.c_double
    lib.dndt.argtypes = [C.c_void_p, C.c_uint64]


def _load(dim, central_gravity, fast):
    key = variant_name(dim, central_gravity, fast)
    if key in _libs:
        return _libs[key]
    path = os.path.join(_REF_DIR, f"libref_{key}.so")
    if not os.path.exists(path):
        raise FileNotFoundError(f"{path} missing: build it with `make -C oracle ref` where /root/reference exists")
    lib = C.CDLL(path)
    api = _Api(lib, "ref_")
    api.create.restype = C.c_void_p
    api.create.argtypes = [C.c_double] * 7 + [_dp, _dp, C.c_double]
    _bind_common(api)
    assert lib.ref_dim() == dim and bool(lib.ref_central_gravity()) == bool(central_gravity)
    _libs[key] = api
    return api


def port_available():
    return os.path.exists(os.path.join(_HERE, "_build", "libmps_oracle.so"))


def _load_port():
    if "port" in _libs:
        return _libs["port"]
    path = os.path.join(_HERE, "_build", "libmps_oracle.so")
    if not os.path.exists(path):
        raise FileNotFoundError(f"{path} missing: build it with `make -C oracle port`")
    lib = C.CDLL(path)
    api = _Api(lib, "orc_")
    api.create.restype = C.c_void_p
    api.create.argtypes = [C.c_int, C.c_int] + [C.c_double] * 7 + [_dp, _dp, C.c_double]
    _bind_common(api)
    api.last_iterations.restype = C.c_uint64
    api.last_iterations.argtypes = [C.c_void_p]
    _libs["port"] = api
    return api


class RefError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(msg)
        self.code = code


class _StepEngine:
    """Stage-level driver shared by both checkers (same calls the upstream gtest fixtures make)."""

    def __init__(self, env, fast=False):
        self.env = env
        self.dim = env.dim
        self.h = None
        self._open(env, fast)

    def _open(self, env, fast):
        raise NotImplementedError

    def close(self):
        if self.h:
            self.lib.destroy(self.h)
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
            raise RefError(rc, self.lib.last_error(self.h).decode())

    def add_particles(self, x, u, p, n, type):
        cnt = len(type)
        self.lib.add_particles(self.h, cnt, np.ascontiguousarray(x, np.float64), np.ascontiguousarray(u, np.float64),
                                   np.ascontiguousarray(p, np.float64), np.ascontiguousarray(n, np.float64),
                                   np.ascontiguousarray(type, np.int32))

    def set_wall_positions(self, ids, x):
        ids = np.ascontiguousarray(ids, np.uint64)
        self.lib.set_wall_positions(self.h, len(ids), ids, np.ascontiguousarray(x, np.float64))

    @property
    def count(self):
        return int(self.lib.count(self.h))

    def state(self):
        n, d = self.count, self.dim
        x = np.empty((n, d)); u = np.empty((n, d)); p = np.empty(n); nd = np.empty(n); t = np.empty(n, np.int32)
        self.lib.get_state(self.h, x, u, p, nd, t)
        return {"x": x, "u": u, "p": p, "n": nd, "type": t}

    def set_state(self, x=None, u=None, p=None, n=None):
        arrs = [None if a is None else np.ascontiguousarray(a, np.float64) for a in (x, u, p, n)]
        self.lib.set_state(self.h, *[None if a is None else a.ctypes.data for a in arrs])

    def env_values(self):
        out = np.empty(10)
        self.lib.get_env(self.h, out)
        keys = ["t", "dt", "n0", "MaxDt", "MaxDx", "R_e", "NeighborLength", "L_0", "Rho", "Nu"]
        return dict(zip(keys, out.tolist()))

    def set_dt(self, dt, advance=True):
        self.lib.set_dt(self.h, dt, int(advance))

    def determine_dt(self):
        return float(self.lib.determine_dt(self.h))

    def stage(self, name):
        self._check(self.lib.stage(self.h, name.encode()))

    def forward(self, steps=1, dt=None):
        done = C.c_uint64(0); sec = C.c_double(0)
        rc = self.lib.forward(self.h, steps, -1.0 if dt is None else dt, C.byref(done), C.byref(sec))
        self.last_seconds = sec.value
        self.last_steps = done.value
        self._check(rc)
        return done.value

    def run_until(self, t_end):
        done = C.c_uint64(0)
        rc = self.lib.run_until(self.h, t_end, C.byref(done))
        self._check(rc)
        return done.value

    def cells(self):
        out = np.empty((self.count, self.dim), np.int64)
        self.lib.get_cells(self.h, out)
        return out

    def grid_capacity(self):
        return int(self.lib.grid_capacity(self.h))

    def neighbors(self):
        n = self.count
        rowptr = np.empty(n + 1, np.uint64)
        self.lib.get_neighbors(self.h, rowptr, None)
        idx = np.empty(int(rowptr[-1]), np.uint64)
        self.lib.get_neighbors(self.h, rowptr, idx.ctypes.data)
        return rowptr, idx

    def csr(self):
        n = self.count
        nnz = int(self.lib.csr_nnz(self.h))
        rowptr = np.empty(n + 1, np.uint32); col = np.empty(nnz, np.uint32); val = np.empty(nnz)
        self.lib.get_csr(self.h, rowptr, col, val)
        return rowptr, col, val

    def set_system(self, rowptr, col, val, b, x0):
        n = len(b)
        self._n_sys = n
        self.lib.set_system(self.h, n, np.ascontiguousarray(rowptr, np.uint32), np.ascontiguousarray(col, np.uint32),
                                np.ascontiguousarray(val, np.float64), np.ascontiguousarray(b, np.float64),
                                np.ascontiguousarray(x0, np.float64))

    def vec(self, name, n=None):
        which = VEC[name]
        n = self.count if n is None else n
        out = np.empty((n, self.dim)) if which >= 7 else np.empty(n)
        self.lib.get_vec(self.h, which, out)
        return out

    def dndt(self, i):
        return float(self.lib.dndt(self.h, int(i)))


class RefComputer(_StepEngine):
    """The reference ``Computer`` (Computer.hpp), unmodified, from oracle/_ref."""
    kind = "reference"

    def _open(self, env, fast):
        self.lib = _load(env.dim, env.central_gravity, fast)
        lo = np.ascontiguousarray(env.min_x, dtype=np.float64)
        hi = np.ascontiguousarray(env.max_x, dtype=np.float64)
        self.h = self.lib.create(env.max_dt, env.courant, env.g, env.rho, env.nu, env.r_e_by_l0, env.l0, lo, hi, env.eps)


class PortComputer(_StepEngine):
    """Our CPU restatement (oracle/mps_oracle.cpp)."""
    kind = "port"

    def _open(self, env, fast):
        self.lib = _load_port()
        lo = np.ascontiguousarray(env.min_x, dtype=np.float64)
        hi = np.ascontiguousarray(env.max_x, dtype=np.float64)
        self.h = self.lib.create(env.dim, int(env.central_gravity), env.max_dt, env.courant, env.g, env.rho, env.nu,
                                 env.r_e_by_l0, env.l0, lo, hi, env.eps)

    def last_iterations(self):
        return int(self.lib.last_iterations(self.h))
