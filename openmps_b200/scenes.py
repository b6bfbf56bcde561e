"""Deterministic synthetic inputs for the MPS hot path (particle blocks + environment constants).

These re-implement, from their published description, the reference's input generators so that the same
particle sets can be produced on a box where /root/reference does not exist:

* ``dambreak2d``      — Benchmark/DamBreak/generate_koshizukaoka1996.py:57-149 (== Benchmark/Sample/Sample.xml for l0 = 8 mm)
* ``static_pressure`` — Benchmark/StaticPressure/generate.py:56-191
* ``central_gravity`` — Benchmark/CentralGravity/generate.py:55-77 (needs the CENTRAL_GRAVITY variant, Computer.hpp:925,981)
* ``dambreak3d``      — our 3-D extension of the K&O tank (no upstream 3-D generator exists); square
                        4L x 4L footprint because of the reference's 3-D Grid bug (Grid.hpp:169,198,232)
* ``lattice``         — the square lattices of the upstream gtest fixtures (e.g. test_ComputerNumberDensity.cpp:82-98)

Coordinates are produced with the same floating-point expression the scripts use (``i * l_0``), and Sample.xml stores
them through Python ``str(float)`` (shortest round-trip), so values are bit-identical to the reference inputs.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field, replace

import numpy as np

FLUID, WALL, DUMMY, DISABLED = 0, 1, 2, 3


@dataclass
class Env:
    """Arguments of the reference's ``Environment`` constructor (Environment.hpp:101-128) + solver eps + variant."""
    dim: int
    max_dt: float          # outputInterval / minStepCountPerOutput (Main.cpp:234)
    courant: float
    g: float
    rho: float
    nu: float
    r_e_by_l0: float
    l0: float
    min_x: tuple
    max_x: tuple
    eps: float = 1e-10
    central_gravity: bool = False

    def scaled(self, **kw):
        return replace(self, **kw)


@dataclass
class Scene:
    env: Env
    x: np.ndarray      # (n, dim) float64
    u: np.ndarray      # (n, dim) float64
    p: np.ndarray      # (n,)     float64
    n: np.ndarray      # (n,)     float64  (particle number density)
    type: np.ndarray   # (n,)     int32    0 fluid, 1 wall, 2 dummy, 3 disabled (Particle.hpp:16-29)
    name: str = ""
    meta: dict = field(default_factory=dict)

    @property
    def count(self):
        return int(self.type.shape[0])


def _scene(env, pts, types, name, **meta):
    x = np.ascontiguousarray(np.asarray(pts, dtype=np.float64).reshape(-1, env.dim))
    t = np.ascontiguousarray(np.asarray(types, dtype=np.int32))
    n = t.shape[0]
    return Scene(env, x, np.zeros_like(x), np.zeros(n), np.zeros(n), t, name, meta)


# ---------------------------------------------------------------------------------------------------------------------
def dambreak2d(l0=8e-3, max_dt=None, eps=1e-10):
    """Koshizuka & Oka (1996) dam break; l0 = 8e-3 reproduces Benchmark/Sample/Sample.xml (1 323 particles).

    ``max_dt`` defaults to the Sample's 5e-4 scaled with resolution (5e-4 * l0 / 8e-3): the reference itself diverges
    at finer l0 with an unscaled cap (SURVEY.md §8a profile note 5).
    """
    L, height, width = 14.6e-2, 2, 4
    l = math.ceil(L / l0)
    h = math.ceil(L * height / l0)
    w = math.ceil(L * width / l0)
    pts, types = [], []

    def add(t, x, z):
        pts.append((x, z)); types.append(t)

    for i in range(0, l):                       # water column
        for j in range(0, h):
            add(FLUID, i * l0, j * l0)
    for i in range(-1, w + 1):                  # floor: 1 wall + 3 dummy layers
        add(WALL, i * l0, -1 * l0)
        add(DUMMY, i * l0, -2 * l0)
        add(DUMMY, i * l0, -3 * l0)
        add(DUMMY, i * l0, -4 * l0)
    for j in range(0, h - 1):                   # left wall
        add(WALL, -1 * l0, j * l0)
        add(DUMMY, -2 * l0, j * l0)
        add(DUMMY, -3 * l0, j * l0)
        add(DUMMY, -4 * l0, j * l0)
    for j in range(0, h - 1):                   # right wall
        add(WALL, (w + 0) * l0, j * l0)
        add(DUMMY, (w + 1) * l0, j * l0)
        add(DUMMY, (w + 2) * l0, j * l0)
        add(DUMMY, (w + 3) * l0, j * l0)
    for i in range(0, 4):                       # wall crest
        z = (h - 1) * l0
        add(WALL, -(i + 1) * l0, z)
        add(WALL, (w + i) * l0, z)
    for i in range(1, 4):                       # bottom corners
        for j in range(-4, 0):
            add(DUMMY, (-1 - i) * l0, j * l0)
            add(DUMMY, (w + i) * l0, j * l0)

    if max_dt is None:
        max_dt = 0.005 / 10 * (l0 / 8e-3)
    env = Env(2, max_dt, 0.1, 9.8, 998.20, 1.004e-6, 2.4, l0,
              (-4 * l0, -4 * l0), (width * L, height * 2 * L), eps)
    return _scene(env, pts, types, f"dambreak2d_l0={l0:g}", L=L)


def dambreak2d_fast(l0, max_dt=None, eps=1e-10):
    """Vectorised ``dambreak2d`` for the million-particle configs (same particle order and values)."""
    L, height, width = 14.6e-2, 2, 4
    l = math.ceil(L / l0); h = math.ceil(L * height / l0); w = math.ceil(L * width / l0)
    f = np.float64
    ii, jj = np.meshgrid(np.arange(l, dtype=f), np.arange(h, dtype=f), indexing="ij")
    water = np.stack([ii.ravel() * l0, jj.ravel() * l0], 1)
    parts = [water]; types = [np.full(l * h, FLUID, np.int32)]

    def block(xs, zs, ts):
        parts.append(np.stack([np.asarray(xs, f), np.asarray(zs, f)], 1)); types.append(np.asarray(ts, np.int32))

    i = np.repeat(np.arange(-1, w + 1, dtype=f), 4)
    k = np.tile(np.array([-1, -2, -3, -4], f), w + 2)
    block(i * l0, k * l0, np.tile(np.array([WALL, DUMMY, DUMMY, DUMMY], np.int32), w + 2))
    j = np.repeat(np.arange(0, h - 1, dtype=f), 4)
    block(np.tile(np.array([-1, -2, -3, -4], f), h - 1) * l0, j * l0, np.tile(np.array([WALL, DUMMY, DUMMY, DUMMY], np.int32), h - 1))
    block(np.tile(np.array([w + 0, w + 1, w + 2, w + 3], f), h - 1) * l0, j * l0, np.tile(np.array([WALL, DUMMY, DUMMY, DUMMY], np.int32), h - 1))
    xs, zs = [], []
    for a in range(0, 4):
        xs += [-(a + 1) * l0, (w + a) * l0]; zs += [(h - 1) * l0] * 2
    block(xs, zs, [WALL] * 8)
    xs, zs = [], []
    for a in range(1, 4):
        for b in range(-4, 0):
            xs += [(-1 - a) * l0, (w + a) * l0]; zs += [b * l0] * 2
    block(xs, zs, [DUMMY] * len(xs))
    if max_dt is None:
        max_dt = 0.005 / 10 * (l0 / 8e-3)
    env = Env(2, max_dt, 0.1, 9.8, 998.20, 1.004e-6, 2.4, l0, (-4 * l0, -4 * l0), (width * L, height * 2 * L), eps)
    x = np.concatenate(parts); t = np.concatenate(types)
    return Scene(env, np.ascontiguousarray(x), np.zeros_like(x), np.zeros(len(t)), np.zeros(len(t)), t, f"dambreak2d_l0={l0:g}", {"L": L})


def static_pressure(l0=1e-3, width=50, height=100, max_dt=None, eps=1e-10):
    """Hydrostatic column (Benchmark/StaticPressure/generate.py:56-191): default 6 040 particles (5 000 fluid)."""
    pts, types = [], []

    def add(t, x, z):
        pts.append((x, z)); types.append(t)

    for i in range(0, width):
        for j in range(0, height):
            add(FLUID, i * l0, j * l0)
    for i in range(-1, width + 1):              # floor
        add(WALL, i * l0, -1 * l0)
        add(DUMMY, i * l0, -2 * l0); add(DUMMY, i * l0, -3 * l0); add(DUMMY, i * l0, -4 * l0)
    for j in range(0, height + 1):              # left wall
        add(WALL, -1 * l0, j * l0)
        add(DUMMY, -2 * l0, j * l0); add(DUMMY, -3 * l0, j * l0); add(DUMMY, -4 * l0, j * l0)
    for j in range(0, height + 1):              # right wall
        add(WALL, (width + 0) * l0, j * l0)
        add(DUMMY, (width + 1) * l0, j * l0); add(DUMMY, (width + 2) * l0, j * l0); add(DUMMY, (width + 3) * l0, j * l0)
    for i in range(1, 4):                       # bottom corners
        for j in range(-4, 0):
            add(DUMMY, (-1 - i) * l0, j * l0)
            add(DUMMY, (width + i) * l0, j * l0)
    arr = np.asarray(pts)
    if max_dt is None:
        max_dt = 0.0005 / 100 * (l0 / 1e-3)
    # the script sets the domain to the bounding box of everything it emitted before the corner blocks; the corners lie
    # inside that box, so the full bounding box is the same
    env = Env(2, max_dt, 0.1, 9.8, 998.20, 1.004e-6, 2.4, l0,
              (float(arr[:, 0].min()), float(arr[:, 1].min())), (float(arr[:, 0].max()), float(arr[:, 1].max())), eps)
    return _scene(env, pts, types, f"static_pressure_l0={l0:g}")


def central_gravity(half=100, l0=0.5e-3, area=2, max_dt=None, eps=1e-10):
    """Square fluid block pulled to the origin (Benchmark/CentralGravity/generate.py); (2*half+1)^2 fluid particles."""
    f = np.float64
    ii, jj = np.meshgrid(np.arange(-half, half + 1, dtype=f), np.arange(-half, half + 1, dtype=f), indexing="ij")
    x = np.ascontiguousarray(np.stack([ii.ravel() * l0, jj.ravel() * l0], 1))
    t = np.full(x.shape[0], FLUID, np.int32)
    if max_dt is None:
        max_dt = 0.01 / 50 * (l0 / 0.5e-3)
    env = Env(2, max_dt, 0.1, 9.8, 998.20, 1.004e-6, 2.4, l0,
              (-area * l0 * half, -area * l0 * half), (area * l0 * half, area * l0 * half), eps, central_gravity=True)
    return Scene(env, x, np.zeros_like(x), np.zeros(len(t)), np.zeros(len(t)), t, f"central_gravity_{2 * half + 1}^2")


def dambreak3d(l0=8e-3, r_e_by_l0=2.4, max_dt=None, eps=1e-10, depth_by_L=4):
    """3-D K&O tank: water L x W x 2L against the x = 0, y-spanning wall; tank footprint 4L x (depth_by_L * L).

    One wall layer + three dummy layers on the floor and the four sides (same recipe as the 2-D script applied on both
    horizontal axes); open top.  The footprint is square by default so the reference's 3-D grid stays valid.
    """
    L, height, width = 14.6e-2, 2, 4
    l = math.ceil(L / l0); h = math.ceil(L * height / l0); w = math.ceil(L * width / l0); d = math.ceil(L * depth_by_L / l0)
    f = np.float64
    parts, types = [], []

    def box(i0, i1, j0, j1, k0, k1, t):
        if i1 <= i0 or j1 <= j0 or k1 <= k0:
            return
        ii, jj, kk = np.meshgrid(np.arange(i0, i1, dtype=f), np.arange(j0, j1, dtype=f), np.arange(k0, k1, dtype=f), indexing="ij")
        parts.append(np.stack([ii.ravel() * l0, jj.ravel() * l0, kk.ravel() * l0], 1))
        types.append(np.full(parts[-1].shape[0], t, np.int32))

    box(0, l, 0, d, 0, h, FLUID)                       # water spans the full depth, like the 2-D column extruded in y
    # shells: layer s = 1 is wall, s = 2..4 dummy; each shell is the boundary of the box [-s, w-1+s] x [-s, d-1+s] x [-s, h-2]
    for s in range(1, 5):
        t = WALL if s == 1 else DUMMY
        x0, x1, y0, y1, zt = -s, w + s, -s, d + s, h - 1     # exclusive upper bounds in x, y; walls rise to z index h-2
        box(x0, x1, y0, y1, -s, -s + 1, t)                   # floor slab
        box(x0, x0 + 1, y0, y1, -s + 1, zt, t)               # x-min side
        box(x1 - 1, x1, y0, y1, -s + 1, zt, t)               # x-max side
        box(x0 + 1, x1 - 1, y0, y0 + 1, -s + 1, zt, t)       # y-min side
        box(x0 + 1, x1 - 1, y1 - 1, y1, -s + 1, zt, t)       # y-max side
    # crest: the top ring (z index h-1) is wall across all four layers, as in the 2-D script's "wall crest"
    for s in range(1, 5):
        x0, x1, y0, y1 = -s, w + s, -s, d + s
        box(x0, x0 + 1, y0, y1, h - 1, h, WALL)
        box(x1 - 1, x1, y0, y1, h - 1, h, WALL)
        box(x0 + 1, x1 - 1, y0, y0 + 1, h - 1, h, WALL)
        box(x0 + 1, x1 - 1, y1 - 1, y1, h - 1, h, WALL)
    x = np.ascontiguousarray(np.concatenate(parts)); t = np.concatenate(types)
    if max_dt is None:
        max_dt = 0.005 / 10 * (l0 / 8e-3)
    env = Env(3, max_dt, 0.1, 9.8, 998.20, 1.004e-6, r_e_by_l0, l0,
              (-4 * l0, -4 * l0, -4 * l0), (width * L, depth_by_L * L, height * 2 * L), eps)
    return Scene(env, x, np.zeros_like(x), np.zeros(len(t)), np.zeros(len(t)), t, f"dambreak3d_l0={l0:g}", {"L": L})


def lattice(dim, num, l0, r_e_by_l0, margin_cells=2.0, courant=0.1, g=9.8, max_dt=1e-2, jitter=0.0, seed=12345,
            types=None, eps=1e-10):
    """Square/cubic lattice of ``num`` particles per axis (the upstream fixtures), optional uniform jitter * l0."""
    f = np.float64
    ax = [np.arange(num, dtype=f)] * dim
    grid = np.meshgrid(*ax, indexing="ij")
    x = np.stack([a.ravel() * l0 for a in grid], 1)
    if jitter:
        rng = np.random.default_rng(seed)
        x = x + rng.uniform(-jitter * l0, jitter * l0, x.shape)
    x = np.ascontiguousarray(x)
    t = np.full(x.shape[0], FLUID, np.int32) if types is None else np.asarray(types, np.int32)
    lo = tuple([-margin_cells * l0 * num] * dim); hi = tuple([margin_cells * l0 * num] * dim)
    env = Env(dim, max_dt, courant, g, 998.2, 1.004e-6, r_e_by_l0, l0, lo, hi, eps)
    return Scene(env, x, np.zeros_like(x), np.zeros(len(t)), np.zeros(len(t)), t, f"lattice{dim}d_{num}")


# ---------------------------------------------------------------------------------------------------------------------
def write_xml(scene, path, start_time=0.0, end_time=1.0, output_interval=5e-3, min_step_count_per_output=None, pretty=True):
    """The run description the reference's driver reads (Main.cpp:186-274), in the layout its generators write
    (Benchmark/DamBreak/generate_koshizukaoka1996.py:47-149: `value` attributes, particle table as CSV text with
    ``str(float)`` numbers).  ``max_dt`` of the scene becomes outputInterval / minStepCountPerOutput."""
    import xml.etree.ElementTree as ET
    from xml.dom import minidom

    env = scene.env
    if min_step_count_per_output is None:
        min_step_count_per_output = max(1, int(round(output_interval / env.max_dt)))
    root = ET.Element("openmps")
    c = ET.SubElement(root, "condition")
    for key, v in (("startTime", start_time), ("endTime", end_time), ("outputInterval", output_interval), ("eps", env.eps)):
        ET.SubElement(c, key).set("value", str(v))
    e = ET.SubElement(root, "environment")
    for key, v in (("l_0", env.l0), ("minStepCountPerOutput", int(min_step_count_per_output)), ("courant", env.courant), ("g", env.g),
                   ("rho", env.rho), ("nu", env.nu), ("r_eByl_0", env.r_e_by_l0), ("surfaceRatio", 0.97)):
        ET.SubElement(e, key).set("value", str(v))
    axes = ("X", "Z") if env.dim == 2 else ("X", "Y", "Z")
    for k, ax in enumerate(axes):
        ET.SubElement(e, "min" + ax).set("value", str(float(env.min_x[k])))
    for k, ax in enumerate(axes):
        ET.SubElement(e, "max" + ax).set("value", str(float(env.max_x[k])))
    cols = "Type, x, z, u, w, p, n\n" if env.dim == 2 else "Type, x, y, z, u, v, w, p, n\n"
    rows = [cols]
    d = env.dim
    for i in range(scene.count):
        vals = [str(int(scene.type[i]))] + [repr(float(v)) for v in scene.x[i]] + [repr(float(v)) for v in scene.u[i]] + \
               [repr(float(scene.p[i])), repr(float(scene.n[i]))]
        rows.append(", ".join(vals) + "\n")
    p = ET.SubElement(root, "particles")
    p.set("type", "csv")
    p.text = "".join(rows)
    text = minidom.parseString(ET.tostring(root)).toprettyxml(indent="\t") if pretty else ET.tostring(root, encoding="unicode")
    with open(path, "w") as f:
        f.write(text)
    return path


def read_result_csv(path):
    """result/particles_%05d.csv (Main.cpp:31-67) -> dict of arrays."""
    with open(path) as f:
        header = [h.strip() for h in f.readline().split(",")]
        data = np.loadtxt(f, delimiter=",", ndmin=2)
    dim = 3 if "y" in header else 2
    col = {h: k for k, h in enumerate(header)}
    xs = ["x", "z"] if dim == 2 else ["x", "y", "z"]
    us = ["u", "w"] if dim == 2 else ["u", "v", "w"]
    return {"type": data[:, col["Type"]].astype(np.int32), "x": data[:, [col[a] for a in xs]], "u": data[:, [col[a] for a in us]],
            "p": data[:, col["p"]], "n": data[:, col["n"]]}
