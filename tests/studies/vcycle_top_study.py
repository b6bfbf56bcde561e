#!/usr/bin/env python
"""Study of the V-cycle's coarse end on the CPU (test infrastructure; uses the CPU restatement to assemble real systems and
tests/mg_model.py, the numpy model of the device's preconditioner, to count PCG iterations with the reference's stopping rule).

Question: on the dam break a less accurate coarsest level gives FEWER iterations, on StaticPressure and
CentralGravity more.  Variants: extra sweeps on the top level, cutting the hierarchy earlier, exact top solve, over-correction
gamma per level.

usage: tests/studies/vcycle_top_study.py [scene ...]
"""
import os
import sys
import time

import numpy as np
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import mg_model  # noqa: E402
from openmps_b200 import scenes  # noqa: E402
from oracle import bind  # noqa: E402

SCENES = {
    "dambreak2d_18k": (lambda: scenes.dambreak2d_fast(1.6e-3), 25),
    "dambreak2d_72k": (lambda: scenes.dambreak2d_fast(8e-4), 5),
    "dambreak2d_72k_dev": (lambda: scenes.dambreak2d_fast(8e-4), 150),
    "static_pressure": (lambda: scenes.static_pressure(), 5),
    "central_gravity_58k": (lambda: scenes.central_gravity(half=120), 5),
    "dambreak3d_123k": (lambda: scenes.dambreak3d(8e-3), 3),
}


def system(name):
    make, steps = SCENES[name]
    sc = make()
    p = bind.PortComputer.from_scene(sc)
    p.forward(steps)
    p.set_dt(p.determine_dt(), True)
    for st in ("search", "density", "ecs", "explicit", "density", "savex", "setppe"):
        p.stage(st)
    n = sc.count
    rp, col, val = p.csr()
    A = sp.csr_matrix((val, col.astype(np.int64), rp.astype(np.int64)), shape=(n, n))
    t = p.state()["type"]
    act = (t != 2) & (t != 3)
    Dm = sp.diags(act.astype(float))
    A = (Dm @ A @ Dm).tocsr()
    b = p.vec("b") * act; x0 = p.vec("x") * act
    cells = p.cells()                                   # (n, dim) cell coordinates, Grid::Block
    ev = p.env_values()
    dim = sc.env.dim
    nl = ev["neighbor_length"] if "neighbor_length" in ev else sc.env.r_e_by_l0 * sc.env.l0 * (1 + 2 * sc.env.courant)
    dims0 = [int(np.ceil((sc.env.max_x[a] - sc.env.min_x[a]) / nl)) + 2 for a in range(dim)]
    key = mg_model.encode(cells.astype(np.int64), dims0)
    row_key = np.where(act, key, -1)
    return sc, A, b, x0, row_key, dims0, act


def variants(A, b, x0, eps, row_key, dims0, act):
    nlev = 12
    levels, cell_of_row = mg_model.host_hierarchy(A, row_key, dims0, nlev, 0.8)
    sizes = [len(l["dinv"]) for l in levels]
    diag = A.diagonal()
    dinv0 = np.where(act & (diag != 0), 1.0 / np.where(diag != 0, diag, 1.0), 0.0)
    ok = cell_of_row >= 0
    crow = cell_of_row[ok]
    ncell = sizes[0]

    def make_M(cycle):
        def M(r):
            r1 = np.bincount(crow, weights=r[ok], minlength=ncell)
            e0 = cycle(r1)
            z = r * dinv0
            z[ok] += e0[crow]
            return z
        return M

    def vc(gammas, top_sweeps, top_cells, exact_top=False):
        def cycle(r0):
            L = 1
            while L < len(levels) and len(levels[L - 1]["dinv"]) > top_cells and "parent" in levels[L - 1]:
                L += 1
            r = [None] * L; e = [None] * L
            r[0] = r0; e[0] = levels[0]["dinv"] * r0
            for l in range(L - 1):
                lo = levels[l]
                res = r[l] - mg_model._apply(lo["S"], lo["nbr"], e[l])
                r[l + 1] = np.bincount(lo["parent"].astype(np.int64), weights=res, minlength=len(levels[l + 1]["dinv"]))
                e[l + 1] = levels[l + 1]["dinv"] * r[l + 1]
            top = levels[L - 1]
            if exact_top:
                nt = len(top["dinv"])
                Md = np.zeros((nt, nt))
                for c in range(nt):
                    for s in range(top["S"].shape[1]):
                        j = top["nbr"][c, s]
                        if j != mg_model.NONE:
                            Md[c, j] += top["S"][c, s]
                e[L - 1] = np.linalg.lstsq(Md, r[L - 1], rcond=None)[0]
            else:
                for _ in range(top_sweeps):
                    e[L - 1] = e[L - 1] + top["dinv"] * (r[L - 1] - mg_model._apply(top["S"], top["nbr"], e[L - 1]))
            for l in range(L - 2, -1, -1):
                lv = levels[l]
                g = gammas[l] if isinstance(gammas, (list, tuple)) else gammas
                et = e[l] + g * e[l + 1][lv["parent"].astype(np.int64)]
                e[l] = et + lv["dinv"] * (r[l] - mg_model._apply(lv["S"], lv["nbr"], et))
            return e[0]
        return cycle

    out = {}
    _, out["plain"] = mg_model.pcg(A, b, x0, eps, lambda r: r, maxit=20000)
    _, out["jacobi"] = mg_model.pcg(A, b, x0, eps, lambda r: r * dinv0, maxit=20000)
    for ts in (0, 2, 4, 8, 16):
        _, out[f"g1.8 top64 sweeps{ts}"] = mg_model.pcg(A, b, x0, eps, make_M(vc(1.8, ts, 64)), maxit=5000)
    _, out["g1.8 top64 exact"] = mg_model.pcg(A, b, x0, eps, make_M(vc(1.8, 0, 64, True)), maxit=5000)
    for tc in (256, 1024):
        _, out[f"g1.8 top{tc} sweeps4"] = mg_model.pcg(A, b, x0, eps, make_M(vc(1.8, 4, tc)), maxit=5000)
    for g in (1.0, 1.4, 1.6, 2.0):
        _, out[f"g{g} top64 sweeps4"] = mg_model.pcg(A, b, x0, eps, make_M(vc(g, 4, 64)), maxit=5000)
    # over-correction only between the finest levels
    for name, gs in (("g[1.8,1.8,1.4,1.2,1..]", [1.8, 1.8, 1.4, 1.2] + [1.0] * 10), ("g[2,1.6,1.3,1.1,1..]", [2.0, 1.6, 1.3, 1.1] + [1.0] * 10),
                     ("g[1.8 x3, 1..]", [1.8] * 3 + [1.0] * 10), ("g[1.8 x2, 1..]", [1.8] * 2 + [1.0] * 10)):
        _, out[f"{name} top64 exact"] = mg_model.pcg(A, b, x0, eps, make_M(vc(gs, 0, 64, True)), maxit=5000)
        _, out[f"{name} top64 sweeps4"] = mg_model.pcg(A, b, x0, eps, make_M(vc(gs, 4, 64)), maxit=5000)
    return sizes, out


if __name__ == "__main__":
    names = sys.argv[1:] or list(SCENES)
    for name in names:
        t0 = time.time()
        sc, A, b, x0, row_key, dims0, act = system(name)
        sizes, out = variants(A, b, x0, sc.env.eps, row_key, dims0, act)
        print(f"== {name}: {sc.count} particles, {int(act.sum())} rows, level sizes {sizes}  ({time.time() - t0:.0f} s)")
        for k, v in out.items():
            print(f"   {k:42s} {v}")
        sys.stdout.flush()
