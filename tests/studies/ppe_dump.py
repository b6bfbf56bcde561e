#!/usr/bin/env python
"""Dump the pressure Poisson system of a developed dam-break state (CPU restatement of the reference) in the GPU's
cell-sorted slot order, for the preconditioner studies.  Test infrastructure: uses the oracle.

usage: tests/studies/ppe_dump.py out.npz [scene=dambreak2d] [l0=8e-4] [steps=40]
"""
import os
import sys

import numpy as np
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from openmps_b200 import scenes  # noqa: E402
from oracle import bind  # noqa: E402


def dump(out, scene="dambreak2d", l0=8e-4, steps=40):
    if scene == "dambreak2d":
        sc = scenes.dambreak2d_fast(l0)
    elif scene == "dambreak3d":
        sc = scenes.dambreak3d(l0)
    elif scene == "central_gravity":
        sc = scenes.central_gravity(int(l0))
    else:
        raise SystemExit("unknown scene")
    eng = bind.PortComputer.from_scene(sc)
    eng.forward(steps)
    eng.set_dt(eng.determine_dt(), True)
    for st in ("search", "density", "ecs", "explicit", "density", "savex", "setppe"):
        eng.stage(st)
    rowptr, col, val = eng.csr()
    n = sc.count
    A = sp.csr_matrix((val, col.astype(np.int64), rowptr.astype(np.int64)), shape=(n, n))
    b = eng.vec("b"); x0 = eng.vec("x")
    st = eng.state()
    t = st["type"]
    cells = eng.cells()                      # (n, dim) cell index per axis at the search of this step
    envv = eng.env_values()
    # slot order: x-major ... z-minor key, ties by original id (stable sort); Disabled go last
    grid_n = [int(v) for v in envv["grid_cells"][: sc.env.dim]] if isinstance(envv, dict) and "grid_cells" in envv else None
    if grid_n is None:
        grid_n = [int(cells[:, k].max()) + 3 for k in range(sc.env.dim)]
    key = np.zeros(n, np.int64)
    for k in range(sc.env.dim):
        key = key * grid_n[k] + cells[:, k]
    key[t == 3] = np.iinfo(np.int64).max
    order = np.argsort(key, kind="stable")
    A = A[order][:, order].tocsr()
    np.savez_compressed(out, indptr=A.indptr, indices=A.indices, data=A.data, b=b[order], x0=x0[order], type=t[order],
                        cells=cells[order], grid_n=np.array(grid_n), eps=sc.env.eps, dim=sc.env.dim, x=st["x"][order])
    print(f"{out}: {n} particles, {A.nnz} entries, grid {grid_n}")


if __name__ == "__main__":
    a = sys.argv
    dump(a[1], a[2] if len(a) > 2 else "dambreak2d", float(a[3]) if len(a) > 3 else 8e-4, int(a[4]) if len(a) > 4 else 40)
