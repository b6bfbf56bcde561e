"""GPU tier: the multigrid preconditioner of the PPE solve (csrc/mps_mg.cu, csrc/mps_cg.cu k_pcg_stream).

(1) the device-built cell hierarchy (compact ids, neighbour / parent tables, Galerkin stencils of every level) equals the one
    computed on the host from the assembled matrix: S_0 = P^T A P with P = rows -> cells, S_{l+1} = P_l^T S_l P_l;
(2) the device solve takes the iteration count of a numpy PCG that applies the same preconditioner with the same tables
    (tests/mg_model.py) — i.e. the kernel applies what mps_mg.cu describes;
(3) both land inside the reference's stopping rule."""
import numpy as np
import pytest
import scipy.sparse as sp

import mg_model
from openmps_b200 import capi, scenes

pytestmark = pytest.mark.gpu


def _assembled(sc, steps):
    g = capi.GpuComputer.from_scene(sc)
    if steps:
        g.forward(steps)
    g.set_dt(g.determine_dt() if steps else sc.env.max_dt, True)
    for st in ("search", "density", "ecs", "explicit", "density", "savex", "setppe"):
        g.stage(st)
    return g


def _slot_system(g, sc):
    """A, b, x0 in the device's slot order + every row's cell key as the sort saw it."""
    n = sc.count
    crp, col, val = g.csr()
    A = sp.csr_matrix((val, col.astype(np.int64), crp.astype(np.int64)), shape=(n, n))
    orig = g.mg_table(-1, 3).astype(np.int64)        # slot -> original id
    A = A[orig][:, orig].tocsr()
    # drop the reference layout's identity rows of Dummy / Disabled particles: the device keeps them empty
    t = g.state()["type"][orig]
    act = (t != 2) & (t != 3)
    Dm = sp.diags(act.astype(float))
    A = (Dm @ A @ Dm).tocsr()
    return A, g.vec("b")[orig], g.vec("x")[orig], orig, act


@pytest.mark.parametrize("make,steps", [(lambda: scenes.dambreak2d_fast(1.6e-3), 25), (lambda: scenes.dambreak3d(l0=1.6e-2), 5)])
def test_hierarchy_equals_host_galerkin_and_solve_matches_model(make, steps, monkeypatch):
    monkeypatch.setenv("MPS_CG_ADAPTIVE", "0")
    sc = make()
    g = _assembled(sc, steps)
    dim = sc.env.dim
    K = 3 ** dim
    A, b, x0, orig, act = _slot_system(g, sc)
    dims0 = [int(v) for v in g.env_info().grid_cells[:dim]]
    crow = g.mg_table(-1, 0)
    keys0 = g.mg_table(0, 0).astype(np.int64)
    row_key = np.where(crow == mg_model.NONE, -1, keys0[np.minimum(crow, len(keys0) - 1).astype(np.int64)])
    # rows of one cell are contiguous and cstart delimits them
    cstart = g.mg_table(-1, 1).astype(np.int64)
    for c in (0, len(keys0) // 2, len(keys0) - 1):
        assert np.all(crow[cstart[c]:cstart[c + 1]] == c)
    assert cstart[-1] == np.count_nonzero(crow != mg_model.NONE)
    levels_dev = []
    l = 0
    while True:
        try:
            key = g.mg_table(l, 0)
        except capi.MpsError:
            break
        lv = {"key": key, "nbr": g.mg_table(l, 1).reshape(-1, K), "S": g.mg_table(l, 4).reshape(-1, K), "dinv": g.mg_table(l, 5)}
        par = g.mg_table(l, 3)
        if len(par):
            lv["parent"] = par
        levels_dev.append(lv)
        l += 1
        if len(key) <= 1:
            break
    levels_host, cell_of_row = mg_model.host_hierarchy(A, row_key, dims0, len(levels_dev), 0.8)
    assert np.array_equal(np.where(crow == mg_model.NONE, -1, crow.astype(np.int64)), cell_of_row)
    for l, (d, h) in enumerate(zip(levels_dev, levels_host)):
        assert np.array_equal(d["key"], h["key"]), f"level {l}: occupied cells"
        assert np.array_equal(d["nbr"], h["nbr"]), f"level {l}: neighbour table"
        if "parent" in h and "parent" in d:
            assert np.array_equal(d["parent"], h["parent"]), f"level {l}: parents"
        scale = np.abs(h["S"]).max()
        assert np.abs(d["S"] - h["S"]).max() <= 1e-12 * scale, f"level {l}: Galerkin stencil (max diff {np.abs(d['S'] - h['S']).max() / scale:.2e})"
        assert np.allclose(d["dinv"], h["dinv"], rtol=1e-11, atol=0), f"level {l}: damped inverse diagonal"
    # 1 / a_ii
    diag = A.diagonal()
    dinv0 = g.mg_table(-1, 2)
    assert np.allclose(dinv0[act & (diag != 0)], 1.0 / diag[act & (diag != 0)], rtol=1e-13) and np.all(dinv0[~act] == 0)

    # the solve: same iteration count as the numpy PCG with the same preconditioner
    def M(r):
        ok = crow != mg_model.NONE
        r1 = np.bincount(crow[ok].astype(np.int64), weights=r[ok], minlength=len(keys0))
        e0, _ = mg_model.vcycle(levels_dev, r1, 1.8, 4, 64)
        z = r * dinv0
        z[ok] += e0[crow[ok].astype(np.int64)]
        return z
    x_model, it_model = mg_model.pcg(A, b, x0, sc.env.eps, M, maxit=5000)
    _, it_plain = mg_model.pcg(A, b, x0, sc.env.eps, lambda r: r, maxit=20000)
    g.stage("solveppe")
    x_dev = g.vec("x")[orig]
    it_dev = g.last_iterations()
    print(f"iterations: device {it_dev}, numpy model {it_model}, plain CG {it_plain}; levels {g.stats().mg_levels}")
    r0 = b - A @ x0; r = b - A @ x_dev
    assert r @ r <= 4.0 * sc.env.eps ** 2 * (r0 @ r0)
    assert abs(it_dev - it_model) <= max(2, it_model // 10), (it_dev, it_model)
    assert it_dev * 4 <= it_plain, (it_dev, it_plain)
    scale = np.abs(x_model).max()
    assert np.abs(x_dev[act] - x_model[act]).max() <= 1e-6 * scale
