#!/usr/bin/env python
"""Second preconditioner study (round 2): multilevel candidates on the cell hierarchy, on systems dumped by ppe_dump.py in the GPU's
cell-sorted slot order.  Reports PCG iterations with the reference's stopping rule (true recursive residual,
||r||^2 < eps^2 ||r0||^2, warm start).  Test infrastructure only.

usage: tests/studies/precond_study2.py dump.npz
"""
import sys
import time

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla


def pcg(A, b, x0, eps, M=None, maxit=None):
    x = x0.copy()
    r = b - A @ x
    rr0 = r @ r
    tol = rr0 * eps * eps
    if tol == 0:
        return x, 0
    z = M(r) if M else r
    p = z.copy()
    rz = r @ z
    n = len(b)
    for it in range(1, (maxit or n) + 1):
        Ap = A @ p
        alpha = rz / (p @ Ap)
        x += alpha * p
        r -= alpha * Ap
        if r @ r < tol:
            return x, it
        z = M(r) if M else r
        rz_new = r @ z
        p = z + (rz_new / rz) * p
        rz = rz_new
    return x, maxit or n


def load(path):
    d = np.load(path)
    n = len(d["b"])
    A = sp.csr_matrix((d["data"], d["indices"], d["indptr"]), shape=(n, n))
    t = d["type"]
    act = np.flatnonzero((t != 2) & (t != 3))
    A = A[act][:, act].tocsr()
    b = d["b"][act]; x0 = d["x0"][act]
    if A.diagonal().mean() < 0:
        A = -A; b = -b
    return A, b, x0, float(d["eps"]), d["cells"][act], d["grid_n"], int(d["dim"]), act


def agg_matrix(ids):
    """piecewise-constant prolongation from aggregate ids (arbitrary integers) -> (P, compact ids)"""
    u, inv = np.unique(ids, return_inverse=True)
    m = len(ids)
    return sp.csr_matrix((np.ones(m), (np.arange(m), inv)), shape=(m, len(u))), inv


def main():
    A, b, x0, eps, cells, grid_n, dim, act = load(sys.argv[1])
    m = A.shape[0]
    d = A.diagonal()
    print(f"{m} active rows, {A.nnz} entries ({A.nnz / m:.1f}/row), dim {dim}, grid {grid_n}")
    results = []

    def run(name, M, note=""):
        t0 = time.perf_counter()
        x, it = pcg(A, b, x0, eps, M, maxit=20000)
        res = np.linalg.norm(b - A @ x) / max(np.linalg.norm(b - A @ x0), 1e-300)
        results.append((name, it))
        print(f"  {name:58s} its {it:6d}  res {res:.1e}  {time.perf_counter() - t0:6.1f}s {note}", flush=True)

    run("plain CG", None)
    run("Jacobi", lambda r: r / d)

    # ---- aggregates ----
    def cell_ids(shift):
        key = np.zeros(m, np.int64)
        for k in range(dim):
            key = key * (int(grid_n[k]) + 2) + (cells[:, k] >> shift)
        return key

    # two-level, exact coarse
    for name, ids in (("16 consecutive rows", np.arange(m) // 16), ("8 consecutive rows", np.arange(m) // 8),
                      ("cells (1x1)", cell_ids(0)), ("cells 2x2", cell_ids(1))):
        P, _ = agg_matrix(ids)
        Ac = (P.T @ A @ P).tocsc()
        lu = spla.splu(Ac)
        run(f"Jacobi + exact coarse [{name}: {P.shape[1]} aggr]", lambda r, P=P, lu=lu: r / d + P @ lu.solve(P.T @ r))

    # ---- multilevel hierarchy on the cell quadtree: level 1 = cells, level l = 2^(l-1) x 2^(l-1) cells ----
    levels = []  # (P_l from level l-1 to l, A_l, d_l)
    ids = cell_ids(0)
    P1, inv = agg_matrix(ids)
    # coordinates of coarse unknowns at level 1
    ccoord = np.zeros((P1.shape[1], dim), np.int64)
    ccoord[inv] = cells
    Al = (P1.T @ A @ P1).tocsr()
    levels.append((P1, Al, Al.diagonal()))
    while Al.shape[0] > 64:
        ccoord2 = ccoord >> 1
        key = np.zeros(len(ccoord), np.int64)
        for k in range(dim):
            key = key * (int(grid_n[k]) + 2) + ccoord2[:, k]
        Pl, inv = agg_matrix(key)
        nc = Pl.shape[1]
        cc = np.zeros((nc, dim), np.int64); cc[inv] = ccoord2
        ccoord = cc
        Al = (Pl.T @ Al @ Pl).tocsr()
        levels.append((Pl, Al, Al.diagonal()))
    print("  hierarchy:", [lv[1].shape[0] for lv in levels], " nnz:", [lv[1].nnz for lv in levels])
    lu_top = spla.splu(levels[-1][1].tocsc())

    def bpx(r, w=1.0, top_exact=True):
        # additive over all levels: z = D0^-1 r + sum_l P..P D_l^-1 P^T..P^T r
        rs = [r]
        for (P, _, _) in levels:
            rs.append(P.T @ rs[-1])
        L = len(levels)
        zc = lu_top.solve(rs[L]) if top_exact else rs[L] / levels[-1][2]
        for l in range(L - 1, -1, -1):
            P = levels[l][0]
            zl = P @ zc
            if l > 0:
                zl = zl + w * rs[l] / levels[l - 1][2]
            zc = zl
        return r / d + zc

    run("BPX additive, all levels Jacobi (w=1), top exact", lambda r: bpx(r))
    run("BPX additive, w=0.5 on coarse levels", lambda r: bpx(r, 0.5))

    # V-cycle on the coarse hierarchy (levels >= 1), additive with fine Jacobi
    def vcycle(l, rl, nu=1, omega=0.7):
        # solve A_l e = rl approximately; l indexes levels[] (0 = cells)
        Al, dl = levels[l][1], levels[l][2]
        if l == len(levels) - 1:
            return lu_top.solve(rl)
        e = omega * rl / dl
        for _ in range(nu - 1):
            e = e + omega * (rl - Al @ e) / dl
        res = rl - Al @ e
        P = levels[l + 1][0]
        e = e + P @ vcycle(l + 1, P.T @ res, nu, omega)
        for _ in range(nu):
            e = e + omega * (rl - Al @ e) / dl
        return e

    for nu, om in ((1, 0.7), (2, 0.7), (1, 0.9)):
        run(f"Jacobi + P1 V({nu},{nu})-cycle(w={om}) on cells hierarchy", lambda r, nu=nu, om=om: r / d + levels[0][0] @ vcycle(0, levels[0][0].T @ r, nu, om))

    # coarse (cells) solved by k Chebyshev-Jacobi iterations
    A1, d1 = levels[0][1], levels[0][2]
    lam_max = spla.eigsh(sp.diags(1 / d1) @ A1, k=1, which="LM", return_eigenvectors=False, tol=1e-3)[0] * 1.05
    for k, frac in ((5, 30.0), (10, 100.0), (20, 400.0)):
        lam_min = lam_max / frac
        theta, delta = (lam_max + lam_min) / 2, (lam_max - lam_min) / 2

        def cheb(rc, k=k, theta=theta, delta=delta):
            rhs = rc / d1
            sigma = theta / delta
            rho = 1.0 / sigma
            z = rhs / theta
            dz = z.copy()
            for _ in range(k - 1):
                res = rhs - (A1 @ z) / d1
                rho_new = 1.0 / (2 * sigma - rho)
                dz = rho_new * rho * dz + (2 * rho_new / delta) * res
                z = z + dz
                rho = rho_new
            return z
        run(f"Jacobi + P1 Chebyshev({k}) on the cell level", lambda r, cheb=cheb: r / d + levels[0][0] @ cheb(levels[0][0].T @ r))

    # two-level on cells + exact on 2x2-cells beneath (three-level, middle level Jacobi, additive)
    def three(r):
        P1 = levels[0][0]
        r1 = P1.T @ r
        z1 = r1 / d1
        if len(levels) > 1:
            P2 = levels[1][0]
            A2 = levels[1][1]
            z1 = z1 + P2 @ three.lu2.solve(P2.T @ r1)
        return r / d + P1 @ z1
    if len(levels) > 1:
        three.lu2 = spla.splu(levels[1][1].tocsc())
        run("Jacobi + cells Jacobi + exact 2x2 cells (additive 3-level)", three)

    # multiplicative V-cycle INCLUDING the fine level (costs 2 extra fine sweeps): yard-stick
    def vfine(r, omega=0.7):
        e = omega * r / d
        res = r - A @ e
        P = levels[0][0]
        e = e + P @ vcycle(0, P.T @ res, 1, omega)
        e = e + omega * (r - A @ e) / d
        return e
    run("full V(1,1) incl. fine level [3 sweeps/iteration]", vfine)


if __name__ == "__main__":
    main()
