#!/usr/bin/env python
"""CPU study for SURVEY.md 8f rank 2 (preconditioned CG): which preconditioner would cut the CG iterations of the MPS pressure
Poisson equation, and by how much, on matrices taken from the CPU restatement of the reference.  It lives under tests/ because it uses the oracle, which only
tests, smoke() and the CPU arms of bench.py may touch.

For one developed state of a dam break: the system A x = b of the step (active rows only), solved with the reference's stopping
rule (||r||^2 < ||r0||^2 eps^2, warm start from the previous pressure) by plain CG and by PCG with candidate preconditioners.
Reported: iterations, and the number of matrix-sized memory sweeps per iteration each one costs on the GPU (the CG kernel is
bound by matrix bytes), i.e. the bound on the speed-up a streaming implementation could reach.

usage: tests/studies/precond_study.py [l0=1.6e-3] [steps=40]
"""
import os
import sys
import time

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from openmps_b200 import scenes  # noqa: E402
from oracle import bind  # noqa: E402


def pcg(A, b, x0, eps, M=None, maxit=None, flexible=False):
    """Textbook PCG with the reference's stopping rule on the TRUE recursive residual (Computer.hpp:1382-1428).
    flexible=True uses the Polak-Ribiere beta, which tolerates a preconditioner that changes from iteration to iteration
    (an inexact inner solve)."""
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
        z_old, r_old = (z, r + alpha * Ap) if flexible else (None, None)
        z = M(r) if M else r
        rz_new = r @ z
        beta = (rz_new - (r_old @ z)) / rz if flexible else rz_new / rz
        p = z + beta * p
        rz = rz_new
    return x, maxit or n


def main():
    l0 = float(sys.argv[1]) if len(sys.argv) > 1 else 1.6e-3
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    sc = scenes.dambreak2d_fast(l0)
    eng = bind.PortComputer.from_scene(sc)
    eng.forward(steps)
    # the system of the NEXT step, assembled stage by stage like ForwardTime (Computer.hpp:1708-1741)
    eng.set_dt(eng.determine_dt(), True)
    for st in ("search", "density", "ecs", "explicit", "density", "savex", "setppe"):
        eng.stage(st)
    rowptr, col, val = eng.csr()
    n = sc.count
    A = sp.csr_matrix((val, col.astype(np.int64), rowptr.astype(np.int64)), shape=(n, n))
    b = eng.vec("b"); x0 = eng.vec("x")
    # active rows: everything but the identity rows of Dummy / Disabled particles
    t = eng.state()["type"]
    act = np.flatnonzero((t != 2) & (t != 3))
    A = A[act][:, act].tocsr(); b = b[act]; x0 = x0[act]
    eps = sc.env.eps
    print(f"dam break l0={l0}: {n} particles, {len(act)} active rows, {A.nnz} entries ({A.nnz / len(act):.1f} per row), after {steps} steps")
    sym = abs(A - A.T).max()
    print(f"max |A - A^T| = {sym:.3e}; diagonal sign: {np.sign(A.diagonal()).min():.0f}..{np.sign(A.diagonal()).max():.0f}")
    if A.diagonal().mean() < 0:
        A = -A; b = -b      # the reference's matrix is negative definite on these rows; CG is invariant, preconditioners want SPD

    d = A.diagonal()
    results = []

    def run(name, M, sweeps, flexible=False):
        t0 = time.perf_counter()
        x, it = pcg(A, b, x0, eps, M, flexible=flexible)
        res = np.linalg.norm(b - A @ x) / max(np.linalg.norm(b - A @ x0), 1e-300)
        results.append((name, it, sweeps, res, time.perf_counter() - t0))
        print(f"  {name:34s} iterations {it:6d}   matrix sweeps / iteration {sweeps:4.1f}   true rel. residual {res:.2e}")

    run("plain CG (the reference)", None, 1.0)
    run("Jacobi", lambda r: r / d, 1.0)
    # symmetric Gauss-Seidel (SSOR, omega = 1): two triangular solves = sequential on a GPU; listed as the quality yard-stick
    L = sp.tril(A, 0).tocsr(); U = sp.triu(A, 0).tocsr()
    run("SSOR(1) [sequential solves]", lambda r: spla.spsolve_triangular(U, d * spla.spsolve_triangular(L, r, lower=True), lower=False), 2.0)
    # Chebyshev polynomial in D^-1 A of degree k: k extra SpMVs per iteration, no dot products, fully parallel
    Dinv = sp.diags(1.0 / d)
    lam_max = spla.eigsh(Dinv @ A, k=1, which="LM", return_eigenvectors=False, tol=1e-3)[0] * 1.05
    for k, frac in ((2, 8.0), (4, 16.0), (8, 32.0)):
        lam_min = lam_max / frac
        theta, delta = (lam_max + lam_min) / 2, (lam_max - lam_min) / 2

        def cheb(r, k=k, theta=theta, delta=delta):
            # k steps of the Chebyshev iteration for (D^-1 A) z = D^-1 r from z = 0
            rhs = r / d
            sigma = theta / delta
            rho = 1.0 / sigma
            z = rhs / theta
            dz = z.copy()
            for _ in range(k - 1):
                res = rhs - (A @ z) / d
                rho_new = 1.0 / (2 * sigma - rho)
                dz = rho_new * rho * dz + (2 * rho_new / delta) * res
                z = z + dz
                rho = rho_new
            return z
        run(f"Chebyshev(D^-1 A), degree {k}", cheb, float(k))
    # two-level: plain aggregation of `agg` consecutive rows (slots are cell-sorted, so consecutive rows are neighbours), Galerkin coarse
    # operator solved exactly, additive with Jacobi
    for agg in (16, 64):
        m = len(act)
        nc = (m + agg - 1) // agg
        P = sp.csr_matrix((np.ones(m), (np.arange(m), np.arange(m) // agg)), shape=(m, nc))
        Ac = (P.T @ A @ P).tocsc()
        lu = spla.splu(Ac)
        run(f"Jacobi + exact coarse solve (1:{agg})", lambda r, P=P, lu=lu: r / d + P @ lu.solve(P.T @ r), 1.0 + 2.0 / agg)
    # the same two-level method with an INEXACT coarse solve: m CG iterations on the coarse operator from a zero guess (what a GPU
    # implementation can afford: a coarse iteration moves ~1/agg of the fine matrix), outer iteration flexible
    for agg, m_inner in ((16, 10), (16, 30), (16, 60), (8, 30)):
        m = len(act)
        nc = (m + agg - 1) // agg
        P = sp.csr_matrix((np.ones(m), (np.arange(m), np.arange(m) // agg)), shape=(m, nc))
        Ac = (P.T @ A @ P).tocsr()
        dc = Ac.diagonal()

        def coarse(rc, Ac=Ac, dc=dc, m_inner=m_inner):
            xc = np.zeros_like(rc)
            r = rc.copy(); z = r / dc; p = z.copy(); rz = r @ z
            for _ in range(m_inner):
                Ap = Ac @ p
                a = rz / (p @ Ap)
                xc += a * p; r -= a * Ap
                z = r / dc; rz_new = r @ z
                p = z + (rz_new / rz) * p; rz = rz_new
            return xc
        cost = 1.0 + m_inner * (Ac.nnz / A.nnz) + 2.0 / agg
        run(f"Jacobi + {m_inner} coarse PCG its (1:{agg}), flexible", lambda r, P=P, coarse=coarse: r / d + P @ coarse(P.T @ r), cost, flexible=True)
    base = results[0][1]
    print("\nupper bound on the CG-kernel speed-up if iterations cost `sweeps` matrix passes each (coarse solves taken as free):")
    for name, it, sweeps, res, sec in results:
        print(f"  {name:34s} {base / (it * sweeps):5.2f}x")


if __name__ == "__main__":
    main()
