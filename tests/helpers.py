"""Shared drivers for the parity tests: run one recorded reference step on any engine (CPU restatement, reference
build or the CUDA library) through the same stage-level calls the upstream gtest fixtures make, and compare."""
import numpy as np

from openmps_b200.scenes import Env, Scene


def env_from_golden(g):
    d = int(g["env_dim"])
    return Env(d, float(g["env_max_dt"]), float(g["env_courant"]), float(g["env_g"]), float(g["env_rho"]), float(g["env_nu"]),
               float(g["env_r_e_by_l0"]), float(g["env_l0"]), tuple(g["env_min_x"].tolist()), tuple(g["env_max_x"].tolist()),
               float(g["env_eps"]), bool(int(g["env_central_gravity"])))


def scene_from_golden(g):
    return Scene(env_from_golden(g), g["in_x"].copy(), g["in_u"].copy(), g["in_p"].copy(), g["in_n"].copy(), g["in_type"].copy())


def open_engine(factory, g):
    """Engine holding the golden input state; walls pinned to the recorded targets (the scene's initial positions)."""
    sc = scene_from_golden(g)
    eng = factory(sc.env)
    eng.add_particles(sc.x, sc.u, sc.p, sc.n, sc.type)
    nonfluid = np.where(sc.type != 0)[0]
    if len(nonfluid):
        eng.set_wall_positions(nonfluid, g["wall_target"][nonfluid])
    return eng


def rel_err(a, b):
    """max |a-b| / max(|b|) over the array (scale-relative, robust to entries that are exactly zero)."""
    a = np.asarray(a, float); b = np.asarray(b, float)
    if a.size == 0:
        return 0.0
    scale = np.abs(b).max()
    d = np.abs(a - b).max()
    return float(d / scale) if scale > 0 else float(d)


def max_ulp(a, b):
    a = np.ascontiguousarray(a, np.float64).ravel(); b = np.ascontiguousarray(b, np.float64).ravel()
    if a.size == 0:
        return 0
    ia = a.view(np.int64).copy(); ib = b.view(np.int64).copy()
    ia[ia < 0] = np.int64(-2**63) - ia[ia < 0]
    ib[ib < 0] = np.int64(-2**63) - ib[ib < 0]
    return int(np.abs(ia.astype(object) - ib.astype(object)).max()) if a.size < 200000 else int(np.abs(ia - ib).max())


def csr_rows(rowptr, col, val):
    return [dict(zip(col[int(rowptr[i]):int(rowptr[i + 1])].tolist(), val[int(rowptr[i]):int(rowptr[i + 1])].tolist()))
            for i in range(len(rowptr) - 1)]


def csr_matvec(rowptr, col, val, x):
    n = len(rowptr) - 1
    rp = np.asarray(rowptr, np.int64)
    rows = np.repeat(np.arange(n), np.diff(rp))
    out = np.zeros(n)
    np.add.at(out, rows, val * x[col])
    return out


class StepReport(dict):
    """stage -> metric values collected while replaying a golden step (kept for printing on failure)."""


def check_iterations(got, want, plain_cg):
    """plain CG (the reference's algorithm, MPS_CG_PRECOND=0) must take the reference's iteration count to within 10 %;
    the preconditioned solve must not need more than that."""
    if plain_cg:
        assert abs(got - want) <= max(3, want // 10), (got, want)
    else:
        assert got <= want + 3, (got, want)


def replay_golden_step(eng, g, exact, resync=True, tol=1e-12, cg_tol=1e-6, plain_cg=True):
    """Replays the recorded reference step stage by stage on ``eng`` and asserts parity after every stage.

    exact=True  : floating-point outputs must be bit-identical (CPU restatement vs reference).
    exact=False : integers (cells, neighbour lists, types, CSR pattern) bit-identical; floats within ``tol`` relative
                  to the field's scale; the CG solution within ``cg_tol`` and inside the reference's stopping rule.
    resync      : after a floating-point stage, overwrite the engine state with the recorded one so that every stage
                  is checked on identical inputs (errors cannot compound or mask each other).
    """
    rep = StepReport()
    dt = float(g["dt"])
    assert eng.determine_dt() == dt if exact else abs(eng.determine_dt() - dt) <= 1e-15 * dt, "DetermineDt"
    eng.set_dt(dt, True)

    def cmp(tag, got, want):
        if exact:
            assert np.array_equal(got, want), f"{tag}: not bit-identical (rel {rel_err(got, want):.3e})"
            rep[tag] = 0.0
        else:
            e = rel_err(got, want)
            rep[tag] = e
            rep[tag + "_ulp"] = max_ulp(got, want)
            assert e <= tol, f"{tag}: relative error {e:.3e} > {tol:.1e}"

    # --- SearchNeighbor: integers, always bit-exact
    eng.stage("search")
    assert np.array_equal(eng.state()["type"], g["search_type"]), "types after search (Disable on leaving the grid)"
    alive = g["search_type"] != 3
    assert np.array_equal(eng.cells()[alive], g["cells"][alive]), "cell ids"
    rp, idx = eng.neighbors()
    assert np.array_equal(rp, g["nbr_rowptr"]), "neighbour counts"
    assert np.array_equal(idx.astype(np.uint32), g["nbr_idx"]), "neighbour lists (order included)"

    eng.stage("density")
    cmp("density1_n", eng.state()["n"], g["density1_n"]); cmp("density1_nws", eng.vec("nWithoutSpp"), g["density1_nws"])
    eng.stage("ecs")
    act = (g["search_type"] != 2) & (g["search_type"] != 3)
    cmp("ecs", eng.vec("ecs")[act], g["ecs"][act])
    eng.stage("explicit")
    s = eng.state()
    cmp("explicit_x", s["x"], g["explicit_x"]); cmp("explicit_u", s["u"], g["explicit_u"])
    if resync and not exact:
        eng.set_state(x=g["explicit_x"], u=g["explicit_u"])
    eng.stage("density")
    cmp("density2_n", eng.state()["n"], g["density2_n"]); cmp("density2_nws", eng.vec("nWithoutSpp"), g["density2_nws"])
    eng.stage("savex")
    eng.stage("setppe")
    rp, col, val = eng.csr()
    assert np.array_equal(np.asarray(rp, np.uint64), g["csr_rowptr"].astype(np.uint64)), "CSR row pointers"
    assert np.array_equal(col, g["csr_col"]), "CSR pattern"
    cmp("csr_val", val, g["csr_val"])
    cmp("ppe_b", eng.vec("b"), g["ppe_b"]); cmp("ppe_x0", eng.vec("x"), g["ppe_x0"])
    eng.stage("solveppe")
    x = eng.vec("x")
    if exact:
        assert np.array_equal(x, g["ppe_x"]), "CG solution"
        assert eng.last_iterations() == int(g["cg_iterations"])
    else:
        rep["cg_x"] = rel_err(x, g["ppe_x"])
        rep["cg_iterations"] = (eng.last_iterations(), int(g["cg_iterations"]))
        assert rep["cg_x"] <= cg_tol, f"CG solution differs by {rep['cg_x']:.3e}"
        # inside the reference's own stopping rule: ||b - A x||^2 < eps^2 ||b - A x0||^2 (Computer.hpp:1386,1408)
        A = (g["csr_rowptr"], g["csr_col"], g["csr_val"])
        r0 = g["ppe_b"] - csr_matvec(*A, g["ppe_x0"]); r = g["ppe_b"] - csr_matvec(*A, x)
        eps = float(g["env_eps"])
        rep["cg_residual_ratio"] = float(np.dot(r, r) / max(np.dot(r0, r0), 1e-300))
        assert np.dot(r, r) <= 4.0 * eps * eps * np.dot(r0, r0) + 1e-300, "true residual outside the stopping tolerance"
        check_iterations(eng.last_iterations(), int(g["cg_iterations"]), plain_cg)
    if exact:
        eng.stage("implicit")
    else:
        eng.stage("pressure")
        p = eng.state()["p"]
        rep["pressure"] = rel_err(p, g["implicit_p"])
        assert rep["pressure"] <= cg_tol
        if resync:
            eng.set_state(p=g["implicit_p"])
        eng.stage("gradient")
    s = eng.state()
    ptol = tol if resync else 1e-7
    if exact:
        cmp("implicit_x", s["x"], g["implicit_x"]); cmp("implicit_u", s["u"], g["implicit_u"]); cmp("implicit_p", s["p"], g["implicit_p"])
    else:
        rep["implicit_x"] = rel_err(s["x"], g["implicit_x"]); rep["implicit_u"] = rel_err(s["u"], g["implicit_u"])
        assert rep["implicit_x"] <= ptol and rep["implicit_u"] <= ptol, rep
        if resync:
            eng.set_state(x=g["implicit_x"], u=g["implicit_u"])
    eng.stage("ds")
    s = eng.state()
    if exact:
        for k in ("x", "u", "p", "n"):
            cmp("out_" + k, s[k], g["out_" + k])
        assert eng.determine_dt() == float(g["next_dt"])
    else:
        rep["out_x"] = rel_err(s["x"], g["out_x"]); rep["out_u"] = rel_err(s["u"], g["out_u"])
        assert rep["out_x"] <= ptol and rep["out_u"] <= ptol, rep
        assert abs(eng.determine_dt() - float(g["next_dt"])) <= 1e-9 * float(g["next_dt"])
    assert np.array_equal(s["type"], g["out_type"])
    return rep
