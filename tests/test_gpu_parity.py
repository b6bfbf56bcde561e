"""GPU tier (-m gpu): the CUDA path, called through the C ABI, against the committed golden fixtures (generated from
the unmodified reference) and against the CPU restatement on fresh seeded inputs.

Bar (BASELINE.json north star): neighbour lists / cell ids bit-exact; number density, explicit forces, positions
<= 1e-10 relative after one step in FP64 (asserted at 1e-12, stage by stage, on identical inputs); the PPE solution
within the CG stopping tolerance; the same error behaviour as the reference's exceptions.
"""
import numpy as np
import pytest

from conftest import golden_names
from helpers import check_iterations, csr_matvec, open_engine, rel_err, replay_golden_step
from openmps_b200 import capi, scenes
from oracle import bind

pytestmark = pytest.mark.gpu

TOL = 1e-12       # asserted; the north star asks for 1e-10
CG_TOL = 1e-6     # solution-space agreement of two CG runs that both satisfy ||r||^2 < eps^2 ||r0||^2 with eps = 1e-10


# The PPE is solved by multigrid-preconditioned CG by default; MPS_CG_PRECOND=0 selects the reference's plain CG
# (Computer.hpp:1359-1429).  Both must land inside the reference's stopping rule; only the plain one is expected to take
# the reference's iteration count.
SOLVERS = [pytest.param("1", id="pcg"), pytest.param("0", id="plain_cg")]


@pytest.mark.parametrize("precond", SOLVERS)
@pytest.mark.parametrize("name", golden_names())
def test_golden_step_stage_by_stage(golden, name, precond, monkeypatch):
    monkeypatch.setenv("MPS_CG_PRECOND", precond)
    g = golden(name)
    eng = open_engine(capi.GpuComputer, g)
    rep = replay_golden_step(eng, g, exact=False, resync=True, tol=TOL, cg_tol=CG_TOL, plain_cg=(precond == "0"))
    print(name, {k: (f"{v:.2e}" if isinstance(v, float) else v) for k, v in rep.items()})
    assert (eng.stats().mg_levels > 0) == (precond == "1")


@pytest.mark.parametrize("precond", SOLVERS)
@pytest.mark.parametrize("name", golden_names())
def test_golden_full_step_without_resync(golden, name, precond, monkeypatch):
    """One ForwardTime(dt) call end to end: errors of all stages and of the CG solve compound."""
    monkeypatch.setenv("MPS_CG_PRECOND", precond)
    g = golden(name)
    eng = open_engine(capi.GpuComputer, g)
    eng.forward(1, dt=float(g["dt"]))
    s = eng.state()
    assert np.array_equal(s["type"], g["out_type"])
    assert rel_err(s["x"], g["out_x"]) <= 1e-10
    assert rel_err(s["u"], g["out_u"]) <= 1e-7      # u carries dt/rho * grad(p): bounded by the CG tolerance
    assert rel_err(s["p"], g["out_p"]) <= CG_TOL
    assert rel_err(s["n"], g["out_n"]) <= 1e-10
    assert abs(eng.determine_dt() - float(g["next_dt"])) <= 1e-7 * float(g["next_dt"])
    check_iterations(eng.last_iterations(), int(g["cg_iterations"]), precond == "0")


def _port_and_gpu(sc):
    return bind.PortComputer.from_scene(sc), capi.GpuComputer.from_scene(sc)


@pytest.mark.parametrize("make,steps,xtol", [
    (lambda: scenes.dambreak2d(), 60, 1e-8),
    (lambda: scenes.static_pressure(width=16, height=24), 30, 1e-8),
    (lambda: scenes.static_pressure(), 30, 1e-8),                      # BASELINE.json configs[0] at its default resolution: 6 040 particles
    (lambda: scenes.central_gravity(half=14), 30, 1e-8),
    (lambda: scenes.dambreak3d(l0=0.035), 6, 1e-8),
    (lambda: scenes.lattice(3, 6, 0.1, 2.1, jitter=0.05, max_dt=1e-3), 5, 1e-8),
])
def test_many_steps_against_cpu_restatement(make, steps, xtol):
    """Free-running trajectories (ForwardTime() with DetermineDt) stay together over many steps."""
    sc = make()
    p, g = _port_and_gpu(sc)
    p.forward(steps); g.forward(steps)
    sp, sg = p.state(), g.state()
    assert np.array_equal(sp["type"], sg["type"])
    assert rel_err(sg["x"], sp["x"]) <= xtol
    assert rel_err(sg["n"], sp["n"]) <= 1e-7
    tp, tg = p.env_values()["t"], g.time()[0]
    assert abs(tp - tg) <= 1e-9 * max(tp, 1e-30)
    # neighbour lists of the last step: same sets even though positions agree only to ~1e-9
    a, b = p.neighbors(), g.neighbors()
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])


def test_seeded_jitter_neighbour_lists_bit_exact_2d_and_3d():
    for dim, num in ((2, 40), (3, 14)):
        sc = scenes.lattice(dim, num, 0.1, 2.1, jitter=0.3, margin_cells=0.6, seed=2024 + dim)
        sc.type[::5] = scenes.WALL; sc.type[2::9] = scenes.DUMMY
        p, g = _port_and_gpu(sc)
        p.set_dt(1e-3, True); g.set_dt(1e-3, True)
        p.stage("search"); g.stage("search")
        assert np.array_equal(p.state()["type"], g.state()["type"])
        alive = p.state()["type"] != 3
        assert np.array_equal(p.cells()[alive], g.cells()[alive])
        a, b = p.neighbors(), g.neighbors()
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
        p.stage("density"); g.stage("density")
        assert rel_err(g.state()["n"], p.state()["n"]) <= TOL
        assert rel_err(g.vec("nWithoutSpp"), p.vec("nWithoutSpp")) <= TOL
        # NeighborDensityVariationSpeed(i) for a few particles (Computer.hpp:838-872)
        rng = np.random.default_rng(5)
        u = rng.normal(0, 0.1, sc.x.shape)
        p.set_state(u=u); g.set_state(u=u)
        for i in rng.integers(0, sc.count, 8):
            assert abs(p.dndt(int(i)) - g.dndt(int(i))) <= 1e-12 * max(1.0, abs(p.dndt(int(i))))


@pytest.mark.parametrize("dim", [2, 3])
def test_search_threshold_is_exact_at_the_boundary(dim):
    """k_search tests r2 < T with T the smallest double whose correctly rounded root is >= NL instead of sqrt(r2) < NL
    (EnvConst::nl2_lim).  Adversarial input: pairs at distances NL * (1 + t) for t from a few ulp to +-3e-3 around the threshold,
    far from the origin, in random directions.  The lists must equal the CPU restatement's (which evaluates the reference's
    R(x_i, x_j) < neighborLength, Computer.hpp:743) entry for entry."""
    rng = np.random.default_rng(77 + dim)
    l0, q = 1e-3, 2.4
    nl = q * l0 * 1.2
    span = np.array([4000 * l0] + [40 * l0] * (dim - 1))   # long in x only, so that the grid stays small
    m = 6000
    anchors = rng.uniform(0.25 * span, 0.9 * span, (m, dim))
    t = np.concatenate([np.linspace(-3e-3, 3e-3, m // 2), rng.choice([-1, 1], m // 4) * 2.0 ** -rng.integers(30, 52, m // 4),
                        rng.normal(0, 3e-4, m - m // 2 - m // 4)])
    d = rng.normal(size=(m, dim)); d /= np.linalg.norm(d, axis=1)[:, None]
    partners = anchors + d * (nl * (1.0 + t))[:, None]
    x = np.ascontiguousarray(np.concatenate([anchors, partners]))
    typ = np.zeros(len(x), np.int32); typ[::7] = scenes.WALL
    env = scenes.Env(dim, 1e-3, 0.1, 9.8, 998.2, 1.004e-6, q, l0, tuple([0.0] * dim), tuple(float(v) for v in span), 1e-10)
    sc = scenes.Scene(env, x, np.zeros_like(x), np.zeros(len(x)), np.zeros(len(x)), typ, "threshold")
    p, g = _port_and_gpu(sc)
    p.set_dt(1e-3, True); g.set_dt(1e-3, True)
    p.stage("search"); g.stage("search")
    a, b = p.neighbors(), g.neighbors()
    assert int(a[0][-1]) > m // 3                       # about half of the engineered pairs are inside
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])


def test_first_disable_event_matches():
    sc = scenes.dambreak2d()
    sc.x[10] = (0.3, 0.6355); sc.u[10] = (0.0, 5.0)
    p, g = _port_and_gpu(sc)
    for k in range(6):
        p.forward(1); g.forward(1)
        sp, sg = p.state(), g.state()
        assert np.array_equal(sp["type"], sg["type"]), f"step {k}"
        assert rel_err(sg["x"], sp["x"]) <= 1e-9 and rel_err(sg["u"], sp["u"]) <= 1e-6
        assert abs(p.determine_dt() - g.determine_dt()) <= 1e-9 * p.determine_dt()
    assert (sg["type"] == 3).sum() == 1


def test_cell_overflow_maps_to_grid_exception():
    sc = scenes.lattice(2, 6, 0.1, 2.1)
    sc.x[:] = sc.x[0] + np.random.default_rng(3).uniform(0, 1e-3, sc.x.shape)
    g = capi.GpuComputer.from_scene(sc)
    with pytest.raises(capi.MpsError) as ei:
        g.stage("search")
    assert ei.value.code == capi.MPS_CELL_OVERFLOW and "Too many particle in a block" in ei.value.message


@pytest.mark.parametrize("precond", SOLVERS)
def test_cell_overflow_stops_a_whole_step(precond, monkeypatch):
    """ForwardTime on over-full cells: the reference throws Grid::Exception out of SearchNeighbor (Grid.hpp:311-318) and nothing
    after the sort runs.  Here the step must end with MPS_CELL_OVERFLOW right after the sort (no assembly / CG on over-full
    cells), the error is per call (not sticky), and the handle stays usable."""
    monkeypatch.setenv("MPS_CG_PRECOND", precond)
    sc = scenes.dambreak2d()
    crowd = np.flatnonzero(sc.type == 0)[:40]
    good_x = sc.x.copy()
    sc.x[crowd] = sc.x[crowd[0]] + np.random.default_rng(4).uniform(0, 1e-4, (len(crowd), 2))   # 40 particles in one cell (capacity 16)
    g = capi.GpuComputer.from_scene(sc)
    launches_before = g.stats().kernel_launches
    for _ in range(2):                                    # the second call must report again, not a stale flag of the first
        with pytest.raises(capi.MpsError) as ei:
            g.forward(1)
        assert ei.value.code == capi.MPS_CELL_OVERFLOW and "Too many particle in a block" in ei.value.message
    assert g.stats().kernel_launches - launches_before < 50, "the failed steps must stop after the sort"
    g.set_state(x=good_x)                                 # un-crowd: the same handle steps normally afterwards
    g.forward(2)
    assert np.isfinite(g.state()["x"]).all() and g.last_iterations() > 0


def _dense_to_csr(A):
    n = A.shape[0]
    rowptr = [0]; col = []; val = []
    for i in range(n):
        for j in range(n):
            if A[i, j] != 0:
                col.append(j); val.append(A[i, j])
        rowptr.append(len(col))
    return np.array(rowptr), np.array(col), np.array(val, float)


def test_conjugate_gradient_known_answers():
    """test_ComputerConjugateGradient.cpp:115-229 through the C ABI (eps = 1e-7, answers to 1e-3 as upstream)."""
    from test_oracle import cg_system, poisson_1d
    env = scenes.lattice(2, 2, 1.0, 2.1).env.scaled(eps=1e-7)
    for case in ("identity", "pm1_4x4", "poisson1d"):
        A, b, want = poisson_1d() if case == "poisson1d" else cg_system(case)
        g = capi.GpuComputer(env)
        g.set_system(*_dense_to_csr(A), b, np.zeros(len(b)))
        g.stage("solveppe")
        got = g.solution()
        assert np.allclose(got, want, rtol=0, atol=1e-3), case
        p = bind.PortComputer(env)
        p.set_system(*_dense_to_csr(A), b, np.zeros(len(b)))
        p.stage("solveppe")
        assert abs(g.last_iterations() - p.last_iterations()) <= 2, case
        assert rel_err(got, p.vec("x", n=len(b))) <= 1e-6


def test_conjugate_gradient_failure_maps_to_computer_exception():
    """A singular inconsistent system cannot converge in n iterations -> Computer::Exception text (Computer.hpp:1424-1428)."""
    env = scenes.lattice(2, 2, 1.0, 2.1).env.scaled(eps=1e-12)
    A = np.array([[1.0, 1.0], [1.0, 1.0]])
    g = capi.GpuComputer(env)
    g.set_system(*_dense_to_csr(A), np.array([1.0, -1.0]), np.zeros(2))
    with pytest.raises(capi.MpsError) as ei:
        g.stage("solveppe")
    assert ei.value.code == capi.MPS_CG_NOT_CONVERGED
    assert "Conjugate Gradient method couldn't solve Pressure Poison Equation" in ei.value.message


def test_empty_and_ragged_inputs(monkeypatch):
    monkeypatch.setenv("MPS_CG_ADAPTIVE", "0")    # frozen CTA split: the two runs below must then agree bit for bit
    env = scenes.dambreak2d().env
    g = capi.GpuComputer(env)
    assert g.count == 0
    g.forward(1)                                  # no particles: a step is a no-op apart from t += dt
    assert g.time()[0] == pytest.approx(env.max_dt)
    # particles added in two ragged batches (AddParticles may be called repeatedly, Computer.hpp:1754-1777)
    sc = scenes.dambreak2d()
    k = 517
    g2 = capi.GpuComputer(env)
    g2.add_particles(sc.x[:k], sc.u[:k], sc.p[:k], sc.n[:k], sc.type[:k])
    g2.add_particles(sc.x[k:], sc.u[k:], sc.p[k:], sc.n[k:], sc.type[k:])
    g1 = capi.GpuComputer.from_scene(sc)
    g1.forward(3); g2.forward(3)
    a, b = g1.state(), g2.state()
    assert all(np.array_equal(a[f], b[f]) for f in a), "batching must not change the result"


def _run_twice(sc, steps):
    outs = []
    for _ in range(2):
        g = capi.GpuComputer.from_scene(sc)
        g.forward(steps)
        outs.append(g.state())
        g.close()
    return outs


def test_run_to_run_determinism(monkeypatch):
    """With the CTA split frozen (MPS_CG_ADAPTIVE=0) every reduction runs in a fixed order: runs are bit-identical."""
    monkeypatch.setenv("MPS_CG_ADAPTIVE", "0")
    outs = _run_twice(scenes.dambreak2d(), 25)
    assert all(np.array_equal(outs[0][f], outs[1][f]) for f in outs[0]), "fixed-order reductions: runs must be bit-identical"


def test_adaptive_split_changes_rounding_only():
    """Default: the CG kernel re-balances its CTAs from its own cycle counters, which regroups the dot-product sums; two runs
    then agree to rounding (the reference's OpenMP reductions are not bit-reproducible either)."""
    outs = _run_twice(scenes.dambreak2d(), 25)
    assert np.array_equal(outs[0]["type"], outs[1]["type"])
    for f, tol in (("x", 1e-12), ("u", 1e-8), ("p", 1e-7), ("n", 1e-12)):
        err = np.abs(outs[0][f] - outs[1][f]).max() / max(np.abs(outs[1][f]).max(), 1e-300)
        assert err <= tol, (f, err)


def test_full_size_properties_1m_dambreak():
    """BASELINE.json config 2 at full size (1 008 104 particles): size-independent properties instead of an oracle diff."""
    sc = scenes.dambreak2d_fast(2.08e-4)
    assert sc.count == 1008104
    g = capi.GpuComputer.from_scene(sc)
    g.set_dt(sc.env.max_dt, True)
    g.stage("search")
    rp, idx = g.neighbors()
    n = sc.count
    rows = np.repeat(np.arange(n, dtype=np.uint64), np.diff(rp).astype(np.int64))
    # (1) symmetry: j in N(i) <=> i in N(j)  (test_ComputerNumberDensity.cpp:209-237); (2) i not in N(i) (:240-260)
    fwd = rows * np.uint64(n) + idx
    bwd = idx * np.uint64(n) + rows
    assert np.array_equal(np.sort(fwd), np.sort(bwd))
    assert not np.any(rows == idx)
    # (3) lattice interior: exactly 24 neighbours within NL = 2.88 l0, and n = n0 there
    g.stage("density")
    st = g.state()
    inner = 350 * 1404 + 700            # a water particle far from every boundary
    assert rp[inner + 1] - rp[inner] == 24
    assert abs(st["n"][inner] - g.env_values()["n0"]) <= 1e-9
    # (4) assemble + solve: symmetric matrix, a_ii = -sum a_ij on interior rows, true residual inside the stopping rule
    g.stage("ecs"); g.stage("explicit"); g.stage("density"); g.stage("savex"); g.stage("setppe")
    crp, col, val = g.csr()
    crows = np.repeat(np.arange(n, dtype=np.int64), np.diff(crp).astype(np.int64))
    import scipy.sparse as sp
    A = sp.csr_matrix((val, col.astype(np.int64), crp.astype(np.int64)), shape=(n, n))
    assert abs(A - A.T).max() == 0.0, "PPE matrix must be exactly symmetric (a_ij is a function of |x_i - x_j|)"
    assert abs(A[inner].sum()) <= 1e-9 * abs(A[inner, inner])
    b = g.vec("b"); x0 = g.vec("x")
    g.stage("solveppe")
    x = g.vec("x")
    r0 = b - A @ x0; r = b - A @ x
    assert r @ r <= 4.0 * sc.env.eps ** 2 * (r0 @ r0)
    st = g.stats()
    # preconditioned by default: a few dozen iterations where plain CG needs well over a thousand (BENCH_r01: 1 251 per step)
    assert 5 < g.last_iterations() < 150 and st.mg_levels >= 4 and st.mg_cells > 100000, (g.last_iterations(), st.mg_levels, st.mg_cells)
    del crows


@pytest.mark.parametrize("make,steps", [(lambda: scenes.dambreak2d_fast(8e-4), 30), (lambda: scenes.dambreak3d(l0=8e-3), 8),
                                        (lambda: scenes.central_gravity(half=120), 20)])
def test_preconditioned_solve_same_answer_far_fewer_sweeps(make, steps, monkeypatch):
    """SURVEY 8f rank 2.  The same assembled system solved by plain CG and by the preconditioned CG from the same guess: both
    inside the reference's stopping rule (true residual, recomputed on the host in FP64), the solutions agree to the CG
    tolerance, and the preconditioned solve makes at least 5x fewer passes over the matrix."""
    import scipy.sparse as sp
    sc = make()
    monkeypatch.setenv("MPS_CG_PRECOND", "1")
    lead = capi.GpuComputer.from_scene(sc)
    lead.forward(steps)                       # one history for both solvers: MPS pressures are sensitive to round-off in it
    state, (t, dt) = lead.state(), lead.time()
    assert not (state["type"] == 3).any()
    lead.close()
    out = {}
    for precond in ("0", "1"):
        monkeypatch.setenv("MPS_CG_PRECOND", precond)
        g = capi.GpuComputer.from_scene(sc)
        g.set_state(x=state["x"], u=state["u"], p=state["p"], n=state["n"])
        g.set_time(t, dt)
        g.set_dt(g.determine_dt(), True)
        for st in ("search", "density", "ecs", "explicit", "density", "savex", "setppe"):
            g.stage(st)
        n = sc.count
        crp, col, val = g.csr()
        A = sp.csr_matrix((val, col.astype(np.int64), crp.astype(np.int64)), shape=(n, n))
        b = g.vec("b"); x0 = g.vec("x")
        g.stage("solveppe")
        x = g.vec("x")
        r0 = b - A @ x0; r = b - A @ x
        assert r @ r <= 4.0 * sc.env.eps ** 2 * (r0 @ r0), (precond, (r @ r) / (r0 @ r0))
        out[precond] = (x, g.last_iterations(), g.stats().mg_levels)
        g.close()
    (x_plain, it_plain, lv_plain), (x_pcg, it_pcg, lv_pcg) = out["0"], out["1"]
    print(f"iterations: plain {it_plain}, preconditioned {it_pcg} on {lv_pcg} levels")
    assert lv_plain == 0 and lv_pcg >= 2
    assert rel_err(x_pcg, x_plain) <= CG_TOL
    assert it_pcg * 5 <= it_plain, (it_pcg, it_plain)


@pytest.mark.parametrize("make,count,lattice_nbrs", [
    (lambda: scenes.central_gravity(half=1000, l0=5e-5), 4004001, 24),     # BASELINE.json configs[2] (C3): 2001^2 particles, 2-D
    (lambda: scenes.dambreak3d(1.36e-3), 12244936, None),                  # BASELINE.json configs[3] (C4): 12.2M particles, 3-D
])
def test_full_size_properties_c3_c4(make, count, lattice_nbrs, monkeypatch):
    """BASELINE.json configs 3 and 4 at FULL size: too big for an oracle diff (the CPU code needs minutes per step), so
    size-independent properties, all evaluated on the device or on O(n) downloads:
    the first step's system solved twice from the same state — by the reference's plain CG (MPS_CG_PRECOND=0) and by the
    preconditioned CG — gives the same pressures to the CG tolerance, both recurrence residuals are inside the stopping rule,
    the preconditioned solve makes >= 5x fewer sweeps; neighbour counts are symmetric in total (sum of counts is even), no
    particle lists itself more often than the cell stencil allows, lattice-interior particles have the lattice count and n = n0."""
    sc = make()
    assert sc.count == count
    out = {}
    for precond in ("0", "1"):
        monkeypatch.setenv("MPS_CG_PRECOND", precond)
        g = capi.GpuComputer.from_scene(sc)
        g.set_dt(sc.env.max_dt, True)
        g.stage("search"); g.stage("density")
        if precond == "1":
            cnt = g.neighbor_counts()
            assert int(cnt.sum()) % 2 == 0                                       # j in N(i) <=> i in N(j)
            assert cnt.max() <= (3 ** sc.env.dim) * g.grid_capacity()
            n_now = g.state()["n"]
            n0 = g.env_values()["n0"]
            if lattice_nbrs is not None:
                # central gravity: every particle sits on the lattice at rest; the centre particle is far from the free surface
                centre = int(np.argmin((sc.x ** 2).sum(axis=1)))
                assert cnt[centre] == lattice_nbrs and abs(n_now[centre] - n0) <= 1e-9 * n0
            assert np.isfinite(n_now).all() and (n_now[sc.type == 0] >= n0 * (1 - 1e-12)).all()   # N = max(n, n0), Computer.hpp:826
        for st in ("ecs", "explicit", "density", "savex", "setppe", "solveppe"):
            g.stage(st)
        x = g.vec("x")
        stt = g.stats()
        assert np.isfinite(x).all()
        assert stt.last_rr <= sc.env.eps ** 2 * stt.last_rr0, (precond, stt.last_rr, stt.last_rr0)
        out[precond] = (x, g.last_iterations(), stt.mg_levels)
        g.close()
    (x_plain, it_plain, lv_plain), (x_pcg, it_pcg, lv_pcg) = out["0"], out["1"]
    print(f"{count} particles: plain CG {it_plain} iterations, preconditioned {it_pcg} on {lv_pcg} levels")
    assert lv_plain == 0 and lv_pcg >= 4
    assert rel_err(x_pcg, x_plain) <= CG_TOL
    assert it_pcg * 5 <= it_plain, (it_pcg, it_plain)


def _wall_motion_positions(base, t, amp, vel, omega, phase, t0, t1):
    tau = min(max(t - t0, 0.0), t1 - t0)
    return base + np.asarray(vel) * tau + np.asarray(amp) * (np.sin(omega * tau + phase) - np.sin(phase))


@pytest.mark.parametrize("make,steps,motion", [
    # a piston: the left wall of the sample tank (Wall + Dummy columns with x < 0) pushed into the water at 5 cm/s with a shake on top
    (lambda: scenes.dambreak2d(), 60, dict(amplitude=(4e-4, 0.0), velocity=(0.05, 0.0), omega=300.0, phase=0.3, t_begin=2e-3, t_end=2.4e-2)),
    # a sloshing tank in 3-D: every non-fluid particle shaken along x
    (lambda: scenes.dambreak3d(l0=0.035), 8, dict(amplitude=(2e-3, 0.0, 5e-4), velocity=(0.0, 0.0, 0.0), omega=400.0, phase=0.0, t_begin=0.0, t_end=1.0)),
])
def test_device_evaluated_wall_motion_matches_host_callback(make, steps, motion):
    """SURVEY 8f rank 4 (Computer.hpp:993,1012-1019): positionWall(i, t, dt) of an analytic motion evaluated inside k_explicit_move
    (mps_set_wall_motion) == the CPU restatement fed, step by step, with the same motion evaluated on the host the way the reference
    calls its callback (t = Environment::T() after SetNextT)."""
    sc = make()
    p, g = _port_and_gpu(sc)
    D = sc.env.dim
    nonfluid = np.flatnonzero(sc.type != 0)
    ids = nonfluid[sc.x[nonfluid, 0] < 0] if D == 2 else nonfluid
    assert len(ids) > 10
    g.set_wall_motion(ids=(ids if D == 2 else None), **motion)
    base = sc.x[ids].copy()
    kw = dict(amp=motion["amplitude"], vel=motion["velocity"], omega=motion["omega"], phase=motion["phase"], t0=motion["t_begin"], t1=motion["t_end"])
    for _ in range(steps):
        dt = p.determine_dt()
        t_next = p.env_values()["t"] + dt
        p.set_wall_positions(ids, _wall_motion_positions(base, t_next, **kw))
        p.forward(1, dt=dt)
        g.forward(1)
    sp, sg = p.state(), g.state()
    assert np.array_equal(sp["type"], sg["type"])
    moved = np.abs(sg["x"][ids] - base).max()
    assert moved > 1e-4                                   # the wall really moved ...
    assert rel_err(sg["x"][ids], sp["x"][ids]) <= 1e-12   # ... to where the host callback puts it (device sin vs libm: an ulp or two)
    assert rel_err(sg["u"][ids], sp["u"][ids]) <= 1e-7    # u = (x - x_prev) / dt of a sub-millimetre move: cancellation, still tiny
    assert rel_err(sg["x"], sp["x"]) <= 1e-8
    assert rel_err(sg["n"], sp["n"]) <= 1e-7
    assert abs(p.env_values()["t"] - g.time()[0]) <= 1e-9 * g.time()[0]
    # clearing the motions freezes the wall where positionWall = wall[] puts it: back at the base positions after one more step
    g.set_wall_motion(clear=True)
    g.forward(1)
    assert rel_err(g.state()["x"][ids], base) <= 1e-15
    p.close(); g.close()
