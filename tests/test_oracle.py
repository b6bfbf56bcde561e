"""CPU tier (no GPU): pins the oracle.

* the CPU restatement (oracle/mps_oracle.cpp) reproduces every committed golden fixture BIT FOR BIT (the fixtures were
  generated from the unmodified reference build, tests/golden/make_golden.py);
* where oracle/_ref is present, restatement == reference build on fresh seeded inputs, stage by stage and over many
  steps, including the first-Disable event;
* the known answers of the upstream gtests hold (test_ComputerNumberDensity.cpp:273, test_ComputerConjugateGradient.cpp:
  115-229, test_ComputerImplicitForces.cpp:247-257,440-448).
"""
import os

import numpy as np
import pytest

from conftest import golden_names
from helpers import open_engine, replay_golden_step
from openmps_b200 import scenes
from oracle import bind

pytestmark = pytest.mark.skipif(not bind.port_available(), reason="oracle/_build/libmps_oracle.so not built (make -C oracle port)")


@pytest.mark.parametrize("name", golden_names())
def test_port_reproduces_golden_bit_exact(golden, name):
    g = golden(name)
    eng = open_engine(bind.PortComputer, g)
    replay_golden_step(eng, g, exact=True)


def _need_ref(dim=2, cg=False):
    if not bind.available(dim, cg):
        pytest.skip("oracle/_ref not built here")


def _same_state(a, b):
    sa, sb = a.state(), b.state()
    return all(np.array_equal(sa[k], sb[k]) for k in sa)


@pytest.mark.parametrize("make,steps", [
    (lambda: scenes.dambreak2d(), 120),
    (lambda: scenes.static_pressure(width=12, height=20), 30),
    (lambda: scenes.central_gravity(half=10), 40),
    (lambda: scenes.dambreak3d(l0=0.04), 8),
    (lambda: scenes.lattice(2, 9, 0.1, 2.1, jitter=0.05, max_dt=1e-3), 15),
])
def test_port_equals_reference_build_over_steps(make, steps):
    sc = make()
    _need_ref(sc.env.dim, sc.env.central_gravity)
    r = bind.RefComputer.from_scene(sc); p = bind.PortComputer.from_scene(sc)
    assert r.env_values() == p.env_values()
    assert r.grid_capacity() == p.grid_capacity()
    for k in range(steps):
        assert r.determine_dt() == p.determine_dt()
        r.forward(1); p.forward(1)
        assert _same_state(r, p), f"state differs after step {k + 1}"
    a, b = r.neighbors(), p.neighbors()
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    a, b = r.csr(), p.csr()
    assert all(np.array_equal(x, y) for x, y in zip(a, b))
    for v in ("x", "b", "ecs", "nWithoutSpp", "du", "originalX"):
        assert np.array_equal(r.vec(v), p.vec(v)), v


def test_port_equals_reference_through_first_disable():
    """A fluid particle thrown out of the domain is disabled, takes the wall branch and enters the dt reduction
    (Computer.hpp:711-715,1012-1019,764-775)."""
    _need_ref()
    sc = scenes.dambreak2d()
    # grid top = MinZ + 29 cells * NL = 0.63616: a lone particle just below it, moving up, leaves on the first step
    sc.x[10] = (0.3, 0.6355)
    sc.u[10] = (0.0, 5.0)
    r = bind.RefComputer.from_scene(sc); p = bind.PortComputer.from_scene(sc)
    seen = False
    for k in range(12):
        r.forward(1); p.forward(1)
        assert _same_state(r, p), f"step {k}"
        seen = seen or bool((r.state()["type"] == 3).any())
    assert seen, "the test input no longer triggers a Disable"


def test_grid_overflow_is_reported_like_the_reference():
    """More particles in one cell than (ceil(NL/l0)+1)^D -> Grid::Exception("Too many particle in a block")."""
    sc = scenes.lattice(2, 6, 0.1, 2.1)
    sc.x[:] = sc.x[0] + np.random.default_rng(3).uniform(0, 1e-3, sc.x.shape)   # 36 particles in one cell, cap = 16
    p = bind.PortComputer.from_scene(sc)
    with pytest.raises(bind.RefError) as ei:
        p.stage("search")
    assert ei.value.code == 2 and "Too many particle in a block" in str(ei.value)
    if bind.available():
        r = bind.RefComputer.from_scene(sc)
        with pytest.raises(bind.RefError) as ei:
            r.stage("search")
        assert ei.value.code == 2


def test_known_answer_number_density():
    # test_ComputerNumberDensity.cpp:263-276: centre of a 7 x 7 lattice, r_e / l0 = 2.1 -> n = 6.539696962 (Koshizuka 2014)
    sc = scenes.lattice(2, 7, 0.1, 2.1)
    p = bind.PortComputer.from_scene(sc)
    p.stage("search"); p.stage("density")
    assert abs(p.state()["n"][24] - 6.539696962) < 1e-5
    assert abs(p.env_values()["n0"] - 6.539696962) < 1e-8


def _dense_to_csr(A):
    n = A.shape[0]
    rowptr = [0]; col = []; val = []
    for i in range(n):
        for j in range(n):
            if A[i, j] != 0:
                col.append(j); val.append(A[i, j])
        rowptr.append(len(col))
    return np.array(rowptr), np.array(col), np.array(val, float)


CG_CASES = {
    # test_ComputerConjugateGradient.cpp:115-144
    "identity": (np.eye(5), np.array([1.0, 2.0, 3.0, 4.0, 5.0]), np.array([1.0, 2.0, 3.0, 4.0, 5.0])),
    # test_ComputerConjugateGradient.cpp:146-182: symmetric INDEFINITE +-1 matrix, answer (2, 4, 6, 8)
    "pm1_4x4": (np.array([[1, 1, 1, 1], [1, 1, -1, -1], [1, -1, 1, -1], [1, -1, -1, 1]], float), None, np.array([2.0, 4.0, 6.0, 8.0])),
}


def cg_system(name):
    A, b, x = CG_CASES[name]
    if b is None:
        b = A @ x
    return A, b, x


def poisson_1d(n=32):
    # test_ComputerConjugateGradient.cpp:186-229: f'' = x on (0, 1), f(0) = f(1) = 0; negative-definite tridiagonal
    h = 1.0 / (n + 1)
    A = np.zeros((n, n))
    for i in range(n):
        A[i, i] = -2.0 / h / h
        if i > 0:
            A[i, i - 1] = 1.0 / h / h
        if i < n - 1:
            A[i, i + 1] = 1.0 / h / h
    xs = (np.arange(n) + 1) * h
    return A, xs.copy(), (xs ** 3 - xs) / 6.0


@pytest.mark.parametrize("case", ["identity", "pm1_4x4", "poisson1d"])
def test_known_answer_conjugate_gradient(case):
    A, b, want = poisson_1d() if case == "poisson1d" else cg_system(case)
    sc = scenes.lattice(2, 2, 1.0, 2.1).env.scaled(eps=1e-7)
    p = bind.PortComputer(sc)
    p.set_system(*_dense_to_csr(A), b, np.zeros(len(b)))
    p.stage("solveppe")
    got = p.vec("x", n=len(b))
    assert np.allclose(got, want, rtol=0, atol=1e-3)
    if bind.available():
        r = bind.RefComputer(sc)
        r.set_system(*_dense_to_csr(A), b, np.zeros(len(b)))
        r.stage("solveppe")
        assert np.array_equal(r.vec("x", n=len(b)), got)


def test_known_answer_ppe_matrix():
    # test_ComputerImplicitForces.cpp:216-258,378-455: a_ij = (5 - D) r_e / n0 / r^3, a_ii = -sum a_ij at the centre
    sc = scenes.lattice(2, 15, 1.0, 5.0, g=9.8, max_dt=1e-2)
    p = bind.PortComputer.from_scene(sc)
    p.set_dt(1e-2, True)
    p.stage("search"); p.stage("density"); p.stage("setppe")
    rp, col, val = p.csr()
    ev = p.env_values()
    i = 7 * 15 + 7
    row = dict(zip(col[rp[i]:rp[i + 1]].tolist(), val[rp[i]:rp[i + 1]].tolist()))
    off = 0.0
    for j, a in row.items():
        if j == i:
            continue
        r = np.linalg.norm(sc.x[j] - sc.x[i])
        assert abs(a - 3 * ev["R_e"] / ev["n0"] / r ** 3) <= 1e-3 * a
        off += a
    assert abs(row[i] + off) <= 1e-9 * abs(row[i])


def test_scene_generator_reproduces_sample_xml():
    path = "/root/reference/Benchmark/Sample/Sample.xml"
    if not os.path.exists(path):
        pytest.skip("reference tree not mounted")
    import xml.etree.ElementTree as ET
    root = ET.parse(path).getroot()
    rows = [l for l in root.find("particles").text.split("\n") if l.strip()][1:]
    a = np.array([[float(v) for v in r.split(",")] for r in rows])
    sc = scenes.dambreak2d()
    assert np.array_equal(a[:, 0].astype(np.int32), sc.type) and np.array_equal(a[:, 1:3], sc.x)
    envx = {c.tag: float(c.get("value")) for c in root.find("environment")}
    assert (envx["minX"], envx["minZ"], envx["maxX"], envx["maxZ"]) == (*sc.env.min_x, *sc.env.max_x)
    assert np.array_equal(scenes.dambreak2d_fast(8e-3).x, sc.x)
