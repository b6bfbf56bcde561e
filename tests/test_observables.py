"""Benchmark observables: known answers on synthetic states, agreement between a state and the CSV the
driver writes from it, and a short physical sanity run of the CPU restatement."""
import math
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from openmps_b200 import observables as ob, scenes  # noqa: E402


def _state(sc):
    return {"type": sc.type, "x": sc.x, "u": sc.u, "p": sc.p, "n": sc.n}


def test_lattice_n0_known_answer():
    # the reference's own known answer: n at the centre of a 7x7 lattice with r_e/l0 = 2.1 is 6.539696962 (test_ComputerNumberDensity.cpp:273-275)
    assert abs(ob.lattice_n0_2d(2.1) - 6.539696962) < 1e-8


def test_dam_break_edge_is_the_column_width_at_rest():
    sc = scenes.dambreak2d()
    L = 0.146
    e = ob.dam_break_edge(_state(sc), L=L, g=9.8, t=0.1)
    fluid_x = sc.x[sc.type == 0, 0]
    assert e["edge"] == fluid_x.max() and abs(e["Z_by_L"] - 1.0) < 0.06     # column width L = 18 particles of 8 mm
    assert e["t_star"] == pytest.approx(0.1 * math.sqrt(2 * 9.8 / L))


def test_central_gravity_perfect_disc():
    # a disc of lattice particles: surface ring detected by n, roundness close to 100 %, centre particle found
    sc = scenes.central_gravity(half=20, l0=1e-3)
    st = _state(sc)
    r = np.sqrt((sc.x ** 2).sum(axis=1))
    st["n"] = np.where(r > r.max() - 1.5e-3, 1.0, ob.lattice_n0_2d(2.4))   # low density on the rim only
    st["p"] = 1000 * 9.8 * (r.max() - r)
    L = math.sqrt(math.pi) * r.max()
    out = ob.central_gravity(st, 2.4, 0.97, L)
    assert out["roundness_percent"] > 90.0
    assert out["p_center"] == pytest.approx(1000 * 9.8 * r.max())
    assert out["p_theoretical"] == pytest.approx(1000 * 9.8 * r.max())


def test_hydrostatic_exact_profile():
    sc = scenes.static_pressure()
    st = _state(sc)
    fluid = sc.type == 0
    h = sc.x[fluid, 1].max()
    st["p"] = np.where(fluid, 998.2 * 9.8 * (h - sc.x[:, 1]), 0.0)
    out = ob.hydrostatic(st, 998.2, 9.8)
    assert out["slope_by_rho_g"] == pytest.approx(1.0, abs=1e-12) and out["max_rel_dev"] < 1e-12


def test_probe_heights_and_wall_pressure():
    l0 = 0.01
    xs = np.arange(0, 1.2, l0)
    pts = [(x, z) for x in xs for z in np.arange(0, 0.3, l0)]
    x = np.array(pts)
    n = np.full(len(x), 6.0)
    t = np.zeros(len(x), np.int32)
    wall = np.array([(-l0, z) for z in np.arange(0.1, 0.25, l0)])
    st = {"type": np.concatenate([t, np.ones(len(wall), np.int32)]), "x": np.vstack([x, wall]), "u": np.zeros((len(x) + len(wall), 2)),
          "p": np.concatenate([np.zeros(len(x)), np.full(len(wall), 500.0)]), "n": np.concatenate([n, np.full(len(wall), 6.0)])}
    out = ob.probe_heights_and_pressure(st, l0, 0.6)
    assert out["h1"] == pytest.approx(0.29) and out["h2"] == pytest.approx(0.29) and out["p2"] == pytest.approx(500.0)


BIN2 = os.path.join(ROOT, "openmps_b200", "bin", "OpenMps")


@pytest.mark.skipif(not os.path.exists(BIN2), reason="driver not built")
def test_observables_from_state_and_from_driver_csv_agree(tmp_path):
    sc = scenes.dambreak2d()
    xml = scenes.write_xml(sc, str(tmp_path / "in.xml"))
    (tmp_path / "result").mkdir()
    r = subprocess.run([BIN2, "--check-io", xml, str(tmp_path / "result")], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0, r.stderr
    back = scenes.read_result_csv(str(tmp_path / "result" / "particles_00000.csv"))
    assert ob.dam_break_edge(back) == pytest.approx(ob.dam_break_edge(_state(sc)), rel=1e-5)


def test_short_dam_break_run_of_the_cpu_restatement_moves_the_edge():
    """Physical sanity of the checker (not a parity test): the column starts collapsing, the edge only advances."""
    from oracle import bind
    sc = scenes.dambreak2d()
    eng = bind.PortComputer.from_scene(sc)
    edges = [ob.dam_break_edge(eng.state())]
    for _ in range(4):
        eng.forward(15)
        edges.append(ob.dam_break_edge(eng.state()))
    assert all(b >= a - 1e-12 for a, b in zip(edges, edges[1:])) and edges[-1] > edges[0]
