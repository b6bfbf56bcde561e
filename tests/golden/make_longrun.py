#!/usr/bin/env python
"""Generates tests/golden/longrun_observables.json: time series of the reference's benchmark observables over LONG runs
(north star: "DamBreak/StaticPressure physical observables must agree over long runs"), computed with the reference build
(oracle/_ref, the unmodified reference headers; falls back to the CPU restatement, which is pinned bit-exact against it).

Every scene is run twice, with the reference's allowable residual (1e-10) and with 1e-11: long MPS runs are chaotic (two correct
solvers diverge through CG round-off), so the second run measures how far an observable may legitimately move; the GPU test
(tests/test_longrun_observables.py) accepts |gpu - reference| <= max(floor, 4 x that spread).

    python tests/golden/make_longrun.py        (about 5 minutes; needs /root/reference for the reference build)
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from openmps_b200 import observables as ob, scenes  # noqa: E402
from oracle import bind  # noqa: E402


def engine(sc):
    if bind.available(sc.env.dim, sc.env.central_gravity, fast=True):
        return bind.RefComputer.from_scene(sc, fast=True), "reference"
    return bind.PortComputer.from_scene(sc), "port"


def bottom_ratio(st, env):
    """mean pressure of the lowest fluid layer over rho g (h - z_bottom)"""
    fl = st["type"] == 0
    z = st["x"][fl, -1]
    zb, h = z.min(), z.max()
    return float(st["p"][fl][z < zb + 0.5 * env.l0].mean() / (env.rho * env.g * (h - zb)))


def dam_break(eps):
    sc = scenes.dambreak2d(eps=eps)
    e, kind = engine(sc)
    rows, steps = [], 0
    for k in range(1, 10):
        steps += e.run_until(0.05 * k)
        rows.append({"t": 0.05 * k, "steps": steps, "edge": ob.dam_break_edge(e.state())})
    return {"kind": kind, "L": sc.meta.get("L", 0.146) if hasattr(sc, "meta") else 0.146, "series": rows}


def static_pressure(eps, total=3000, every=250):
    sc = scenes.static_pressure(eps=eps)
    e, kind = engine(sc)
    rows = []
    for k in range(total // every):
        e.forward(every)
        st = e.state()
        h = ob.hydrostatic(st, sc.env.rho, sc.env.g)
        rows.append({"steps": (k + 1) * every, "slope_by_rho_g": h["slope_by_rho_g"], "h": h["h"], "bottom_ratio": bottom_ratio(st, sc.env),
                     "p_max": float(st["p"].max())})
    return {"kind": kind, "series": rows}


CG_HALF, CG_BETA = 20, 0.97


def central_gravity(eps, total=1500, every=250):
    sc = scenes.central_gravity(half=CG_HALF, eps=eps)
    e, kind = engine(sc)
    L = (2 * CG_HALF + 1) * sc.env.l0
    rows = []
    for k in range(total // every):
        e.forward(every)
        st = e.state()
        o = ob.central_gravity(st, sc.env.r_e_by_l0, CG_BETA, L)
        # check_result.py's surface set (n < beta n0) is EMPTY in the reference's default build: with MPS_SPP the stored number
        # density of a free-surface particle includes its virtual neighbours and never drops below n0 (Computer.hpp:820-831), so
        # the script's roundness is undefined there (recorded as null).  The extent of the drop, max |x| / R, is recorded instead.
        rows.append({"steps": (k + 1) * every, "roundness_percent": None if np.isnan(o["roundness_percent"]) else o["roundness_percent"],
                     "surface_particles": int((st["n"] < ob.lattice_n0_2d(sc.env.r_e_by_l0) * CG_BETA).sum()),
                     "r_max_by_R": float(np.sqrt((st["x"] ** 2).sum(axis=1)).max() / o["R"]),
                     "p_center": o["p_center"], "p_theoretical": o["p_theoretical"]})
    return {"kind": kind, "half": CG_HALF, "beta": CG_BETA, "L": L, "series": rows}


def main():
    path = os.path.join(ROOT, "tests", "golden", "longrun_observables.json")
    out = {"generator": "tests/golden/make_longrun.py", "eps": [1e-10, 1e-11]}
    only = sys.argv[1:]          # e.g. `make_longrun.py central_gravity` regenerates one section of an existing file
    if only and os.path.exists(path):
        with open(path) as f:
            out = json.load(f)
    for name, fn in (("dam_break_sample", dam_break), ("static_pressure", static_pressure), ("central_gravity", central_gravity)):
        if only and name not in only:
            continue
        t0 = time.perf_counter()
        out[name] = {"reference": fn(1e-10), "perturbed": fn(1e-11)}
        print(name, f"{time.perf_counter() - t0:.1f} s", flush=True)
    with open(path, "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
