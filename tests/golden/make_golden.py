"""Generates the committed golden fixtures tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref).

Run here (build container, where /root/reference exists):

    make -C oracle ref port && OMP_NUM_THREADS=1 python tests/golden/make_golden.py

Each fixture records one reference time step, stage by stage, from a given input state: the outputs of
SearchNeighbor, ComputeNeighborDensities, ComputeErrorCorrection, ComputeExplicitForces, the second density pass,
SetPressurePoissonEquation, SolvePressurePoissonEquation, the pressure write-back + ModifyByPressureGradient and
DynamicStabilize (Computer.hpp:1700-1742), plus the next DetermineDt.  OMP_NUM_THREADS=1 makes ViennaCL's dot
products sequential, hence reproducible (viennacl/linalg/host_based/vector_operations.hpp:44-45,540).
The CG iteration count is taken from the bit-identical restatement (the reference does not expose it).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from openmps_b200 import scenes  # noqa: E402
from oracle.bind import PortComputer, RefComputer  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def env_arrays(env):
    return {
        "env_dim": np.int32(env.dim), "env_max_dt": env.max_dt, "env_courant": env.courant, "env_g": env.g,
        "env_rho": env.rho, "env_nu": env.nu, "env_r_e_by_l0": env.r_e_by_l0, "env_l0": env.l0,
        "env_min_x": np.asarray(env.min_x, np.float64), "env_max_x": np.asarray(env.max_x, np.float64),
        "env_eps": env.eps, "env_central_gravity": np.int32(env.central_gravity),
    }


def record_step(name, scene, warm_steps, wall_move=None):
    ref = RefComputer.from_scene(scene)
    port = PortComputer.from_scene(scene)
    if warm_steps:
        ref.forward(warm_steps)
        port.forward(warm_steps)
    rec = dict(env_arrays(scene.env))
    st = ref.state()
    assert all(np.array_equal(st[k], port.state()[k]) for k in st), "port diverged from reference during warm-up"
    ev = ref.env_values()
    rec.update({"in_" + k: v for k, v in st.items()})
    rec["in_t"] = ev["t"]
    rec["wall_target"] = scene.x.copy()
    if wall_move is not None:
        ids, newx = wall_move
        ref.set_wall_positions(ids, newx); port.set_wall_positions(ids, newx)
        rec["wall_target"][ids] = newx
    dt = ref.determine_dt()
    rec["dt"] = dt
    for e in (ref, port):
        e.set_dt(dt, True)

    def both(stage):
        ref.stage(stage); port.stage(stage)

    both("search")
    rec["cells"] = ref.cells()
    rp, idx = ref.neighbors()
    rec["nbr_rowptr"] = rp; rec["nbr_idx"] = idx.astype(np.uint32)
    rec["search_type"] = ref.state()["type"]
    both("density")
    rec["density1_n"] = ref.state()["n"]; rec["density1_nws"] = ref.vec("nWithoutSpp")
    both("ecs")
    rec["ecs"] = ref.vec("ecs")
    both("explicit")
    s = ref.state(); rec["explicit_x"] = s["x"]; rec["explicit_u"] = s["u"]
    both("density")
    rec["density2_n"] = ref.state()["n"]; rec["density2_nws"] = ref.vec("nWithoutSpp")
    both("savex")
    both("setppe")
    r, c, v = ref.csr()
    rec["csr_rowptr"] = r; rec["csr_col"] = c; rec["csr_val"] = v
    rec["ppe_b"] = ref.vec("b"); rec["ppe_x0"] = ref.vec("x")
    both("solveppe")
    rec["ppe_x"] = ref.vec("x")
    assert np.array_equal(rec["ppe_x"], port.vec("x")), "port CG differs from reference CG"
    rec["cg_iterations"] = np.int64(port.last_iterations())
    both("implicit")   # re-assembles and re-solves identically, then P = max(x, 0) and the pressure-gradient move
    s = ref.state(); rec["implicit_x"] = s["x"]; rec["implicit_u"] = s["u"]; rec["implicit_p"] = s["p"]
    both("ds")
    s = ref.state()
    for k, v_ in s.items():
        rec["out_" + k] = v_
    rec["next_dt"] = ref.determine_dt()
    ps = port.state()
    assert all(np.array_equal(s[k], ps[k]) for k in s), "port differs from reference at end of step"
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **rec)
    print(f"{name}: n={scene.count} warm={warm_steps} dt={dt:.3e} iters={int(rec['cg_iterations'])} "
          f"disabled={(s['type'] == 3).sum()} -> {os.path.getsize(path) / 1024:.0f} KiB")


def main():
    record_step("sample_rest", scenes.dambreak2d(), 0)
    record_step("sample_dev400", scenes.dambreak2d(), 400)
    record_step("static_small_dev50", scenes.static_pressure(width=20, height=30), 50)
    record_step("central_gravity_dev50", scenes.central_gravity(half=12), 50)
    record_step("dambreak3d_dev10", scenes.dambreak3d(l0=0.035), 10)

    # jittered lattice with particles outside the domain (disabled on the first search) and exactly on cell faces
    lat = scenes.lattice(2, 12, 0.1, 2.1, jitter=0.05, margin_cells=0.5, g=9.8, max_dt=1e-3)
    nl = 2.1 * 0.1 * 1.2
    lat.x[5] = (lat.env.min_x[0] + 3 * nl, lat.env.min_x[1] + 2 * nl)      # exactly on a cell corner
    lat.x[17] = (lat.env.min_x[0] - 1e-9, 0.3)                               # just below MinX -> disabled
    lat.x[40] = (lat.env.max_x[0] + 2.5 * nl, 0.2)                           # beyond the two spare cells -> disabled
    lat.x[41] = (lat.env.max_x[0] + 0.5 * nl, 0.2)                           # inside the spare cells -> kept
    lat.type[::7] = scenes.WALL
    lat.type[3::11] = scenes.DUMMY
    rng = np.random.default_rng(7)
    lat.u[:] = rng.normal(0, 0.05, lat.u.shape)
    lat.p[:] = rng.uniform(0, 100, lat.p.shape)
    record_step("lattice_edge_cases", lat, 0)

    # moving wall: one wall particle is told to move this step (positionWall callback, Computer.hpp:1012-1019)
    mv = scenes.dambreak2d()
    ids = np.where(mv.type == scenes.WALL)[0][:5]
    record_step("sample_moving_wall", mv, 20, wall_move=(ids, mv.x[ids] + np.array([1e-4, 0.0])))


if __name__ == "__main__":
    main()
