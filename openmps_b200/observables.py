"""Physical observables of the reference's benchmarks as functions of a particle state.

The reference computes them in plotting scripts that read result/particles_%05d.csv; here the same definitions work on a state
dict ({"type", "x", "u", "p", "n"}, as returned by GpuComputer.state(), a CPU checker, or scenes.read_result_csv) so that
long-run physical agreement is an assertion, not a figure:

* ``dam_break_edge``            Benchmark/DamBreak/koshizukaoka1996_edge.py:14-20, :55-63
* ``probe_heights_and_pressure`` Benchmark/DamBreak/zhouetal1999.py:15-39
* ``central_gravity``           Benchmark/CentralGravity/check_result.py:14-57
* ``hydrostatic``               Benchmark/StaticPressure has only a generator; the observable is p(z) against rho g (h - z)
"""
from __future__ import annotations

import math

import numpy as np

FLUID, WALL, DUMMY, DISABLED = 0, 1, 2, 3


def lattice_n0_2d(r_e_by_l0):
    """n0 as the checker scripts compute it (check_result.py:14-23: lattice units, square window of ceil(r_e))."""
    size = int(math.ceil(r_e_by_l0))
    n0 = 0.0
    for y in range(-size, size + 1):
        for x in range(-size, size + 1):
            r = math.sqrt(x * x + y * y)
            if 0 < r < r_e_by_l0:
                n0 += r_e_by_l0 / r - 1
    return n0


def dam_break_edge(state, L=None, g=None, t=None):
    """Leading edge of the collapsing column: max x over fluid particles (koshizukaoka1996_edge.py:18).
    With L (and g, t) also the plotted dimensionless pair (t sqrt(2 g / L), Z / L) (:55-63)."""
    fluid = state["type"] == FLUID
    edge = float(state["x"][fluid, 0].max()) if fluid.any() else float("nan")
    if L is None:
        return edge
    return {"edge": edge, "Z_by_L": edge / L, "t_star": None if (g is None or t is None) else t * math.sqrt(2 * g / L)}


def probe_heights_and_pressure(state, l0, min_n, x_h1=2020e-3 - 1525e-3, x_h2=2020e-3 - 1028e-3, z_p2=160e-3, d=90e-3):
    """Zhou et al. (1999) probes (zhouetal1999.py:26-39): water height at two stations = highest particle with n > min_n
    within l0/2 of the station; wall pressure = mean p of Wall particles with x < 0 within d/2 of z_p2."""
    x, z, p, n, t = state["x"][:, 0], state["x"][:, -1], state["p"], state["n"], state["type"]

    def height(xs):
        m = (np.abs(x - xs) < l0 / 2) & (n > min_n)
        return float(max(0.0, z[m].max())) if m.any() else 0.0

    m = (x < 0) & (t == WALL) & (np.abs(z - z_p2) < d / 2)
    return {"h1": height(x_h1), "h2": height(x_h2), "p2": float(p[m].mean()) if m.any() else float("nan")}


def central_gravity(state, r_e_by_l0, beta, L):
    """Roundness [%] and centre pressure of the self-gravitating drop (check_result.py:26-57): surface particles are those
    with n < beta n0; roundness = 1 - (r_max - r_min) / R with R = L / sqrt(pi); the centre pressure is p of the particle
    nearest the origin; the theoretical value is 1000 * 9.8 * R (:61-62)."""
    r = np.sqrt((state["x"] ** 2).sum(axis=1))
    R = L / math.sqrt(math.pi)
    surface = r[state["n"] < lattice_n0_2d(r_e_by_l0) * beta]
    roundness = float("nan") if surface.size == 0 else 1.0 - float(surface.max() - surface.min()) / R
    return {"roundness_percent": 100.0 * roundness, "p_center": float(state["p"][int(np.argmin(r))]), "p_theoretical": 1000 * 9.8 * R, "R": R}


def hydrostatic(state, rho, g, surface_z=None):
    """Hydrostatic column: least-squares slope of p against depth over fluid particles below the free-surface layer, relative
    to rho g, and the largest deviation of p from rho g (h - z) relative to rho g h."""
    fluid = state["type"] == FLUID
    z, p = state["x"][fluid, -1], state["p"][fluid]
    h = float(z.max()) if surface_z is None else float(surface_z)
    depth = h - z
    inner = p > 0                      # free-surface particles carry the Dirichlet value 0
    if inner.sum() < 2:
        return {"slope_by_rho_g": float("nan"), "max_rel_dev": float("nan"), "h": h}
    slope = float(np.polyfit(depth[inner], p[inner], 1)[0])
    dev = float(np.abs(p[inner] - rho * g * depth[inner]).max() / (rho * g * h))
    return {"slope_by_rho_g": slope / (rho * g), "max_rel_dev": dev, "h": h}


# ---- the same observables from the device reductions of mps_observe (csrc/mps_observe.cu): no state download ----------------
def device_dam_break_edge(gpu):
    """dam_break_edge() of the state resident on the GPU (one reduction pass, 24 doubles back)."""
    return gpu.observe()["edge_x"]


def device_probes(gpu, min_n, x_h1=2020e-3 - 1525e-3, x_h2=2020e-3 - 1028e-3, z_p2=160e-3, d=90e-3):
    o = gpu.observe(x_h1=x_h1, x_h2=x_h2, min_n=min_n, z_p2=z_p2, d=d)
    return {"h1": o["h1"], "h2": o["h2"], "p2": o["p2_sum"] / o["p2_count"] if o["p2_count"] else float("nan")}


def device_central_gravity(gpu, r_e_by_l0, beta, L):
    o = gpu.observe(surface_n=lattice_n0_2d(r_e_by_l0) * beta)
    R = L / math.sqrt(math.pi)
    roundness = float("nan") if o["surface_count"] == 0 else 1.0 - (o["r_max_surface"] - o["r_min_surface"]) / R
    return {"roundness_percent": 100.0 * roundness, "p_center": o["center_p"], "p_theoretical": 1000 * 9.8 * R, "R": R}


def device_hydrostatic(gpu, rho, g, surface_z=None):
    """hydrostatic() from the moments the device returns: slope = (N S_dp - S_d S_p) / (N S_dd - S_d^2)."""
    h = gpu.observe()["top_z"] if surface_z is None else float(surface_z)
    o = gpu.observe(rho_g=rho * g, surface_z=h)
    N = o["inner"]
    if N < 2:
        return {"slope_by_rho_g": float("nan"), "max_rel_dev": float("nan"), "h": h}
    # centred moments (the raw form cancels badly when the depths are close together)
    md, mp = o["sum_d"] / N, o["sum_p"] / N
    slope = (o["sum_dp"] - N * md * mp) / (o["sum_dd"] - N * md * md)
    return {"slope_by_rho_g": slope / (rho * g), "max_rel_dev": o["max_dev"] / (rho * g * h), "h": h}
