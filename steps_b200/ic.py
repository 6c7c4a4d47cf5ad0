"""Synthetic initial conditions of the shapes BASELINE.json names (no IC files ship with the reference).

Geometry follows the reference's glass generator
(tools/GenerateIC_for_3DSphericalGlassMaking/GenerateIC_for_3DSphericalGlassMaking.py:36-97, :252,
:272-294, :362-375; yamlfiles/GlassIC_Nr224_Nhp32_D1860.yaml): a constant-resolution core inside
R_crit plus tan-spaced shells of NSHELL particles out to R_sim, core first then shells outward.
Everything is generated in-process from fixed seeds so that the oracle and the GPU engine see
bit-identical arrays.  See SURVEY.md 8(d).
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np

from .api import Globals, PI, UNIT_T, UNIT_V, TOPO_R3, TOPO_S1R2_LOOKUP, TOPO_S1R2_NOLOOKUP, TOPO_T3, calculate_softening_length

H0_KMS = 67.66
OMEGA_M = 0.3111
OMEGA_L = 0.6889


@dataclass
class IC:
    g: Globals
    x: np.ndarray  # [3N] AoS
    v: np.ndarray  # [3N] AoS
    name: str


def _rlim(i, d_s, n_r, last):
    return d_s * np.tan(i * np.pi / (2.0 * (n_r + last)))


def _random_directions(rng, n):
    u = rng.normal(size=(n, 3))
    u /= np.linalg.norm(u, axis=1)[:, None]
    return u


def _base_globals(topology, REAL, n) -> Globals:
    g = Globals(topology=topology, REAL=REAL, N=n)
    g.COSMOLOGY, g.COMOVING_INTEGRATION = 1, 1
    g.H0 = H0_KMS / UNIT_V
    g.Omega_m, g.Omega_lambda, g.Omega_r, g.Omega_b = OMEGA_M, OMEGA_L, 0.0, 0.0
    g.a_start = 1.0 / 64.0
    g.ACC_PARAM = 0.005  # examples/LCDM_SP_1860_com_VOI100.param
    g.h_min = 2e-5 / UNIT_T
    g.h_max = 0.03125 / UNIT_T
    return g


def compactified_r3(n_total: int, n_r: int, n_shell: int, seed: int, REAL=np.float64, d_s: float = 105.0,
                    r_sim: float = 930.0266, r_crit: float = 82.25, perturb: float = 0.05, name: str = "r3") -> IC:
    """Compactified R^3 zoom-in load: n_total particles = core + (n_r - i_crit)*n_shell shell particles."""
    rng = np.random.default_rng(seed)
    last = n_r * math.pi / (2 * math.atan(r_sim / d_s)) - n_r
    i_crit = int(math.atan(r_crit / d_s) * (2.0 * (n_r + last)) / math.pi)
    n_out = (n_r - i_crit) * n_shell
    n_in = n_total - n_out
    if n_in <= 0:
        raise ValueError(f"n_total={n_total} too small for {n_r - i_crit} shells of {n_shell}")
    g = _base_globals(TOPO_R3, REAL, n_total)
    g.Rsim = r_sim
    rho_mean = g.rho_crit * OMEGA_M
    r_core = _rlim(i_crit, d_s, n_r, last)
    pos = np.empty((n_total, 3))
    mass = np.empty(n_total)
    # core: uniform in the ball
    pos[:n_in] = _random_directions(rng, n_in) * (r_core * rng.random(n_in) ** (1.0 / 3.0))[:, None]
    mass[:n_in] = rho_mean * (4.0 * math.pi / 3.0 * r_core**3) / n_in
    # shells outward
    for k, j in enumerate(range(i_crit, n_r)):
        r0, r1 = _rlim(j, d_s, n_r, last), _rlim(j + 1, d_s, n_r, last)
        sl = slice(n_in + k * n_shell, n_in + (k + 1) * n_shell)
        rad = 0.5 * (r0 + r1) + (r1 - r0) * (rng.random(n_shell) - 0.5)
        pos[sl] = _random_directions(rng, n_shell) * rad[:, None]
        mass[sl] = rho_mean * (4.0 * math.pi / 3.0 * (r1**3 - r0**3)) / n_shell
    # small Gaussian perturbation, sigma = perturb * local spacing, to mimic an LCDM IC
    spacing = np.cbrt(mass / rho_mean)
    pos += rng.normal(size=pos.shape) * (perturb * spacing)[:, None]
    vel = rng.normal(size=pos.shape) * (50.0 / math.sqrt(g.a_start) / UNIT_V)
    g.M = np.ascontiguousarray(mass, dtype=REAL)
    g.ParticleRadi = float(np.cbrt(mass[0] / rho_mean)) / 40.0  # generator's own recommendation (...py:550)
    g.set_background()
    calculate_softening_length(g)
    return IC(g, np.ascontiguousarray(pos.reshape(-1), dtype=REAL), np.ascontiguousarray(vel.reshape(-1), dtype=REAL), name)


def t3_lattice(n_side: int, seed: int, REAL=np.float64, L: float = 100.0, is_periodic: int = 2, name: str = "t3") -> IC:
    """T^3: n_side^3 particles on a cubic lattice in [0,L)^3 displaced by sigma = 0.1*L/n_side, equal masses."""
    rng = np.random.default_rng(seed)
    n = n_side**3
    g = _base_globals(TOPO_T3, REAL, n)
    g.L, g.IS_PERIODIC = L, is_periodic
    ax = (np.arange(n_side) + 0.5) * (L / n_side)
    pos = np.stack(np.meshgrid(ax, ax, ax, indexing="ij"), axis=-1).reshape(-1, 3)
    pos = pos + rng.normal(size=pos.shape) * (0.1 * L / n_side)
    pos = np.mod(pos, L)
    mass = np.full(n, OMEGA_M * g.rho_crit * L**3 / n)  # passes the 1e-5 mass check of main.cc:1349
    vel = rng.normal(size=pos.shape) * (50.0 / math.sqrt(g.a_start) / UNIT_V)
    g.M = np.ascontiguousarray(mass, dtype=REAL)
    g.ParticleRadi = (L / n_side) / 40.0
    g.set_background()
    calculate_softening_length(g)
    x = np.ascontiguousarray(pos.reshape(-1), dtype=REAL)
    # keep strictly inside [0, L) after the cast to REAL
    x[x >= REAL(L)] = REAL(0.0)
    return IC(g, x, np.ascontiguousarray(vel.reshape(-1), dtype=REAL), name)


def s1r2_cylinder(n_total: int, n_r: int, n_shell: int, seed: int, REAL=np.float64, lookup: bool = False, is_periodic: int = 2,
                  L: float = 100.0, r_sim: float = 500.0, d_s: float = 45.0, r_crit: float = 75.0, perturb: float = 0.05,
                  name: str = "s1r2") -> IC:
    """S^1xR^2 slab: constant-resolution core rho < r_crit plus tan-spaced cylindrical shells (...py:80-97), z uniform in [0, L)."""
    rng = np.random.default_rng(seed)
    last = n_r * math.pi / (2 * math.atan(r_sim / d_s)) - n_r
    i_crit = int(math.atan(r_crit / d_s) * (2.0 * (n_r + last)) / math.pi)
    n_out = (n_r - i_crit) * n_shell
    n_in = n_total - n_out
    if n_in <= 0:
        raise ValueError(f"n_total={n_total} too small for {n_r - i_crit} shells of {n_shell}")
    g = _base_globals(TOPO_S1R2_LOOKUP if lookup else TOPO_S1R2_NOLOOKUP, REAL, n_total)
    g.L, g.Rsim, g.IS_PERIODIC = L, r_sim, is_periodic
    g.RADIAL_FORCE_TABLE_SIZE = 500
    rho_mean = g.rho_crit * OMEGA_M
    r_core = _rlim(i_crit, d_s, n_r, last)
    pos = np.empty((n_total, 3))
    mass = np.empty(n_total)
    phi = rng.random(n_total) * 2 * math.pi
    rad = np.empty(n_total)
    rad[:n_in] = r_core * np.sqrt(rng.random(n_in))
    mass[:n_in] = rho_mean * (math.pi * r_core**2 * L) / n_in
    for k, j in enumerate(range(i_crit, n_r)):
        r0, r1 = _rlim(j, d_s, n_r, last), _rlim(j + 1, d_s, n_r, last)
        sl = slice(n_in + k * n_shell, n_in + (k + 1) * n_shell)
        rad[sl] = 0.5 * (r0 + r1) + (r1 - r0) * (rng.random(n_shell) - 0.5)
        mass[sl] = rho_mean * (math.pi * (r1**2 - r0**2) * L) / n_shell
    pos[:, 0] = rad * np.cos(phi)
    pos[:, 1] = rad * np.sin(phi)
    pos[:, 2] = rng.random(n_total) * L
    spacing = np.cbrt(mass / rho_mean)
    pos += rng.normal(size=pos.shape) * (perturb * np.minimum(spacing, 0.2 * L))[:, None]
    pos[:, 2] = np.mod(pos[:, 2], L)
    vel = rng.normal(size=pos.shape) * (50.0 / math.sqrt(g.a_start) / UNIT_V)
    g.M = np.ascontiguousarray(mass, dtype=REAL)
    g.ParticleRadi = float(np.cbrt(mass[0] / rho_mean)) / 40.0
    g.set_background()
    calculate_softening_length(g)
    x = np.ascontiguousarray(pos.reshape(-1), dtype=REAL)
    z = x[2::3]
    z[z >= REAL(L)] = REAL(0.0)
    return IC(g, x, np.ascontiguousarray(vel.reshape(-1), dtype=REAL), name)


def random_sphere(n: int, seed: int, REAL=np.float64, two_species: bool = True, radius: float = 10.0, cosmology: int = 1,
                  comoving: int = 1, particle_radii: float = 0.2) -> IC:
    """small generic test load: uniform ball, two mass species, softening large enough that all three
    branches of force_softening (forces.cc:52-87) are exercised."""
    rng = np.random.default_rng(seed)
    g = _base_globals(TOPO_R3, REAL, n)
    g.COSMOLOGY, g.COMOVING_INTEGRATION = cosmology, comoving
    g.Rsim = radius
    pos = _random_directions(rng, n) * (radius * rng.random(n) ** (1.0 / 3.0))[:, None]
    mass = np.full(n, 1.0)
    if two_species:
        mass[n // 2:] = 8.0
    rho_mean = g.rho_crit * OMEGA_M
    mass *= rho_mean * (4.0 * math.pi / 3.0 * radius**3) / mass.sum()
    vel = rng.normal(size=pos.shape) * 0.5
    g.M = np.ascontiguousarray(mass, dtype=REAL)
    g.ParticleRadi = particle_radii
    g.set_background()
    calculate_softening_length(g)
    return IC(g, np.ascontiguousarray(pos.reshape(-1), dtype=REAL), np.ascontiguousarray(vel.reshape(-1), dtype=REAL), "sphere")


# the BASELINE.json configurations
def config_c1(REAL=np.float64) -> IC:
    return compactified_r3(32768, 64, 432, 20241, REAL, name="C1 compactified R^3 N=32768")


def config_c2(REAL=np.float64, n_total: int = 2_000_000) -> IC:
    return compactified_r3(n_total, 224, 14000, 20242, REAL, name=f"C2 compactified R^3 zoom-in N={n_total}")


def config_c5(REAL=np.float32) -> IC:
    return compactified_r3(16_777_216, 640, 38400, 20245, REAL, name="C5 compactified R^3 N=16777216 FP32")
