"""CPU: the plain-C oracle port (oracle/steps_oracle.c) is pinned against
 (a) golden vectors produced by the unmodified reference (tests/golden, tools/make_golden.py),
 (b) the reference itself when oracle/_ref is present,
 (c) analytic known answers."""
import math

import numpy as np
import pytest

from helpers import attach_t3_table, load_golden, needs_t3_table, rel_err
from oracle import pyport, pyref
from steps_b200 import ic

FORCE_CASES = ["r3_f64_comoving", "r3_f64_noncomoving", "r3_f64_nocosmo", "r3_f32_comoving", "r3_f64_zoom", "t3_f64_quasi",
               "t3_f64_ewald", "t3_f32_ewald", "s1r2nl_f64_images", "s1r2nl_f64_quasi", "s1r2nl_f32_images", "s1r2_f64_lookup",
               "s1r2_f64_lookup_quasi", "s1r2_f64_lookup_cic", "s1r2_f64_lookup_ngp"]


def tol_for(g):
    # FP64: the port differs from the reference build only by -ffast-math reassociation and pow/sqrt forms
    return 1e-11 if g.REAL == np.float64 else 2e-4


@pytest.mark.parametrize("name", FORCE_CASES)
def test_port_matches_golden_forces(name):
    g, d = load_golden(name)
    if needs_t3_table(g) and not attach_t3_table(g, d):
        pytest.skip("T^3 Ewald table needs oracle/_ref")
    F = pyport.forces(g, d["x"], 0, g.N - 1)
    assert rel_err(F, d["F"]).max() < tol_for(g)
    lo, hi = int(d["sub_lo"]), int(d["sub_hi"])
    Fs = pyport.forces(g, d["x"], lo, hi)
    assert Fs.shape == d["Fsub"].shape
    assert rel_err(Fs, d["Fsub"]).max() < tol_for(g)
    # sub-range output is indexed relative to ID_min
    assert np.array_equal(Fs, F[3 * lo: 3 * (hi + 1)])


def test_port_softening_kernel_golden():
    d = np.load(pyport.HERE + "/../tests/golden/scalars.npz")
    for r, b, w in zip(d["r"], d["beta"], d["w64"]):
        assert math.isclose(pyport.force_softening(r, b), w, rel_tol=1e-13, abs_tol=0)
    for r, b, w in zip(d["r"], d["beta"], d["w32"]):
        assert math.isclose(pyport.force_softening(np.float32(r), np.float32(b), np.float32), w, rel_tol=2e-6)


def test_softening_continuity_and_limits():
    # forces.cc:52-87: continuous at beta/2 (-> 5.0667/beta^3 = 76/15) and at beta (-> 1/beta^3); r=0 -> 32/(3 beta^3)
    for beta in (0.1, 1.0, 7.0):
        eps = 1e-9 * beta
        a, b = pyport.force_softening(0.5 * beta - eps, beta), pyport.force_softening(0.5 * beta + eps, beta)
        assert math.isclose(a, b, rel_tol=1e-6)
        assert math.isclose(a, 76.0 / 15.0 / beta**3, rel_tol=1e-6)
        a, b = pyport.force_softening(beta - eps, beta), pyport.force_softening(beta + eps, beta)
        assert math.isclose(a, b, rel_tol=1e-6)
        assert math.isclose(b, 1.0 / beta**3, rel_tol=1e-6)
        assert math.isclose(pyport.force_softening(0.0, beta), 32.0 / 3.0 / beta**3, rel_tol=1e-14)


def test_friedmann_and_hubble_golden():
    d = np.load(pyport.HERE + "/../tests/golden/scalars.npz")
    g = ic.random_sphere(8, 1).g
    a = float(d["fr_a"][0])
    for h, a_next, H_next in zip(d["fr_h"], d["fr_a"][1:], d["fr_H"][1:]):
        a = pyport.friedmann_step(g, a, float(h))
        assert math.isclose(a, a_next, rel_tol=1e-14)
        assert math.isclose(pyport.hubble(g, a), H_next, rel_tol=1e-14)


def test_two_body_known_answer():
    c = ic.random_sphere(2, 5, cosmology=0)
    g = c.g
    g.M[:] = [2.0, 3.0]
    g.SOFT_LENGTH[:] = [0.01, 0.02]
    x = np.array([0.0, 0, 0, 3.0, 4.0, 0.0])
    F = pyport.forces(g, x, 0, 1).reshape(2, 3)
    # F_i = m_j d / r^3 (G=1), r = 5
    np.testing.assert_allclose(F[0], 3.0 * np.array([3, 4, 0.0]) / 125.0, rtol=1e-15)
    np.testing.assert_allclose(F[1], -2.0 * np.array([3, 4, 0.0]) / 125.0, rtol=1e-15)


def test_momentum_conservation_and_uniform_sphere():
    # sum_i m_i F_i = 0 without the background term; uniform sphere: pair force ~ -mass_in_unit_sphere * x
    c = ic.random_sphere(2000, 9, two_species=False, cosmology=0)
    g = c.g
    F = pyport.forces(g, c.x, 0, g.N - 1).reshape(-1, 3)
    P = (g.M[:, None] * F).sum(axis=0)
    scale = np.abs(g.M[:, None] * F).sum()
    assert np.abs(P).max() < 1e-12 * scale
    # uniform sphere: the mean radial pair force equals -mass_in_unit_sphere * r (what the background term cancels)
    c2 = ic.random_sphere(2000, 9, two_species=False, cosmology=1, comoving=1)
    X = c2.x.reshape(-1, 3)
    inner = np.linalg.norm(X, axis=1) < 0.6 * c2.g.Rsim
    ratio = (F[inner] * X[inner]).sum() / (c2.g.mass_in_unit_sphere * (X[inner] ** 2).sum())
    assert abs(ratio + 1.0) < 0.1


def test_empty_self_and_coincident():
    c = ic.random_sphere(1, 3, cosmology=0)
    F = pyport.forces(c.g, c.x, 0, 0)
    assert np.array_equal(F, np.zeros(3))  # only the self pair: w(0)*0 = 0
    c = ic.random_sphere(4, 3, cosmology=0)
    c.x[3:6] = c.x[0:3]  # coincident pair: r = 0 hits the inner branch, contributes 0
    F = pyport.forces(c.g, c.x, 0, 3)
    assert np.isfinite(F).all()


KDK_CASES = ["kdk_r3_f64", "kdk_r3_f32", "kdk_t3_f64", "kdk_s1r2nl_f64"]


@pytest.mark.parametrize("name", KDK_CASES)
def test_port_kdk_matches_reference_step(name):
    """KDK with the port's halves + the reference's h sequence reproduces the reference's own step()"""
    g, d = load_golden(name)
    x, v = d["x"].copy(), d["v"].copy()
    F = pyport.forces(g, x, 0, g.N - 1)
    tol = 1e-10 if g.REAL == np.float64 else 5e-4
    assert rel_err(F, d["F0"]).max() < tol
    a = g.a_start
    H = pyport.hubble(g, a)
    e0 = pyport.kick_errmax(g, v, F, a, H, 0.0, do_kick=0)
    assert math.isclose(math.sqrt(2 * g.ACC_PARAM / e0), d["h_seq"][0], rel_tol=tol)
    for k, h in enumerate(d["h_seq"][:-1]):
        pyport.kick_drift(g, x, v, F, a, H, float(h))
        F = pyport.forces(g, x, 0, g.N - 1)
        a = pyport.friedmann_step(g, a, float(h))
        H = pyport.hubble(g, a)
        e = pyport.kick_errmax(g, v, F, a, H, float(h), do_kick=1)
        assert math.isclose(e, d["errmax_seq"][k], rel_tol=10 * tol)
        assert math.isclose(a, d["a_seq"][k], rel_tol=1e-13)
    scale = max(g.Rsim, g.L)
    assert np.abs(x - d["x1"]).max() / scale < (1e-12 if g.REAL == np.float64 else 1e-5) * len(d["h_seq"])
    assert rel_err(v, d["v1"]).max() < 100 * tol


@pytest.mark.skipif(not pyref.available("r3_f64"), reason="oracle/_ref not built")
def test_port_matches_live_reference_r3():
    c = ic.compactified_r3(3000, 16, 120, 77, d_s=20.0, r_sim=150.0, r_crit=25.0)
    g = c.g
    r = pyref.Reference("r3_f64")
    r.configure(g)
    g.mass_in_unit_sphere = r.scalars()["mass_in_unit_sphere"]
    assert np.allclose(r.softening(), g.SOFT_LENGTH, rtol=1e-15)
    Fr = r.forces(c.x, 0, g.N - 1, 0)
    Fp = pyport.forces(g, c.x, 0, g.N - 1)
    assert rel_err(Fp, Fr).max() < 1e-11
