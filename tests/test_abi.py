"""CPU: the C-ABI library loads, exports every symbol include/steps_b200.h declares, its host-side
helpers agree with the oracle, and compute calls fail loudly (no fallback) without a GPU."""
import ctypes as C
import math
import os
import re

import numpy as np
import pytest

import steps_b200 as sb
from helpers import load_golden
from oracle import pyport
from steps_b200 import _lib, ic

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "steps_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(steps_b200_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = C.CDLL(_lib.LIB_PATH)
    names = header_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"{n} declared in steps_b200.h but not exported"
    # and the Python binding table covers the same set
    assert sorted(_lib.SYMBOLS) == names


def test_abi_version_and_struct_size():
    lib = _lib.load()
    assert lib.steps_b200_abi_version() == _lib.ABI_VERSION
    assert C.sizeof(_lib.CParams) == 10 * 4 + 5 * 8 + 2 * 8


def test_no_gpu_means_loud_failure_not_fallback():
    lib = _lib.load()
    if lib.steps_b200_device_count() > 0:
        pytest.skip("GPU present")
    c = ic.random_sphere(64, 1)
    F = np.full(3 * 64, 7.0)
    with pytest.raises(sb.StepsError, match="no CUDA device"):
        sb.forces(c.g, c.x, F, 0, 63)
    assert c.g.ForceError is True  # reference error convention (forces_cuda.cu:970-974)
    assert np.all(F == 7.0)  # nothing was computed on the CPU
    with pytest.raises(sb.StepsError, match="no CUDA device"):
        sb.Engine(c.g, 0)


def test_bad_arguments_are_rejected():
    c = ic.random_sphere(64, 1)
    F = np.zeros(3 * 64)
    g = c.g
    g.IS_PERIODIC = 2  # R^3 build refuses periodic boundary flags (main.cc:728-733)
    with pytest.raises(sb.StepsError, match="IS_PERIODIC"):
        sb.forces(g, c.x, F, 0, 63)
    g.IS_PERIODIC = 0
    with pytest.raises(sb.StepsError):
        sb.forces_periodic(g, c.x, F, 0, 63)  # wrong entry point for this build
    with pytest.raises(TypeError):
        sb.forces(g, c.x.astype(np.float32), F, 0, 63)
    with pytest.raises(ValueError):
        sb.forces(g, c.x[:-3], F, 0, 63)


def test_partition_contiguous_balanced():
    for n, p in [(10, 3), (2_000_000, 8), (7, 8), (32768, 4), (1, 1)]:
        hi_prev = 0
        sizes = []
        for r in range(p):
            lo, hi = sb.partition(n, p, r)
            assert lo == hi_prev and hi >= lo
            hi_prev = hi
            sizes.append(hi - lo)
        assert hi_prev == n and max(sizes) - min(sizes) <= 1


def test_softening_length_matches_oracle_and_golden():
    for REAL in (np.float64, np.float32):
        c = ic.compactified_r3(2048, 16, 60, 5, REAL, d_s=20.0, r_sim=150.0, r_crit=25.0)
        s, mmin, rp = pyport.softening(c.g.M, c.g.ParticleRadi)
        assert np.allclose(c.g.SOFT_LENGTH, s, rtol=1e-15 if REAL == np.float64 else 1e-6)
        assert math.isclose(c.g.M_min, mmin, rel_tol=1e-15) and math.isclose(c.g.rho_part, rp, rel_tol=1e-6)
        # s_i = ParticleRadi * cbrt(M_i / M_min)
        assert np.allclose(c.g.SOFT_LENGTH, c.g.ParticleRadi * np.cbrt(c.g.M / c.g.M_min), rtol=1e-6)
    g, d = load_golden("r3_f64_zoom")  # softening computed by the reference itself
    g2 = sb.Globals(REAL=np.float64, N=g.N, ParticleRadi=g.ParticleRadi, M=g.M.copy())
    sb.calculate_softening_length(g2)
    assert np.allclose(g2.SOFT_LENGTH, g.SOFT_LENGTH, rtol=1e-15)


def test_host_scalars_match_golden():
    d = np.load(os.path.join(ROOT, "tests", "golden", "scalars.npz"))
    g = ic.random_sphere(8, 1).g
    a = float(d["fr_a"][0])
    for h, a_next, H_next in zip(d["fr_h"], d["fr_a"][1:], d["fr_H"][1:]):
        a = sb.friedmann_solver_step(g, a, float(h))
        assert math.isclose(a, a_next, rel_tol=1e-14)
        assert math.isclose(sb.CALCULATE_Hubble_param(g, a), H_next, rel_tol=1e-14)
    lib = _lib.load()
    # main.cc:1834-1842
    assert lib.steps_b200_next_timestep(0.005, 1e8, 1e-7, 1e-3) == pytest.approx(math.sqrt(2 * 0.005 / 1e8))
    assert lib.steps_b200_next_timestep(0.005, 1e20, 1e-7, 1e-3) == 1e-7
    assert lib.steps_b200_next_timestep(0.005, 1e-9, 1e-7, 1e-3) == 1e-3
    # main.cc:1843-1846: outputs scheduled in time -> the step ends 1e-9 h_min past the next output time; scheduled in redshift -> untouched
    f = lib.steps_b200_next_timestep_to_output
    assert f(0.005, 1e-9, 1e-7, 1e-3, 0.9995, 1.0, 0) == pytest.approx(1.0 - 0.9995 + 1e-16, rel=1e-12)
    assert f(0.005, 1e-9, 1e-7, 1e-3, 0.9995, 1.0, 1) == 1e-3
    assert f(0.005, 1e-9, 1e-7, 1e-3, 0.5, 1.0, 0) == 1e-3


def test_ic_shapes():
    c = ic.config_c1()
    g = c.g
    assert g.N == 32768 and c.x.shape == (3 * 32768,)
    r = np.linalg.norm(c.x.reshape(-1, 3), axis=1)
    assert r.max() < 1.05 * g.Rsim
    # mass consistency check of main.cc:1332-1368 (R^3: total mass = rho_mean * V(R_sim))
    rho_mean = g.rho_crit * g.Omega_m
    assert math.isclose(g.M.sum(), rho_mean * 4 / 3 * math.pi * g.Rsim**3, rel_tol=1e-9)
    t = ic.t3_lattice(8, 3)
    assert t.x.min() >= 0 and t.x.max() < t.g.L
    assert math.isclose(t.g.M.sum() / t.g.L**3, t.g.rho_crit * t.g.Omega_m, rel_tol=1e-12)
    s = ic.s1r2_cylinder(4096, 24, 100, 4)
    assert s.x[2::3].min() >= 0 and s.x[2::3].max() < s.g.L


def test_dropin_shim_exports_the_reference_symbols_and_fails_loudly_without_gpu():
    """The drop-in build (reference TUs minus forces.cc/step.cc plus steps_b200/csrc/shim/*.cc) must define the
    reference's C++-mangled entry points and, without a GPU, set ForceError instead of computing."""
    import subprocess

    from oracle import pyref

    if not pyref.available("r3_f64", shim=True):
        pytest.skip("oracle/_ref shim build needs /root/reference")
    path = os.path.join(pyref.REF_DIR, "libsteps_shim_r3_f64.so")
    syms = subprocess.run(["nm", "-D", "--defined-only", path], capture_output=True, text=True, check=True).stdout
    for mangled in ("_Z6forcesPdS_ii", "_Z21recalculate_softeningv", "_Z4stepPdS_S_", "_Z16calculate_init_hv"):
        assert mangled in syms, mangled
    if _lib.load().steps_b200_device_count() > 0:
        return
    c = ic.random_sphere(128, 3)
    r = pyref.Reference("r3_f64", shim=True)
    r.configure(c.g)
    with pytest.raises(RuntimeError, match="ForceError"):
        r.forces(c.x, 0, 127)


def test_ewald_space_counts_and_t3_defaults():
    """host side of the T^3 Ewald table producer: the lattice enumeration of ewald_space() (ewald_space.cc:118-196) and the
    IS_PERIODIC -> (grid, cuts) mapping of main.cc:425-446.  Counts must fit the reference's own arrays ([739], [11459])."""
    from steps_b200 import _lib, api

    lib = _lib.load()
    assert lib.steps_b200_ewald_space_count(3.6) == 179      # rel_cut 2.6 + 1  (the reference build here reports last index 178)
    assert lib.steps_b200_ewald_space_count(10.0) == 4139    # rec_cut 8 + 2
    assert lib.steps_b200_ewald_space_count(5.6) <= 739 and lib.steps_b200_ewald_space_count(14.0) <= 11459
    assert lib.steps_b200_ewald_space_count(1.0) == 1        # only the origin
    d = api.t3_ewald_defaults(2, 100.0)
    assert d == {"ngrid": 63, "alpha": 0.02, "rel_cut": 2.6, "rec_cut": 8.0}
    assert api.t3_ewald_defaults(4, 50.0)["ngrid"] == 255
    with pytest.raises(sb.StepsError):
        api.t3_ewald_defaults(1, 100.0)


def test_s1r2_ewald_defaults():
    """main.cc:575-605 restated: table dimensions and Ewald parameters of the S^1xR^2 lookup build"""
    from steps_b200 import api

    d = api.s1r2_ewald_defaults(2, 20.0, 30.0)
    assert (d["nz"], d["nrho"], d["nmax"], d["mmax"]) == (128, 432, 4, 10)
    assert d["rho_max"] == 67.5 and abs(d["alpha"] - 0.787875 / 20.0) < 1e-18
    d = api.s1r2_ewald_defaults(5, 100.0, 500.0)
    assert (d["nz"], d["nmax"], d["mmax"]) == (512, 7, 14)
