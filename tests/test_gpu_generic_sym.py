"""GPU (B200): the action-reaction kernel of the table-lookup topologies (pair_generic_sym.cuh): T^3 with the tricubic Ewald
correction and the S^1xR^2 lookup build, against the reference's own forces, against the one-sided kernel, ragged sizes,
multi-pass, KDK steps and every rank of a multi-GPU job played on one GPU.

First run on a B200 at the start of round 2 (profiles/r2a_*.log): all green; part of the default `-m gpu` suite since."""
import ctypes as C
import os

import numpy as np
import pytest

import steps_b200 as sb
from helpers import rel_err
from oracle import pyref
from steps_b200 import _lib, ic

pytestmark = [pytest.mark.gpu]
TOL64 = 1e-12


def make(case, size, seed, REAL=np.float64, is_periodic=2):
    if case == "t3":
        return ic.t3_lattice(size, seed, REAL, L=30.0, is_periodic=is_periodic)
    return ic.s1r2_cylinder(size, 24, max(1, size // 40), seed, REAL, lookup=True, is_periodic=is_periodic, L=20.0, r_sim=30.0, d_s=8.0, r_crit=10.0)


def reference_forces(c):
    g = c.g
    v = pyref.VARIANT[(g.topology, 8 if g.REAL == np.float64 else 4)]
    if not pyref.available(v):
        pytest.skip("tables need oracle/_ref")
    r = pyref.Reference(v)
    r.configure(g, 400)
    r.build_tables()
    r.export_tables(g)
    g.mass_in_unit_sphere = r.scalars()["mass_in_unit_sphere"]
    return r.forces(c.x, 0, g.N - 1, 0)


def engine_forces(c, symmetric):
    g = c.g
    eng = sb.Engine(g, 0)
    eng.set_symmetric(symmetric)
    eng.upload(c.x, c.v)
    eng.forces()
    F = eng.download_forces(0, g.N - 1)
    used = eng.symmetric
    shape = eng.launch_shape(0, g.N - 1)
    eng.close()
    return F, used, shape


@pytest.mark.parametrize("case,size", [("t3", 14), ("s1r2", 3000)])
def test_generic_sym_vs_reference_and_one_sided(case, size):
    c = make(case, size, 61)
    Fo = reference_forces(c)
    F1, used1, _ = engine_forces(c, False)
    F2, used2, shape = engine_forces(c, True)
    assert not used1 and used2, "the action-reaction path must actually be the one that ran"
    e, e1 = rel_err(F2, Fo), rel_err(F2, F1)
    print(f"{case} sym N={c.g.N} shape={shape}: vs reference |dF|/|F| p99 {np.percentile(e, 99):.2e} max {e.max():.2e}; "
          f"vs one-sided p99 {np.percentile(e1, 99):.2e} max {e1.max():.2e}")
    assert np.isfinite(F2).all()
    assert np.percentile(e, 99) < TOL64
    assert e.max() < 50 * TOL64


def test_generic_sym_t3_nearest_image_only():
    """IS_PERIODIC = 1: no table, quasi-periodic nearest-image sum"""
    c = make("t3", 12, 63, is_periodic=1)
    Fo = reference_forces(c)
    F, used, _ = engine_forces(c, True)
    assert used
    e = rel_err(F, Fo)
    assert np.percentile(e, 99) < TOL64 and e.max() < 50 * TOL64


@pytest.mark.parametrize("case,size", [("t3", 8), ("t3", 9), ("t3", 11), ("s1r2", 513), ("s1r2", 777), ("s1r2", 2381)])
def test_generic_sym_ragged_sizes(case, size):
    """partial last i-block (256) and partial last j-tile (128): the padded records must never be evaluated"""
    c = make(case, size, 200 + size)
    Fo = reference_forces(c)
    F, used, _ = engine_forces(c, True)
    assert used
    e = rel_err(F, Fo)
    assert np.isfinite(F).all()
    assert np.percentile(e, 99) < TOL64 and e.max() < 50 * TOL64


def test_generic_sym_too_small_falls_back():
    c = make("t3", 6, 5)  # 216 particles: one i-block
    Fo = reference_forces(c)
    F, used, _ = engine_forces(c, True)
    assert not used
    assert rel_err(F, Fo).max() < 50 * TOL64


def test_generic_sym_softened_and_coincident_pairs_across_blocks():
    c = make("t3", 12, 13)
    g = c.g
    x = c.x.reshape(-1, 3)
    x[5] = x[1200]                                              # r = 0 across blocks
    x[100] = x[900] + 1e-3 * g.SOFT_LENGTH[100]                 # deep inside the softening radius, across blocks
    x[:] = np.mod(x, 30.0)
    Fo = reference_forces(c)
    F, used, _ = engine_forces(c, True)
    assert used and np.isfinite(F).all()
    e = rel_err(F, Fo)
    assert np.percentile(e, 99) < TOL64 and e.max() < 50 * TOL64


def test_generic_sym_f32():
    c = make("t3", 12, 17, np.float32)
    Fo = reference_forces(c)
    F1, _, _ = engine_forces(c, False)
    F2, used, _ = engine_forces(c, True)
    assert used
    e_ref, e_one = rel_err(F2, Fo), rel_err(F2, F1)
    print(f"t3 fp32 sym: vs reference |dF|/|F| p99 {np.percentile(e_ref, 99):.2e}; vs one-sided p99 {np.percentile(e_one, 99):.2e}")
    assert np.percentile(e_ref, 99) < 1e-4 and np.percentile(e_one, 99) < 1e-4


def test_generic_sym_deterministic_and_multi_pass(monkeypatch):
    c = make("t3", 13, 21)
    reference_forces(c)  # table into c.g
    F1, used, _ = engine_forces(c, True)
    F2, _, _ = engine_forces(c, True)
    assert used and np.array_equal(F1, F2)
    monkeypatch.setenv("STEPS_B200_SYM_GPART_MB", "1")
    F3, _, _ = engine_forces(c, True)
    assert np.abs(F3 - F1).max() / np.abs(F1).max() < 1e-13
    P = (c.g.M[:, None] * F1.reshape(-1, 3)).sum(axis=0)  # T^3 has no background term: the pair sum conserves momentum
    assert np.abs(P).max() < 1e-11 * np.abs(c.g.M[:, None] * F1.reshape(-1, 3)).sum()


def test_generic_sym_kdk_steps_match_one_sided_engine():
    c = make("t3", 12, 43)
    reference_forces(c)
    g = c.g
    out = []
    for symmetric in (False, True):
        eng = sb.Engine(g, 0)
        eng.set_symmetric(symmetric)
        eng.upload(c.x, c.v)
        eng.forces()
        h = eng.calculate_init_h()
        errs = []
        for _ in range(4):
            errs.append(eng.step(h))
            h = eng.next_h()
        assert eng.symmetric == symmetric
        out.append((eng.download(), errs))
        eng.close()
    (x1, v1, F1), e1 = out[0]
    (x2, v2, F2), e2 = out[1]
    assert np.allclose(e1, e2, rtol=1e-9)
    assert np.abs(x1 - x2).max() / g.L < 1e-12


@pytest.mark.parametrize("nranks", [2, 3, 4])
def test_generic_sym_multi_rank_on_one_gpu(nranks):
    c = make("t3", 14, 42)
    Fo = reference_forces(c)
    g = c.g
    lib = _lib.load()
    engines, fsyms, ranges = [], [], []
    for r in range(nranks):
        eng = sb.Engine(g, 0)
        _lib.check(lib.steps_b200_engine_debug_set_rank(eng._h, r, nranks, 1))
        assert eng.symmetric
        eng.i_lo, eng.i_hi = eng.range()
        ranges.append((eng.i_lo, eng.i_hi))
        eng.upload(c.x, c.v)
        eng.forces()
        n_pad = C.c_int()
        _lib.check(lib.steps_b200_engine_debug_fsym(eng._h, None, None, C.byref(n_pad)))
        f = np.empty(3 * n_pad.value)
        _lib.check(lib.steps_b200_engine_debug_fsym(eng._h, f.ctypes.data, None, None))
        engines.append(eng)
        fsyms.append(f)
    assert ranges[0][0] == 0 and ranges[-1][1] == g.N and all(ranges[k][1] == ranges[k + 1][0] for k in range(nranks - 1))
    total = np.sum(fsyms, axis=0)
    F = np.empty(3 * g.N)
    for eng, (lo, hi) in zip(engines, ranges):
        _lib.check(lib.steps_b200_engine_debug_fsym(eng._h, None, total.ctypes.data, None))
        F[3 * lo: 3 * hi] = eng.download_forces(lo, hi - 1)
        eng.close()
    e = rel_err(F, Fo)
    print(f"t3 sym, {nranks} ranks on one GPU: |dF|/|F| p99 {np.percentile(e, 99):.2e} max {e.max():.2e}")
    assert np.percentile(e, 99) < TOL64 and e.max() < 50 * TOL64
