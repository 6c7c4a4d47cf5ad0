"""GPU (B200): the action-reaction kernel of the S^1xR^2 NOLOOKUP image sum (pair_s1r2_sym.cuh) against the reference's
image loop, against the one-sided tuned kernel, for every image count, with softened pairs, pairs at the image cut, z outside
the box (fallback), several passes and every rank of a multi-GPU job played on one GPU.

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
L = 20.0


def cylinder(n, seed, is_periodic=2):
    return ic.s1r2_cylinder(n, 24, max(1, n // 40), seed, lookup=False, is_periodic=is_periodic, L=L, r_sim=60.0, d_s=10.0, r_crit=15.0)


def reference_forces(c, radial_accuracy=400):
    g = c.g
    v = pyref.VARIANT[(g.topology, 8)]
    if not pyref.available(v):
        pytest.skip("tables need oracle/_ref")
    r = pyref.Reference(v)
    r.configure(g, radial_accuracy)
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


@pytest.mark.parametrize("is_periodic", [2, 3, 4])
def test_s1r2_sym_all_image_counts_vs_reference_and_one_sided(is_periodic):
    c = cylinder(4000, 170 + is_periodic, is_periodic)
    x = c.x.reshape(-1, 3)
    x[1] = x[0] + [1e-3, -2e-3, 1.5e-3]           # deep inside the softening length, same block
    x[3000] = x[2] + [2e-3, 1e-3, -1e-3]          # ... across blocks (two-sided slow path)
    x[3001] = x[4] + [0.0, 0.0, 0.4 * L]          # dz at the +-M selection threshold region, across blocks
    x[3002] = x[6] + [0.01, 0.0, 0.6 * L]         # dz at the +-(M-1) cut region, across blocks
    x[:, 2] = np.mod(x[:, 2], L)
    x[8, 2] = 5e-4                                # a pair softened through the m = -1 image: z just above 0 and just below L
    x[3003] = [x[8, 0] + 1e-3, x[8, 1], L - 5e-4]
    Fo = reference_forces(c)
    F1, used1, _ = engine_forces(c, False)
    F2, used2, shape = engine_forces(c, True)
    assert not used1 and used2, "the action-reaction path must actually be the one that ran"
    e, e1 = rel_err(F2, Fo), rel_err(F2, F1)
    print(f"s1r2 sym IS_PERIODIC={is_periodic} shape={shape}: vs reference |dF|/|F| p99 {np.percentile(e, 99):.2e} max {e.max():.2e}; "
          f"vs one-sided max {e1.max():.2e}")
    assert np.isfinite(F2).all()
    assert np.percentile(e, 99) < TOL64
    assert e.max() < 50 * TOL64


@pytest.mark.parametrize("n", [768, 769, 1000, 1537, 2381, 5000])
def test_s1r2_sym_ragged_sizes(n):
    c = cylinder(n, 250 + n)
    Fo = reference_forces(c)
    F, used, _ = engine_forces(c, True)
    assert used
    e = rel_err(F, Fo)
    assert np.isfinite(F).all()
    assert np.percentile(e, 99) < TOL64 and e.max() < 50 * TOL64


def test_s1r2_sym_too_small_falls_back():
    # one i-block of 384 (the default shape): the engine keeps the one-sided kernel
    c = ic.s1r2_cylinder(380, 24, 4, 5, lookup=False, is_periodic=2, L=L, r_sim=60.0, d_s=10.0, r_crit=15.0)
    Fo = reference_forces(c)
    F, used, _ = engine_forces(c, True)
    assert not used
    assert rel_err(F, Fo).max() < 50 * TOL64


def test_s1r2_sym_z_outside_box_takes_the_one_sided_launch():
    c = cylinder(3000, 77)
    x = c.x.reshape(-1, 3)
    x[::7, 2] += L
    x[3::11, 2] -= 2 * L
    Fo = reference_forces(c)
    g = c.g
    eng = sb.Engine(g, 0)
    eng.set_symmetric(True)
    assert eng.symmetric
    eng.upload(c.x, c.v)
    eng.forces()
    e = rel_err(eng.download_forces(0, g.N - 1), Fo)
    assert np.percentile(e, 99) < TOL64 and e.max() < 50 * TOL64
    x[:, 2] = np.mod(x[:, 2], L)  # back inside the box on the same engine: the action-reaction kernel again
    Fo = reference_forces(c)
    eng.upload(c.x, c.v)
    eng.forces()
    e = rel_err(eng.download_forces(0, g.N - 1), Fo)
    eng.close()
    assert np.percentile(e, 99) < TOL64 and e.max() < 50 * TOL64


def test_s1r2_sym_deterministic_multi_pass_and_momentum(monkeypatch):
    c = cylinder(6000, 21)
    reference_forces(c)  # builds the radial table into c.g
    F1, used, _ = engine_forces(c, True)
    F2, _, _ = engine_forces(c, True)
    assert used and np.array_equal(F1, F2)
    monkeypatch.setenv("STEPS_B200_SYM_GPART_MB", "1")
    F3, _, _ = engine_forces(c, True)
    assert np.abs(F3 - F1).max() / np.abs(F1).max() < 1e-14
    # the pairwise part conserves momentum along z (the background term acts on x, y only: forces.cc:1385-1396)
    Pz = (c.g.M * F1.reshape(-1, 3)[:, 2]).sum()
    assert abs(Pz) < 1e-12 * np.abs(c.g.M * F1.reshape(-1, 3)[:, 2]).sum()


def test_s1r2_sym_kdk_steps_match_one_sided_engine():
    c = cylinder(5000, 43)
    reference_forces(c)  # builds the radial table into c.g
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
    assert np.allclose(e1, e2, rtol=1e-10)
    assert np.abs(x1 - x2).max() / g.Rsim < 1e-13


@pytest.mark.parametrize("nranks", [2, 3, 4])
def test_s1r2_sym_multi_rank_on_one_gpu(nranks):
    c = cylinder(8000, 42)
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
    print(f"s1r2 sym, {nranks} ranks on one GPU: |dF|/|F| p99 {np.percentile(e, 99):.2e} max {e.max():.2e}")
    assert np.percentile(e, 99) < TOL64 and e.max() < 50 * TOL64
