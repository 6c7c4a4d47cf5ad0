"""GPU (B200): the action-reaction R^3 FP64 path (pair_r3_sym.cuh) -- every unordered pair evaluated once and applied
to both particles -- against the oracle, against the one-sided kernel, and through the KDK step.  Same force law and
tolerances as test_gpu_parity.py (north_star: 1e-12 FP64)."""
import os

import numpy as np
import pytest

import steps_b200 as sb
from helpers import noise_err, rel_err
from oracle import pyport
from steps_b200 import ic

pytestmark = pytest.mark.gpu
TOL64 = 1e-12


def engine_forces(c, symmetric, x=None):
    g = c.g
    eng = sb.Engine(g, 0)
    eng.set_symmetric(symmetric)
    eng.upload(c.x if x is None else x, c.v)
    eng.forces()
    F = eng.download_forces(0, g.N - 1)
    used = eng.symmetric
    shape = eng.launch_shape(0, g.N - 1)
    eng.close()
    return F, used, shape


def test_sym_zoom_geometry_vs_oracle_and_one_sided():
    c = ic.compactified_r3(20000, 64, 250, 42, d_s=105.0)
    g = c.g
    Fo = pyport.forces(g, c.x, 0, g.N - 1)
    S = pyport.force_norms(g, c.x, 0, g.N - 1)
    F1, used1, _ = engine_forces(c, False)
    F2, used2, shape = engine_forces(c, True)
    assert not used1 and used2, "symmetric path must actually be the one that ran"
    ne, re_ = noise_err(F2, Fo, S), rel_err(F2, Fo)
    print(f"sym N={g.N} shape={shape}: max |dF|/sum|f| = {ne.max():.3e}; |dF|/|F| p99 {np.percentile(re_, 99):.2e} max {re_.max():.2e}; "
          f"vs one-sided max |dF|/sum|f| = {noise_err(F2, F1, S).max():.3e}")
    assert np.isfinite(F2).all()
    assert ne.max() < TOL64
    assert np.percentile(re_, 99) < TOL64
    assert noise_err(F2, F1, S).max() < 1e-13


@pytest.mark.parametrize("n", [1536, 1537, 2381, 3072, 5000, 7777])
def test_sym_ragged_sizes(n):
    """partial last i-block, partial last j-tile, exactly two blocks"""
    c = ic.random_sphere(n, 50 + n)
    F, used, _ = engine_forces(c, True)
    assert used
    Fo = pyport.forces(c.g, c.x, 0, n - 1)
    assert np.isfinite(F).all()
    assert np.abs(F - Fo).max() / np.abs(Fo).max() < 1e-13
    assert rel_err(F, Fo).max() < 1e-11


def test_sym_too_small_falls_back_to_one_sided_kernel():
    c = ic.random_sphere(700, 5)
    F, used, _ = engine_forces(c, True)
    assert not used  # fewer than two i-blocks: the engine keeps the one-sided CUDA kernel
    Fo = pyport.forces(c.g, c.x, 0, 699)
    assert rel_err(F, Fo).max() < TOL64


def test_sym_coincident_and_softened_pairs_across_blocks():
    """exact-branch slow path on BOTH sides: coincident particles and softened pairs whose partners sit in different i-blocks"""
    c = ic.random_sphere(4000, 13, cosmology=0)
    g = c.g
    c.x[3 * 5: 3 * 5 + 3] = c.x[3 * 3000: 3 * 3000 + 3]        # r = 0 across blocks (block 0 vs block 3)
    c.x[3 * 900: 3 * 900 + 3] = c.x[3 * 901: 3 * 901 + 3]      # r = 0 inside a block
    c.x[3 * 100: 3 * 100 + 3] = c.x[3 * 2500: 3 * 2500 + 3] + 1e-3 * g.SOFT_LENGTH[100]  # deep inside the softening radius
    F, used, _ = engine_forces(c, True)
    assert used
    Fo = pyport.forces(g, c.x, 0, g.N - 1)
    assert np.isfinite(F).all()
    assert rel_err(F, Fo).max() < TOL64
    # softening larger than the system: every pair takes a softened branch, on both sides
    g.SOFT_LENGTH = np.full(g.N, 40.0)
    F, _, _ = engine_forces(c, True)
    Fo = pyport.forces(g, c.x, 0, g.N - 1)
    assert rel_err(F, Fo).max() < TOL64
    # mixed: per-tile thresholds differ strongly
    g.SOFT_LENGTH = np.where(np.arange(g.N) % 2 == 0, 3.0, 1e-4)
    F, _, _ = engine_forces(c, True)
    Fo = pyport.forces(g, c.x, 0, g.N - 1)
    assert rel_err(F, Fo).max() < TOL64


def test_sym_deterministic_and_multi_pass(monkeypatch):
    c = ic.random_sphere(9000, 21)
    F1, used, _ = engine_forces(c, True)
    F2, _, _ = engine_forces(c, True)
    assert used and np.array_equal(F1, F2)  # fixed summation order: bitwise reproducible
    # bound the j-side partial buffer to a few rows: the evaluation runs in several passes; same rows, same order of addition
    monkeypatch.setenv("STEPS_B200_SYM_GPART_MB", "1")
    F3, _, _ = engine_forces(c, True)
    Fo = pyport.forces(c.g, c.x, 0, c.g.N - 1)
    assert rel_err(F3, Fo).max() < TOL64
    assert np.abs(F3 - F1).max() / np.abs(F1).max() < 1e-14


def test_sym_momentum_conservation():
    c = ic.random_sphere(30000, 31, cosmology=0)
    g = c.g
    F, used, _ = engine_forces(c, True)
    assert used
    F = F.reshape(-1, 3)
    P = (g.M[:, None] * F).sum(axis=0)
    assert np.abs(P).max() < 1e-12 * np.abs(g.M[:, None] * F).sum()


def test_sym_kdk_steps_match_one_sided_engine():
    c = ic.compactified_r3(12000, 64, 150, 43, d_s=105.0)
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
    S = pyport.force_norms(g, x1, 0, g.N - 1)
    assert noise_err(F2, F1, S).max() < 1e-12


_ORACLE = {}


def _oracle_cached(c):
    """oracle forces and force scales of an input, computed once per test process (several tests share the N = 20 000 zoom input)"""
    g = c.g
    key = (g.N, np.asarray(c.x).tobytes(), np.asarray(g.M).tobytes(), np.asarray(g.SOFT_LENGTH).tobytes(), g.COSMOLOGY, g.COMOVING_INTEGRATION)
    if key not in _ORACLE:
        _ORACLE[key] = (pyport.forces(c.g, c.x, 0, c.g.N - 1), pyport.force_norms(c.g, c.x, 0, c.g.N - 1))
    return _ORACLE[key]


@pytest.mark.parametrize("nranks", [2, 3, 4, 8])
def test_sym_multi_rank_rules_and_kernel_on_one_gpu(nranks):
    """one GPU plays every rank of a P-GPU job in turn (rows, rules, pair kernel, row reduction of each rank); the host
    stands in for the all-reduce of the j-side sums.  Leaves only the NCCL call itself to the >= 2 GPU tests."""
    import ctypes as C

    from steps_b200 import _lib

    c = ic.compactified_r3(20000, 64, 250, 42, d_s=105.0)
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
    Fo, S = _oracle_cached(c)
    ne = noise_err(F, Fo, S)
    print(f"{nranks} ranks on one GPU: max |dF|/sum|f| = {ne.max():.3e}")
    assert np.isfinite(F).all()
    assert ne.max() < TOL64
    assert np.percentile(rel_err(F, Fo), 99) < TOL64


def test_sym_large_n_properties():
    """N = 400k zoom geometry: sampled rows against the oracle, momentum, agreement with the one-sided kernel"""
    n = 400_000
    c = ic.compactified_r3(n, 224, max(1, int(0.854 * n / 122)), 20242)
    g = c.g
    F1, _, _ = engine_forces(c, False)
    F2, used, shape = engine_forces(c, True)
    assert used
    rows = np.unique(np.concatenate([np.arange(0, 64), np.random.default_rng(5).integers(0, n, 192), np.arange(n - 64, n)]))
    for lo in (0, n // 2, n - 256):
        Fo = pyport.forces(g, c.x, lo, lo + 255)
        S = pyport.force_norms(g, c.x, lo, lo + 255)
        ne = noise_err(F2[3 * lo: 3 * (lo + 256)], Fo, S)
        assert ne.max() < TOL64
    a, b = F1.reshape(-1, 3), F2.reshape(-1, 3)
    d = np.linalg.norm(a - b, axis=1) / np.maximum(np.linalg.norm(a, axis=1), 1e-300)
    print(f"sym vs one-sided at N={n} shape={shape}: |dF|/|F| p50 {np.median(d):.2e} p99 {np.percentile(d, 99):.2e} max {d.max():.2e}")
    assert np.percentile(d, 99) < TOL64
    assert rows.size > 0
