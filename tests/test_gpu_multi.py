"""GPU, >= 2 devices: the i-partitioned multi-GPU paths against the single-GPU result.
  - steps_b200_forces_multi_*: the stateless call split over n_GPU devices (replaces forces_cuda(x,F,n_GPU,...))
  - steps_b200_group_*: n resident engines in one process, NCCL position all-gather + errmax all-reduce per KDK step
Skipped on a single-GPU box (the round-end GPU tier); run with `gpurun --gpus 2`."""
import ctypes as C

import numpy as np
import pytest

import steps_b200 as sb
from steps_b200 import _lib, ic

pytestmark = pytest.mark.gpu


def ndev():
    return _lib.load().steps_b200_device_count()


@pytest.fixture(autouse=True)
def _need2():
    if ndev() < 2:
        pytest.skip("needs >= 2 GPUs")


@pytest.mark.parametrize("n", [5000, 5001, 1023])
def test_stateless_multi_gpu_split_matches_single(n):
    c = ic.random_sphere(n, 7)
    g = c.g
    F1 = np.zeros(3 * n)
    sb.forces(g, c.x, F1, 0, n - 1)
    g.n_GPU = min(ndev(), 4)
    F2 = np.full(3 * n, np.nan)
    sb.forces(g, c.x, F2, 0, n - 1)
    # different i-blocking and j-chunking per device: equal to rounding, not bitwise
    scale = np.abs(F1).max()
    assert np.abs(F1 - F2).max() / scale < 1e-13
    lo, hi = 100, n - 37
    F3 = np.full(3 * (hi - lo + 1), np.nan)
    sb.forces(g, c.x, F3, lo, hi)
    assert np.abs(F3 - F1[3 * lo: 3 * (hi + 1)]).max() / scale < 1e-13


class Group:
    def __init__(self, g, n_gpu):
        self.lib = _lib.load()
        self.h = C.c_void_p()
        p = g.cparams()
        _lib.check(self.lib.steps_b200_group_create(C.byref(self.h), C.byref(p), 8 if g.REAL == np.float64 else 4, n_gpu, 0))
        self.g = g

    def close(self):
        self.lib.steps_b200_group_destroy(self.h)


@pytest.mark.parametrize("REAL", [np.float64, np.float32])
def test_group_kdk_matches_single_engine(REAL):
    n = 6001
    c = ic.random_sphere(n, 19, REAL)
    g = c.g
    # single engine
    eng = sb.Engine(g, 0)
    eng.upload(c.x, c.v)
    eng.forces()
    h = eng.calculate_init_h()
    e1 = [eng.step(h) for _ in range(3)]
    x1, v1, F1 = eng.download()
    eng.close()
    # group of 2..4 engines
    k = min(ndev(), 4)
    grp = Group(g, k)
    lib = grp.lib
    assert lib.steps_b200_group_size(grp.h) == k
    _lib.check(lib.steps_b200_group_upload(grp.h, c.x.ctypes.data, c.v.ctypes.data, g.M.ctypes.data, g.SOFT_LENGTH.ctypes.data, None))
    _lib.check(lib.steps_b200_group_forces(grp.h))
    a0 = g.a_start
    H0 = sb.CALCULATE_Hubble_param(g, a0)
    em = C.c_double()
    _lib.check(lib.steps_b200_group_init_errmax(grp.h, a0, H0, C.byref(em)))
    h2 = (2 * g.ACC_PARAM / em.value) ** 0.5
    tol = 1e-12 if REAL == np.float64 else 1e-4
    assert abs(h2 - h) <= tol * h
    a, H = a0, H0
    e2 = []
    for _ in range(3):
        an = sb.friedmann_solver_step(g, a, h)
        Hn = sb.CALCULATE_Hubble_param(g, an)
        _lib.check(lib.steps_b200_group_kdk_step(grp.h, h, a, H, an, Hn, C.byref(em)))
        a, H = an, Hn
        e2.append(em.value)
    x2, v2, F2 = (np.empty(3 * n, dtype=REAL) for _ in range(3))
    _lib.check(lib.steps_b200_group_download(grp.h, x2.ctypes.data, v2.ctypes.data, F2.ctypes.data))
    grp.close()
    assert np.allclose(e1, e2, rtol=1e-10 if REAL == np.float64 else 1e-3)
    assert np.abs(x1 - x2).max() / g.Rsim < (1e-13 if REAL == np.float64 else 1e-5)
    assert np.abs(F1 - F2).max() / np.abs(F1).max() < (1e-12 if REAL == np.float64 else 1e-4)
    assert np.abs(v1 - v2).max() / np.abs(v1).max() < (1e-12 if REAL == np.float64 else 1e-4)


def test_library_calls_restore_the_callers_current_device():
    """a host program's current CUDA device must survive every library call (bench.py's ranks, MPI ranks with their own GPU)"""
    import torch

    torch.cuda.set_device(1)
    try:
        c = ic.random_sphere(2000, 3)
        F = np.zeros(3 * 2000)
        sb.forces(c.g, c.x, F, 0, 1999)  # stateless call, runs on $STEPS_B200_DEVICE (default 0)
        assert torch.cuda.current_device() == 1
        eng = sb.Engine(c.g, 0)
        eng.upload(c.x, c.v)
        eng.forces()
        eng.sync()
        assert torch.cuda.current_device() == 1
        t = torch.zeros(4, device="cuda")
        assert t.device.index == 1
        eng.close()
        assert torch.cuda.current_device() == 1
    finally:
        torch.cuda.set_device(0)


def _group_forces_and_steps(c, k, nsteps=2):
    g = c.g
    REAL = g.REAL
    n = g.N
    grp = Group(g, k)
    lib = grp.lib
    _lib.check(lib.steps_b200_group_upload(grp.h, c.x.ctypes.data, c.v.ctypes.data, g.M.ctypes.data, g.SOFT_LENGTH.ctypes.data, None))
    _lib.check(lib.steps_b200_group_forces(grp.h))
    sym = [bool(lib.steps_b200_engine_is_symmetric(C.c_void_p(lib.steps_b200_group_engine(grp.h, d)))) for d in range(k)]
    F0 = np.empty(3 * n, dtype=REAL)
    _lib.check(lib.steps_b200_group_download(grp.h, None, None, F0.ctypes.data))
    a = g.a_start
    H = sb.CALCULATE_Hubble_param(g, a)
    em = C.c_double()
    _lib.check(lib.steps_b200_group_init_errmax(grp.h, a, H, C.byref(em)))
    h = (2 * g.ACC_PARAM / em.value) ** 0.5
    errs = []
    for _ in range(nsteps):
        an = sb.friedmann_solver_step(g, a, h)
        Hn = sb.CALCULATE_Hubble_param(g, an)
        _lib.check(lib.steps_b200_group_kdk_step(grp.h, h, a, H, an, Hn, C.byref(em)))
        a, H = an, Hn
        errs.append(em.value)
    x, v, F = (np.empty(3 * n, dtype=REAL) for _ in range(3))
    _lib.check(lib.steps_b200_group_download(grp.h, x.ctypes.data, v.ctypes.data, F.ctypes.data))
    grp.close()
    return sym, F0, h, errs, x, v, F


@pytest.mark.parametrize("k", [2, 4, 8])
@pytest.mark.parametrize("case", ["r3_f64", "r3_f32", "s1r2nl"])
def test_group_action_reaction_over_k_gpus(case, k):
    """the action-reaction evaluation split over k GPUs (ring assignment of block pairs, NCCL all-gather of positions and
    all-reduce of the j-side sums) against one GPU: initial forces, the first time step and two KDK steps"""
    if ndev() < k:
        pytest.skip(f"needs {k} GPUs")
    if case == "r3_f64":
        c = ic.compactified_r3(40000, 64, 500, 52, d_s=105.0)
    elif case == "r3_f32":
        c = ic.compactified_r3(40000, 64, 500, 53, np.float32, d_s=105.0)
    else:
        from oracle import pyref

        if not pyref.available("s1r2nl_f64"):
            pytest.skip("radial table needs oracle/_ref")
        c = ic.s1r2_cylinder(24000, 24, 600, 54, lookup=False, is_periodic=2, L=20.0, r_sim=60.0, d_s=10.0, r_crit=15.0)
        sb.get_cylindrical_force_table(c.g, 400, 0)
    tol = 1e-12 if c.g.REAL == np.float64 else 2e-5
    sym1, F0a, ha, ea, xa, va, Fa = _group_forces_and_steps(c, 1)
    symk, F0b, hb, eb, xb, vb, Fb = _group_forces_and_steps(c, k)
    assert all(sym1) and all(symk), "the action-reaction path must be the one that runs"
    scale = np.abs(F0a.astype(np.float64)).max()
    assert np.abs(F0a.astype(np.float64) - F0b).max() / scale < tol
    assert abs(ha - hb) <= (1e-12 if c.g.REAL == np.float64 else 1e-4) * ha
    assert np.allclose(ea, eb, rtol=1e-10 if c.g.REAL == np.float64 else 1e-3)
    assert np.abs(xa.astype(np.float64) - xb).max() / c.g.Rsim < (1e-13 if c.g.REAL == np.float64 else 1e-5)
    assert np.abs(Fa.astype(np.float64) - Fb).max() / scale < tol
