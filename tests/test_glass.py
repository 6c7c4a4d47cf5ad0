"""CPU tier: the GLASS_MAKING build of step().  The plain-C restatement (oracle/steps_oracle_impl.h:
oracle_glass_kick_drift / oracle_glass_kick_errmax, G = -1 and the diagnostics of step.cc:143-148, :270-303) is pinned against the
UNMODIFIED reference compiled with -DGLASS_MAKING (oracle/_ref/libsteps_ref_r3_f64_glass.so): same h sequence, positions and
velocities after the steps, errmax per step, and the eight statistics the reference appends to Glass_logfile.dat.  The CUDA side
(glass_kernels.cuh) is compared with this port in tests/test_gpu_glass.py."""
import math

import numpy as np
import pytest

from helpers import rel_err
from oracle import pyport, pyref
from steps_b200 import ic

UNIT_V = 20.738652969925447  # global_variables.h:17


def glass_ic(n=400, seed=7):
    c = ic.random_sphere(n, seed)
    c.v[:] = 0.0  # main.cc:1240-1254: a glass-making run starts from rest
    return c


def port_glass_run(g, x, v, hs):
    """KDK steps of the glass build with the port's halves; returns errmax and the statistics of every step in the argument
    order of Log_write_glass (inputoutput.cc:974): F_mean, Fmax, A_mean, A_max, dmean, dmax, V_mean, V_max"""
    F = pyport.forces(g, x, 0, g.N - 1)
    a = g.a_start
    H = pyport.hubble(g, a)
    errs, stats = [], []
    for h in hs:
        dsum, dmax = pyport.glass_kick_drift(g, x, v, F, a, H, float(h))
        F = pyport.forces(g, x, 0, g.N - 1)
        a = pyport.friedmann_step(g, a, float(h))
        H = pyport.hubble(g, a)
        e, (fs, fm, as_, am, vs, vm) = pyport.glass_kick_errmax(g, v, F, a, H, float(h))
        errs.append(e)
        stats.append([fs / g.N, fm, as_ / g.N, am, dsum / g.N, dmax, vs / g.N, vm])
    return np.array(errs), np.array(stats), F


@pytest.mark.skipif(not pyref.available("r3_f64_glass"), reason="oracle/_ref glass variant not built")
def test_glass_port_matches_reference_glass_build(tmp_path):
    c = glass_ic()
    g = c.g
    r = pyref.Reference("r3_f64_glass")
    assert r.is_glass
    r.configure(g)
    r.set_out_dir(str(tmp_path))
    g.mass_in_unit_sphere = r.scalars()["mass_in_unit_sphere"]
    h = r.kdk_begin(c.x, c.v)
    hs, errs_ref = [], []
    for _ in range(4):
        hs.append(h)
        h, out = r.kdk_step(h)
        errs_ref.append(out["errmax"])
    x_ref, v_ref, F_ref = r.kdk_state()
    log = r.glass_log()
    assert log.shape == (4, 13)

    x, v = c.x.copy(), c.v.copy()
    e0 = pyport.kick_errmax(g, v, pyport.forces(g, x, 0, g.N - 1), g.a_start, pyport.hubble(g, g.a_start), 0.0, do_kick=0)
    assert math.isclose(math.sqrt(2 * g.ACC_PARAM / e0), hs[0], rel_tol=1e-10)
    errs, stats, F = port_glass_run(g, x, v, hs)
    assert np.allclose(errs, errs_ref, rtol=1e-9)
    assert np.abs(x - x_ref).max() / g.Rsim < 1e-12 * len(hs)
    assert rel_err(v, v_ref).max() < 1e-8
    assert rel_err(F, F_ref).max() < 1e-9
    # gravity is repulsive in this build: from rest the particles move apart
    assert (np.linalg.norm(x.reshape(-1, 3), axis=1).mean() > np.linalg.norm(c.x.reshape(-1, 3), axis=1).mean())
    # the logfile prints %.15f (velocities in km/s): compare at that resolution
    ref_stats = log[:, 5:13].copy()
    ref_stats[:, 6:8] /= UNIT_V
    assert np.allclose(stats, ref_stats, rtol=1e-9, atol=2e-15)


def test_glass_statistics_are_what_they_say():
    """the port's statistics against a direct numpy evaluation of their definitions"""
    c = glass_ic(300, 11)
    g = c.g
    x, v = c.x.copy(), c.v.copy()
    F0 = pyport.forces(g, x, 0, g.N - 1)
    a0, h = g.a_start, 1e-4
    H0 = pyport.hubble(g, a0)
    dsum, dmax = pyport.glass_kick_drift(g, x, v, F0, a0, H0, h)
    acc0 = -F0 * a0 ** -3 - 2 * H0 * c.v
    v_half = c.v + acc0 * h / 2
    disp = np.linalg.norm((v_half * h).reshape(-1, 3), axis=1)
    assert math.isclose(dsum, disp.sum(), rel_tol=1e-12) and math.isclose(dmax, disp.max(), rel_tol=1e-12)
    assert np.allclose(x, c.x + v_half * h, rtol=0, atol=1e-13 * g.Rsim)
    F1 = pyport.forces(g, x, 0, g.N - 1)
    a1 = pyport.friedmann_step(g, a0, h)
    H1 = pyport.hubble(g, a1)
    e, (fs, fm, as_, am, vs, vm) = pyport.glass_kick_errmax(g, v, F1, a1, H1, h)
    acc1 = -F1 * a1 ** -3 - 2 * H1 * v_half
    A = np.linalg.norm(acc1.reshape(-1, 3), axis=1)
    Fa = np.linalg.norm(F1.reshape(-1, 3), axis=1)
    V = np.linalg.norm((v_half + acc1 * h / 2).reshape(-1, 3), axis=1)
    assert math.isclose(e, (A / g.SOFT_LENGTH).max(), rel_tol=1e-12)
    for got, want in ((fs, Fa.sum()), (fm, Fa.max()), (as_, A.sum()), (am, A.max()), (vs, V.sum()), (vm, V.max())):
        assert math.isclose(got, want, rel_tol=1e-11)
