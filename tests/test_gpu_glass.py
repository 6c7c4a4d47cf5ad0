"""GPU (B200): glass-making mode of the resident engine (glass_kernels.cuh: G = -1 and the diagnostics of the reference's
-DGLASS_MAKING step(), step.cc:107-148, :254-303) against the CPU port pinned to the reference's glass build (tests/test_glass.py),
in both precisions and with every rank of a multi-GPU job played on one GPU.

First run on a B200 at the start of round 2 (profiles/r2a_*.log): all green; part of the default `-m gpu` suite since."""
import os

import numpy as np
import pytest

import steps_b200 as sb
from helpers import rel_err
from oracle import pyport, pyref
from steps_b200 import ic
from test_glass import port_glass_run

pytestmark = [pytest.mark.gpu]
KEYS = ("F_mean", "Fmax", "A_mean", "A_max", "dmean", "dmax", "V_mean", "V_max")


@pytest.mark.parametrize("REAL,n", [(np.float64, 3000), (np.float64, 777), (np.float32, 3000)])
def test_glass_engine_matches_port(REAL, n):
    c = ic.random_sphere(n, 7, REAL)
    c.v[:] = 0
    g = c.g
    eng = sb.Engine(g, 0)
    eng.set_glass_making(True)
    eng.upload(c.x, c.v)
    eng.forces()
    h = eng.calculate_init_h()
    hs, errs, stats = [], [], []
    for _ in range(4):
        hs.append(h)
        errs.append(eng.step(h))
        s = eng.glass_stats()
        stats.append([s[k] for k in KEYS])
        h = eng.next_h()
    x1, v1, F1 = eng.download()
    eng.close()
    x, v = c.x.copy(), c.v.copy()
    errs_p, stats_p, F = port_glass_run(g, x, v, hs)
    f64 = REAL == np.float64
    tol = 1e-10 if f64 else 2e-4
    print(f"glass {REAL.__name__} N={n}: errmax rel diff {np.abs(np.array(errs) / errs_p - 1).max():.2e}, "
          f"stats rel diff {np.abs(np.array(stats) / stats_p - 1).max():.2e}, dx/Rsim {np.abs(x1.astype(np.float64) - x).max() / g.Rsim:.2e}")
    assert np.allclose(errs, errs_p, rtol=tol)
    assert np.allclose(np.array(stats), stats_p, rtol=tol)
    assert np.abs(x1.astype(np.float64) - x).max() / g.Rsim < (1e-12 if f64 else 1e-5) * len(hs)
    assert rel_err(v1, v).max() < 100 * tol


def test_glass_mode_off_is_the_ordinary_step():
    c = ic.random_sphere(2000, 9)
    g = c.g
    out = []
    for toggle in (False, True):
        eng = sb.Engine(g, 0)
        if toggle:  # on and off again: the ordinary step must come back
            eng.set_glass_making(True)
            eng.set_glass_making(False)
        eng.upload(c.x, c.v)
        eng.forces()
        h = eng.calculate_init_h()
        e = eng.step(h)
        out.append((eng.download(), e))
        with pytest.raises(Exception):
            eng.glass_stats()
        eng.close()
    assert out[0][1] == out[1][1] and all(np.array_equal(p, q) for p, q in zip(out[0][0], out[1][0]))


@pytest.mark.skipif(not (pyref.available("r3_f64_glass") and pyref.available("r3_f64_glass", shim=True)), reason="oracle/_ref glass builds not present")
def test_dropin_glass_build_writes_the_reference_logfile(tmp_path):
    """the reference compiled with -DGLASS_MAKING against the same build with forces.cc/step.cc replaced by the shim TUs:
    same h sequence, same state, same Glass_logfile.dat rows (to the %.15f the reference prints)"""
    c = ic.random_sphere(1500, 21)
    c.v[:] = 0
    g = c.g
    logs, states = [], []
    for shim in (False, True):
        r = pyref.Reference("r3_f64_glass", shim=shim)
        r.configure(g)
        d = tmp_path / ("shim" if shim else "ref")
        d.mkdir()
        r.set_out_dir(str(d))
        h = r.kdk_begin(c.x, c.v)
        for _ in range(3):
            h, _out = r.kdk_step(h)
        states.append(r.kdk_state())
        logs.append(r.glass_log())
    (x0, v0, F0), (x1, v1, F1) = states
    assert np.abs(x0 - x1).max() / g.Rsim < 1e-12 * 3
    assert rel_err(v1, v0).max() < 1e-8
    assert logs[0].shape == logs[1].shape == (3, 13)
    assert np.allclose(logs[1], logs[0], rtol=1e-9, atol=2e-15)
