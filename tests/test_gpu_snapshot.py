"""GPU (B200): the asynchronous ASCII snapshot of the resident engines (snapshot_io.h): the file written in the
background while the engine goes on stepping holds the state of the moment of the call, byte for byte what the host formatter --
itself byte-identical to the reference's writer (tests/test_snapshot_io.py) -- makes of a synchronous download at that moment.

First run on a B200 at the start of round 2 (profiles/r2a_*.log): all green; part of the default `-m gpu` suite since."""
import ctypes as C
import os

import numpy as np
import pytest

import steps_b200 as sb
from steps_b200 import _lib, ic

pytestmark = [pytest.mark.gpu]


@pytest.mark.parametrize("REAL", [np.float64, np.float32])
def test_async_snapshot_holds_the_state_of_the_call(tmp_path, REAL):
    n = 20000
    c = ic.random_sphere(n, 19, REAL)
    g = c.g
    lib = _lib.load()
    grp = C.c_void_p()
    p = g.cparams()
    rb = 8 if REAL == np.float64 else 4
    _lib.check(lib.steps_b200_group_create(C.byref(grp), C.byref(p), rb, 1, 0))
    _lib.check(lib.steps_b200_group_upload(grp, c.x.ctypes.data, c.v.ctypes.data, g.M.ctypes.data, g.SOFT_LENGTH.ctypes.data, None))
    _lib.check(lib.steps_b200_group_forces(grp))
    a0 = g.a_start
    H0 = sb.CALCULATE_Hubble_param(g, a0)
    em = C.c_double()
    _lib.check(lib.steps_b200_group_init_errmax(grp, a0, H0, C.byref(em)))
    h = (2 * g.ACC_PARAM / em.value) ** 0.5
    a1 = sb.friedmann_solver_step(g, a0, h)
    _lib.check(lib.steps_b200_group_kdk_step(grp, h, a0, H0, a1, sb.CALCULATE_Hubble_param(g, a1), C.byref(em)))
    # state of this moment, synchronously, for the comparison
    x1, v1 = np.empty(3 * n, dtype=REAL), np.empty(3 * n, dtype=REAL)
    _lib.check(lib.steps_b200_group_download(grp, x1.ctypes.data, v1.ctypes.data, None))
    path = str(tmp_path / "async.dat")
    _lib.check(lib.steps_b200_group_snapshot_ascii_async(grp, path.encode(), 0.6766, a1, 0))
    # keep stepping while the writer works: the file must not see these steps
    a, H = a1, sb.CALCULATE_Hubble_param(g, a1)
    for _ in range(3):
        an = sb.friedmann_solver_step(g, a, h)
        Hn = sb.CALCULATE_Hubble_param(g, an)
        _lib.check(lib.steps_b200_group_kdk_step(grp, h, a, H, an, Hn, C.byref(em)))
        a, H = an, Hn
    _lib.check(lib.steps_b200_group_snapshot_wait(grp))
    want = str(tmp_path / "sync.dat")
    _lib.check(lib.steps_b200_snapshot_ascii_host(want.encode(), x1.ctypes.data, v1.ctypes.data, np.ascontiguousarray(g.M, dtype=REAL).ctypes.data, n, rb,
                                                  0.6766, a1, 0, 2))
    lib.steps_b200_group_destroy(grp)
    assert open(path, "rb").read() == open(want, "rb").read()
