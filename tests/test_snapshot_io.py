"""CPU tier: ASCII snapshots.  steps_b200_snapshot_ascii_host -- the formatter behind the asynchronous snapshot
of the resident engines (snapshot_io.h) -- against the reference's own write_ascii_snapshot (inputoutput.cc:826-909, run through
oracle/_ref): the files must be byte-identical, in both precisions, in both unit conventions, for any worker count."""
import ctypes as C
import os

import numpy as np
import pytest

from oracle import pyref
from steps_b200 import _lib, ic

UNIT_V = 20.738652969925447


def ours(path, c, a, h0_dimless, zero_v=0, nthreads=0):
    g = c.g
    lib = _lib.load()
    x, v, M = (np.ascontiguousarray(t, dtype=g.REAL) for t in (c.x, c.v, g.M))
    _lib.check(lib.steps_b200_snapshot_ascii_host(path.encode(), x.ctypes.data, v.ctypes.data, M.ctypes.data, g.N, 8 if g.REAL == np.float64 else 4,
                                                  h0_dimless, a, zero_v, nthreads))
    return open(path, "rb").read()


@pytest.mark.parametrize("variant,REAL", [("r3_f64", np.float64), ("r3_f32", np.float32)])
@pytest.mark.parametrize("h0_units", [0, 1])
def test_ascii_snapshot_is_byte_identical_to_the_reference_writer(tmp_path, variant, REAL, h0_units):
    if not pyref.available(variant):
        pytest.skip("oracle/_ref not built")
    c = ic.random_sphere(3001, 17, REAL)
    g = c.g
    r = pyref.Reference(variant)
    r.configure(g)
    a, t_next = 0.37, 0.0123
    ref_path = r.write_ascii_snapshot(str(tmp_path), c.x, c.v, a, t_next, h0_units)
    ref = open(ref_path, "rb").read()
    h0_dimless = g.H0 * UNIT_V / 100.0 if h0_units else 1.0
    got1 = ours(str(tmp_path / "ours1.dat"), c, a, h0_dimless, nthreads=1)
    got4 = ours(str(tmp_path / "ours4.dat"), c, a, h0_dimless, nthreads=4)
    assert got1 == got4, "the worker count must not change the file"
    if got1 != ref:
        # not byte-identical: show how far apart the numbers are before failing
        A = np.array([[float(t) for t in ln.split()] for ln in got1.decode().splitlines()])
        B = np.array([[float(t) for t in ln.split()] for ln in ref.decode().splitlines()])
        diff = np.abs(A - B).max()
        pytest.fail(f"files differ: max abs difference of a printed number {diff:.3e}")
    assert ref.count(b"\n") == g.N


def test_glass_build_prints_zero_velocities(tmp_path):
    if not pyref.available("r3_f64_glass"):
        pytest.skip("oracle/_ref glass variant not built")
    c = ic.random_sphere(500, 3)
    r = pyref.Reference("r3_f64_glass")
    r.configure(c.g)
    ref = open(r.write_ascii_snapshot(str(tmp_path), c.x, c.v, 0.5, 0.02, 0), "rb").read()
    got = ours(str(tmp_path / "ours.dat"), c, 0.5, 1.0, zero_v=1)
    assert got == ref


def test_bad_arguments_fail_loudly(tmp_path):
    lib = _lib.load()
    assert lib.steps_b200_snapshot_ascii_host(None, None, None, None, 0, 8, 1.0, 1.0, 0, 0) != 0
    x = np.zeros(3)
    assert lib.steps_b200_snapshot_ascii_host(b"/nonexistent_dir/x.dat", x.ctypes.data, x.ctypes.data, x.ctypes.data, 1, 8, 1.0, 1.0, 0, 1) != 0
    assert b"cannot open" in lib.steps_b200_last_error()
