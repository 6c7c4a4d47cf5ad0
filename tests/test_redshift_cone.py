"""Redshift-cone output (reference: write_redshift_cone, inputoutput.cc:314-405, ASCII branch).
CPU tier: the host formatter steps_b200_redshift_cone_ascii_host writes, for the particles the reference selects, byte for byte the
lines the unmodified reference (oracle/_ref) writes -- both precisions, both unit conventions, successive radial bins with the
persistent IN_CONE flags, and the end-of-run mode with its shell search.
GPU tier: the device selection (cone_select_kernel through the group API) picks exactly those particles, keeps its flags across calls,
and the file written from its rows is the reference's file."""
import ctypes as C
import os

import numpy as np
import pytest

from oracle import pyref
from steps_b200 import _lib, ic

PI = C.POINTER(C.c_int)
PD = C.POINTER(C.c_double)


def _case(REAL, n=3000, seed=5):
    c = ic.random_sphere(n, seed, REAL, radius=100.0)
    limits = np.array([90.0, 70.0, 50.0, 30.0, 10.0, 0.0])           # r_bin_limits: comoving distances, descending
    zlist = np.array([0.05, 0.04, 0.03, 0.02, 0.01, 0.0])            # out_list: output redshifts
    return c, limits, zlist


def _distance(x, REAL):
    q = np.asarray(x, dtype=REAL).reshape(-1, 3)
    a, b, c = q[:, 0] * q[:, 0], q[:, 1] * q[:, 1], q[:, 2] * q[:, 2]
    return np.sqrt((a + b) + c).astype(np.float64)


def _rows(c, sel, REAL):
    rows = np.zeros((sel.size, 8), dtype=REAL)
    rows[:, 0:3] = c.x.reshape(-1, 3)[sel]
    rows[:, 3:6] = c.v.reshape(-1, 3)[sel]
    rows[:, 6] = c.g.M[sel]
    return np.ascontiguousarray(rows), np.ascontiguousarray(sel.astype(np.int32))


def _format(path, rows, idx, REAL, h0_dimless, all_, limits, zlist, z_index):
    lib = _lib.load()
    _lib.check(lib.steps_b200_redshift_cone_ascii_host(path.encode(), rows.ctypes.data, idx.ctypes.data_as(PI), idx.size, 8 if REAL == np.float64 else 4,
                                                       h0_dimless, all_, limits.ctypes.data_as(PD), limits.size, zlist.ctypes.data_as(PD), z_index))


@pytest.mark.parametrize("REAL,variant", [(np.float64, "r3_f64"), (np.float32, "r3_f32")])
@pytest.mark.parametrize("h0_units,n", [(0, 3000), (1, 3000), (0, 20000)])
def test_formatter_matches_the_reference_writer(tmp_path, REAL, variant, h0_units, n):
    if not pyref.available(variant):
        pytest.skip("needs oracle/_ref")
    c, limits, zlist = _case(REAL, n=n, seed=5 if n == 3000 else 9)
    g = c.g
    r = pyref.Reference(variant)
    r.configure(g)
    ref_dir, our = str(tmp_path / "ref") + "/", str(tmp_path / "ours.dat")
    os.makedirs(ref_dir)
    # H0_dimless as inputoutput.cc:326-333 forms it: REAL H0*UNIT_V/100 with H0_INDEPENDENT_UNITS, else 1
    h0_dimless = float(REAL(g.H0 * 20.738652969925447 / 100.0)) if h0_units else 1.0
    D = _distance(c.x, REAL)
    in_cone = np.zeros(g.N, dtype=bool)
    for z_index in (0, 1, 3):  # three successive radial bins: limits[z_index + 1] <= D
        path = r.write_redshift_cone(ref_dir, c.x, c.v, limits, zlist, z_index, 0, 0, h0_units, reset=(z_index == 0), t_next=zlist[z_index])
        sel = np.flatnonzero((limits[z_index + 1] <= D) & ~in_cone)
        in_cone[sel] = True
        rows, idx = _rows(c, sel, REAL)
        _format(our, rows, idx, REAL, h0_dimless, 0, limits, zlist, z_index)
        assert sel.size > 0
    # the end of the run: everything still outside, with the shell search of the reference (z_index = 3: bins 4, 5 are searched)
    r.write_redshift_cone(ref_dir, c.x, c.v, limits, zlist, 3, 0, 1, h0_units, reset=False, t_next=zlist[3])
    sel = np.flatnonzero(~in_cone)
    rows, idx = _rows(c, sel, REAL)
    _format(our, rows, idx, REAL, h0_dimless, 1, limits, zlist, 3)
    a, b = open(path, "rb").read(), open(our, "rb").read()
    assert len(a) > 100 * n and a.count(b"\n") == g.N
    assert a == b, "redshift cone file differs from the reference's"


def test_formatter_argument_checks(tmp_path):
    lib = _lib.load()
    lim = np.array([1.0, 0.0])
    assert lib.steps_b200_redshift_cone_ascii_host(str(tmp_path / "x").encode(), None, None, 0, 8, 1.0, 0, lim.ctypes.data_as(PD), 2, lim.ctypes.data_as(PD), 5) != 0
    assert lib.steps_b200_redshift_cone_ascii_host(str(tmp_path / "x").encode(), None, None, 0, 8, 1.0, 0, lim.ctypes.data_as(PD), 2, lim.ctypes.data_as(PD), 0) == 0
    assert os.path.getsize(tmp_path / "x") == 0


@pytest.mark.gpu
@pytest.mark.parametrize("REAL,variant", [(np.float64, "r3_f64"), (np.float32, "r3_f32")])
def test_device_selection_reproduces_the_reference_file(tmp_path, REAL, variant):
    if not pyref.available(variant):
        pytest.skip("needs oracle/_ref")
    c, limits, zlist = _case(REAL, n=20000, seed=9)
    g = c.g
    lib = _lib.load()
    grp = C.c_void_p()
    p = g.cparams()
    _lib.check(lib.steps_b200_group_create(C.byref(grp), C.byref(p), 8 if REAL == np.float64 else 4, 1, 0))
    _lib.check(lib.steps_b200_group_upload(grp, c.x.ctypes.data, c.v.ctypes.data, g.M.ctypes.data, g.SOFT_LENGTH.ctypes.data, None))
    r = pyref.Reference(variant)
    r.configure(g)
    ref_dir, our = str(tmp_path / "ref") + "/", str(tmp_path / "ours.dat")
    os.makedirs(ref_dir)
    D = _distance(c.x, REAL)
    in_cone = np.zeros(g.N, dtype=bool)
    cnt = C.c_int()
    for z_index, all_ in ((0, 0), (2, 0), (2, 0), (3, 1)):  # the third call finds nothing new: the flags persist
        path = r.write_redshift_cone(ref_dir, c.x, c.v, limits, zlist, z_index, 0, all_, 0, reset=(z_index == 0), t_next=zlist[z_index])
        _lib.check(lib.steps_b200_group_cone_select(grp, float(limits[z_index + 1]), all_, C.byref(cnt)))
        want = np.flatnonzero(~in_cone) if all_ else np.flatnonzero((limits[z_index + 1] <= D) & ~in_cone)
        in_cone[want] = True
        idx = np.empty(cnt.value, dtype=np.int32)
        rows = np.empty((cnt.value, 8), dtype=REAL)
        _lib.check(lib.steps_b200_group_cone_rows(grp, rows.ctypes.data, idx.ctypes.data_as(PI)))
        assert np.array_equal(idx, want.astype(np.int32)), "the device selected other particles than the reference's rule"
        assert np.array_equal(rows[:, 0:3], c.x.reshape(-1, 3)[want]) and np.array_equal(rows[:, 6], g.M[want])
        _lib.check(lib.steps_b200_group_cone_write_ascii(grp, our.encode(), 1.0, all_, limits.ctypes.data_as(PD), limits.size, zlist.ctypes.data_as(PD), z_index))
    assert in_cone.all()
    assert open(path, "rb").read() == open(our, "rb").read()
    _lib.check(lib.steps_b200_group_cone_reset(grp))
    _lib.check(lib.steps_b200_group_cone_select(grp, 0.0, 1, C.byref(cnt)))
    assert cnt.value == g.N
    lib.steps_b200_group_destroy(grp)
