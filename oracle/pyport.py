"""ctypes wrapper over oracle/liboracle.so -- the plain-C restatement (steps_oracle.c).
TEST INFRASTRUCTURE: used only by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs (when oracle/_ref is unavailable)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "liboracle.so")


class OracleParams(C.Structure):
    _fields_ = [("topology", C.c_int), ("n", C.c_int), ("cosmology", C.c_int), ("comoving", C.c_int), ("is_periodic", C.c_int),
                ("interp_order", C.c_int), ("table_dim0", C.c_int), ("table_dim1", C.c_int), ("radial_size", C.c_int),
                ("nthreads", C.c_int), ("L", C.c_double), ("Rsim", C.c_double), ("mass_in_unit_sphere", C.c_double),
                ("H0", C.c_double), ("Omega_lambda", C.c_double), ("ewald_table", C.c_void_p), ("radial_table", C.c_void_p)]


_lib = None


def build() -> None:
    subprocess.run(["make", "-C", HERE, "port"], check=True, stdout=subprocess.DEVNULL)


def load() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            build()
        _lib = C.CDLL(LIB)
        _lib.oracle_force_softening_f64.restype = C.c_double
        _lib.oracle_force_softening_f64.argtypes = [C.c_double, C.c_double]
        _lib.oracle_force_softening_f32.restype = C.c_float
        _lib.oracle_force_softening_f32.argtypes = [C.c_float, C.c_float]
        for f in ("oracle_kick_errmax_f64", "oracle_kick_errmax_f32", "oracle_glass_kick_errmax_f64", "oracle_glass_kick_errmax_f32",
                  "oracle_friedmann_step", "oracle_hubble"):
            getattr(_lib, f).restype = C.c_double
        _lib.oracle_friedmann_step.argtypes = [C.c_double] * 7
        _lib.oracle_hubble.argtypes = [C.c_double] * 6
    return _lib


def _params(g, nthreads=0, keep=None):
    p = OracleParams()
    p.topology, p.n, p.cosmology, p.comoving, p.is_periodic = g.topology, g.N, g.COSMOLOGY, g.COMOVING_INTEGRATION, g.IS_PERIODIC
    p.interp_order = g.EWALD_INTERPOLATION_ORDER
    p.nthreads = nthreads
    p.L, p.Rsim, p.mass_in_unit_sphere, p.H0, p.Omega_lambda = g.L, g.Rsim, g.mass_in_unit_sphere, g.H0, g.Omega_lambda
    tab = g.T3_EWALD_FORCE_TABLE if g.topology == 1 else g.S1R2_EWALD_FORCE_TABLE if g.topology == 2 else None
    if tab is not None:
        t = np.ascontiguousarray(tab, dtype=g.REAL)
        keep.append(t)
        p.ewald_table = t.ctypes.data
        p.table_dim0 = g.N_EWALD_FORCE_GRID if g.topology == 1 else g.Nrho_EWALD_FORCE_GRID
        p.table_dim1 = g.N_EWALD_FORCE_GRID if g.topology == 1 else g.Nz_EWALD_FORCE_GRID
    if g.RADIAL_FORCE_TABLE is not None:
        t = np.ascontiguousarray(g.RADIAL_FORCE_TABLE, dtype=g.REAL)
        keep.append(t)
        p.radial_table = t.ctypes.data
        p.radial_size = t.shape[0]
    return p


def _sfx(g):
    return "_f64" if g.REAL == np.float64 else "_f32"


def forces(g, x, id_min, id_max, nthreads=0) -> np.ndarray:
    lib, keep = load(), []
    p = _params(g, nthreads, keep)
    x = np.ascontiguousarray(x, dtype=g.REAL)
    F = np.zeros(3 * (id_max - id_min + 1), dtype=g.REAL)
    getattr(lib, "oracle_forces" + _sfx(g))(C.byref(p), x.ctypes.data_as(C.c_void_p), g.M.ctypes.data_as(C.c_void_p),
                                             g.SOFT_LENGTH.ctypes.data_as(C.c_void_p), F.ctypes.data_as(C.c_void_p), id_min, id_max)
    return F


def force_norms(g, x, id_min, id_max, nthreads=0) -> np.ndarray:
    """sum_j |f_ij| per i (float64; nearest image in the periodic topologies): the scale of the reference's own summation noise
    """
    lib, keep = load(), []
    p = _params(g, nthreads, keep)
    x = np.ascontiguousarray(x, dtype=np.float64)
    S = np.zeros(id_max - id_min + 1, dtype=np.float64)
    lib.oracle_force_norms_f64(C.byref(p), x.ctypes.data_as(C.c_void_p), np.asarray(g.M, dtype=np.float64).ctypes.data_as(C.c_void_p),
                               np.asarray(g.SOFT_LENGTH, dtype=np.float64).ctypes.data_as(C.c_void_p), S.ctypes.data_as(C.c_void_p),
                               id_min, id_max)
    return S


def force_softening(r, beta, REAL=np.float64) -> float:
    lib = load()
    return float(lib.oracle_force_softening_f64(r, beta) if REAL == np.float64 else lib.oracle_force_softening_f32(r, beta))


def softening(M, particle_radii):
    lib = load()
    M = np.ascontiguousarray(M)
    s = np.empty_like(M)
    if M.dtype == np.float64:
        mm, rp = C.c_double(), C.c_double()
        lib.oracle_softening_f64(M.ctypes.data_as(C.c_void_p), M.size, C.c_double(particle_radii), s.ctypes.data_as(C.c_void_p), C.byref(mm), C.byref(rp))
    else:
        mm, rp = C.c_float(), C.c_float()
        lib.oracle_softening_f32(M.ctypes.data_as(C.c_void_p), M.size, C.c_float(particle_radii), s.ctypes.data_as(C.c_void_p), C.byref(mm), C.byref(rp))
    return s, mm.value, rp.value


def kick_drift(g, x, v, F, a, hubble, h) -> None:
    lib, keep = load(), []
    p = _params(g, 0, keep)
    getattr(lib, "oracle_kick_drift" + _sfx(g))(C.byref(p), x.ctypes.data_as(C.c_void_p), v.ctypes.data_as(C.c_void_p),
                                                 F.ctypes.data_as(C.c_void_p), C.c_double(a), C.c_double(hubble), C.c_double(h))


def kick_errmax(g, v, F, a, hubble, h, do_kick=1) -> float:
    lib, keep = load(), []
    p = _params(g, 0, keep)
    return float(getattr(lib, "oracle_kick_errmax" + _sfx(g))(C.byref(p), v.ctypes.data_as(C.c_void_p), F.ctypes.data_as(C.c_void_p),
                                                               g.SOFT_LENGTH.ctypes.data_as(C.c_void_p), C.c_double(a), C.c_double(hubble),
                                                               C.c_double(h), do_kick))


def glass_kick_drift(g, x, v, F, a, hubble, h):
    """GLASS_MAKING first half (step.cc:129-181 with G = -1): returns (sum of displacements, max displacement)"""
    lib, keep = load(), []
    p = _params(g, 0, keep)
    out = (C.c_double * 2)()
    getattr(lib, "oracle_glass_kick_drift" + _sfx(g))(C.byref(p), x.ctypes.data_as(C.c_void_p), v.ctypes.data_as(C.c_void_p),
                                                       F.ctypes.data_as(C.c_void_p), C.c_double(a), C.c_double(hubble), C.c_double(h), out)
    return out[0], out[1]


def glass_kick_errmax(g, v, F, a, hubble, h):
    """GLASS_MAKING second half (step.cc:254-303 with G = -1): returns errmax, (sum|F|, max|F|, sum|A|, max|A|, sum|v|, max|v|)"""
    lib, keep = load(), []
    p = _params(g, 0, keep)
    out = (C.c_double * 6)()
    e = float(getattr(lib, "oracle_glass_kick_errmax" + _sfx(g))(C.byref(p), v.ctypes.data_as(C.c_void_p), F.ctypes.data_as(C.c_void_p),
                                                                  g.SOFT_LENGTH.ctypes.data_as(C.c_void_p), C.c_double(a), C.c_double(hubble),
                                                                  C.c_double(h), out))
    return e, tuple(out)


def friedmann_step(g, a0, h) -> float:
    return float(load().oracle_friedmann_step(g.H0, g.Omega_m, g.Omega_r, g.Omega_lambda, g.Omega_k, a0, h))


def hubble(g, a) -> float:
    return float(load().oracle_hubble(g.H0, g.Omega_m, g.Omega_r, g.Omega_lambda, g.Omega_k, a))
