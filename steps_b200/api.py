"""Host-side mirror of the reference's force/step interface, over the C ABI.

Names, argument meaning and error behaviour follow the reference so that tests read like calls
into StePS (all citations are StePS/src/...):

  ``forces(g, x, F, ID_min, ID_max)``            forces.cc:510 / forces_cuda.cu:74
  ``forces_periodic(g, x, F, ID_min, ID_max)``   forces.cc:776 / forces_cuda.cu:169
  ``forces_periodic_z(g, x, F, ID_min, ID_max)`` forces.cc:1221 / forces_cuda.cu:457
  ``calculate_softening_length(g)``              utils.cc:59
  ``Engine.step(h)`` / ``calculate_init_h``      step.cc:100 / step.cc:35

``g`` (:class:`Globals`) stands for the C++ globals the reference functions read at link time
(global_variables.h:42-147).  Arrays are numpy, AoS ``x[3*i+k]``, dtype = the build's REAL.
"""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass, field
from typing import Optional

import numpy as np

from . import _lib
from ._lib import CCosmo, CParams, StepsError, TOPO_R3, TOPO_S1R2_LOOKUP, TOPO_S1R2_NOLOOKUP, TOPO_T3, check

UNIT_T = 47.14829951063323  # global_variables.h:17
UNIT_V = 20.738652969925447  # global_variables.h:18
PI = 3.14159265358979323846264338327950288419716939937510


@dataclass
class Globals:
    """The reference's globals on the force/step path (same names, global_variables.h)."""

    topology: int = TOPO_R3  # which compile-time build is mirrored (-DPERIODIC / -DPERIODIC_Z [...NOLOOKUP])
    REAL: type = np.float64  # -DUSE_SINGLE_PRECISION -> np.float32
    N: int = 0
    COSMOLOGY: int = 1
    COMOVING_INTEGRATION: int = 1
    IS_PERIODIC: int = 0
    L: float = 0.0
    Rsim: float = 0.0
    H0: float = 0.0  # internal units (HubbleConstant / UNIT_V, read_paramfile.cc:401-403)
    Omega_m: float = 0.0
    Omega_lambda: float = 0.0
    Omega_r: float = 0.0
    Omega_b: float = 0.0
    ParticleRadi: float = 0.0
    ACC_PARAM: float = 0.0
    h_min: float = 0.0
    h_max: float = 0.0
    a_start: float = 1.0
    EWALD_INTERPOLATION_ORDER: int = 4
    RADIAL_FORCE_TABLE_SIZE: int = 0
    M: Optional[np.ndarray] = None
    SOFT_LENGTH: Optional[np.ndarray] = None
    M_min: float = 0.0
    rho_part: float = 0.0
    mass_in_unit_sphere: float = 0.0
    T3_EWALD_FORCE_TABLE: Optional[np.ndarray] = None
    N_EWALD_FORCE_GRID: int = 0
    S1R2_EWALD_FORCE_TABLE: Optional[np.ndarray] = None
    Nrho_EWALD_FORCE_GRID: int = 0
    Nz_EWALD_FORCE_GRID: int = 0
    RADIAL_FORCE_TABLE: Optional[np.ndarray] = None
    ForceError: bool = False
    n_GPU: int = 1  # devices driven by one call (the reference's argv[2], main.cc:1186-1195)
    _keep: list = field(default_factory=list, repr=False)

    @property
    def Omega_k(self) -> float:
        return 1.0 - self.Omega_m - self.Omega_lambda - self.Omega_r

    @property
    def rho_crit(self) -> float:  # main.cc:1262
        return 3.0 * self.H0 * self.H0 / (8.0 * PI)

    def set_background(self) -> None:
        """mass_in_unit_sphere as main.cc:1269/1288 (S^1xR^2), :1313 (R^3); 0 otherwise."""
        self.mass_in_unit_sphere = 0.0
        if self.COSMOLOGY == 1 and self.COMOVING_INTEGRATION == 1:
            if self.topology in (TOPO_S1R2_LOOKUP, TOPO_S1R2_NOLOOKUP):
                self.mass_in_unit_sphere = float(self.REAL(2.0 * PI * self.rho_crit * self.Omega_m))
            elif self.topology == TOPO_R3:
                self.mass_in_unit_sphere = float(self.REAL(4.0 * PI * self.rho_crit * self.Omega_m / 3.0))

    def cparams(self) -> CParams:
        p = CParams()
        p.abi_version = _lib.ABI_VERSION
        p.topology = self.topology
        p.n = self.N
        p.cosmology = self.COSMOLOGY
        p.comoving = self.COMOVING_INTEGRATION
        p.is_periodic = self.IS_PERIODIC
        p.s1r2_interp_order = self.EWALD_INTERPOLATION_ORDER
        p.L = self.L
        p.Rsim = self.Rsim
        p.mass_in_unit_sphere = self.mass_in_unit_sphere
        p.H0 = self.H0
        p.Omega_lambda = self.Omega_lambda
        p.radial_table_size = 0
        self._keep.clear()
        if self.topology == TOPO_T3 and self.T3_EWALD_FORCE_TABLE is not None:
            t = np.ascontiguousarray(self.T3_EWALD_FORCE_TABLE, dtype=self.REAL)
            self._keep.append(t)
            p.ewald_table = t.ctypes.data
            p.table_dim0 = self.N_EWALD_FORCE_GRID
            p.table_dim1 = self.N_EWALD_FORCE_GRID
        if self.topology == TOPO_S1R2_LOOKUP and self.S1R2_EWALD_FORCE_TABLE is not None:
            t = np.ascontiguousarray(self.S1R2_EWALD_FORCE_TABLE, dtype=self.REAL)
            self._keep.append(t)
            p.ewald_table = t.ctypes.data
            p.table_dim0 = self.Nrho_EWALD_FORCE_GRID
            p.table_dim1 = self.Nz_EWALD_FORCE_GRID
        if self.RADIAL_FORCE_TABLE is not None:
            t = np.ascontiguousarray(self.RADIAL_FORCE_TABLE, dtype=self.REAL)
            self._keep.append(t)
            p.radial_table = t.ctypes.data
            p.radial_table_size = int(t.shape[0])
        return p

    def ccosmo(self) -> CCosmo:
        return CCosmo(self.H0, self.Omega_m, self.Omega_r, self.Omega_lambda, self.Omega_k)


def _real_bytes(g: Globals) -> int:
    return 8 if g.REAL == np.float64 else 4


def _arr(a: np.ndarray, g: Globals, n: int, name: str) -> np.ndarray:
    if not isinstance(a, np.ndarray) or a.dtype != g.REAL or not a.flags.c_contiguous:
        raise TypeError(f"{name} must be a C-contiguous numpy array of dtype {np.dtype(g.REAL)}")
    if a.size < n:
        raise ValueError(f"{name} has {a.size} elements, needs {n}")
    return a


def calculate_softening_length(g: Globals) -> None:
    """utils.cc:59-82: sets g.M_min, g.rho_part, g.SOFT_LENGTH from g.M and g.ParticleRadi."""
    lib = _lib.load()
    M = _arr(g.M, g, g.N, "M")
    soft = np.empty(g.N, dtype=g.REAL)
    if g.REAL == np.float64:
        mm, rp = C.c_double(), C.c_double()
        check(lib.steps_b200_softening_f64(M.ctypes.data, g.N, g.ParticleRadi, soft.ctypes.data, C.byref(mm), C.byref(rp)))
    else:
        mm, rp = C.c_float(), C.c_float()
        check(lib.steps_b200_softening_f32(M.ctypes.data, g.N, g.ParticleRadi, soft.ctypes.data, C.byref(mm), C.byref(rp)))
    g.SOFT_LENGTH, g.M_min, g.rho_part = soft, float(mm.value), float(rp.value)


def _forces_any(g: Globals, topo_ok: tuple, x: np.ndarray, F: np.ndarray, ID_min: int, ID_max: int) -> None:
    if g.topology not in topo_ok:
        raise StepsError("this entry point does not exist in a build of that topology (step.cc:191-197)")
    lib = _lib.load()
    n_out = 3 * (ID_max - ID_min + 1)
    x = _arr(x, g, 3 * g.N, "x")
    F = _arr(F, g, n_out, "F")
    M = _arr(g.M, g, g.N, "M")
    s = _arr(g.SOFT_LENGTH, g, g.N, "SOFT_LENGTH")
    p = g.cparams()
    if g.n_GPU > 1:
        fn = lib.steps_b200_forces_multi_f64 if g.REAL == np.float64 else lib.steps_b200_forces_multi_f32
        rc = fn(C.byref(p), x.ctypes.data, M.ctypes.data, s.ctypes.data, F.ctypes.data, ID_min, ID_max, g.n_GPU)
    else:
        fn = lib.steps_b200_forces_f64 if g.REAL == np.float64 else lib.steps_b200_forces_f32
        rc = fn(C.byref(p), x.ctypes.data, M.ctypes.data, s.ctypes.data, F.ctypes.data, ID_min, ID_max)
    if rc != 0:
        g.ForceError = True  # reference convention: forces_cuda.cu:970-974 + main.cc:1851-1856
        check(rc)


def forces(g: Globals, x: np.ndarray, F: np.ndarray, ID_min: int, ID_max: int) -> None:
    """R^3: ``void forces(REAL*x, REAL*F, int ID_min, int ID_max)``; F is overwritten, index relative to ID_min."""
    _forces_any(g, (TOPO_R3,), x, F, ID_min, ID_max)


def forces_periodic(g: Globals, x: np.ndarray, F: np.ndarray, ID_min: int, ID_max: int) -> None:
    """T^3: ``void forces_periodic(REAL*x, REAL*F, int ID_min, int ID_max)``."""
    _forces_any(g, (TOPO_T3,), x, F, ID_min, ID_max)


def forces_periodic_z(g: Globals, x: np.ndarray, F: np.ndarray, ID_min: int, ID_max: int) -> None:
    """S^1xR^2: ``void forces_periodic_z(REAL*x, REAL*F, int ID_min, int ID_max)``."""
    _forces_any(g, (TOPO_S1R2_LOOKUP, TOPO_S1R2_NOLOOKUP), x, F, ID_min, ID_max)


def force_entry(g: Globals):
    """the entry point step.cc:191-197 selects for this build"""
    return {TOPO_R3: forces, TOPO_T3: forces_periodic}.get(g.topology, forces_periodic_z)


def friedmann_solver_step(g: Globals, a0: float, h: float) -> float:
    """friedmann_solver.cc:100-159 (COSMOPARAM=0)"""
    c = g.ccosmo()
    return _lib.load().steps_b200_friedmann_step(C.byref(c), a0, h)


def CALCULATE_Hubble_param(g: Globals, a: float) -> float:
    """friedmann_solver.cc:161-164"""
    c = g.ccosmo()
    return _lib.load().steps_b200_hubble(C.byref(c), a)


def partition(n: int, nranks: int, rank: int) -> tuple:
    lo, hi = C.c_int(), C.c_int()
    _lib.load().steps_b200_partition(n, nranks, rank, C.byref(lo), C.byref(hi))
    return lo.value, hi.value


def t3_ewald_defaults(is_periodic: int, L: float) -> dict:
    """main.cc:425-446: grid size, alpha and cuts of the T^3 Ewald table for IS_PERIODIC = 2, 3, 4"""
    ng, al, rel, rec = C.c_int(), C.c_double(), C.c_double(), C.c_double()
    check(_lib.load().steps_b200_t3_ewald_defaults(is_periodic, L, C.byref(ng), C.byref(al), C.byref(rel), C.byref(rec)))
    return {"ngrid": ng.value, "alpha": al.value, "rel_cut": rel.value, "rec_cut": rec.value}


def calculate_t3_ewald_lookup_table(g: Globals, device: int = 0) -> np.ndarray:
    """ewald_space.cc:288-383 on the GPU: builds T3_EWALD_FORCE_TABLE for g.IS_PERIODIC / g.L (setup of main.cc:425-494) and
    stores it, with N_EWALD_FORCE_GRID, in g (REAL of the build).  Returns the FP64 table [Ngrid, Ngrid, Ngrid, 3]."""
    d = t3_ewald_defaults(g.IS_PERIODIC, g.L)
    n = d["ngrid"]
    tab = np.empty(n * n * n * 3, dtype=np.float64)
    check(_lib.load().steps_b200_t3_ewald_table_f64(n, g.L, d["alpha"], d["rel_cut"], d["rec_cut"], tab.ctypes.data, device))
    g.N_EWALD_FORCE_GRID = n
    g.T3_EWALD_FORCE_TABLE = np.ascontiguousarray(tab, dtype=g.REAL)
    return tab.reshape(n, n, n, 3)


def s1r2_ewald_defaults(is_periodic: int, L: float, Rsim: float) -> dict:
    """main.cc:575-605: dimensions and Ewald parameters of the S^1xR^2 lookup table"""
    nrho, nz, nmax, mmax = C.c_int(), C.c_int(), C.c_int(), C.c_int()
    rho_max, alpha = C.c_double(), C.c_double()
    check(_lib.load().steps_b200_s1r2_ewald_defaults(is_periodic, L, Rsim, C.byref(nrho), C.byref(nz), C.byref(rho_max), C.byref(alpha),
                                                     C.byref(nmax), C.byref(mmax)))
    return {"nrho": nrho.value, "nz": nz.value, "rho_max": rho_max.value, "alpha": alpha.value, "nmax": nmax.value, "mmax": mmax.value}


def calculate_S1R2ewald_correction_table(g: Globals, device: int = 0) -> np.ndarray:
    """ewald_space.cc:754-798 on the GPU: builds S1R2_EWALD_FORCE_TABLE for g.IS_PERIODIC / g.L / g.Rsim (setup of main.cc:562-705),
    stores it with Nrho_/Nz_EWALD_FORCE_GRID in g (REAL of the build).  Returns the FP64 table [Nrho, Nz, 2]."""
    d = s1r2_ewald_defaults(g.IS_PERIODIC, g.L, g.Rsim)
    tab = np.empty(d["nrho"] * d["nz"] * 2, dtype=np.float64)
    check(_lib.load().steps_b200_s1r2_ewald_table_f64(d["nrho"], d["nz"], d["rho_max"], g.L, d["alpha"], d["nmax"], d["mmax"],
                                                      tab.ctypes.data, device))
    g.Nrho_EWALD_FORCE_GRID, g.Nz_EWALD_FORCE_GRID = d["nrho"], d["nz"]
    g.S1R2_EWALD_FORCE_TABLE = np.ascontiguousarray(tab, dtype=g.REAL)
    return tab.reshape(d["nrho"], d["nz"], 2)


def get_cylindrical_force_table(g: Globals, accuracy: int = 7500, device: int = 0) -> np.ndarray:
    """utils.cc:162-228 on the GPU: RADIAL_FORCE_TABLE for g (S^1xR^2), with the Lz the reference passes at main.cc:1263-1310:
    L/2 in the quasi-periodic mode (IS_PERIODIC == 1), L * (IS_PERIODIC + 1 - 0.4) in the NOLOOKUP image-sum mode."""
    Lz = 0.5 * g.L if g.IS_PERIODIC == 1 else g.L * ((g.IS_PERIODIC + 1) - 0.4)
    tab = np.empty(g.RADIAL_FORCE_TABLE_SIZE, dtype=np.float64)
    check(_lib.load().steps_b200_radial_force_table_f64(g.Rsim, Lz, g.RADIAL_FORCE_TABLE_SIZE, accuracy, tab.ctypes.data, device))
    g.RADIAL_FORCE_TABLE = np.ascontiguousarray(tab, dtype=g.REAL)
    return tab


def sym_rules(n: int, nranks: int, rank: int, ib_size: int):
    """host-only rule builder of the action-reaction path: -> (i_lo, i_hi, rules[nb, 16]) or None"""
    nb_max = (n + ib_size - 1) // ib_size + 1
    out = np.zeros((nb_max, 16), dtype=np.int32)
    lo, hi = C.c_int(), C.c_int()
    nb = _lib.load().steps_b200_sym_rules(n, nranks, rank, ib_size, C.byref(lo), C.byref(hi), out.ctypes.data_as(C.POINTER(C.c_int)), nb_max)
    if nb < 0:
        return None
    return lo.value, hi.value, out[:nb]


def fma_peak(device: int, real_bytes: int) -> tuple:
    """measured FMA-pipe TFLOP/s (2 flop/FMA) and implied SM clock (MHz) -- the roofline denominator"""
    tf, mhz = C.c_double(), C.c_double()
    check(_lib.load().steps_b200_fma_peak(device, real_bytes, C.byref(tf), C.byref(mhz)))
    return tf.value, mhz.value


def fma_peak_sustained(device: int, real_bytes: int, seconds: float = 2.0) -> float:
    """FMA-pipe TFLOP/s of the same microbenchmark run back to back for `seconds` (power-capped clocks)"""
    tf = C.c_double()
    check(_lib.load().steps_b200_fma_peak_sustained(device, real_bytes, seconds, C.byref(tf)))
    return tf.value


class Engine:
    """Device-resident x, v, F, M, s + KDK stepping (replaces step(), step.cc:100-312).

    One Engine per process per GPU.  For multi-GPU runs every rank creates one and calls
    :meth:`comm_init` with the 128-byte NCCL id produced by rank 0 (:func:`nccl_unique_id`).
    """

    def __init__(self, g: Globals, device: int = 0):
        self.g = g
        self.lib = _lib.load()
        self._h = C.c_void_p()
        p = g.cparams()
        check(self.lib.steps_b200_engine_create(C.byref(self._h), C.byref(p), _real_bytes(g), device))
        self.a = g.a_start if g.COSMOLOGY == 1 else 1.0
        self.T = 0.0
        self.Hubble_param = CALCULATE_Hubble_param(g, self.a) if (g.COSMOLOGY == 1 and g.COMOVING_INTEGRATION == 1) else 0.0
        self.errmax = 0.0
        self.i_lo, self.i_hi = 0, g.N

    def close(self) -> None:
        if self._h:
            self.lib.steps_b200_engine_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @staticmethod
    def nccl_unique_id() -> bytes:
        buf = C.create_string_buffer(128)
        check(_lib.load().steps_b200_nccl_unique_id(buf))
        return buf.raw

    def comm_init(self, unique_id: bytes, rank: int, nranks: int) -> None:
        buf = C.create_string_buffer(unique_id, 128)
        check(self.lib.steps_b200_engine_comm_init(self._h, buf, rank, nranks))
        self.i_lo, self.i_hi = self.range()

    def range(self) -> tuple:
        """rows [i_lo, i_hi) this engine owns under the partition in force (i-block aligned in symmetric mode)"""
        lo, hi = C.c_int(), C.c_int()
        check(self.lib.steps_b200_engine_range(self._h, C.byref(lo), C.byref(hi)))
        return lo.value, hi.value

    def set_symmetric(self, on: bool) -> None:
        """action-reaction evaluation of the R^3 FP64 path (before comm_init; every rank the same choice)"""
        check(self.lib.steps_b200_engine_set_symmetric(self._h, 1 if on else 0))
        self.i_lo, self.i_hi = self.range()

    @property
    def symmetric(self) -> bool:
        return bool(self.lib.steps_b200_engine_is_symmetric(self._h))

    def upload(self, x: np.ndarray, v: Optional[np.ndarray] = None) -> None:
        g = self.g
        x = _arr(x, g, 3 * g.N, "x")
        vp = _arr(v, g, 3 * g.N, "v").ctypes.data if v is not None else None
        check(self.lib.steps_b200_engine_upload(self._h, x.ctypes.data, vp, _arr(g.M, g, g.N, "M").ctypes.data,
                                                _arr(g.SOFT_LENGTH, g, g.N, "SOFT_LENGTH").ctypes.data))

    def upload_x(self, x: np.ndarray) -> None:
        check(self.lib.steps_b200_engine_upload_x(self._h, _arr(x, self.g, 3 * self.g.N, "x").ctypes.data))

    def forces(self, ID_min: Optional[int] = None, ID_max: Optional[int] = None) -> None:
        lo = self.i_lo if ID_min is None else ID_min
        hi = self.i_hi - 1 if ID_max is None else ID_max
        check(self.lib.steps_b200_engine_forces(self._h, lo, hi))

    def download_forces(self, ID_min: int, ID_max: int) -> np.ndarray:
        F = np.empty(3 * (ID_max - ID_min + 1), dtype=self.g.REAL)
        check(self.lib.steps_b200_engine_download_forces(self._h, F.ctypes.data, ID_min, ID_max))
        return F

    def download(self, want_x=True, want_v=True, want_F=True):
        n3 = 3 * self.g.N
        x = np.empty(n3, dtype=self.g.REAL) if want_x else None
        v = np.empty(n3, dtype=self.g.REAL) if want_v else None
        F = np.empty(n3, dtype=self.g.REAL) if want_F else None
        check(self.lib.steps_b200_engine_download(self._h, x.ctypes.data if want_x else None, v.ctypes.data if want_v else None,
                                                  F.ctypes.data if want_F else None))
        return x, v, F

    def calculate_init_h(self) -> float:
        """step.cc:35-98: needs forces() first; returns sqrt(2*ACC_PARAM/errmax)."""
        e = C.c_double()
        check(self.lib.steps_b200_engine_init_errmax(self._h, self.a, self.Hubble_param, C.byref(e)))
        self.errmax = e.value
        return math.pow(2 * self.g.ACC_PARAM / self.errmax, 0.5)

    def step(self, h: float) -> float:
        """one KDK step of length h (step.cc:100-312 + main.cc:1712); returns errmax."""
        g = self.g
        self.T += h
        a_new, H_new = self.a, self.Hubble_param
        if g.COSMOLOGY == 1 and g.COMOVING_INTEGRATION == 1:  # step.cc:233-242
            a_new = friedmann_solver_step(g, self.a, h)
            H_new = CALCULATE_Hubble_param(g, a_new)
        e = C.c_double()
        check(self.lib.steps_b200_engine_kdk_step(self._h, h, self.a, self.Hubble_param, a_new, H_new, C.byref(e)))
        self.a, self.Hubble_param, self.errmax = a_new, H_new, e.value
        return self.errmax

    def set_glass_making(self, on: bool) -> None:
        """the arithmetic of a -DGLASS_MAKING build: G = -1 (global_variables.h:19-23) and the diagnostics of step.cc:143-148, :270-303.
        The caller zeroes the velocities before upload(), as main.cc:1240-1254 does."""
        check(self.lib.steps_b200_engine_set_glass_making(self._h, 1 if on else 0))

    def glass_stats(self) -> dict:
        """diagnostics of the last step() in glass-making mode, named as the arguments of Log_write_glass (inputoutput.cc:974)"""
        out = (C.c_double * 8)()
        check(self.lib.steps_b200_engine_glass_stats(self._h, out))
        return dict(zip(("F_mean", "Fmax", "A_mean", "A_max", "dmean", "dmax", "V_mean", "V_max"), out))

    def next_h(self) -> float:
        """main.cc:1834-1842"""
        return self.lib.steps_b200_next_timestep(self.g.ACC_PARAM, self.errmax, self.g.h_min, self.g.h_max)

    def timings(self) -> tuple:
        f, s = C.c_double(), C.c_double()
        check(self.lib.steps_b200_engine_timings(self._h, C.byref(f), C.byref(s)))
        return f.value, s.value

    def pair_kernel_ms(self) -> float:
        m = C.c_double()
        check(self.lib.steps_b200_engine_pair_kernel_ms(self._h, C.byref(m)))
        return m.value

    def mark(self, slot: int) -> None:
        check(self.lib.steps_b200_engine_mark(self._h, slot))

    def elapsed_ms(self, slot_a: int, slot_b: int) -> float:
        m = C.c_double()
        check(self.lib.steps_b200_engine_elapsed_ms(self._h, slot_a, slot_b, C.byref(m)))
        return m.value

    def launch_count(self) -> int:
        return int(self.lib.steps_b200_engine_launch_count(self._h))

    def launch_shape(self, ID_min: int, ID_max: int) -> dict:
        out = (C.c_int * 4)()
        check(self.lib.steps_b200_engine_launch_shape(self._h, ID_min, ID_max, out))
        return {"i_per_cta": out[0], "j_chunks": out[1], "ctas": out[2], "j_tile": out[3]}

    def sync(self) -> None:
        check(self.lib.steps_b200_engine_sync(self._h))
