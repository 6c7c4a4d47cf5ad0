"""ctypes wrapper over oracle/_ref/libsteps_ref_<variant>.so -- the UNMODIFIED reference compiled by
oracle/Makefile (see ref_harness.cc).  TEST INFRASTRUCTURE: imported only by tests/, tools/ that
generate golden vectors, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")

VARIANT = {  # (topology, real_bytes) -> variant name
    (0, 8): "r3_f64", (0, 4): "r3_f32", (1, 8): "t3_f64", (1, 4): "t3_f32",
    (2, 8): "s1r2_f64", (3, 8): "s1r2nl_f64", (3, 4): "s1r2nl_f32",
}
# the S^1xR^2 lookup build exists once per EWALD_INTERPOLATION_ORDER (a compile-time macro of the reference): 4 = TSC, 2 = CIC, 0 = NGP
S1R2_ORDER_VARIANT = {4: "s1r2_f64", 2: "s1r2cic_f64", 0: "s1r2ngp_f64"}


def variant_for(g) -> str:
    rb = 8 if g.REAL == np.float64 else 4
    if g.topology == 2 and rb == 8:
        return S1R2_ORDER_VARIANT[int(g.EWALD_INTERPOLATION_ORDER)]
    return VARIANT[(g.topology, rb)]


class SrefConfig(C.Structure):
    _fields_ = [("cosmology", C.c_int), ("comoving", C.c_int), ("is_periodic", C.c_int), ("radial_table_size", C.c_int),
                ("radial_accuracy", C.c_int), ("pad_", C.c_int), ("L", C.c_double), ("Rsim", C.c_double), ("H0", C.c_double),
                ("Omega_m", C.c_double), ("Omega_lambda", C.c_double), ("Omega_r", C.c_double), ("Omega_b", C.c_double),
                ("particle_radii", C.c_double), ("acc_param", C.c_double), ("h_min", C.c_double), ("h_max", C.c_double),
                ("a_start", C.c_double)]


def _kind(shim: bool, cuda: bool) -> str:
    return "shim" if shim else ("refcuda" if cuda else "ref")


def available(variant: str, shim: bool = False, cuda: bool = False) -> bool:
    return os.path.exists(os.path.join(REF_DIR, f"libsteps_{_kind(shim, cuda)}_{variant}.so"))


class Reference:
    """One loaded variant of the reference.  Not re-entrant (the reference is a bag of globals)."""

    def __init__(self, variant: str, shim: bool = False, cuda: bool = False):
        """shim=True loads the DROP-IN build instead: the same reference TUs minus forces.cc/step.cc plus
        steps_b200/csrc/shim/*.cc linked against libstepsb200.so (needs a GPU to compute anything).
        cuda=True loads the reference's OWN CUDA build (-DUSE_CUDA, forces_cuda.cu compiled unmodified for sm_100a; `make -C oracle
        refcuda`): the kernel to beat and a second oracle; needs a GPU and set_n_gpu(>= 1) before forces()."""
        path = os.path.join(REF_DIR, f"libsteps_{_kind(shim, cuda)}_{variant}.so")
        if not os.path.exists(path):
            raise FileNotFoundError(f"{path}: build with `make -C oracle ref` (needs /root/reference)")
        self.lib = C.CDLL(path)
        self.variant = variant
        self.real_bytes = self.lib.sref_real_bytes()
        self.REAL = np.float64 if self.real_bytes == 8 else np.float32
        self.creal = C.c_double if self.real_bytes == 8 else C.c_float
        self.topology = self.lib.sref_topology()
        self.shim = bool(self.lib.sref_is_shim())
        assert self.shim == shim
        if not shim and not cuda:
            self.lib.sref_force_softening.restype = self.creal
            self.lib.sref_force_softening.argtypes = [self.creal, self.creal]
        self.lib.sref_table.restype = C.c_void_p
        self.lib.sref_kdk_begin.restype = C.c_double
        self.lib.sref_kdk_step.restype = C.c_double
        self.lib.sref_kdk_step.argtypes = [C.c_double, C.POINTER(C.c_double)]
        self.lib.sref_friedmann_step.restype = C.c_double
        self.lib.sref_friedmann_step.argtypes = [C.c_double, C.c_double]
        self.lib.sref_hubble.restype = C.c_double
        self.lib.sref_hubble.argtypes = [C.c_double]
        self.N = 0

    @classmethod
    def for_globals(cls, g) -> "Reference":
        return cls(variant_for(g))

    def configure(self, g, radial_accuracy: int = 7500) -> None:
        """copy a steps_b200.api.Globals into the reference's globals and run its own setup"""
        c = SrefConfig(g.COSMOLOGY, g.COMOVING_INTEGRATION, g.IS_PERIODIC, g.RADIAL_FORCE_TABLE_SIZE, radial_accuracy, 0, g.L, g.Rsim,
                       g.H0, g.Omega_m, g.Omega_lambda, g.Omega_r, g.Omega_b, g.ParticleRadi, g.ACC_PARAM, g.h_min, g.h_max, g.a_start)
        assert self.lib.sref_configure(C.byref(c)) == 0
        M = np.ascontiguousarray(g.M, dtype=self.REAL)
        assert self.lib.sref_set_particles(g.N, M.ctypes.data_as(C.c_void_p)) == 0
        self.N = g.N

    def softening(self) -> np.ndarray:
        s = np.empty(self.N, dtype=self.REAL)
        self.lib.sref_get_softening(s.ctypes.data_as(C.c_void_p))
        return s

    def scalars(self) -> dict:
        out = (C.c_double * 4)()
        self.lib.sref_get_scalars(out)
        return {"M_min": out[0], "rho_part": out[1], "mass_in_unit_sphere": out[2], "DE": out[3]}

    def force_softening(self, r: float, beta: float) -> float:
        return float(self.lib.sref_force_softening(r, beta))

    def build_tables(self) -> None:
        assert self.lib.sref_build_tables() == 0

    def table(self, which: int):
        dims = (C.c_int * 2)()
        ptr = self.lib.sref_table(which, dims)
        if not ptr:
            return None, (0, 0)
        if which == 1:
            n = dims[0]
        elif self.topology == 1:
            n = dims[0] ** 3 * 3
        else:
            n = dims[0] * dims[1] * 2
        arr = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(self.creal)), shape=(n,)).copy()
        return arr, (dims[0], dims[1])

    def export_tables(self, g) -> None:
        """hand the reference-built tables to a Globals so the engine uses value-identical inputs"""
        if self.topology == 1:
            t, d = self.table(0)
            g.T3_EWALD_FORCE_TABLE, g.N_EWALD_FORCE_GRID = t, d[0]
        if self.topology in (2, 3):
            t, d = self.table(1)
            g.RADIAL_FORCE_TABLE = t
        if self.topology == 2:
            t, d = self.table(0)
            g.S1R2_EWALD_FORCE_TABLE, g.Nrho_EWALD_FORCE_GRID, g.Nz_EWALD_FORCE_GRID = t, d[0], d[1]

    def forces(self, x: np.ndarray, id_min: int, id_max: int, nthreads: int = 0) -> np.ndarray:
        x = np.ascontiguousarray(x, dtype=self.REAL)
        F = np.zeros(3 * (id_max - id_min + 1), dtype=self.REAL)
        rc = self.lib.sref_forces(x.ctypes.data_as(C.c_void_p), F.ctypes.data_as(C.c_void_p), id_min, id_max, nthreads)
        if rc != 0:
            raise RuntimeError("reference ForceError flag set (forces_cuda.cu:970-974 convention)")
        return F

    def set_n_gpu(self, n: int) -> None:
        self.lib.sref_set_n_gpu(n)

    def kdk_begin(self, x, v, nthreads: int = 0) -> float:
        x = np.ascontiguousarray(x, dtype=self.REAL)
        v = np.ascontiguousarray(v, dtype=self.REAL)
        return float(self.lib.sref_kdk_begin(x.ctypes.data_as(C.c_void_p), v.ctypes.data_as(C.c_void_p), nthreads))

    def kdk_step(self, h: float):
        out = (C.c_double * 4)()
        hn = float(self.lib.sref_kdk_step(h, out))
        return hn, {"errmax": out[0], "a": out[1], "H": out[2], "T": out[3]}

    def kdk_state(self):
        n3 = 3 * self.N
        x, v, F = (np.empty(n3, dtype=self.REAL) for _ in range(3))
        self.lib.sref_kdk_state(x.ctypes.data_as(C.c_void_p), v.ctypes.data_as(C.c_void_p), F.ctypes.data_as(C.c_void_p))
        return x, v, F

    def write_redshift_cone(self, out_dir: str, x, v, limits, zlist, z_index: int, delta_z_index: int = 0, all_: int = 0,
                            h0_independent_units: int = 0, reset: bool = False, t_next: float = 0.0) -> str:
        """the reference's own write_redshift_cone (inputoutput.cc:314-405, ASCII branch): appends to <out_dir>redshift_cone.dat and keeps
        its IN_CONE flags between calls unless reset; returns the path of the file"""
        x = np.ascontiguousarray(x, dtype=self.REAL)
        v = np.ascontiguousarray(v, dtype=self.REAL)
        lim = np.ascontiguousarray(limits, dtype=np.float64)
        zl = np.ascontiguousarray(zlist, dtype=np.float64)
        if not out_dir.endswith("/"):
            out_dir += "/"
        self.lib.sref_write_redshift_cone.argtypes = [C.c_char_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                                      C.c_int, C.c_int, C.c_int, C.c_double]
        self.lib.sref_write_redshift_cone(out_dir.encode(), x.ctypes.data_as(C.c_void_p), v.ctypes.data_as(C.c_void_p), lim.ctypes.data_as(C.c_void_p),
                                          lim.size, zl.ctypes.data_as(C.c_void_p), zl.size, z_index, delta_z_index, all_, h0_independent_units,
                                          1 if reset else 0, t_next)
        return out_dir + "redshift_cone.dat"

    def write_ascii_snapshot(self, out_dir: str, x, v, a: float, t_next: float, h0_independent_units: int = 0) -> str:
        """the reference's own write_ascii_snapshot (inputoutput.cc:826-909); returns the path of the file it wrote"""
        if not out_dir.endswith(os.sep):
            out_dir += os.sep
        x = np.ascontiguousarray(x, dtype=self.REAL)
        v = np.ascontiguousarray(v, dtype=self.REAL)
        self.lib.sref_write_ascii_snapshot.argtypes = [C.c_char_p, C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_int]
        self.lib.sref_write_ascii_snapshot(out_dir.encode(), x.ctypes.data_as(C.c_void_p), v.ctypes.data_as(C.c_void_p), a, t_next, h0_independent_units)
        unit_t = 47.14829951063323  # global_variables.h:16
        return os.path.join(out_dir, "t%d.dat" % int(round(100 * t_next * unit_t)))

    @property
    def is_glass(self) -> bool:
        return bool(self.lib.sref_is_glass())

    def set_out_dir(self, path: str) -> None:
        """directory (with trailing separator) the reference writes its log files to (OUT_DIR)"""
        if not path.endswith(os.sep):
            path += os.sep
        self.lib.sref_set_out_dir(path.encode())
        self._out_dir = path

    def glass_log(self) -> np.ndarray:
        """rows of <OUT_DIR>Glass_logfile.dat written by the GLASS_MAKING step() (Log_write_glass, inputoutput.cc:974-1060):
        columns 6..13 = Mean(F) Max(F) Mean(A) Max(A) Mean(disp) Max(disp) Mean(v)[km/s] Max(v)[km/s], printed with %.15f"""
        rows = [ln.split() for ln in open(os.path.join(self._out_dir, "Glass_logfile.dat")) if ln.strip() and not ln.startswith("#")]
        return np.array([[float(v) for v in r] for r in rows])

    def friedmann_step(self, a0: float, h: float) -> float:
        return float(self.lib.sref_friedmann_step(a0, h))

    def hubble(self, a: float) -> float:
        return float(self.lib.sref_hubble(a))
