"""shared test helpers: golden-fixture loading and the parity statistics of SURVEY.md H2"""
import os

import numpy as np

from steps_b200.api import Globals

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    """-> (Globals, dict of arrays).  T^3 Ewald tables are not stored (6 MB): see t3_table()."""
    d = dict(np.load(os.path.join(GOLDEN, name + ".npz")))
    REAL = np.float64 if int(d["real_bytes"]) == 8 else np.float32
    g = Globals(topology=int(d["topology"]), REAL=REAL, N=int(d["N"]))
    for k in ("COSMOLOGY", "COMOVING_INTEGRATION", "IS_PERIODIC", "EWALD_INTERPOLATION_ORDER", "RADIAL_FORCE_TABLE_SIZE",
              "N_EWALD_FORCE_GRID", "Nrho_EWALD_FORCE_GRID", "Nz_EWALD_FORCE_GRID"):
        setattr(g, k, int(d[k]))
    for k in ("L", "Rsim", "H0", "Omega_m", "Omega_lambda", "Omega_r", "Omega_b", "ParticleRadi", "ACC_PARAM", "h_min", "h_max",
              "a_start", "mass_in_unit_sphere", "M_min", "rho_part"):
        setattr(g, k, float(d[k]))
    g.M = np.ascontiguousarray(d["M"], dtype=REAL)
    g.SOFT_LENGTH = np.ascontiguousarray(d["SOFT_LENGTH"], dtype=REAL)
    if "RADIAL_FORCE_TABLE" in d:
        g.RADIAL_FORCE_TABLE = np.ascontiguousarray(d["RADIAL_FORCE_TABLE"], dtype=REAL)
    if "S1R2_EWALD_FORCE_TABLE" in d:
        g.S1R2_EWALD_FORCE_TABLE = np.ascontiguousarray(d["S1R2_EWALD_FORCE_TABLE"], dtype=REAL)
    return g, d


def needs_t3_table(g):
    return g.topology == 1 and g.IS_PERIODIC >= 2


def attach_t3_table(g, d):
    """rebuild the T^3 Ewald table with the reference's own builder (oracle/_ref travels to the GPU box)
    and check it against the checksum stored in the fixture.  Returns False if _ref is unavailable."""
    from oracle import pyref

    variant = pyref.VARIANT[(1, 8 if g.REAL == np.float64 else 4)]
    if not pyref.available(variant):
        return False
    r = pyref.Reference(variant)
    r.configure(g)
    r.build_tables()
    r.export_tables(g)
    t = g.T3_EWALD_FORCE_TABLE
    tol = 1e-9 if g.REAL == np.float64 else 1e-3
    assert abs(t.sum(dtype=np.float64) - float(d["T3_table_sum"])) <= tol * float(d["T3_table_abs_sum"])
    probe = t[:: max(1, t.size // 997)][:997]
    assert np.allclose(probe, d["T3_table_probe"], rtol=1e-10 if g.REAL == np.float64 else 1e-4, atol=1e-12)
    return True


def rel_err(F, Fref):
    """per-particle |dF_i| / |F_i|"""
    a, b = np.asarray(F, dtype=np.float64).reshape(-1, 3), np.asarray(Fref, dtype=np.float64).reshape(-1, 3)
    return np.linalg.norm(a - b, axis=1) / np.maximum(np.linalg.norm(b, axis=1), 1e-300)


def noise_err(F, Fref, S):
    """per-particle |dF_i| / sum_j |f_ij| -- the scale of the reference's own rounding noise"""
    a, b = np.asarray(F, dtype=np.float64).reshape(-1, 3), np.asarray(Fref, dtype=np.float64).reshape(-1, 3)
    return np.linalg.norm(a - b, axis=1) / S
