"""debug aid (>= 2 GPUs): the resident group on k devices for random_sphere(n): forces, init_errmax, one KDK step; prints finiteness and errmax"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

import steps_b200 as sb  # noqa: E402
from steps_b200 import _lib, ic  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 6001
k = int(sys.argv[2]) if len(sys.argv) > 2 else 2
c = ic.random_sphere(n, 19)
g = c.g
lib = _lib.load()
grp = C.c_void_p()
p = g.cparams()
_lib.check(lib.steps_b200_group_create(C.byref(grp), C.byref(p), 8, k, 0))
_lib.check(lib.steps_b200_group_upload(grp, c.x.ctypes.data, c.v.ctypes.data, g.M.ctypes.data, g.SOFT_LENGTH.ctypes.data, None))
_lib.check(lib.steps_b200_group_forces(grp))
x, v, F = (np.empty(3 * n) for _ in range(3))
_lib.check(lib.steps_b200_group_download(grp, x.ctypes.data, v.ctypes.data, F.ctypes.data))
for d in range(k):
    e = C.c_void_p(lib.steps_b200_group_engine(grp, d))
    lo, hi = C.c_int(), C.c_int()
    lib.steps_b200_engine_range(e, C.byref(lo), C.byref(hi))
    print("engine", d, "rows", lo.value, hi.value, "sym", lib.steps_b200_engine_is_symmetric(e), "F finite", np.isfinite(F[3 * lo.value:3 * hi.value]).all(),
          "max|F|", np.abs(F[3 * lo.value:3 * hi.value]).max(), "v finite", np.isfinite(v[3 * lo.value:3 * hi.value]).all(), "max|v|", np.abs(v[3 * lo.value:3 * hi.value]).max())
em = C.c_double()
a0 = g.a_start
H0 = sb.CALCULATE_Hubble_param(g, a0)
_lib.check(lib.steps_b200_group_init_errmax(grp, a0, H0, C.byref(em)))
print("errmax", em.value)
eng = sb.Engine(g, 0)
eng.upload(c.x, c.v)
eng.forces()
F1 = eng.download_forces(0, n - 1)
print("single h", eng.calculate_init_h(), "group h", (2 * g.ACC_PARAM / em.value) ** 0.5, "max rel dF", np.abs(F - F1).max() / np.abs(F1).max())
lib.steps_b200_group_destroy(grp)
