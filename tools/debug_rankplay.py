"""debug aid: one GPU plays every rank of a P-GPU action-reaction job (as tests/test_gpu_sym.py does) for a given n; prints per-rank
ranges, the launch shape, whether F is finite and the error against the single-engine result."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

import steps_b200 as sb  # noqa: E402
from steps_b200 import _lib, ic  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 6001
P = int(sys.argv[2]) if len(sys.argv) > 2 else 2
REAL = np.float32 if len(sys.argv) > 3 and sys.argv[3] == "f32" else np.float64
c = ic.random_sphere(n, 19, REAL)
g = c.g
lib = _lib.load()
e0 = sb.Engine(g, 0)
e0.upload(c.x, c.v)
e0.forces()
F1 = e0.download_forces(0, n - 1)
print("single engine: symmetric", e0.symmetric, "shape", e0.launch_shape(0, n - 1), "finite", np.isfinite(F1).all())
e0.close()
engines, fsyms, ranges = [], [], []
for r in range(P):
    eng = sb.Engine(g, 0)
    _lib.check(lib.steps_b200_engine_debug_set_rank(eng._h, r, P, 1))
    eng.i_lo, eng.i_hi = eng.range()
    ranges.append((eng.i_lo, eng.i_hi))
    eng.upload(c.x, c.v)
    eng.forces()
    n_pad = C.c_int()
    _lib.check(lib.steps_b200_engine_debug_fsym(eng._h, None, None, C.byref(n_pad)))
    f = np.empty(3 * n_pad.value, dtype=REAL)
    _lib.check(lib.steps_b200_engine_debug_fsym(eng._h, f.ctypes.data, None, None))
    print("rank", r, "rows", ranges[-1], "symmetric", eng.symmetric, "shape", eng.launch_shape(eng.i_lo, eng.i_hi - 1), "fsym finite", np.isfinite(f).all(),
          "max|fsym|", np.abs(f).max())
    engines.append(eng)
    fsyms.append(f)
total = np.sum(fsyms, axis=0).astype(REAL)
F = np.empty(3 * n, dtype=REAL)
for eng, (lo, hi) in zip(engines, ranges):
    _lib.check(lib.steps_b200_engine_debug_fsym(eng._h, None, total.ctypes.data, None))
    F[3 * lo: 3 * hi] = eng.download_forces(lo, hi - 1)
    print("rank rows", lo, hi, "finite", np.isfinite(F[3 * lo:3 * hi]).all())
    eng.close()
d = np.abs(F.astype(np.float64) - F1).max() / np.abs(F1).max()
bad = np.flatnonzero(~np.isfinite(F))
print("max rel diff vs single engine", d, "non-finite entries", bad.size, bad[:10] // 3)
