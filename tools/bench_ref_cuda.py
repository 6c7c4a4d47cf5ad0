"""development aid / evidence: the reference's OWN CUDA force path (forces_cuda.cu compiled unmodified for sm_100a,
oracle/_ref/libsteps_refcuda_<variant>.so) timed on the same GPU as ours, on the same input, and compared with ours row by row.
It is the kernel this repository is here to beat ("also time the reference forces_cuda.cu built for sm_100a").
The reference's call copies x, M, SOFT_LENGTH to the device and F back on every call (forces_cuda.cu:878-1117), so its time is an
end-to-end time; ours is measured through the stateless host-buffer C-ABI call for the same rows (same H2D/D2H inside).
usage: bench_ref_cuda.py [case] [rows]    case: c2 (default) | c2:<N> | t3:<n_side> | s1r2nl:<N>;  rows = i-rows timed (default 65536)"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

import steps_b200 as sb  # noqa: E402
from oracle import pyref  # noqa: E402
from steps_b200 import ic  # noqa: E402


def main():
    case = sys.argv[1] if len(sys.argv) > 1 else "c2"
    rows = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
    kind, _, arg = case.partition(":")
    if kind == "c2":
        n = int(arg) if arg else 2_000_000
        c = ic.config_c2() if n == 2_000_000 else ic.compactified_r3(n, 224, max(1, int(0.854 * n / 122)), 20242)
    elif kind == "t3":
        c = ic.t3_lattice(int(arg), 20243, L=100.0, is_periodic=2)
    elif kind == "s1r2nl":
        n = int(arg)
        c = ic.s1r2_cylinder(n, 224, max(1, int(0.8 * n / 200)), 20244, lookup=False, is_periodic=2)
    else:
        raise SystemExit(f"unknown case {case}")
    g = c.g
    variant = pyref.VARIANT[(g.topology, 8 if g.REAL == np.float64 else 4)]
    if not pyref.available(variant, cuda=True):
        raise SystemExit(f"oracle/_ref/libsteps_refcuda_{variant}.so not built (make -C oracle refcuda)")
    r = pyref.Reference(variant, cuda=True)
    r.configure(g)
    r.set_n_gpu(1)
    if g.topology != 0:
        r.build_tables()
        r.export_tables(g)
    g.mass_in_unit_sphere = r.scalars()["mass_in_unit_sphere"]
    rows = min(rows, g.N)
    lo = (g.N - rows) // 2
    hi = lo + rows - 1
    r.forces(c.x, lo, min(lo + 255, hi), 0)  # warm-up: context, first allocation
    t0 = time.perf_counter()
    F_ref = r.forces(c.x, lo, hi, 0)
    t_ref = time.perf_counter() - t0
    F = np.empty(3 * rows, dtype=g.REAL)
    sb.force_entry(g)(g, c.x, F, lo, min(lo + 255, hi))  # warm-up
    t0 = time.perf_counter()
    sb.force_entry(g)(g, c.x, F, lo, hi)
    t_ours = time.perf_counter() - t0
    a, b = F.astype(np.float64).reshape(-1, 3), np.asarray(F_ref, dtype=np.float64).reshape(-1, 3)
    e = np.linalg.norm(a - b, axis=1) / np.maximum(np.linalg.norm(b, axis=1), 1e-300)
    pairs = float(rows) * g.N
    print(json.dumps({"case": case, "N": int(g.N), "rows": rows, "reference_cuda_s": t_ref, "reference_cuda_pairs_per_s": pairs / t_ref,
                      "ours_subrange_call_s": t_ours, "ours_pairs_per_s": pairs / t_ours, "speedup": t_ref / t_ours,
                      "ours_vs_reference_cuda_rel_p99": float(np.percentile(e, 99)), "ours_vs_reference_cuda_rel_max": float(e.max()),
                      "note": "ours = one-sided kernel (a sub-range call cannot use the action-reaction path); whole-range throughput is in bench.py"}))


if __name__ == "__main__":
    main()
