"""development aid: throughput of the pair kernels of the periodic topologies (and C1/C5 shapes) on one GPU.
Input tables (T^3 Ewald, S^1xR^2 radial/Ewald) are built by the reference's own builders through oracle/_ref --
this is a measuring tool, not product code.
usage: topo_bench.py case[,case...]   cases: t3:<n_side>[:random|:cellsort]  s1r2nl:<N>  s1r2:<N>  r3f32:<N>  c1"""
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


def tables(g, radial_accuracy=7500):
    key = (g.topology, 8 if g.REAL == np.float64 else 4)
    v = pyref.VARIANT[key]
    if not pyref.available(v):
        raise SystemExit(f"oracle/_ref variant {v} not built")
    r = pyref.Reference(v)
    r.configure(g, radial_accuracy)
    t0 = time.perf_counter()
    r.build_tables()
    r.export_tables(g)
    g.mass_in_unit_sphere = r.scalars()["mass_in_unit_sphere"]
    return time.perf_counter() - t0


def run(case):
    kind, _, arg = case.partition(":")
    tb = 0.0
    if kind == "t3":
        # t3:<n_side>[:random|:cellsort]  particle order: as generated (lattice, z fastest), shuffled, or shuffled then sorted by table cell
        ns, _, order = arg.partition(":")
        ns = int(ns)
        c = ic.t3_lattice(ns, 20243, L=100.0, is_periodic=2, name=f"T^3 {ns}^3 {order or 'lattice'} order")
        if order in ("random", "cellsort"):
            rng = np.random.default_rng(7)
            perm = rng.permutation(c.g.N)
            if order == "cellsort":
                ng = 63  # the table grid of IS_PERIODIC = 2
                cell = np.floor(c.x.reshape(-1, 3)[perm] / (100.0 / ng)).astype(np.int64) % ng
                perm = perm[np.argsort((cell[:, 0] * ng + cell[:, 1]) * ng + cell[:, 2], kind="stable")]
            c.x[:] = c.x.reshape(-1, 3)[perm].reshape(-1)
            c.v[:] = c.v.reshape(-1, 3)[perm].reshape(-1)
            c.g.M = np.ascontiguousarray(c.g.M[perm])
            c.g.SOFT_LENGTH = np.ascontiguousarray(c.g.SOFT_LENGTH[perm])
        elif order:
            raise SystemExit(f"unknown particle order {order}")
        tb = tables(c.g)
        evals = 1
    elif kind == "s1r2nl":
        n = int(arg)
        c = ic.s1r2_cylinder(n, 224, max(1, int(0.8 * n / 200)), 20244, lookup=False, is_periodic=2, name=f"S1xR2 NOLOOKUP N={n}")
        tb = tables(c.g)
        evals = 7
    elif kind == "s1r2":
        n = int(arg)
        c = ic.s1r2_cylinder(n, 224, max(1, int(0.8 * n / 200)), 20244, lookup=True, is_periodic=2, name=f"S1xR2 lookup N={n}")
        tb = tables(c.g)
        evals = 1
    elif kind == "r3f32":
        n = int(arg)
        c = ic.compactified_r3(n, 224, max(1, int(0.854 * n / 122)), 20245, np.float32, name=f"R^3 FP32 N={n}")
        evals = 1
    elif kind == "c1":
        c = ic.config_c1()
        evals = 1
    else:
        raise SystemExit(f"unknown case {case}")
    g = c.g
    eng = sb.Engine(g, 0)
    eng.upload(c.x, c.v)
    ms = []
    for _ in range(3):
        eng.forces()
        eng.sync()
        ms.append(eng.pair_kernel_ms())
    F = eng.download_forces(0, min(g.N, 8) - 1)
    best = min(ms[1:])
    print(json.dumps({"case": case, "sym": bool(eng.symmetric), "name": c.name, "N": int(g.N), "pair_kernel_ms": best, "force_ms": eng.timings()[0], "pairs_per_s": g.N * float(g.N) / (best * 1e-3),
                      "image_evals_per_pair": evals, "shape": eng.launch_shape(0, g.N - 1), "table_build_s": round(tb, 2), "F0": [float(v) for v in F[:3]]}), flush=True)
    eng.close()


if __name__ == "__main__":
    for cs in sys.argv[1].split(","):
        run(cs)
