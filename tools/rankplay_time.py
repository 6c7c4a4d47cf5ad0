"""evidence: per-rank pair-phase time of a P-GPU action-reaction job measured on ONE GPU (the engine plays rank r: rows, rules and
launch plan of that rank; no collectives) -- what a P-GPU step costs per rank apart from the NCCL calls.
usage: rankplay_time.py <c2|c5> <P> <rank>[,rank...] [N]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

import steps_b200 as sb  # noqa: E402
from steps_b200 import _lib, ic  # noqa: E402

cfg, P = sys.argv[1], int(sys.argv[2])
ranks = [int(r) for r in sys.argv[3].split(",")]
n = int(sys.argv[4]) if len(sys.argv) > 4 else 0
if cfg == "c2":
    c = ic.config_c2() if not n else ic.compactified_r3(n, 224, max(1, int(0.854 * n / 122)), 20242)
else:
    c = ic.config_c5() if not n else ic.compactified_r3(n, 224, max(1, int(0.854 * n / 122)), 20245, np.float32)
g = c.g
lib = _lib.load()
for r in ranks:
    eng = sb.Engine(g, 0)
    _lib.check(lib.steps_b200_engine_debug_set_rank(eng._h, r, P, 1))
    eng.i_lo, eng.i_hi = eng.range()
    eng.upload(c.x, c.v)
    ms = []
    for _ in range(2):
        eng.forces()
        eng.sync()
        ms.append(eng.pair_kernel_ms())
    n_i = eng.i_hi - eng.i_lo
    print(json.dumps({"config": cfg, "N": int(g.N), "ranks": P, "rank": r, "rows": [eng.i_lo, eng.i_hi], "pair_phase_ms": min(ms),
                      "interactions_per_s_of_this_rank": n_i * float(g.N) / (min(ms) * 1e-3), "shape": eng.launch_shape(eng.i_lo, eng.i_hi - 1)}), flush=True)
    eng.close()
