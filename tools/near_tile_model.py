"""CPU model: which fraction of the (i-block, j-tile) units each rank of a P-GPU action-reaction job evaluates takes the CHECKED loop
(tile not provably outside the softening radius of the block: the kernel's far/near test, pair_r3_sym.cuh, applied here with the bounds
of whole i-blocks instead of warps -- a lower bound of 'far').  Explains why rank 0 of an 8-GPU C2 job is the slowest.
usage: near_tile_model.py [P] [N]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

from steps_b200 import api, ic  # noqa: E402

P = int(sys.argv[1]) if len(sys.argv) > 1 else 8
n = int(sys.argv[2]) if len(sys.argv) > 2 else 2_000_000
c = ic.config_c2() if n == 2_000_000 else ic.compactified_r3(n, 224, max(1, int(0.854 * n / 122)), 20242)
g = c.g
TJ, IB = 128, 768
x = c.x.reshape(-1, 3)
s = g.SOFT_LENGTH


def bounds(size):
    nt = (g.N + size - 1) // size
    lo, hi, rlo, rhi, sm = (np.empty((nt, 3)), np.empty((nt, 3)), np.empty(nt), np.empty(nt), np.empty(nt))
    r = np.linalg.norm(x, axis=1)
    for t in range(nt):
        sl = slice(t * size, min(g.N, (t + 1) * size))
        lo[t], hi[t] = x[sl].min(axis=0), x[sl].max(axis=0)
        rlo[t], rhi[t], sm[t] = r[sl].min(), r[sl].max(), s[sl].max()
    return lo, hi, rlo, rhi, sm


tlo, thi, trlo, trhi, tsm = bounds(TJ)
blo, bhi, brlo, brhi, bsm = bounds(IB)
for rank in range(P):
    i_lo, i_hi, rules = api.sym_rules(g.N, P, rank, IB)
    units = near = 0
    for b, ru in enumerate(rules):
        gb = i_lo // IB + b
        rngs = [(int(ru[0]), int(ru[1]))] + [(int(ru[3 + k]), int(ru[8 + k])) for k in range(int(ru[2]))]
        for lo_t, hi_t in rngs:
            if lo_t >= hi_t:
                continue
            gap = np.maximum(np.maximum(blo[gb] - thi[lo_t:hi_t], tlo[lo_t:hi_t] - bhi[gb]), 0.0)
            gap2 = (gap * gap).sum(axis=1)
            rg = np.maximum(brlo[gb] - trhi[lo_t:hi_t], trlo[lo_t:hi_t] - brhi[gb])
            bb = (bsm[gb] + tsm[lo_t:hi_t]) * 1.000001
            far = (gap2 > bb * bb) | (rg > bb)
            units += hi_t - lo_t
            near += int((~far).sum())
    print(f"rank {rank} of {P}: rows [{i_lo}, {i_hi})  units {units}  checked {near} = {100.0 * near / units:.2f} %", flush=True)
