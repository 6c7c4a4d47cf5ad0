"""development aid: how many distinct 128-byte lines (L1 wavefronts) and 32-byte sectors one warp-wide table load of the T^3 tricubic
gather touches, for the two lane mappings: (a) one-sided kernel -- 32 consecutive i-particles against ONE j; (b) action-reaction kernel --
lane l holds i0+l and meets j0+((l+s) mod 32) (systolic visit).  Input: the C3 load (128^3 lattice, sigma = 0.1 spacing, L = 100, 63^3 table;
smaller lattice by argument) or a randomly ordered copy of it.  The model counts, per (row a,b; point c; component k) load instruction,
the lines the 32 lane addresses fall into; DESIGN.md section 3.4 compares it with ncu's l1tex wavefront counters.
usage: t3_wavefront_model.py [n_side] [samples]"""
import sys

import numpy as np


def cells(d, L, N):
    h = L / N
    u = (d + 0.5 * L) / h - 0.5
    return (np.floor(u).astype(np.int64)) % N


def count(xi, xj, L, N):
    """xi, xj: [32, 3] lane positions; returns (mean lines, mean sectors) per load instruction over the 64 x 3 loads of a step"""
    d = xj - xi
    d -= L * np.round(d / L)  # nearest image
    c = np.stack([cells(d[:, k], L, N) for k in range(3)], axis=1)  # [32, 3] base cells
    lines, sectors, n = 0, 0, 0
    for a in range(4):
        for b in range(4):
            for cc in range(4):
                ix, iy, iz = (c[:, 0] - 1 + a) % N, (c[:, 1] - 1 + b) % N, (c[:, 2] - 1 + cc) % N
                base = ((ix * N + iy) * N + iz) * 24  # byte address of the cell's 3 doubles
                for k in range(3):
                    addr = base + 8 * k
                    lines += np.unique(addr // 128).size
                    sectors += np.unique(addr // 32).size
                    n += 1
    return lines / n, sectors / n


def main():
    ns = int(sys.argv[1]) if len(sys.argv) > 1 else 128
    samples = int(sys.argv[2]) if len(sys.argv) > 2 else 200
    L, N = 100.0, 63
    rng = np.random.default_rng(20243)
    g = np.arange(ns)
    pos = np.stack(np.meshgrid(g, g, g, indexing="ij"), axis=-1).reshape(-1, 3) * (L / ns)  # z fastest, as ic.t3_lattice
    pos = np.mod(pos + rng.normal(size=pos.shape) * 0.1 * L / ns, L)
    n = pos.shape[0]
    for order in ("lattice order", "random order"):
        p = pos if order == "lattice order" else pos[rng.permutation(n)]
        res = {"one-sided": [], "action-reaction": []}
        for _ in range(samples):
            i0 = 32 * rng.integers(0, n // 32)
            j0 = 32 * rng.integers(0, n // 32)
            s = rng.integers(0, 32)
            xi = p[i0:i0 + 32]
            res["one-sided"].append(count(xi, np.repeat(p[j0 + s][None, :], 32, axis=0), L, N))
            res["action-reaction"].append(count(xi, p[j0 + (np.arange(32) + s) % 32], L, N))
        for k, v in res.items():
            v = np.array(v)
            print(f"{order:14s} {k:16s}: {v[:, 0].mean():5.2f} lines, {v[:, 1].mean():5.2f} sectors per warp-wide load "
                  f"({192 * v[:, 0].mean() / 32:5.1f} wavefronts per pair)")


if __name__ == "__main__":
    main()
