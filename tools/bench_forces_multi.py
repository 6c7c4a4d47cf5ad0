"""evidence: the stateless multi-GPU drop-in call steps_b200_forces_multi_f64(params, x, M, soft, F, 0, N-1, n_gpu, 0) -- what the
forces() shim issues when the reference's n_GPU > 1 -- at C2 (or --n), host buffers in and out, wall clock per call.
Since round 2 a whole-range call runs the action-reaction kernel on every device (cached NCCL group) instead of n_gpu one-sided
sub-range launches; STEPS_B200_MULTI_ONESIDED=1 gives the old behaviour for comparison.
usage: bench_forces_multi.py <n_gpu> [N]"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

import steps_b200 as sb  # noqa: E402
from steps_b200 import ic  # noqa: E402


def main():
    n_gpu = int(sys.argv[1])
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 2_000_000
    c = ic.config_c2() if n == 2_000_000 else ic.compactified_r3(n, 224, max(1, int(0.854 * n / 122)), 20242)
    g = c.g
    g.n_GPU = n_gpu
    F = np.empty(3 * g.N)
    sb.forces(g, c.x, F, 0, g.N - 1)  # first call: engines, communicator, buffers
    ts = []
    for _ in range(3):
        t0 = time.perf_counter()
        sb.forces(g, c.x, F, 0, g.N - 1)
        ts.append(time.perf_counter() - t0)
    t = min(ts)
    peak = sb.fma_peak_sustained(0, 8, 1.0)
    pairs = float(g.N) * g.N
    # the same rows from one device as the yardstick of correctness
    g1 = c.g
    g1.n_GPU = 1
    lo, hi = g.N // 2, g.N // 2 + 4095
    F1 = np.empty(3 * 4096)
    sb.forces(g1, c.x, F1, lo, hi)
    d = np.abs(F1 - F[3 * lo:3 * (hi + 1)]).max() / np.abs(F1).max()
    print(json.dumps({"call": "steps_b200_forces_multi_f64 (whole range)", "one_sided_split": bool(os.environ.get("STEPS_B200_MULTI_ONESIDED")), "n_gpu": n_gpu,
                      "N": int(g.N), "seconds_per_call": t, "pairs_per_s": pairs / t, "tflops_20flop": 20 * pairs / t / 1e12,
                      "frac_of_fp64_peak_all_gpus": 20 * pairs / t / 1e12 / (peak * n_gpu), "peak_tflops_one_gpu": peak,
                      "max_rel_diff_vs_one_gpu_rows": float(d)}))


if __name__ == "__main__":
    main()
