"""development aid: the FP32 R^3 pair kernels on one GPU at the C5 shape scaled to N particles -- the one-sided kernel and
every compiled shape of the action-reaction kernel (STEPS_B200_SYM_F32_VARIANT), with the accuracy of sampled rows
against FP64 truth from the oracle (checker only).
usage: sweep_f32_sym.py [N] [variant,variant,...]   ('one' = one-sided kernel)"""
import copy
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def child(n, which):
    import numpy as np

    import steps_b200 as sb
    from oracle import pyport
    from steps_b200 import ic

    c = ic.compactified_r3(n, 224, max(1, int(0.854 * n / 122)), 20245, np.float32)
    g = c.g
    eng = sb.Engine(g, 0)
    eng.set_symmetric(which != "one")
    eng.upload(c.x, c.v)
    ms = []
    for _ in range(3):
        eng.forces()
        eng.sync()
        ms.append(eng.pair_kernel_ms())
    used = bool(eng.symmetric)
    shape = eng.launch_shape(0, n - 1)
    g64 = copy.copy(g)
    g64.REAL = np.float64
    g64.M, g64.SOFT_LENGTH = g.M.astype(np.float64), g.SOFT_LENGTH.astype(np.float64)
    x64 = c.x.astype(np.float64)
    errs = []
    for lo in (0, n // 2, n - 256):
        F = eng.download_forces(lo, lo + 255).astype(np.float64).reshape(-1, 3)
        Ft = pyport.forces(g64, x64, lo, lo + 255).reshape(-1, 3)
        S = pyport.force_norms(g64, x64, lo, lo + 255)
        errs.append(float((np.linalg.norm(F - Ft, axis=1) / S).max()))
    eng.close()
    best = min(ms[1:])
    print(json.dumps({"kernel": which, "sym": used, "N": n, "ms": best, "pairs_per_s": n * float(n) / (best * 1e-3), "shape": shape,
                      "max_dF_over_sum_f_rows": errs}))


if __name__ == "__main__":
    if len(sys.argv) > 3 and sys.argv[1] == "child":
        child(int(sys.argv[2]), sys.argv[3])
    else:
        n = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
        variants = sys.argv[2].split(",") if len(sys.argv) > 2 else ["one", "0", "1", "2"]
        for v in variants:
            env = dict(os.environ)
            if v != "one":
                env["STEPS_B200_SYM_F32_VARIANT"] = v
            r = subprocess.run([sys.executable, __file__, "child", str(n), v], env=env, capture_output=True, text=True)
            print(r.stdout.strip().splitlines()[-1] if r.stdout.strip() else "FAILED " + r.stderr[-400:], flush=True)
