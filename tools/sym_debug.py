"""development aid: where does the action-reaction path (pair_r3_sym.cuh) differ from the one-sided kernel?
Prints, per problem size, the error of the symmetric evaluation against the one-sided one by i-block, so that a wrong
i-side sum (own block), a wrong j-side sum (lower blocks) and a wrong diagonal can be told apart from one GPU run."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import steps_b200 as sb
from steps_b200 import ic


def run(c, symmetric):
    eng = sb.Engine(c.g, 0)
    eng.set_symmetric(symmetric)
    eng.upload(c.x, c.v)
    eng.forces()
    F = eng.download_forces(0, c.g.N - 1)
    used, shape = eng.symmetric, eng.launch_shape(0, c.g.N - 1)
    ms = eng.pair_kernel_ms()
    eng.close()
    return F.reshape(-1, 3), used, shape, ms


def main():
    ib = int(os.environ.get("SYM_IB", "768"))
    for n, kind in ((1537, "sphere"), (5000, "sphere"), (20000, "zoom"), (100000, "zoom")):
        c = ic.random_sphere(n, 7) if kind == "sphere" else ic.compactified_r3(n, 64, max(1, n // 80), 42, d_s=105.0)
        try:
            F1, _, _, ms1 = run(c, False)
            F2, used, shape, ms2 = run(c, True)
        except Exception as ex:  # noqa: BLE001
            print(json.dumps({"n": n, "error": str(ex)}), flush=True)
            return
        scale = np.linalg.norm(F1, axis=1).max()
        d = np.linalg.norm(F2 - F1, axis=1) / scale
        nb = (n + ib - 1) // ib
        per_block = [float(d[b * ib:(b + 1) * ib].max()) for b in range(nb)]
        worst = int(np.argmax(d))
        print(json.dumps({"n": n, "kind": kind, "sym_used": used, "shape": shape, "ms_one_sided": ms1, "ms_sym": ms2,
                          "max_err_over_maxF": float(d.max()), "worst_i": worst, "finite": bool(np.isfinite(F2).all()),
                          "per_block_max_err_first8": per_block[:8], "per_block_max_err_last4": per_block[-4:],
                          "F1_worst": F1[worst].tolist(), "F2_worst": F2[worst].tolist()}), flush=True)


if __name__ == "__main__":
    main()
