"""development aid: time every compiled shape of the tuned R^3 FP64 pair kernel (STEPS_B200_F64_VARIANT) on one GPU"""
import os, subprocess, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

def child(n):
    import numpy as np
    import steps_b200 as sb
    from steps_b200 import ic
    REAL = np.float32 if os.environ.get("SWEEP_REAL", "f64") == "f32" else np.float64
    c = ic.compactified_r3(n, 224, max(1, int(0.854 * n / 122)), 20242, REAL)
    eng = sb.Engine(c.g, 0)
    if os.environ.get("SWEEP_SYM"):
        eng.set_symmetric(True)
    eng.upload(c.x, c.v)
    ms = []
    for _ in range(4):
        eng.forces(); eng.sync(); ms.append(eng.pair_kernel_ms())
    F = eng.download_forces(0, 9)
    print(json.dumps({"sym": bool(eng.symmetric), "sym_variant": os.environ.get("STEPS_B200_SYM_VARIANT", "0"), "variant": int(os.environ.get("STEPS_B200_F32_VARIANT" if REAL == np.float32 else "STEPS_B200_F64_VARIANT", "0")), "real": REAL.__name__, "ms": min(ms[1:]), "pairs_per_s": n * float(n) / (min(ms[1:]) * 1e-3),
                      "shape": eng.launch_shape(0, n - 1), "F0": float(F[0])}))

if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "child":
        child(int(sys.argv[2]))
    else:
        n = int(sys.argv[1]) if len(sys.argv) > 1 else 400000
        variants = sys.argv[2].split(",") if len(sys.argv) > 2 else [str(i) for i in range(12)]
        for v in variants:
            env = dict(os.environ, STEPS_B200_F64_VARIANT=v, STEPS_B200_F32_VARIANT=v)
            r = subprocess.run([sys.executable, __file__, "child", str(n)], env=env, capture_output=True, text=True)
            print(r.stdout.strip().splitlines()[-1] if r.stdout.strip() else "FAILED " + r.stderr[-300:], flush=True)
