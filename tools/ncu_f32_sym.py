"""development aid: two force evaluations of the FP32 action-reaction kernel at N (default 400k, one pass each) -- the
target of `ncu --set full -k regex:force_r3_f32_sym -s 1 -c 1`."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

import steps_b200 as sb  # noqa: E402
from steps_b200 import ic  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 400_000
c = ic.compactified_r3(n, 224, max(1, int(0.854 * n / 122)), 20245, np.float32)
eng = sb.Engine(c.g, 0)
eng.upload(c.x, c.v)
for _ in range(2):
    eng.forces()
    eng.sync()
print("symmetric", eng.symmetric, "pair kernel ms", eng.pair_kernel_ms(), "shape", eng.launch_shape(0, n - 1))
eng.close()
