"""quick GPU sanity script (development aid): parity of the R^3 FP64 kernel + a timing"""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import steps_b200 as sb
from steps_b200 import ic
from oracle import pyport

print("devices", sb._lib.load().steps_b200_device_count())
print("fp64 peak", sb.fma_peak(0, 8), "fp32 peak", sb.fma_peak(0, 4))
for n in (1000, 4096, 20000):
    c = ic.random_sphere(n, 11)
    g = c.g
    F = np.zeros(3 * n)
    sb.forces(g, c.x, F, 0, n - 1)
    Fo = pyport.forces(g, c.x, 0, n - 1)
    S = pyport.force_norms(g, c.x, 0, n - 1)
    d = np.linalg.norm((F - Fo).reshape(-1, 3), axis=1)
    print(n, "max dF/S", (d / S).max(), "max dF/|F|", (d / np.linalg.norm(Fo.reshape(-1, 3), axis=1)).max())
c = ic.config_c1()
g = c.g
eng = sb.Engine(g, 0)
eng.upload(c.x, c.v)
for _ in range(3):
    eng.forces(); eng.sync()
print("C1 force ms", eng.timings()[0], eng.launch_shape(0, g.N - 1), "pairs/s", g.N**2 / (eng.timings()[0] * 1e-3))
n = int(os.environ.get("NBIG", "262144"))
c = ic.compactified_r3(n, 224, 600, 5)
g = c.g
eng2 = sb.Engine(g, 0)
eng2.upload(c.x, c.v)
for _ in range(2):
    eng2.forces(); eng2.sync()
ms = eng2.timings()[0]
print(n, "force ms", ms, eng2.launch_shape(0, g.N - 1), "pairs/s %.4e" % (g.N**2 / (ms * 1e-3)), "TF(20) %.2f" % (20 * g.N**2 / (ms * 1e-3) / 1e12))
