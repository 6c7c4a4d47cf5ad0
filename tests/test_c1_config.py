"""BASELINE.json configs[0]: compactified R^3, N = 32768, FP64 -- one force evaluation + 10 KDK steps, the case the reference
itself runs on the CPU.  tests/golden/c1_kdk_f64.npz holds the UNMODIFIED reference's outputs for exactly this run
(tools/make_golden_c1.py, generated in the authoring container with oracle/_ref), sampled on every 16th particle.

CPU tier: the seeded input generator still produces the inputs the fixture was made from, and the plain-C port reproduces the
reference's initial forces on the sampled rows.  GPU tier: the resident engine runs the whole case."""
import os

import numpy as np
import pytest

from helpers import GOLDEN, rel_err
from oracle import pyport
from steps_b200 import ic


@pytest.fixture(scope="module")
def c1():
    d = dict(np.load(os.path.join(GOLDEN, "c1_kdk_f64.npz")))
    c = ic.config_c1()
    c.g.mass_in_unit_sphere = float(d["mass_in_unit_sphere"])
    return c, d


def test_c1_inputs_are_the_ones_the_fixture_was_made_from(c1):
    c, d = c1
    g = c.g
    assert g.N == int(d["N"]) == 32768
    assert np.isclose(c.x.sum(), float(d["x_in_sum"]), rtol=0, atol=1e-9 * float(d["x_in_abs_sum"]))
    assert np.isclose(np.abs(c.x).sum(), float(d["x_in_abs_sum"]), rtol=1e-13)
    assert np.isclose(np.abs(c.v).sum(), float(d["v_in_abs_sum"]), rtol=1e-13)
    assert np.isclose(g.M.sum(), float(d["M_sum"]), rtol=1e-13)
    assert np.isclose(g.SOFT_LENGTH.sum(), float(d["soft_sum"]), rtol=1e-12)


def test_c1_port_initial_forces_on_sampled_rows(c1):
    c, d = c1
    g = c.g
    idx = d["idx"][::8]  # 256 rows x 32768 pairs
    F0 = d["F0"].reshape(-1, 3)[::8]
    for i, f in zip(idx, F0):
        Fi = pyport.forces(g, c.x, int(i), int(i))
        assert rel_err(Fi, f).max() < 1e-10


@pytest.mark.gpu
def test_c1_force_evaluation_and_ten_kdk_steps_on_the_gpu(c1):
    import steps_b200 as sb

    c, d = c1
    g = c.g
    idx = d["idx"]
    sel = (3 * idx[:, None] + np.arange(3)[None, :]).reshape(-1)
    eng = sb.Engine(g, 0)
    eng.upload(c.x, c.v)
    eng.forces()
    F0 = eng.download_forces(0, g.N - 1)
    e0 = rel_err(F0[sel], d["F0"])
    assert np.percentile(e0, 99) < 1e-12 and e0.max() < 1e-10
    assert np.isclose(np.abs(F0).sum(), float(d["F0_abs_sum"]), rtol=1e-12)
    h = eng.calculate_init_h()
    hs, errs = [h], []
    for _ in range(int(d["nsteps"])):
        errs.append(eng.step(h))
        h = eng.next_h()
        hs.append(h)
    x1, v1, F1 = eng.download()
    eng.close()
    dx = np.abs(x1[sel] - d["x1"]).max() / g.Rsim
    print(f"C1: h sequence max rel diff {np.abs(np.array(hs) / d['h_seq'] - 1).max():.2e}, errmax {np.abs(np.array(errs) / d['errmax_seq'] - 1).max():.2e}, "
          f"max |dx|/Rsim after 10 steps {dx:.2e}, |dF0|/|F| p99 {np.percentile(e0, 99):.2e}")
    assert np.allclose(hs, d["h_seq"], rtol=1e-9)
    assert np.allclose(errs, d["errmax_seq"], rtol=1e-8)
    assert dx < 1e-12 * int(d["nsteps"])
    assert np.isclose(np.abs(x1).sum(), float(d["x1_abs_sum"]), rtol=1e-12)
