"""GPU (B200): the reference's OWN CUDA force path (forces_cuda.cu, compiled unmodified for sm_100a: oracle/Makefile target
refcuda) as a second oracle next to its CPU path: our forces against it on the same inputs, for the three topologies.
SURVEY.md 8c lists the semantic differences between the reference's two implementations (1/r^3 by pow vs division, ...): they agree
to rounding, so the tolerance is the one of the CPU oracle.

First run on a B200 at the start of round 2 (profiles/r2a_*.log): all green; part of the default `-m gpu` suite since."""
import os

import numpy as np
import pytest

import steps_b200 as sb
from helpers import rel_err
from oracle import pyref
from steps_b200 import ic

pytestmark = [pytest.mark.gpu]


def cases():
    yield "r3_f64", ic.compactified_r3(20000, 64, 250, 42, d_s=105.0), 1e-12
    yield "r3_f32", ic.compactified_r3(12000, 64, 150, 43, np.float32, d_s=105.0), 1e-4
    yield "t3_f64", ic.t3_lattice(12, 61, L=30.0, is_periodic=2), 1e-12
    yield "s1r2nl_f64", ic.s1r2_cylinder(3000, 24, 80, 62, lookup=False, is_periodic=2, L=20.0, r_sim=60.0, d_s=10.0, r_crit=15.0), 1e-12


@pytest.mark.parametrize("variant,c,tol", list(cases()), ids=lambda v: v if isinstance(v, str) else "")
def test_ours_against_the_reference_cuda_kernels(variant, c, tol):
    if not pyref.available(variant, cuda=True):
        pytest.skip("oracle/_ref reference CUDA build not present")
    g = c.g
    r = pyref.Reference(variant, cuda=True)
    r.configure(g, 400)
    r.set_n_gpu(1)
    if g.topology != 0:
        r.build_tables()
        r.export_tables(g)
    g.mass_in_unit_sphere = r.scalars()["mass_in_unit_sphere"]
    F_ref = r.forces(c.x, 0, g.N - 1, 0)
    F = np.empty(3 * g.N, dtype=g.REAL)
    sb.force_entry(g)(g, c.x, F, 0, g.N - 1)
    e = rel_err(F, F_ref)
    print(f"{variant}: ours vs the reference's CUDA kernel |dF|/|F| p99 {np.percentile(e, 99):.2e} max {e.max():.2e}")
    assert np.isfinite(F_ref).all()
    assert np.percentile(e, 99) < tol
