"""CPU tier: the per-pair functions of the table-lookup topologies (steps_b200/csrc/pair_generic.cuh, pair_generic_sym.cuh)
executed on the HOST through tests/hostcheck/pair_host_check.cu, with the reference's own tables (oracle/_ref).  Checks the
three numerical claims the action-reaction kernel of these topologies rests on, pair by pair on 200 000 random pairs (uniform
in the box, one in seven closer than a softening length, one in eleven exactly at the |d| = L/2 wrap boundary):
  (1) factoring the mass out of the reference's pair term changes nothing,
  (2) t(-d) = -t(d) to the rounding of the table cell coordinate, far inside the 1e-12 tolerance of the path,
  (3) the lean T^3 arithmetic agrees with the reference's operation-by-operation form equally well.
The functions are __host__ __device__; the product only calls them from CUDA kernels."""
import os
import shutil
import subprocess

import numpy as np
import pytest

from oracle import pyref
from steps_b200 import ic

HERE = os.path.dirname(os.path.abspath(__file__))
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
N_PAIRS = 200_000


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    if not (os.path.exists(NVCC) or shutil.which("nvcc")):
        pytest.skip("nvcc needed to compile the host harness")
    out = str(tmp_path_factory.mktemp("hostcheck") / "pair_host_check")
    nvcc = NVCC if os.path.exists(NVCC) else shutil.which("nvcc")
    subprocess.run([nvcc, "-O2", "-std=c++17", "-o", out, os.path.join(HERE, "hostcheck", "pair_host_check.cu")], check=True, capture_output=True)
    return out


def _run(harness, args):
    r = subprocess.run([harness, *[repr(a) if isinstance(a, float) else str(a) for a in args]], check=True, capture_output=True, text=True)
    return [float(v) for v in r.stdout.split()]


@pytest.fixture(scope="module")
def t3_table(tmp_path_factory):
    if not pyref.available("t3_f64"):
        pytest.skip("reference table builder needs oracle/_ref")
    g = ic.t3_lattice(4, 5, L=30.0, is_periodic=2).g
    r = pyref.Reference("t3_f64")
    r.configure(g, 400)
    r.build_tables()
    r.export_tables(g)
    path = str(tmp_path_factory.mktemp("tables") / "t3.bin")
    np.asarray(g.T3_EWALD_FORCE_TABLE, dtype=np.float64).tofile(path)
    return path, g.N_EWALD_FORCE_GRID, 30.0


@pytest.mark.parametrize("soft", [0.05, 1.0])
def test_t3_pair_numerics(harness, t3_table, soft):
    path, n, L = t3_table
    e_factor, e_antisym, e_lean = _run(harness, ["t3", path, n, L, 2, N_PAIRS, 7, soft])
    print(f"T^3 soft={soft}: mass factoring {e_factor:.2e}, antisymmetry {e_antisym:.2e}, lean arithmetic {e_lean:.2e} (max per-pair relative)")
    assert e_factor == 0.0
    assert e_antisym < 5e-13
    assert e_lean < 5e-13


def test_t3_nearest_image_only_pair_numerics(harness, t3_table):
    path, n, L = t3_table
    e_factor, e_antisym, _ = _run(harness, ["t3", path, n, L, 1, N_PAIRS, 9, 0.05])
    assert e_factor < 1e-15 and e_antisym == 0.0


def test_s1r2_lookup_pair_numerics(harness, tmp_path):
    if not pyref.available("s1r2_f64"):
        pytest.skip("reference table builder needs oracle/_ref")
    g = ic.s1r2_cylinder(3000, 24, 80, 63, lookup=True, is_periodic=2, L=20.0, r_sim=30.0, d_s=8.0, r_crit=10.0).g
    r = pyref.Reference("s1r2_f64")
    r.configure(g, 400)
    r.build_tables()
    r.export_tables(g)
    path = str(tmp_path / "s1r2.bin")
    np.asarray(g.S1R2_EWALD_FORCE_TABLE, dtype=np.float64).tofile(path)
    for order in (0, 2, 4):
        e_factor, e_antisym, _ = _run(harness, ["s1r2", path, g.Nrho_EWALD_FORCE_GRID, g.Nz_EWALD_FORCE_GRID, 2.25 * g.Rsim, g.L, order, N_PAIRS, 11, 0.03])
        print(f"S^1xR^2 lookup order {order}: mass factoring {e_factor:.2e}, antisymmetry {e_antisym:.2e}")
        assert e_factor == 0.0
        assert e_antisym < 5e-13


def _state_file(tmp_path, c):
    g = c.g
    path = str(tmp_path / "state.bin")
    np.concatenate([np.asarray(c.x, dtype=np.float64), np.asarray(g.M, dtype=np.float64), np.asarray(g.SOFT_LENGTH, dtype=np.float64)]).tofile(path)
    return path


@pytest.mark.parametrize("lean", [0, 1])
def test_t3_forces_formed_pair_symmetrically_match_the_reference(harness, tmp_path, lean):
    """the whole force evaluation as the action-reaction kernel forms it (every unordered pair once, applied to both particles),
    on the host, against the reference's own forces_periodic(): the per-pair antisymmetry error does not accumulate past 1e-12"""
    if not pyref.available("t3_f64"):
        pytest.skip("reference needs oracle/_ref")
    c = ic.t3_lattice(10, 31, L=30.0, is_periodic=2)
    g = c.g
    x = c.x.reshape(-1, 3)
    x[7] = x[500] + 1e-3 * g.SOFT_LENGTH[7]  # a softened pair
    x[:] = np.mod(x, 30.0)
    r = pyref.Reference("t3_f64")
    r.configure(g, 400)
    r.build_tables()
    r.export_tables(g)
    Fo = r.forces(c.x, 0, g.N - 1, 0)
    tab = str(tmp_path / "t3.bin")
    np.asarray(g.T3_EWALD_FORCE_TABLE, dtype=np.float64).tofile(tab)
    out = str(tmp_path / "F.bin")
    subprocess.run([harness, "forces_t3", tab, str(g.N_EWALD_FORCE_GRID), "30.0", "2", _state_file(tmp_path, c), str(g.N), str(lean), out], check=True)
    F = np.fromfile(out, dtype=np.float64)
    a, b = F.reshape(-1, 3), np.asarray(Fo, dtype=np.float64).reshape(-1, 3)
    e = np.linalg.norm(a - b, axis=1) / np.linalg.norm(b, axis=1)
    print(f"T^3 N={g.N} pair-symmetric host evaluation (lean={lean}) vs reference: |dF|/|F| p99 {np.percentile(e, 99):.2e} max {e.max():.2e}")
    assert np.percentile(e, 99) < 1e-12 and e.max() < 5e-11


def test_s1r2_lookup_forces_formed_pair_symmetrically_match_the_reference(harness, tmp_path):
    if not pyref.available("s1r2_f64"):
        pytest.skip("reference needs oracle/_ref")
    c = ic.s1r2_cylinder(900, 24, 20, 63, lookup=True, is_periodic=2, L=20.0, r_sim=30.0, d_s=8.0, r_crit=10.0)
    g = c.g
    r = pyref.Reference("s1r2_f64")
    r.configure(g, 400)
    r.build_tables()
    r.export_tables(g)
    g.mass_in_unit_sphere = r.scalars()["mass_in_unit_sphere"]
    Fo = np.asarray(r.forces(c.x, 0, g.N - 1, 0), dtype=np.float64).reshape(-1, 3)
    tab = str(tmp_path / "s1r2.bin")
    np.asarray(g.S1R2_EWALD_FORCE_TABLE, dtype=np.float64).tofile(tab)
    out = str(tmp_path / "F.bin")
    subprocess.run([harness, "forces_s1r2", tab, str(g.Nrho_EWALD_FORCE_GRID), str(g.Nz_EWALD_FORCE_GRID), repr(2.25 * g.Rsim), repr(g.L), "4",
                    _state_file(tmp_path, c), str(g.N), out], check=True)
    F = np.fromfile(out, dtype=np.float64).reshape(-1, 3)
    # the reference adds a background term to x and y only (forces.cc:1385-1396; the harness forms the pair sum alone): compare z
    ez = np.abs(F[:, 2] - Fo[:, 2]) / np.linalg.norm(Fo, axis=1)
    print(f"S^1xR^2 lookup N={g.N} pair-symmetric host evaluation vs reference: |dF_z|/|F| p99 {np.percentile(ez, 99):.2e} max {ez.max():.2e}")
    assert np.percentile(ez, 99) < 1e-12 and ez.max() < 5e-11
