"""CPU tier: the restated table-producer algorithms (steps_b200/csrc/ewald_t3.cuh, ewald_s1r2.cuh,
radial_table.cuh) executed on the HOST through a test harness and compared with the reference's own builders (oracle/_ref).
The functions are __host__ __device__; the product only launches them as CUDA kernels (GPU tier: tests/test_gpu_parity.py)."""
import os
import shutil
import subprocess

import numpy as np
import pytest

from oracle import pyref
from steps_b200 import api, ic

HERE = os.path.dirname(os.path.abspath(__file__))
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    if not (os.path.exists(NVCC) or shutil.which("nvcc")):
        pytest.skip("nvcc needed to compile the host harness")
    out = str(tmp_path_factory.mktemp("hostcheck") / "table_host_check")
    nvcc = NVCC if os.path.exists(NVCC) else shutil.which("nvcc")
    subprocess.run([nvcc, "-O2", "-std=c++17", "-Xcompiler", "-fopenmp", "-o", out, os.path.join(HERE, "hostcheck", "table_host_check.cu")],
                   check=True, capture_output=True)
    return out


def _run(harness, args, tmp_path):
    out = str(tmp_path / "t.bin")
    subprocess.run([harness, *[repr(a) if isinstance(a, float) else str(a) for a in args], out], check=True)
    return np.fromfile(out, dtype=np.float64)


def test_t3_ewald_table_host_path_matches_reference(harness, tmp_path):
    if not pyref.available("t3_f64"):
        pytest.skip("reference builder needs oracle/_ref")
    L = 30.0
    g = ic.t3_lattice(4, 5, L=L, is_periodic=2).g
    r = pyref.Reference("t3_f64")
    r.configure(g, 400)
    r.build_tables()
    r.export_tables(g)
    n = g.N_EWALD_FORCE_GRID
    ref = np.asarray(g.T3_EWALD_FORCE_TABLE, dtype=np.float64)
    d = api.t3_ewald_defaults(2, L)
    assert d["ngrid"] == n
    tab = _run(harness, ["t3", n, L, d["rel_cut"], d["rec_cut"]], tmp_path)
    h = L / n
    assert tab.shape == ref.shape and np.isfinite(tab).all()
    assert np.abs(tab - ref).max() <= 5e-14 / (h * h)
    c = n // 2
    assert np.all(tab.reshape(n, n, n, 3)[c, c, c] == 0.0)  # D(0) = 0: the centre point must not see a rounding residue as r


def test_s1r2_ewald_table_host_path_matches_reference(harness, tmp_path):
    if not pyref.available("s1r2_f64"):
        pytest.skip("reference builder needs oracle/_ref")
    g = ic.s1r2_cylinder(3000, 24, 80, 63, lookup=True, is_periodic=2, L=20.0, r_sim=30.0, d_s=8.0, r_crit=10.0).g
    r = pyref.Reference("s1r2_f64")
    r.configure(g, 400)
    r.build_tables()
    r.export_tables(g)
    ref = np.asarray(g.S1R2_EWALD_FORCE_TABLE, dtype=np.float64)
    d = api.s1r2_ewald_defaults(2, g.L, g.Rsim)
    assert (d["nrho"], d["nz"]) == (g.Nrho_EWALD_FORCE_GRID, g.Nz_EWALD_FORCE_GRID)
    tab = _run(harness, ["s1r2", d["nrho"], d["nz"], d["rho_max"], g.L, d["alpha"], d["nmax"], d["mmax"]], tmp_path)
    dz = g.L / d["nz"]
    assert tab.shape == ref.shape and np.isfinite(tab).all()
    assert np.abs(tab - ref).max() <= 1e-15 * 4 / dz**2


@pytest.mark.parametrize("is_periodic", [1, 2])
def test_radial_force_table_host_path_matches_reference(harness, tmp_path, is_periodic):
    if not pyref.available("s1r2nl_f64"):
        pytest.skip("reference builder needs oracle/_ref")
    g = ic.s1r2_cylinder(3000, 24, 80, 62, lookup=False, is_periodic=is_periodic, L=20.0, r_sim=60.0, d_s=10.0, r_crit=15.0).g
    r = pyref.Reference("s1r2nl_f64")
    r.configure(g, 400)
    r.build_tables()
    r.export_tables(g)
    ref = np.asarray(g.RADIAL_FORCE_TABLE, dtype=np.float64)
    Lz = 0.5 * g.L if is_periodic == 1 else g.L * ((is_periodic + 1) - 0.4)
    tab = _run(harness, ["radial", g.Rsim, Lz, ref.size, 400], tmp_path)
    assert np.isfinite(tab).all() and np.abs(tab / ref - 1).max() < 1e-13
