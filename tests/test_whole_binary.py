"""Whole-binary drop-in (SURVEY.md 8c mode (i)): `oracle/_ref/StePS_ref_r3_f64` is the UNMODIFIED reference built as its own executable;
`oracle/_ref/StePS_b200_r3_f64` is the same main.cc, parameter reader, I/O and Friedmann solver with forces.cc and step.cc replaced by
the two shim TUs and libstepsb200.so (oracle/Makefile target `exe`) -- run like `StePS_CUDA <paramfile> <nGPU>`.
CPU tier: the reference binary runs the case; the drop-in binary, without a GPU, fails the way the reference's error convention says
(message, ForceError, main loop ends) -- no CPU fallback.  GPU tier: both binaries run the same parameter file and ASCII IC and their
snapshots and logfiles agree."""
import os
import subprocess

import numpy as np
import pytest

from steps_b200 import ic

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "..", "oracle", "_ref", "StePS_ref_r3_f64")
B200 = os.path.join(HERE, "..", "oracle", "_ref", "StePS_b200_r3_f64")
UNIT_V, UNIT_T = 20.738652969925447, 47.14829951063323

needs_binaries = pytest.mark.skipif(not (os.path.exists(REF) and os.path.exists(B200)), reason="oracle/_ref executables not built (make -C oracle exe)")


def write_case(d, n=400, seed=5, a_factor=1.06):
    """ASCII IC (x y z vx vy vz M per line, inputoutput.cc:142-175) + parameter file (read_paramfile.cc) of a short comoving LCDM run"""
    c = ic.random_sphere(n, seed)
    g = c.g
    x, v = c.x.reshape(-1, 3), c.v.reshape(-1, 3)
    with open(os.path.join(d, "ic.dat"), "w") as f:
        for i in range(g.N):
            f.write("\t".join("%.16f" % t for t in (*x[i], *(v[i] * UNIT_V * np.sqrt(g.a_start)), g.M[i])) + "\n")
    for sub in ("ref", "b200"):
        os.makedirs(os.path.join(d, sub), exist_ok=True)
        with open(os.path.join(d, sub + ".param"), "w") as f:
            f.write(f"""Cosmological parameters:
------------------------
Omega_b		{g.Omega_b}
Omega_lambda	{g.Omega_lambda}
Omega_m   {g.Omega_m}
Omega_r		{g.Omega_r}
HubbleConstant		{g.H0 * UNIT_V}
a_start		{g.a_start}
a_max		{g.a_start * a_factor}

Simulation parameters:
-----------------------
COSMOLOGY	1
IS_PERIODIC	0
COMOVING_INTEGRATION	1
L_BOX		{2 * g.Rsim}
R_SIM		{g.Rsim}
IC_FILE 	{os.path.join(d, 'ic.dat')}
IC_FORMAT	0
OUT_DIR		{os.path.join(d, sub)}/
OUT_LST		./none.txt
OUTPUT_TIME_VARIABLE	0
OUTPUT_FORMAT	0
REDSHIFT_CONE	0
MIN_REDSHIFT	0.0
ACC_PARAM	{g.ACC_PARAM}
STEP_MIN		{g.h_min * UNIT_T}
STEP_MAX           {g.h_max * UNIT_T}
PARTICLE_RADII    {g.ParticleRadi}
FIRST_T_OUT	100.0
H_OUT		100.0
N_PARTICLE	{g.N}
""")
    return c


def run(exe, param, *args):
    os.chmod(exe, 0o755)
    return subprocess.run([exe, param, *args], capture_output=True, text=True, timeout=300)


def snapshot(d, sub):
    files = [f for f in os.listdir(os.path.join(d, sub)) if f.startswith("t") and f.endswith(".dat")]
    assert len(files) == 1, files
    return np.loadtxt(os.path.join(d, sub, files[0]))


@needs_binaries
def test_reference_binary_runs_the_case(tmp_path):
    d = str(tmp_path)
    c = write_case(d)
    r = run(REF, os.path.join(d, "ref.param"))
    assert r.returncode == 0 and "The simulation ended" in r.stdout
    s = snapshot(d, "ref")
    assert s.shape == (c.g.N, 7) and np.isfinite(s).all()
    assert r.stdout.count("KDK Leapfrog integration...done.") >= 3  # several steps were taken


@needs_binaries
def test_dropin_binary_without_a_gpu_fails_loudly():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present: the drop-in binary runs (GPU tier)")
    import tempfile

    with tempfile.TemporaryDirectory() as d:
        write_case(d, n=200)
        r = run(B200, os.path.join(d, "b200.param"), "1")
    out = r.stdout + r.stderr
    assert "no CUDA device available: libstepsb200 has no CPU fallback" in out
    assert "Fatal error has been detected in the force calculation" in out  # main.cc:1851-1856: the reference's own error path
    assert "KDK Leapfrog integration...done." not in out                    # no step was faked on the host


@needs_binaries
@pytest.mark.gpu
def test_dropin_binary_reproduces_the_reference_binary(tmp_path):
    d = str(tmp_path)
    c = write_case(d, n=2000, seed=9)
    g = c.g
    r0 = run(REF, os.path.join(d, "ref.param"))
    r1 = run(B200, os.path.join(d, "b200.param"), "1")
    assert r0.returncode == 0 and r1.returncode == 0, r1.stdout[-2000:] + r1.stderr[-2000:]
    assert "Fatal error" not in r1.stdout + r1.stderr
    s0, s1 = snapshot(d, "ref"), snapshot(d, "b200")
    assert s0.shape == s1.shape == (g.N, 7)
    nsteps = r0.stdout.count("KDK Leapfrog integration...done.")
    assert r1.stdout.count("KDK Leapfrog integration...done.") == nsteps
    dx = np.abs(s1[:, :3] - s0[:, :3]).max() / g.Rsim
    dv = np.abs(s1[:, 3:6] - s0[:, 3:6]).max() / np.abs(s0[:, 3:6]).max()
    print(f"whole binary, {nsteps} steps, N={g.N}: max |dx|/Rsim {dx:.2e}, max |dv|/max|v| {dv:.2e}")
    assert dx < 1e-12 * max(1, nsteps) and dv < 1e-10
    assert np.array_equal(s1[:, 6], s0[:, 6])
    # Logfile.dat: one row per step (time, scale factor, ..., errmax-derived step): same numbers
    l0 = np.loadtxt(os.path.join(d, "ref", "Logfile.dat"), comments="#", ndmin=2)
    l1 = np.loadtxt(os.path.join(d, "b200", "Logfile.dat"), comments="#", ndmin=2)
    assert l0.shape == l1.shape and np.allclose(l1, l0, rtol=1e-8, atol=1e-12)
