"""Generate tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref, built from /root/reference).

Run in the authoring container:  python tools/make_golden.py
The fixtures pin the plain-C oracle port (tests/test_oracle.py, CPU) and the CUDA path
(tests/test_gpu_parity.py, GPU box -- where /root/reference does not exist).
Each fixture stores every input the force path reads (positions, masses, softening lengths, scalar
globals, tables where small) and the reference's outputs.
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

from oracle.pyref import Reference  # noqa: E402
from steps_b200 import ic  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")

SCALARS = ["topology", "N", "COSMOLOGY", "COMOVING_INTEGRATION", "IS_PERIODIC", "L", "Rsim", "H0", "Omega_m", "Omega_lambda",
           "Omega_r", "Omega_b", "ParticleRadi", "ACC_PARAM", "h_min", "h_max", "a_start", "EWALD_INTERPOLATION_ORDER",
           "RADIAL_FORCE_TABLE_SIZE", "mass_in_unit_sphere", "M_min", "rho_part", "N_EWALD_FORCE_GRID", "Nrho_EWALD_FORCE_GRID",
           "Nz_EWALD_FORCE_GRID"]


def pack(c, **extra):
    g = c.g
    d = {k: np.asarray(getattr(g, k)) for k in SCALARS}
    d["real_bytes"] = np.asarray(8 if g.REAL == np.float64 else 4)
    d["x"], d["v"], d["M"], d["SOFT_LENGTH"] = c.x, c.v, g.M, g.SOFT_LENGTH
    if g.RADIAL_FORCE_TABLE is not None:
        d["RADIAL_FORCE_TABLE"] = g.RADIAL_FORCE_TABLE
    d.update(extra)
    return d


def setup_ref(c, radial_accuracy=300):
    g = c.g
    r = Reference.for_globals(g)
    r.configure(g, radial_accuracy)
    # take the reference's own derived scalars / arrays so both sides see identical inputs
    s = r.scalars()
    g.mass_in_unit_sphere = s["mass_in_unit_sphere"]
    g.M_min, g.rho_part = s["M_min"], s["rho_part"]
    g.SOFT_LENGTH = r.softening()
    if g.topology != 0:
        r.build_tables()
        r.export_tables(g)
    return r


def force_case(name, c, store_table=True, radial_accuracy=300):
    r = setup_ref(c, radial_accuracy)
    g = c.g
    F = r.forces(c.x, 0, g.N - 1, 1)
    lo, hi = g.N // 3, g.N // 3 + 40  # a sub-range call: index relative to ID_min
    Fsub = r.forces(c.x, lo, hi, 1)
    extra = dict(F=F, sub_lo=np.asarray(lo), sub_hi=np.asarray(hi), Fsub=Fsub)
    if store_table and g.topology == 2 and g.S1R2_EWALD_FORCE_TABLE is not None:
        extra["S1R2_EWALD_FORCE_TABLE"] = g.S1R2_EWALD_FORCE_TABLE
    if g.topology == 1 and g.T3_EWALD_FORCE_TABLE is not None:
        t = g.T3_EWALD_FORCE_TABLE
        # the 63^3 table is 6 MB: store a checksum and a few entries; tests rebuild it with oracle/_ref
        extra["T3_table_sum"] = np.asarray(t.sum(dtype=np.float64))
        extra["T3_table_abs_sum"] = np.asarray(np.abs(t).sum(dtype=np.float64))
        extra["T3_table_probe"] = t[:: max(1, t.size // 997)][:997].copy()
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **pack(c, **extra))
    print(name, "N", g.N, "|F| max", np.abs(F).max())


def kdk_case(name, c, nsteps):
    r = setup_ref(c)
    h = r.kdk_begin(c.x, c.v, 1)
    hs, errs, aa, HH = [h], [], [], []
    x0, v0, F0 = r.kdk_state()
    for _ in range(nsteps):
        h, st = r.kdk_step(h)
        hs.append(h)
        errs.append(st["errmax"]); aa.append(st["a"]); HH.append(st["H"])
    x1, v1, F1 = r.kdk_state()
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **pack(c, F0=F0, x1=x1, v1=v1, F1=F1, h_seq=np.asarray(hs),
                                                                  errmax_seq=np.asarray(errs), a_seq=np.asarray(aa), H_seq=np.asarray(HH)))
    print(name, "steps", nsteps, "h", hs[:3], "errmax", errs[:2])


def scalar_cases():
    c = ic.random_sphere(64, 1)
    r = Reference("r3_f64"); r.configure(c.g)
    r32 = Reference("r3_f32"); c32 = ic.random_sphere(64, 1, np.float32); r32.configure(c32.g)
    betas = np.array([0.05, 0.4, 1.0, 3.7])
    fr = np.array([0.0, 1e-6, 0.1, 0.25, 0.4999, 0.5, 0.5001, 0.75, 0.9999, 1.0, 1.0001, 1.5, 10.0, 1e3])
    rr = (betas[:, None] * fr[None, :]).reshape(-1)
    bb = np.repeat(betas, fr.size)
    w64 = np.array([r.force_softening(a, b) for a, b in zip(rr, bb)])
    w32 = np.array([r32.force_softening(np.float32(a), np.float32(b)) for a, b in zip(rr, bb)], dtype=np.float32)
    # Friedmann RK4 + Hubble
    a_seq, H_seq = [c.g.a_start], [r.hubble(c.g.a_start)]
    hs = np.array([1e-4, 3e-4, 6e-4, 6.6e-4, 6.6e-4, 2e-3, -1e-3])
    for h in hs:
        a_seq.append(r.friedmann_step(a_seq[-1], h)); H_seq.append(r.hubble(a_seq[-1]))
    np.savez_compressed(os.path.join(OUT, "scalars.npz"), r=rr, beta=bb, w64=w64, w32=w32, fr_h=hs, fr_a=np.asarray(a_seq),
                        fr_H=np.asarray(H_seq), H0=np.asarray(c.g.H0), Omega_m=np.asarray(c.g.Omega_m),
                        Omega_lambda=np.asarray(c.g.Omega_lambda), Omega_r=np.asarray(c.g.Omega_r))
    print("scalars", w64[:4])


def order_cases():
    """S^1xR^2 lookup build with the other two interpolation orders (EWALD_INTERPOLATION_ORDER = 2 is the reference template's
    default, Template-LinuxGCC-Makefile:35; 0 = NGP): forces_cuda.cu:188-283 / ewald_space.cc:803-960"""
    kw = dict(L=20.0, r_sim=10.0, d_s=4.0, r_crit=3.0)
    for tag, order, seed in (("cic", 2, 114), ("ngp", 0, 115)):
        c = ic.s1r2_cylinder(320, 12, 20, seed, lookup=True, is_periodic=2, **kw)
        c.g.EWALD_INTERPOLATION_ORDER = order
        force_case(f"s1r2_f64_lookup_{tag}", c)


def main():
    os.makedirs(OUT, exist_ok=True)
    if "orders" in sys.argv[1:]:
        return order_cases()
    scalar_cases()
    order_cases()
    force_case("r3_f64_comoving", ic.random_sphere(300, 101))
    force_case("r3_f64_noncomoving", ic.random_sphere(257, 102, comoving=0))
    force_case("r3_f64_nocosmo", ic.random_sphere(129, 103, cosmology=0))
    force_case("r3_f32_comoving", ic.random_sphere(300, 104, np.float32))
    c = ic.compactified_r3(1024, 16, 40, 105, d_s=20.0, r_sim=150.0, r_crit=25.0)
    force_case("r3_f64_zoom", c)
    force_case("t3_f64_quasi", ic.t3_lattice(6, 106, L=12.0, is_periodic=1))
    force_case("t3_f64_ewald", ic.t3_lattice(6, 107, L=12.0, is_periodic=2))
    force_case("t3_f32_ewald", ic.t3_lattice(5, 108, np.float32, L=12.0, is_periodic=2))
    kw = dict(L=20.0, r_sim=10.0, d_s=4.0, r_crit=3.0)
    force_case("s1r2nl_f64_images", ic.s1r2_cylinder(320, 12, 20, 109, lookup=False, is_periodic=2, **kw))
    force_case("s1r2nl_f64_quasi", ic.s1r2_cylinder(320, 12, 20, 110, lookup=False, is_periodic=1, **kw))
    force_case("s1r2nl_f32_images", ic.s1r2_cylinder(320, 12, 20, 111, np.float32, lookup=False, is_periodic=3, **kw))
    force_case("s1r2_f64_lookup", ic.s1r2_cylinder(320, 12, 20, 112, lookup=True, is_periodic=2, **kw))
    force_case("s1r2_f64_lookup_quasi", ic.s1r2_cylinder(320, 12, 20, 113, lookup=True, is_periodic=1, **kw))
    kdk_case("kdk_r3_f64", ic.random_sphere(200, 120), 10)
    kdk_case("kdk_r3_f32", ic.random_sphere(200, 121, np.float32), 10)
    kdk_case("kdk_t3_f64", ic.t3_lattice(5, 122, L=12.0, is_periodic=1), 6)
    kdk_case("kdk_s1r2nl_f64", ic.s1r2_cylinder(240, 12, 15, 123, lookup=False, is_periodic=2, **kw), 6)


if __name__ == "__main__":
    main()
