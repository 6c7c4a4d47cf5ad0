// steps_b200_forces_shim.cc -- drop-in replacement TU for StePS/src/forces_cuda.cu (or forces.cc).
//
// Compile it INSIDE the StePS source tree with the build's own -D flags (it includes StePS's own
// mpi.h and global_variables.h, nothing of StePS is copied here) and link libstepsb200.so instead
// of forces_cuda.o / forces.o.  It defines, with the reference's exact C++ signatures,
//     void forces(REAL*, REAL*, int, int)             (no topology flag)   forces_cuda.cu:74-78
//     void forces_periodic(REAL*, REAL*, int, int)    (-DPERIODIC)         forces_cuda.cu:169-173
//     void forces_periodic_z(REAL*, REAL*, int, int)  (-DPERIODIC_Z ...)   forces_cuda.cu:457-461
//     void recalculate_softening()                                         forces_cuda.cu:868-875
// so step.cc:191-197 and main.cc:1581-1607 link against it unchanged.  The globals the reference
// kernels read at link time are packed into steps_b200_params on every call (they are scalars and
// pointers; the device buffers are cached between calls, the tables are uploaded again on every call,
// as the reference does, because a caller may hand in different contents at the same address).
//
// Error convention of the reference (forces_cuda.cu:970-974, main.cc:1851-1856): message on stderr,
// ForceError = true, return.  There is no CPU fallback.
#include <cmath>
#include <cstdio>
#include "mpi.h"
#include "global_variables.h"
#include "steps_b200.h"

namespace {
steps_b200_params pack_globals() {
    steps_b200_params p{};
    p.abi_version = STEPS_B200_ABI_VERSION;
    p.n = N;
    p.cosmology = COSMOLOGY;
    p.comoving = COMOVING_INTEGRATION;
    p.is_periodic = IS_PERIODIC;
    p.L = (double)L;
    p.Rsim = (double)Rsim;
    p.mass_in_unit_sphere = (double)mass_in_unit_sphere;
    p.H0 = H0;
    p.Omega_lambda = Omega_lambda;
#if defined(PERIODIC)
    p.topology = STEPS_TOPO_T3;
    p.ewald_table = T3_EWALD_FORCE_TABLE;
    p.table_dim0 = p.table_dim1 = N_EWALD_FORCE_GRID;
#elif defined(PERIODIC_Z) && !defined(PERIODIC_Z_NOLOOKUP)
    p.topology = STEPS_TOPO_S1R2_LOOKUP;
    p.ewald_table = S1R2_EWALD_FORCE_TABLE;
    p.table_dim0 = Nrho_EWALD_FORCE_GRID;
    p.table_dim1 = Nz_EWALD_FORCE_GRID;
#if defined(EWALD_INTERPOLATION_ORDER)
    p.s1r2_interp_order = EWALD_INTERPOLATION_ORDER;
#else
    p.s1r2_interp_order = 4;
#endif
#elif defined(PERIODIC_Z)
    p.topology = STEPS_TOPO_S1R2_NOLOOKUP;
#else
    p.topology = STEPS_TOPO_R3;
#endif
#if defined(PERIODIC_Z)
    p.radial_table = RADIAL_FORCE_TABLE;
    p.radial_table_size = RADIAL_FORCE_TABLE_SIZE;
#endif
    return p;
}

void run(REAL *xx, REAL *FF, int ID_min, int ID_max) {
    const steps_b200_params p = pack_globals();
    const int ngpu = n_GPU > 0 ? n_GPU : 1;
#ifdef USE_SINGLE_PRECISION
    const int rc = steps_b200_forces_multi_f32(&p, xx, M, SOFT_LENGTH, FF, ID_min, ID_max, ngpu);
#else
    const int rc = steps_b200_forces_multi_f64(&p, xx, M, SOFT_LENGTH, FF, ID_min, ID_max, ngpu);
#endif
    if (rc != 0) {
        fprintf(stderr, "MPI task %i: steps_b200 force calculation failed: %s\n", rank, steps_b200_last_error());
        ForceError = true;
    }
}
}  // namespace

#if defined(PERIODIC)
void forces_periodic(REAL *x, REAL *F, int ID_min, int ID_max) { run(x, F, ID_min, ID_max); }
#elif defined(PERIODIC_Z)
void forces_periodic_z(REAL *x, REAL *F, int ID_min, int ID_max) { run(x, F, ID_min, ID_max); }
#else
void forces(REAL *x, REAL *F, int ID_min, int ID_max) { run(x, F, ID_min, ID_max); }
#endif

// per-step refresh of the two softening scalars (forces.cc:43-50 == forces_cuda.cu:868-875); the
// per-particle lengths SOFT_LENGTH[] are constant over a run
void recalculate_softening() {
    beta = ParticleRadi;
    if (COSMOLOGY == 1) rho_part = M_min / (4.0 * pi * pow(beta, 3.0) / 3.0);
}
