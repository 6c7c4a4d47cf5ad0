// steps_b200_step_shim.cc -- drop-in replacement TU for StePS/src/step.cc: the device-resident KDK step.
//
// Compile it INSIDE the StePS source tree with the build's own -D flags, instead of step.cc, and link
// libstepsb200.so.  It defines, with the reference's exact C++ signatures,
//     double calculate_init_h()                  step.cc:35-98
//     void   step(REAL* x, REAL* v, REAL* F)     step.cc:100-312
// main.cc calls them unchanged (main.cc:1651, :1713).  Positions, velocities and forces live in HBM
// between calls (n_GPU devices of this process, i-partitioned, NCCL position all-gather per step);
// the host copies x, v, F that main.cc's log / snapshot / redshift-cone code reads are refreshed
// after every step (3 x 3N x sizeof(REAL) of D2H -- milliseconds next to an O(N^2) force evaluation;
// set STEPS_B200_LAZY_HOST_STATE=1 to refresh them only every call of steps_b200_shim_sync_host()).
//
// Scope: one MPI rank (numtasks == 1) driving n_GPU devices, the reference's `StePS_CUDA <param> <nGPU>`
// mode.  A -DGLASS_MAKING build switches the engines to that build's arithmetic (G = -1) and reproduces its diagnostics
// (step.cc:143-148, :270-303): the eight statistics are reduced on the device, printed and passed to Log_write_glass as the
// reference does (the device kernels of this mode have not run on a GPU yet, see glass_kernels.cuh).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <omp.h>
#include "mpi.h"
#include "global_variables.h"
#include "steps_b200.h"

void recalculate_softening();
#ifdef GLASS_MAKING
void Log_write_glass(REAL F_mean, REAL Fmax, REAL A_mean, REAL A_max, REAL dmean, REAL dmax, REAL V_mean, REAL V_max);  // inputoutput.cc:974 (declared in step.cc:32)
#endif

namespace {
steps_b200_group *g_group = nullptr;
bool g_lazy = false;

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

bool fail(const char *what) {
    fprintf(stderr, "MPI task %i: steps_b200 %s failed: %s\n", rank, what, steps_b200_last_error());
    ForceError = true;
    return false;
}

// first use: create the engines and move the host state (x, v and the initial F of main.cc:1581-1607) to HBM
bool ensure_resident(REAL *xx, REAL *vv, REAL *FF) {
    if (g_group) return true;
    if (numtasks != 1) {
        fprintf(stderr, "steps_b200 step shim: one MPI rank drives all GPUs of the box (numtasks must be 1, got %d)\n", numtasks);
        ForceError = true;
        return false;
    }
    const steps_b200_params p = pack_globals();
    const char *dev = getenv("STEPS_B200_DEVICE");
    const char *lazy = getenv("STEPS_B200_LAZY_HOST_STATE");
    g_lazy = lazy && atoi(lazy) != 0;
    if (steps_b200_group_create(&g_group, &p, (int)sizeof(REAL), n_GPU > 0 ? n_GPU : 1, dev ? atoi(dev) : 0)) return fail("engine creation");
#ifdef GLASS_MAKING
    if (steps_b200_group_set_glass_making(g_group, 1)) return fail("glass-making mode");
#endif
    // STEPS_B200_SPATIAL_ORDER=<cells per axis>: keep the resident copy sorted by cell (IC files come in arbitrary order; the table gathers
    // of the periodic builds touch 6x fewer cache lines for neighbours in space).  x, v, F of main.cc stay in the file's order.
    if (const char *so = getenv("STEPS_B200_SPATIAL_ORDER"))
        if (atoi(so) > 0 && steps_b200_group_set_spatial_order(g_group, atoi(so))) return fail("spatial order");
    if (steps_b200_group_upload(g_group, xx, vv, M, SOFT_LENGTH, FF)) return fail("state upload");
    // the initial force evaluation of main.cc went through the stateless forces() call, whose cached engines (particle state, partial-sum
    // and row buffers) are not needed any more: the resident group owns the devices from here on
    steps_b200_release_cached();
    // a snapshot still being written when the process ends must reach the disk: join the writer thread at exit
    static bool registered = false;
    if (!registered) {
        registered = true;
        atexit([] {
            if (g_group) steps_b200_group_snapshot_wait(g_group);
        });
    }
    return true;
}
}  // namespace

// explicit refresh of the host copies when STEPS_B200_LAZY_HOST_STATE=1 (call before writing a snapshot)
extern "C" void steps_b200_shim_sync_host() {
    if (g_group && steps_b200_group_download(g_group, x, v, F)) fail("state download");
}

// the resident group (NULL before the first calculate_init_h()/step()): lets a caller that wants snapshots without stalling the GPUs replace
//     write_ascii_snapshot(x, v);                                     (main.cc, OUTPUT_FORMAT == 0)
// by  steps_b200_group_snapshot_ascii_async(steps_b200_shim_group(), filename, H0_dimless, a, glass_making);
// -- same bytes in the file (tests/test_snapshot_io.py), written by a background thread while the next steps run
extern "C" steps_b200_group *steps_b200_shim_group() { return g_group; }

// drop the resident state (a new run in the same process; StePS itself never needs this)
extern "C" void steps_b200_shim_reset() {
    if (g_group) steps_b200_group_destroy(g_group);
    g_group = nullptr;
}

double calculate_init_h() {
    if (!ensure_resident(x, v, F)) return h_min;
    double e = 0.0;
    // wraps the positions into the box on the device as step.cc:42-71 does on the host
    if (steps_b200_group_init_errmax(g_group, a, Hubble_param, &e)) {
        fail("calculate_init_h");
        return h_min;
    }
    if (steps_b200_group_download(g_group, x, nullptr, nullptr)) fail("state download");
    errmax = (REAL)e;
    const double h0 = (double)pow(2 * ACC_PARAM / errmax, 0.5);
    if (COSMOLOGY == 1) {
        if (h0 * UNIT_T >= 1.0) printf("Initial timestep length calculated. h_start=%fGy\n", h0 * UNIT_T);
        else printf("Initial timestep length calculated. h_start=%fMy\n", h0 * UNIT_T * 1000.0);
    } else {
        printf("Initial timestep length calculated. h_start=%f\n", h0);
    }
    return h0;
}

void step(REAL *xx, REAL *vv, REAL *FF) {
    const double t0 = omp_get_wtime();
    if (!ensure_resident(xx, vv, FF)) return;
    printf("KDK Leapfrog integration (device resident, %d GPU)...\n", steps_b200_group_size(g_group));
    // scale factor and Hubble parameter after this step: host scalars, exactly step.cc:230-252
    const double a_old = a, H_old = Hubble_param;
    double a_new = a, H_new = Hubble_param;
    if (COSMOLOGY == 1 && COMOVING_INTEGRATION == 1) {
        a_new = friedmann_solver_step(a, h);
        H_new = CALCULATE_Hubble_param(a_new);
    }
    double e = 0.0;
    if (steps_b200_group_kdk_step(g_group, h, a_old, H_old, a_new, H_new, &e)) {
        fail("KDK step");
        return;
    }
    if (COSMOLOGY == 1) {
        if (COMOVING_INTEGRATION == 1) {
            a = a_new;
            recalculate_softening();
            a_tmp = a;
            Hubble_param = H_new;
            Decel_param = CALCULATE_decel_param(a);
            Omega_m_eff = Omega_m * pow(a, -3) * pow(H0 / Hubble_param, 2);
        } else {
            a_tmp = T;
        }
    } else {
        a_tmp = T;
    }
    errmax = (REAL)e;
    if (!g_lazy && steps_b200_group_download(g_group, xx, vv, FF)) fail("state download");
    printf("KDK Leapfrog integration...done.\n");
#ifdef GLASS_MAKING
    {
        // step.cc:287-302: the statistics of this step, reduced on the device (means over all N particles)
        double gs[8];
        if (steps_b200_group_glass_stats(g_group, gs)) {
            fail("glass-making statistics");
            return;
        }
        const REAL F_mean = (REAL)gs[0], Fmax = (REAL)gs[1], A_mean = (REAL)gs[2], A_max = (REAL)gs[3], dmean = (REAL)gs[4], dmax = (REAL)gs[5],
                   V_mean = (REAL)gs[6], V_max = (REAL)gs[7];
        if (dmax > 1.0)
            printf("Glass making:\tF_max=%e\tA_max = %e\n\t\tdisp-mean=%fMpc\tdisp-maximum = %fMpc\n\t\tV_mean = %e\tV_max = %e\n", Fmax, A_max, dmean, dmax, V_mean, V_max);
        else
            printf("Glass making:\tF_max=%e\tA_max = %e\n\t\tdisp-mean=%fkpc\tdisp-maximum = %fkpc\n\t\tV_mean = %e\tV_max = %e\n", Fmax, A_max, dmean * 1000, dmax * 1000, V_mean, V_max);
        Log_write_glass(F_mean, Fmax, A_mean, A_max, dmean, dmax, V_mean, V_max);
    }
#endif
    printf("Timestep wall-clock time = %fs\n", omp_get_wtime() - t0);
}
