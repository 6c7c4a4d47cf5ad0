/* oracle/ref_harness.cc -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Thin extern "C" driver around the UNMODIFIED reference sources
 * (/root/reference/StePS/src/{main,forces,utils,ewald_space,step,friedmann_solver,...}.cc),
 * which are compiled where they lie by oracle/Makefile into oracle/_ref/libsteps_ref_<variant>.so.
 * Nothing from the reference is copied: this file only *sets the reference's own globals*
 * (declared in its global_variables.h, defined in its main.cc:37-146) and *calls its own
 * functions* (forces / forces_periodic / forces_periodic_z, calculate_softening_length,
 * table builders, step, calculate_init_h, friedmann_solver_*).
 *
 * One shared object per compile-time variant of the reference (Template-LinuxGCC-Makefile:23-47):
 *   r3_f64, r3_f32            : (no topology flag) [+ -DUSE_SINGLE_PRECISION]
 *   t3_f64, t3_f32            : -DPERIODIC                 (CPU interpolation order forced to 4 =
 *                               what the reference CUDA path always uses, main.cc:498-502)
 *   s1r2_f64                  : -DPERIODIC_Z               (lookup table, EWALD_INTERPOLATION_ORDER)
 *   s1r2nl_f64, s1r2nl_f32    : -DPERIODIC_Z -DPERIODIC_Z_NOLOOKUP
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs load it.
 */
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <omp.h>
#include "mpi.h"
#include "global_variables.h"

/* ---- prototypes of reference functions (defined in the reference TUs) ---- */
void calculate_softening_length(REAL *SOFT_LENGTH, REAL *M, int N);   /* utils.cc:59 */
REAL force_softening(REAL r, REAL beta);                              /* forces.cc:52 */
void step(REAL *x, REAL *v, REAL *F);                                 /* step.cc:100 */
double friedmann_solver_start(double a0, double t0, double h, double a_start);
#if defined(PERIODIC)
void forces_periodic(REAL *x, REAL *F, int ID_min, int ID_max);       /* forces.cc:776 */
int ewald_space(REAL R, int ewald_index[][4]);                        /* ewald_space.cc:118 */
void calculate_t3_ewald_lookup_table(int Ngrid, REAL L, REAL alpha, int realspace_el, int recspace_el,
                                     int realspace_ewald_index[][4], int recspace_ewald_index[][4],
                                     REAL rel_cut, REAL rec_cut, REAL *T3_EWALD_FORCE_TABLE);
#elif defined(PERIODIC_Z)
void forces_periodic_z(REAL *x, REAL *F, int ID_min, int ID_max);     /* forces.cc:1221 */
void get_cylindrical_force_table(REAL *FORCE_TABLE, REAL R, REAL Lz, int TABLE_SIZE, int RADIAL_FORCE_ACCURACY);
void set_REAL_array_to_zero(REAL *array, int N);
#if !defined(PERIODIC_Z_NOLOOKUP)
void calculate_S1R2ewald_correction_table(int Nrho, int Nz, REAL rho_max, REAL Lz, REAL alpha, int nmax, int mmax,
                                          REAL *&S1R2_EWALD_FORCE_TABLE);
#endif
#else
void forces(REAL *x, REAL *F, int ID_min, int ID_max);                /* forces.cc:510 */
#endif

/* globals that main.cc defines but global_variables.h does not declare */
extern double a_prev;

void write_ascii_snapshot(REAL *x, REAL *v);  /* inputoutput.cc:826 (C++ linkage, declared in main.cc) */
void write_redshift_cone(REAL *x, REAL *v, double *limits, int z_index, int delta_z_index, int ALL);  /* inputoutput.cc:314 */

extern "C" {

struct sref_config {
    int cosmology;             /* COSMOLOGY */
    int comoving;              /* COMOVING_INTEGRATION */
    int is_periodic;           /* IS_PERIODIC */
    int radial_table_size;     /* RADIAL_FORCE_TABLE_SIZE (S1xR2) */
    int radial_accuracy;       /* RADIAL_FORCE_ACCURACY   (S1xR2) */
    int pad_;
    double L;                  /* L_BOX */
    double Rsim;               /* R_SIM */
    double H0;                 /* already divided by UNIT_V (read_paramfile.cc:401-403) */
    double Omega_m, Omega_lambda, Omega_r, Omega_b;
    double particle_radii;     /* PARTICLE_RADII */
    double acc_param;          /* ACC_PARAM */
    double h_min, h_max;       /* internal time units */
    double a_start;
};

int sref_real_bytes(void) { return (int)sizeof(REAL); }

/* 0 = R^3, 1 = T^3, 2 = S^1xR^2 lookup, 3 = S^1xR^2 NOLOOKUP */
int sref_topology(void)
{
#if defined(PERIODIC)
    return 1;
#elif defined(PERIODIC_Z) && !defined(PERIODIC_Z_NOLOOKUP)
    return 2;
#elif defined(PERIODIC_Z)
    return 3;
#else
    return 0;
#endif
}

static int g_alloc_n = 0;

/* Mirrors the initialisation order of main.cc: parameters (read_paramfile.cc), rho_crit and
 * mass_in_unit_sphere (main.cc:1256-1313), softening (main.cc:1411 -> utils.cc:59-82). */
int sref_configure(const sref_config *c)
{
    numtasks = 1; rank = 0; n_GPU = 0; ForceError = false;
    COSMOLOGY = c->cosmology;
    COMOVING_INTEGRATION = c->comoving;
    IS_PERIODIC = c->is_periodic;
    L = (REAL)c->L;
    Rsim = (REAL)c->Rsim;
    H0 = c->H0;
    Omega_m = c->Omega_m; Omega_lambda = c->Omega_lambda; Omega_r = c->Omega_r; Omega_b = c->Omega_b;
    Omega_dm = Omega_m - Omega_b;
    Omega_k = 1. - Omega_m - Omega_lambda - Omega_r;
    ParticleRadi = (REAL)c->particle_radii;
    beta = ParticleRadi;
    ACC_PARAM = (REAL)c->acc_param;
    h_min = c->h_min; h_max = c->h_max;
    a_start = c->a_start; a = a_start; a_tmp = a; a_max = 1.0;
    if (COSMOLOGY == 0) a = 1;
    T = 0.0; t_next = 0.0;
    rho_crit = 3.0 * H0 * H0 / (8.0 * pi);
    mass_in_unit_sphere = 0;
    if (COSMOLOGY == 1 && COMOVING_INTEGRATION == 1) {
#if defined(PERIODIC_Z)
        mass_in_unit_sphere = (REAL)(2.0 * pi * rho_crit * Omega_m);          /* main.cc:1269,1288 */
#elif !defined(PERIODIC)
        mass_in_unit_sphere = (REAL)(4.0 * pi * rho_crit * Omega_m / 3.0);    /* main.cc:1313 */
#endif
    }
    Hubble_param = 0.0;
    if (COSMOLOGY == 1 && COMOVING_INTEGRATION == 1) Hubble_param = CALCULATE_Hubble_param(a);
#if defined(PERIODIC_Z)
    RADIAL_FORCE_TABLE_SIZE = c->radial_table_size;
    RADIAL_FORCE_ACCURACY = c->radial_accuracy;
#endif
    return 0;
}

#ifdef STEPS_SHIM_BUILD
void steps_b200_shim_reset();   /* steps_b200_step_shim.cc */
#endif

int sref_set_particles(int n, const REAL *masses)
{
#ifdef STEPS_SHIM_BUILD
    steps_b200_shim_reset();
#endif
    if (g_alloc_n) { free(M); free(SOFT_LENGTH); free(x); free(v); free(F); }
    N = n; N_mpi_thread = n; ID_MPI_min = 0; ID_MPI_max = n - 1;
    M = (REAL *)malloc(sizeof(REAL) * n);
    SOFT_LENGTH = (REAL *)malloc(sizeof(REAL) * n);
    x = (REAL *)malloc(sizeof(REAL) * 3 * (size_t)n);
    v = (REAL *)malloc(sizeof(REAL) * 3 * (size_t)n);
    F = (REAL *)malloc(sizeof(REAL) * 3 * (size_t)n);
    if (!M || !SOFT_LENGTH || !x || !v || !F) return -2;
    g_alloc_n = n;
    memcpy(M, masses, sizeof(REAL) * n);
    calculate_softening_length(SOFT_LENGTH, M, N);
    return 0;
}

void sref_get_softening(REAL *out) { memcpy(out, SOFT_LENGTH, sizeof(REAL) * N); }

/* scalars the force path reads: M_min, rho_part, mass_in_unit_sphere, DE=H0^2*Omega_lambda */
void sref_get_scalars(double *out4)
{
    out4[0] = (double)M_min; out4[1] = (double)rho_part; out4[2] = (double)mass_in_unit_sphere;
    out4[3] = (double)((REAL)H0 * H0 * Omega_lambda);
}

#if !defined(STEPS_SHIM_BUILD) && !defined(USE_CUDA)
REAL sref_force_softening(REAL r, REAL b) { return force_softening(r, b); }
#endif
/* 1 when forces()/step() of this library are the steps_b200 shims (drop-in build), 0 for the pure reference */
int sref_is_shim(void)
{
#ifdef STEPS_SHIM_BUILD
    return 1;
#else
    return 0;
#endif
}
/* n_GPU of the reference (argv[2], main.cc:1186-1195): devices one process drives */
void sref_set_n_gpu(int n) { n_GPU = n; }

/* Builds the topology's lookup tables with the reference's own builders, following the call
 * sites in main.cc:412-534 (T^3), :562-726 and :1263-1310 (S^1xR^2).  Returns 0 on success. */
int sref_build_tables(void)
{
#if defined(PERIODIC)
    if (IS_PERIODIC > 1) {
        static int real_idx[739][4];
        static int rec_idx[11459][4];
        REAL rel_cut, rec_cut;
        double alpha = 2.0 / L;
        if (IS_PERIODIC == 2) { rel_cut = 2.6; rec_cut = 8.0; N_EWALD_FORCE_GRID = 63; }
        else if (IS_PERIODIC == 3) { rel_cut = 3.6; rec_cut = 10.0; N_EWALD_FORCE_GRID = 127; }
        else { rel_cut = 4.6; rec_cut = 12.0; N_EWALD_FORCE_GRID = 255; }
        size_t ng = (size_t)N_EWALD_FORCE_GRID;
        T3_EWALD_FORCE_TABLE = (REAL *)malloc(ng * ng * ng * 3 * sizeof(REAL));
        if (!T3_EWALD_FORCE_TABLE) return -2;
        int nreal = ewald_space(rel_cut + 1, real_idx);
        int nrec = ewald_space(rec_cut + 2, rec_idx);
        calculate_t3_ewald_lookup_table(N_EWALD_FORCE_GRID, L, alpha, nreal, nrec, real_idx, rec_idx, rel_cut, rec_cut,
                                        T3_EWALD_FORCE_TABLE);
    } else {
        N_EWALD_FORCE_GRID = 1;
        T3_EWALD_FORCE_TABLE = (REAL *)calloc(3, sizeof(REAL));
    }
#elif defined(PERIODIC_Z)
    RADIAL_FORCE_TABLE = (REAL *)malloc(RADIAL_FORCE_TABLE_SIZE * sizeof(REAL));
#if defined(PERIODIC_Z_NOLOOKUP)
    ewald_max = IS_PERIODIC + 1;
    ewald_cut = ((REAL)ewald_max) - 0.4;
    if (IS_PERIODIC == 1) get_cylindrical_force_table(RADIAL_FORCE_TABLE, Rsim, 0.5 * L, RADIAL_FORCE_TABLE_SIZE, RADIAL_FORCE_ACCURACY);
    else get_cylindrical_force_table(RADIAL_FORCE_TABLE, Rsim, L * ewald_cut, RADIAL_FORCE_TABLE_SIZE, RADIAL_FORCE_ACCURACY);
#else
    if (IS_PERIODIC == 1) get_cylindrical_force_table(RADIAL_FORCE_TABLE, Rsim, 0.5 * L, RADIAL_FORCE_TABLE_SIZE, RADIAL_FORCE_ACCURACY);
    else set_REAL_array_to_zero(RADIAL_FORCE_TABLE, RADIAL_FORCE_TABLE_SIZE);
    if (IS_PERIODIC > 1) {
        double alpha; int rel_cut, rec_cut;
        if (IS_PERIODIC == 2) { rel_cut = 4; rec_cut = 10; alpha = 0.787875 / L; Nz_EWALD_FORCE_GRID = 128; }
        else if (IS_PERIODIC == 3) { rel_cut = 5; rec_cut = 12; alpha = 0.71805 / L; Nz_EWALD_FORCE_GRID = 256; }
        else { rel_cut = IS_PERIODIC + 2; rec_cut = IS_PERIODIC + 9; alpha = 0.6642 / L; Nz_EWALD_FORCE_GRID = 512; }
        Nrho_EWALD_FORCE_GRID = (int)floor(((REAL)Nz_EWALD_FORCE_GRID) * EWALD_LOOKUP_TABLE_RADIAL_EXTENT_FACTOR * Rsim / L);
        S1R2_EWALD_FORCE_TABLE = (REAL *)malloc((size_t)Nz_EWALD_FORCE_GRID * Nrho_EWALD_FORCE_GRID * 2 * sizeof(REAL));
        if (!S1R2_EWALD_FORCE_TABLE) return -2;
        calculate_S1R2ewald_correction_table(Nrho_EWALD_FORCE_GRID, Nz_EWALD_FORCE_GRID,
                                             EWALD_LOOKUP_TABLE_RADIAL_EXTENT_FACTOR * Rsim, L, alpha, rel_cut, rec_cut,
                                             S1R2_EWALD_FORCE_TABLE);
    } else {
        Nz_EWALD_FORCE_GRID = 1; Nrho_EWALD_FORCE_GRID = 1;
        S1R2_EWALD_FORCE_TABLE = (REAL *)calloc(2, sizeof(REAL));
    }
#endif
#endif
    return 0;
}

/* table access: which = 0 -> main Ewald table (T^3: [Ng^3*3], S1R2 lookup: [Nrho*Nz*2]);
 *               which = 1 -> RADIAL_FORCE_TABLE.  dims[0..1] receive the grid dims. */
const REAL *sref_table(int which, int *dims)
{
    dims[0] = dims[1] = 0;
#if defined(PERIODIC)
    if (which == 0) { dims[0] = N_EWALD_FORCE_GRID; dims[1] = N_EWALD_FORCE_GRID; return T3_EWALD_FORCE_TABLE; }
#elif defined(PERIODIC_Z)
    if (which == 1) { dims[0] = RADIAL_FORCE_TABLE_SIZE; return RADIAL_FORCE_TABLE; }
#if !defined(PERIODIC_Z_NOLOOKUP)
    if (which == 0) { dims[0] = Nrho_EWALD_FORCE_GRID; dims[1] = Nz_EWALD_FORCE_GRID; return S1R2_EWALD_FORCE_TABLE; }
#endif
#endif
    (void)which;
    return nullptr;
}

/* The hot path: the reference's own force entry point for this build's topology
 * (step.cc:191-197).  xin: AoS [3N]; Fout: [3*(id_max-id_min+1)], overwritten. */
int sref_forces(const REAL *xin, REAL *Fout, int id_min, int id_max, int nthreads)
{
    if (nthreads > 0) omp_set_num_threads(nthreads);
    N_mpi_thread = id_max - id_min + 1;
    ID_MPI_min = id_min; ID_MPI_max = id_max;
    REAL *xx = const_cast<REAL *>(xin);
#if defined(PERIODIC)
    forces_periodic(xx, Fout, id_min, id_max);
#elif defined(PERIODIC_Z)
    forces_periodic_z(xx, Fout, id_min, id_max);
#else
    forces(xx, Fout, id_min, id_max);
#endif
    return ForceError ? 1 : 0;
}

/* ---- KDK: the reference's own step() (step.cc:100-312) on its own global x,v,F ---- */

/* load state, evaluate initial forces (main.cc:1581-1607), return calculate_init_h() (step.cc:35-98) */
double sref_kdk_begin(const REAL *x0, const REAL *v0, int nthreads)
{
    if (nthreads > 0) omp_set_num_threads(nthreads);
    memcpy(x, x0, sizeof(REAL) * 3 * (size_t)N);
    memcpy(v, v0, sizeof(REAL) * 3 * (size_t)N);
    N_mpi_thread = N; ID_MPI_min = 0; ID_MPI_max = N - 1;
#ifdef STEPS_SHIM_BUILD
    steps_b200_shim_reset();
#endif
    a = a_start; a_tmp = a; T = 0.0;
    if (COSMOLOGY == 0) a = 1;
    if (COSMOLOGY == 1 && COMOVING_INTEGRATION == 1) Hubble_param = CALCULATE_Hubble_param(a);
    else Hubble_param = 0.0;
    sref_forces(x, F, 0, N - 1, nthreads);
    return calculate_init_h();
}

/* one step with timestep hh; afterwards applies the reference's timestep rule (main.cc:1834-1846,
 * without the output-time clamp) and returns the NEXT h.  out[0..3] = errmax, a, Hubble_param, T. */
double sref_kdk_step(double hh, double *out4)
{
    h = hh;
    T = T + h;
    step(x, v, F);
    out4[0] = (double)errmax; out4[1] = a; out4[2] = Hubble_param; out4[3] = T;
    double hn = (double)pow(2 * ACC_PARAM / errmax, 0.5);
    if (hn < h_min) hn = h_min; else if (hn > h_max) hn = h_max;
    return hn;
}

void sref_kdk_state(REAL *xo, REAL *vo, REAL *Fo)
{
    if (xo) memcpy(xo, x, sizeof(REAL) * 3 * (size_t)N);
    if (vo) memcpy(vo, v, sizeof(REAL) * 3 * (size_t)N);
    if (Fo) memcpy(Fo, F, sizeof(REAL) * 3 * (size_t)N);
}

/* the reference's own ASCII snapshot writer (inputoutput.cc:826-909) on the state given here; the file is <dir>t<round(100 t_next UNIT_T)>.dat
 * (COSMOLOGY = 1, OUTPUT_TIME_VARIABLE = 0) */
int sref_write_ascii_snapshot(const char *dir, const REAL *x0, const REAL *v0, double a_now, double t_next_now, int h0_independent_units)
{
    strncpy(OUT_DIR, dir, sizeof(OUT_DIR) - 1);
    OUT_DIR[sizeof(OUT_DIR) - 1] = 0;
    memcpy(x, x0, sizeof(REAL) * 3 * (size_t)N);
    memcpy(v, v0, sizeof(REAL) * 3 * (size_t)N);
    a = a_now;
    t_next = t_next_now;
    OUTPUT_TIME_VARIABLE = 0;
    H0_INDEPENDENT_UNITS = h0_independent_units;
    write_ascii_snapshot(x, v);
    return 0;
}

/* the reference's own redshift-cone writer (inputoutput.cc:314-405, ASCII branch) on the state given here: appends to
 * <dir>redshift_cone.dat and updates IN_CONE; reset_flags != 0 starts from an empty cone.  limits[n_limits] = r_bin_limits (descending),
 * zlist[n_limits - 1...] = out_list as main.cc holds them. */
int sref_write_redshift_cone(const char *dir, const REAL *x0, const REAL *v0, const double *limits, int n_limits, const double *zlist, int n_z,
                             int z_index, int delta_z_index, int all, int h0_independent_units, int reset_flags, double t_next_now)
{
    static double *own_limits = NULL, *own_z = NULL;
    strncpy(OUT_DIR, dir, sizeof(OUT_DIR) - 1);
    OUT_DIR[sizeof(OUT_DIR) - 1] = 0;
    memcpy(x, x0, sizeof(REAL) * 3 * (size_t)N);
    memcpy(v, v0, sizeof(REAL) * 3 * (size_t)N);
    OUTPUT_FORMAT = 0;
    H0_INDEPENDENT_UNITS = h0_independent_units;
    t_next = t_next_now;
    free(own_limits);
    free(own_z);
    own_limits = (double *)malloc(sizeof(double) * (size_t)n_limits);
    own_z = (double *)malloc(sizeof(double) * (size_t)n_z);
    memcpy(own_limits, limits, sizeof(double) * (size_t)n_limits);
    memcpy(own_z, zlist, sizeof(double) * (size_t)n_z);
    out_list = own_z;
    out_list_size = n_z;
    if (reset_flags || !IN_CONE) {
        delete[] IN_CONE;
        IN_CONE = new bool[N];
        for (int i = 0; i < N; i++) IN_CONE[i] = false;
    }
    write_redshift_cone(x, v, own_limits, z_index, delta_z_index, all);
    return 0;
}

/* GLASS_MAKING builds: step() appends its diagnostics to <OUT_DIR>Glass_logfile.dat (Log_write_glass, inputoutput.cc:974) */
int sref_is_glass(void)
{
#ifdef GLASS_MAKING
    return 1;
#else
    return 0;
#endif
}
void sref_set_out_dir(const char *dir)
{
    strncpy(OUT_DIR, dir, sizeof(OUT_DIR) - 1);
    OUT_DIR[sizeof(OUT_DIR) - 1] = 0;
}

double sref_friedmann_step(double a0, double hh) { return friedmann_solver_step(a0, hh); }
double sref_hubble(double aa) { return CALCULATE_Hubble_param(aa); }

} /* extern "C" */
