/* steps_oracle.h -- TEST INFRASTRUCTURE (CPU checker), not product code.
 * Plain-C restatement of the reference's direct-summation force path and KDK step.
 * Parity pinning: validated against oracle/_ref (the unmodified reference compiled here) and the
 * golden vectors in tests/golden/ generated from it (tools/make_golden.py); see tests/test_oracle.py.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may use it. */
#ifndef STEPS_ORACLE_H
#define STEPS_ORACLE_H
#ifdef __cplusplus
extern "C" {
#endif

typedef struct oracle_params {
    int topology;      /* 0 R^3, 1 T^3, 2 S^1xR^2 lookup, 3 S^1xR^2 NOLOOKUP */
    int n;
    int cosmology, comoving, is_periodic;
    int interp_order;  /* EWALD_INTERPOLATION_ORDER of the S^1xR^2 lookup build (0/2/4); T^3 uses 4 */
    int table_dim0, table_dim1, radial_size;
    int nthreads;      /* OpenMP threads, 0 = default */
    double L, Rsim, mass_in_unit_sphere, H0, Omega_lambda;
    const void *ewald_table;   /* REAL */
    const void *radial_table;  /* REAL */
} oracle_params;

double oracle_force_softening_f64(double r, double beta);
float oracle_force_softening_f32(float r, float beta);
void oracle_softening_f64(const double *M, int n, double particle_radii, double *soft, double *M_min, double *rho_part);
void oracle_softening_f32(const float *M, int n, float particle_radii, float *soft, float *M_min, float *rho_part);
/* F[3*(i-id_min)+k], overwritten */
void oracle_forces_f64(const oracle_params *p, const double *x, const double *M, const double *soft, double *F, int id_min, int id_max);
void oracle_forces_f32(const oracle_params *p, const float *x, const float *M, const float *soft, float *F, int id_min, int id_max);
/* sum_j |f_ij| per particle (for the noise-normalised parity statistic of SURVEY.md H2); R^3 only */
void oracle_force_norms_f64(const oracle_params *p, const double *x, const double *M, const double *soft, double *S, int id_min, int id_max);
/* KDK halves (step.cc:128-181, :254-269): kick(h/2)+drift(h)+wrap; kick(h/2)+errmax. */
void oracle_kick_drift_f64(const oracle_params *p, double *x, double *v, const double *F, double a, double hubble, double h);
void oracle_kick_drift_f32(const oracle_params *p, float *x, float *v, const float *F, double a, double hubble, double h);
double oracle_kick_errmax_f64(const oracle_params *p, double *v, const double *F, const double *soft, double a, double hubble, double h, int do_kick);
double oracle_kick_errmax_f32(const oracle_params *p, float *v, const float *F, const float *soft, double a, double hubble, double h, int do_kick);
/* GLASS_MAKING build of step() (G = -1, step.cc:107-148, :254-303): see steps_oracle_impl.h */
void oracle_glass_kick_drift_f64(const oracle_params *p, double *x, double *v, const double *F, double a, double hubble, double h, double *out2);
void oracle_glass_kick_drift_f32(const oracle_params *p, float *x, float *v, const float *F, double a, double hubble, double h, double *out2);
double oracle_glass_kick_errmax_f64(const oracle_params *p, double *v, const double *F, const double *soft, double a, double hubble, double h, double *out6);
double oracle_glass_kick_errmax_f32(const oracle_params *p, float *v, const float *F, const float *soft, double a, double hubble, double h, double *out6);
double oracle_friedmann_step(double H0, double Om, double Or, double Ol, double Ok, double a0, double h);
double oracle_hubble(double H0, double Om, double Or, double Ol, double Ok, double a);

#ifdef __cplusplus
}
#endif
#endif
