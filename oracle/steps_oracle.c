/* steps_oracle.c -- TEST INFRASTRUCTURE (CPU checker), not product code.  See steps_oracle.h. */
#include <math.h>
#include <stddef.h>
#include <omp.h>
#include "steps_oracle.h"

#define ORACLE_PI 3.14159265358979323846264338327950288419716939937510

static int imodp_c(int i, int n) { int r = i % n; return (r < 0) ? (r + n) : r; }

#define REAL double
#define SFX _f64
#include "steps_oracle_impl.h"
#undef REAL
#undef SFX

#define REAL float
#define SFX _f32
#include "steps_oracle_impl.h"
#undef REAL
#undef SFX

/* sum_j |f_ij| (R^3, far+soft exact) -- denominator of the noise-normalised parity statistic */
void oracle_force_norms_f64(const oracle_params *p, const double *x, const double *M, const double *soft, double *S, int id_min, int id_max)
{
    const int N = p->n;
    if (p->nthreads > 0) omp_set_num_threads(p->nthreads);
#pragma omp parallel for schedule(dynamic, 16)
    for (int i = id_min; i <= id_max; i++) {
        double s = 0;
        for (int j = 0; j < N; j++) {
            double dx = x[3 * j] - x[3 * i], dy = x[3 * j + 1] - x[3 * i + 1], dz = x[3 * j + 2] - x[3 * i + 2];
            /* periodic topologies: the nearest image sets the scale (T^3: all axes, forces.cc:800-812; S^1xR^2: z, forces.cc:1320-1323) */
            if (p->topology == 1) {
                if (fabs(dx) > 0.5 * p->L) dx -= copysign(p->L, dx);
                if (fabs(dy) > 0.5 * p->L) dy -= copysign(p->L, dy);
            }
            if (p->topology != 0 && fabs(dz) > 0.5 * p->L) dz -= copysign(p->L, dz);
            double r = sqrt(dx * dx + dy * dy + dz * dz);
            s += fabs(M[j] * oracle_force_softening_f64(r, soft[i] + soft[j])) * r;
        }
        S[i - id_min] = s;
    }
}

/* friedmann_solver.cc:100-159 friedmann_solver_step (COSMOPARAM == 0, LCDM) */
double oracle_friedmann_step(double H0, double Om, double Or, double Ol, double Ok, double a0, double h)
{
    double b = a0, j, l, m, n, k1, k2, k3, k4, K;
    if (fabs(Ok) < 1e-9) {
        j = Om * pow(b, -3.0) + Or * pow(b, -4.0) + Ol;
        k1 = b * H0 * sqrt(j);
        l = Om * pow(b + h * k1 / 2.0, -3.0) + Or * pow(b + h * k1 / 2, -4.0) + Ol;
        k2 = (b + h * k1 / 2.0) * H0 * sqrt(l);
        m = Om * pow(b + h * k2 / 2.0, -3.0) + Or * pow(b + h * k2 / 2, -4.0) + Ol;
        k3 = (b + h * k2 / 2.0) * H0 * sqrt(m);
        n = Om * pow(b + h * k3, -3.0) + Or * pow(b + h * k3, -4.0) + Ol;
        k4 = (b + h * k3) * H0 * sqrt(n);
        K = h * (k1 + k2 * 2.0 + k3 * 2.0 + k4) / 6.0;
        if (j < 0 || l < 0 || m < 0 || n < 0) b = -1; else b += K;
    } else {
        int collapse = H0 > 0 ? 0 : 1;
        j = Om * pow(b, -3.0) + Or * pow(b, -4.0) + Ol + Ok * pow(b, -2.0);
        k1 = b * H0 * sqrt(fabs(j));
        l = Om * pow(b + h * k1 / 2.0, -3.0) + Or * pow(b + h * k1 / 2.0, -4.0) + Ol + Ok * pow(b + h * k1 / 2.0, -2.0);
        k2 = (b + h * k1 / 2.0) * H0 * sqrt(fabs(l));
        m = Om * pow(b + h * k2 / 2.0, -3.0) + Or * pow(b + h * k2 / 2.0, -4.0) + Ol + Ok * pow(b + h * k2 / 2.0, -2.0);
        k3 = (b + h * k2 / 2.0) * H0 * sqrt(fabs(m));
        n = Om * pow(b + h * k3, -3.0) + Or * pow(b + h * k3, -4.0) + Ol + Ok * pow(b + h * k3, -2.0);
        k4 = (b + h * k3) * H0 * sqrt(fabs(n));
        if (j < 0 && l < 0 && m < 0 && n < 0) collapse = 1;
        K = h * (k1 + k2 * 2.0 + k3 * 2.0 + k4) / 6.0;
        b = collapse == 0 ? b + K : b - K;
    }
    return b;
}

/* friedmann_solver.cc:161-164 CALCULATE_Hubble_param */
double oracle_hubble(double H0, double Om, double Or, double Ol, double Ok, double a)
{
    return H0 * sqrt(Om * pow(a, -3) + Or * pow(a, -4) + Ol + Ok * pow(a, -2));
}
