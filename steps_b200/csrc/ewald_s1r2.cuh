// ewald_s1r2.cuh -- GPU builder of the S^1 x R^2 (rho, z) Ewald force-correction lookup table.
//
// Replaces calculate_S1R2ewald_correction_table() + S1R2ewald_force_pair() + exp_erfc_product() + fast_erfcx() + wrap_dz()
// (StePS/src/ewald_space.cc:754-798, :618-752, :563-593, :537-561, :597-608; set up by main.cc:562-605), Ewald variant
// (the default build: PERIODIC_Z_RSPACELOOKUP not defined).  For cell (ir, iz): rho = ir * rho_max/(Nrho-1),
// z = (iz + 1/2) Lz/Nz - Lz/2,
//     F_periodic = real-space image sum over n = -nmax..nmax with the erfc/Gaussian bracket
//                + k-space sum over m = 1..mmax of the singly periodic kernel B(rho) = e^{k rho} erfc(k/2a + a rho) + e^{-k rho} erfc(k/2a - a rho)
//                + zero mode  -2/(L rho) (1 - e^{-a^2 rho^2})                      (Tornberg 2015; Shamshirgar & Tornberg 2017)
//     D = F_periodic - F_newton(nearest image),   table[ir][iz] = (D_rho, D_z).
// One thread per cell, terms in the reference's order, so entries differ from the reference's only in the last bits of
// erfc/exp/sin/cos (CUDA math library vs glibc).
#pragma once
#include <cmath>
#include <cuda_runtime.h>

namespace steps {

struct S1R2EwaldParams {
    int nrho, nz, nmax, mmax;
    double rho_max, Lz, alpha;
};

// exp(x^2) erfc(x) as the reference evaluates it (ewald_space.cc:537-561: the 5-term rational fit for 0 < x <= 1e10)
__host__ __device__ inline double s1r2_erfcx_fit(double x) {
    if (x < 0) return 2.0 * exp(x * x) - s1r2_erfcx_fit(-x);
    if (x == 0) return 1.0;
    if (x > 1e10) return 0.564189583547756286 / x;
    const double t = 1.0 / (1.0 + 0.3275911 * x);
    return t * (0.254829592 + t * (-0.284496736 + t * (1.421413741 + t * (-1.453152027 + t * 1.061405429))));
}

// exp(+-k rho) erfc(k/2a +- a rho) without overflow (ewald_space.cc:563-593)
__host__ __device__ inline double s1r2_exp_erfc(double km, double rho, double alpha, bool plus) {
    const double ka = km / (2.0 * alpha);
    const double ar = alpha * rho;
    const double arg = plus ? (ka + ar) : (ka - ar);
    const double common = -(ka * ka + ar * ar);
    if (arg > 5.0) return exp(common) * s1r2_erfcx_fit(arg);
    if (arg < -5.0) return 2.0 * exp(-km * rho) - exp(common) * s1r2_erfcx_fit(-arg);
    return exp(plus ? (km * rho) : (-km * rho)) * erfc(arg);
}

// nearest image in z, (-Lz/2, Lz/2]   (ewald_space.cc:597-608)
__host__ __device__ inline double s1r2_wrap(double dz, double Lz) {
    const double half = 0.5 * Lz;
    if (dz > half) dz -= Lz * floor((dz + half) / Lz);
    if (dz <= -half) dz -= Lz * floor((dz - half) / Lz);
    return dz;
}

// (D_rho, D_z) of one table cell (ewald_space.cc:754-798 with dx = rho, dy = 0)
__host__ __device__ inline void s1r2_ewald_cell(int ir, int iz, const S1R2EwaldParams &p, double &Drho, double &Dz) {
    const double pi = 3.14159265358979323846;
    const double sqrtpi = 1.7724538509055160272981674833411;
    const double drho = p.rho_max / (double)(p.nrho - 1 > 1 ? p.nrho - 1 : 1);
    const double dzc = p.Lz / (double)p.nz;
    const double rho = (double)ir * drho;
    const double z = ((double)iz + 0.5) * dzc - 0.5 * p.Lz;
    const double Lz = p.Lz, alpha = p.alpha;
    const double dz = s1r2_wrap(z, Lz);
    const double rho2 = rho * rho;
    // real space
    double Fxr = 0.0, Fzr = 0.0;
    for (int n = -p.nmax; n <= p.nmax; ++n) {
        const double dzn = dz + (double)n * Lz;
        const double r2 = rho2 + dzn * dzn;
        if (r2 < 1e-18) continue;
        const double r = sqrt(r2);
        const double invr3 = 1.0 / (r2 * r);
        const double ar = alpha * r;
        const double coeff = (erfc(ar) + (2.0 / sqrtpi) * ar * exp(-ar * ar)) * invr3;
        Fxr -= rho * coeff;
        Fzr -= dzn * coeff;
    }
    // k space
    double Frk = 0.0, Fzk = 0.0;
    const double invL = 1.0 / Lz;
    for (int m = 1; m <= p.mmax; ++m) {
        const double km = (2.0 * pi * m) * invL;
        const double tp = s1r2_exp_erfc(km, rho, alpha, true);
        const double tm = s1r2_exp_erfc(km, rho, alpha, false);
        const double B = tp + tm;
        const double dBdrho = km * (tp - tm);
        Frk += (2.0 * invL) * cos(km * dz) * dBdrho;
        Fzk -= (2.0 * invL) * km * sin(km * dz) * B;
    }
    if (rho > 1e-12) Frk += -(2.0 * invL / rho) * (1.0 - exp(-alpha * alpha * rho2));
    const double Fxk = (rho > 0) ? (Frk * rho / rho) : 0.0;
    const double Frho_periodic = Fxr + Fxk;
    const double Fz_periodic = Fzr + Fzk;
    // nearest-image Newton
    const double dzw = s1r2_wrap(z, Lz);
    const double r2 = rho * rho + dzw * dzw;
    double Frho_newt = 0.0, Fz_newt = 0.0;
    if (r2 > 0.0) {
        const double r = sqrt(r2);
        const double invr3 = 1.0 / (r * r * r);
        Frho_newt = -rho * invr3;
        Fz_newt = -dzw * invr3;
    }
    Drho = Frho_periodic - Frho_newt;
    Dz = Fz_periodic - Fz_newt;
}

__global__ void s1r2_ewald_table_kernel(const S1R2EwaldParams p, double *__restrict__ table) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= p.nrho * p.nz) return;
    const int ir = t / p.nz, iz = t - ir * p.nz;
    double a, b;
    s1r2_ewald_cell(ir, iz, p, a, b);
    table[2 * (size_t)t] = a;
    table[2 * (size_t)t + 1] = b;
}

}  // namespace steps
