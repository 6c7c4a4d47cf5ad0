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

// The periodic force on a unit mass at (rho, z) from a unit line of images along z, in three parts with their own loops, then the
// Newtonian force of the nearest image is taken away (ewald_space.cc:618-798 with dx = rho, dy = 0).  Each part keeps the
// reference's term order; the parts are separate functions so that the table builder reads as the formula in the header.
struct SlabForce {
    double radial, axial;
};

// screened Newtonian pull of the images n = -nmax .. nmax
__host__ __device__ inline SlabForce s1r2_image_part(double rho, double dz, const S1R2EwaldParams &p) {
    const double two_over_sqrtpi = 2.0 / 1.7724538509055160272981674833411;
    SlabForce f{0.0, 0.0};
    for (int n = -p.nmax; n <= p.nmax; ++n) {
        const double zn = dz + (double)n * p.Lz;
        const double d2 = rho * rho + zn * zn;
        if (d2 < 1e-18) continue;
        const double d = sqrt(d2);
        const double ad = p.alpha * d;
        const double screened = (erfc(ad) + two_over_sqrtpi * ad * exp(-ad * ad)) * (1.0 / (d2 * d));
        f.radial -= rho * screened;
        f.axial -= zn * screened;
    }
    return f;
}

// the modes m = 1 .. mmax of the singly periodic kernel and, off the axis, the zero mode
__host__ __device__ inline SlabForce s1r2_mode_part(double rho, double dz, const S1R2EwaldParams &p) {
    const double two_pi = 2.0 * 3.14159265358979323846;
    const double per_length = 1.0 / p.Lz;
    const double weight = 2.0 * per_length;
    SlabForce f{0.0, 0.0};
    for (int m = 1; m <= p.mmax; ++m) {
        const double k = (two_pi * m) * per_length;
        const double grow = s1r2_exp_erfc(k, rho, p.alpha, true);
        const double decay = s1r2_exp_erfc(k, rho, p.alpha, false);
        f.radial += weight * cos(k * dz) * (k * (grow - decay));
        f.axial -= weight * k * sin(k * dz) * (grow + decay);
    }
    if (rho > 1e-12) f.radial += -(weight / rho) * (1.0 - exp(-p.alpha * p.alpha * (rho * rho)));
    // (the reference projects the radial part on x with dx / rho = rho / rho: exactly 1 off the axis, nothing on it)
    if (!(rho > 0)) f.radial = 0.0;
    else f.radial = f.radial * rho / rho;
    return f;
}

// plain Newtonian pull of the nearest image
__host__ __device__ inline SlabForce s1r2_nearest_part(double rho, double dz) {
    const double d2 = rho * rho + dz * dz;
    if (!(d2 > 0.0)) return SlabForce{0.0, 0.0};
    const double d = sqrt(d2);
    const double inv3 = 1.0 / (d * d * d);
    return SlabForce{-rho * inv3, -dz * inv3};
}

// (D_rho, D_z) of table cell (ir, iz): nodes in rho, cell centres in z
__host__ __device__ inline void s1r2_ewald_cell(int ir, int iz, const S1R2EwaldParams &p, double &Drho, double &Dz) {
    const double node = p.rho_max / (double)(p.nrho - 1 > 1 ? p.nrho - 1 : 1);
    const double rho = (double)ir * node;
    const double dz = s1r2_wrap(((double)iz + 0.5) * (p.Lz / (double)p.nz) - 0.5 * p.Lz, p.Lz);
    const SlabForce images = s1r2_image_part(rho, dz, p);
    const SlabForce modes = s1r2_mode_part(rho, dz, p);
    const SlabForce nearest = s1r2_nearest_part(rho, dz);
    Drho = (images.radial + modes.radial) - nearest.radial;
    Dz = (images.axial + modes.axial) - nearest.axial;
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
