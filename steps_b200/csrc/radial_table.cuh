// radial_table.cuh -- GPU builder of RADIAL_FORCE_TABLE, the radial background-force factor of the S^1 x R^2 cylinder
// (SURVEY.md 8f.1).  Replaces get_cylindrical_force_table() (StePS/src/utils.cc:162-228; called at main.cc:1263-1310 with
// Lz = L/2 in the quasi-periodic mode and Lz = L * ewald_cut in the NOLOOKUP image-sum mode).
//
// table[i], i = 1..size-1: a = R i/(size-1) (a = R for the last entry),
//     table[i] = (1 / (pi a)) * integral_0^R [ f1(Y; a) + f2(Y; a) ] dY          (trapezoid rule, `accuracy` equal steps)
//     f1 = -2 ln( (sqrt(Lz^2 + R^2 + a^2 + 2 a s) + Lz) / (sqrt(Lz^2 + R^2 + a^2 - 2 a s) + Lz) ),   s = sqrt(R^2 - Y^2)
//     f2 =    ln( (R^2 + a^2 + 2 a s) / (R^2 + a^2 - 2 a s) )
// with the a >= R branch of the reference (f2 dropped, + pi R added) for the last entry, and table[0] extrapolated linearly
// from table[1], table[2] at the abscissae the reference uses (2R/size, 4R/size).  One thread per entry; the trapezoid sum
// runs in the reference's order inside the thread.
#pragma once
#include <cmath>
#include <cuda_runtime.h>

namespace steps {

__host__ __device__ inline double radial_integrand(double a, double R, double Lz, double Y, bool edge) {
    const double s = sqrt(R * R - Y * Y);
    if (edge) {
        // utils.cc:189-194 (a = R): only the f1-type term, with R^2 + a^2 = 2 R^2
        return -2.0 * log((sqrt(Lz * Lz + 2.0 * R * R + 2.0 * R * s) + Lz) / (sqrt(Lz * Lz + 2.0 * R * R - 2.0 * R * s) + Lz));
    }
    const double f1 = -2.0 * log((sqrt(Lz * Lz + R * R + a * a + 2.0 * a * s) + Lz) / (sqrt(Lz * Lz + R * R + a * a - 2.0 * a * s) + Lz));
    const double f2 = log((R * R + a * a + 2.0 * a * s) / (R * R + a * a - 2.0 * a * s));
    return f1 + f2;
}

// entry i >= 1 of the table (utils.cc:172-221)
__host__ __device__ inline double radial_table_entry(int i, double R, double Lz, int size, int accuracy) {
    const double pi = 3.14159265358979323846;
    const double step = R / (double)accuracy;
    double a = (i == size - 1) ? R : R / (double)(size - 1) * i;
    const bool edge = a >= R;
    if (edge) a = R;
    double total = 0.0;
    for (int j = 1; j <= accuracy; ++j) {
        const double prev = radial_integrand(a, R, Lz, (j - 1) * step, edge);
        const double cur = radial_integrand(a, R, Lz, j * step, edge);
        total += step * (prev + cur) * 0.5;
    }
    return edge ? (2.0 * (total + pi * R) / (2.0 * pi * R)) : (2.0 * total / (2.0 * pi * a));
}

__global__ void radial_table_kernel(double R, double Lz, int size, int accuracy, double *__restrict__ table) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < 1 || i >= size) return;
    table[i] = radial_table_entry(i, R, Lz, size, accuracy);
}

// table[0] (utils.cc:223-227): straight line through (2R/size, table[1]) and (4R/size, table[2]) evaluated at 0
inline double radial_table_origin(double R, int size, double t1, double t2) {
    const double X1 = 2.0 * R / (double)size, X2 = 4.0 * R / (double)size;
    const double A = (t2 - t1) / (X2 - X1);
    const double B = t1 - A * X1;
    return A * 0.0 + B;
}

}  // namespace steps
