// cone_select.cuh -- redshift-cone output: which particles enter the light cone at an output time, selected on
// the device, formatted on the host.
//
// What the reference does (write_redshift_cone, inputoutput.cc:314-405, ASCII branch; called from main.cc:1785 when a radial bin
// of the cone is reached): on rank 0, over all N particles on the host, D_i = |x_i| (comoving distance from the observer at the
// origin); a particle with limits[z_index + 1] <= D_i that is not yet in the cone is appended to <OUT_DIR>redshift_cone.dat
// (x, v, M, D, z, index) and flagged IN_CONE for the rest of the run; at the end of the run (ALL != 0) every particle still outside is
// written with the redshift of the shell its distance falls into.
// With the particle state resident in HBM that scan would need x and v of all N particles on the host at every output.  Here the
// flags live on the device, one kernel per engine scans its own rows and compacts the selected particles (warp-aggregated append),
// and only those rows travel; the host puts them in index order (the reference's order) and formats the same bytes.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <string>
#include <vector>
#include <cuda_runtime.h>

namespace steps {

constexpr int CONE_ROW = 8;                        // reals per compacted row: x y z vx vy vz M (one of padding)
constexpr double CONE_UNIT_V = 20.738652969925447;  // km/s per internal velocity unit (global_variables.h:18)

// |x| as the reference forms it: products and sums in REAL, rounded one by one (no contraction), square root in REAL.  The order of the
// two additions is the one the reference's build (g++ -O3 -ffast-math, Template-LinuxGCC-Makefile:81) ends up with, found by matching its
// printed digits on 20 000 particles in both precisions: (x^2 + y^2) + z^2 in the radial-bin branch -- the one whose value decides the
// selection -- and x^2 + (y^2 + z^2) in the end-of-run branch (`tail` = true; the value is only printed there).
__host__ __device__ __forceinline__ double cone_distance(double x, double y, double z, bool tail = false) {
#ifdef __CUDA_ARCH__
    const double a = __dmul_rn(x, x), b = __dmul_rn(y, y), c = __dmul_rn(z, z);
    return sqrt(tail ? __dadd_rn(a, __dadd_rn(b, c)) : __dadd_rn(__dadd_rn(a, b), c));
#else
    volatile double a = x * x, b = y * y, c = z * z;
    volatile double s = tail ? b + c : a + b;
    return std::sqrt(tail ? a + s : s + c);
#endif
}
__host__ __device__ __forceinline__ float cone_distance(float x, float y, float z, bool tail = false) {
#ifdef __CUDA_ARCH__
    const float a = __fmul_rn(x, x), b = __fmul_rn(y, y), c = __fmul_rn(z, z);
    return sqrtf(tail ? __fadd_rn(a, __fadd_rn(b, c)) : __fadd_rn(__fadd_rn(a, b), c));
#else
    volatile float a = x * x, b = y * y, c = z * z;
    volatile float s = tail ? b + c : a + b;
    return std::sqrt(tail ? a + s : s + c);
#endif
}

#ifdef __CUDACC__
// rows [lo, hi) of one engine: select, flag, append.  The order of the appended rows is arbitrary (the host sorts by index).
template <typename T>
__global__ void cone_select_kernel(const T *__restrict__ x, const T *__restrict__ v, const T *__restrict__ m, unsigned char *__restrict__ in_cone,
                                   int lo, int hi, double r_min, int all, T *__restrict__ rows, int *__restrict__ idx, int *__restrict__ counter) {
    const int i = lo + blockIdx.x * blockDim.x + threadIdx.x;
    bool take = false;
    T xi = 0, yi = 0, zi = 0;
    if (i < hi && !in_cone[i]) {
        xi = x[3 * (size_t)i];
        yi = x[3 * (size_t)i + 1];
        zi = x[3 * (size_t)i + 2];
        take = all != 0 || r_min <= (double)cone_distance(xi, yi, zi);
    }
    const unsigned ballot = __ballot_sync(0xffffffffu, take);
    if (ballot == 0) return;
    const int lane = threadIdx.x & 31;
    int base = 0;
    if (lane == __ffs(ballot) - 1) base = atomicAdd(counter, __popc(ballot));
    base = __shfl_sync(0xffffffffu, base, __ffs(ballot) - 1);
    if (take) {
        const size_t o = (size_t)(base + __popc(ballot & ((1u << lane) - 1u)));
        T *__restrict__ r = rows + o * CONE_ROW;
        r[0] = xi; r[1] = yi; r[2] = zi;
        r[3] = v[3 * (size_t)i]; r[4] = v[3 * (size_t)i + 1]; r[5] = v[3 * (size_t)i + 2];
        r[6] = m[i]; r[7] = 0;
        idx[o] = i;
        in_cone[i] = 1;
    }
}
#endif

// The lines of write_redshift_cone for `count` selected particles given in ascending index order.  all == 0: every line carries
// out_list[z_index] and the distance itself; all != 0: the redshift of the first radial bin after z_index whose limit the distance
// reaches -- with the reference's loop `j = z_index; while (j++) { if (limits[j] <= D) {...; break;} }`, i.e. no search at all when
// z_index == 0 and a z that carries over from the previous particle when no bin matches -- and the distance times H0_dimless.
template <typename T>
static std::string cone_format(const T *rows, const int *index, int count, double h0_dimless_d, int all, const double *limits, int n_limits,
                               const double *out_list, int z_index) {
    std::string out;
    out.reserve((size_t)count * 230);
    char buf[640];
    const T h0 = (T)h0_dimless_d;
    double z_write = out_list[z_index];
    for (int k = 0; k < count; ++k) {
        const T *r = rows + (size_t)k * CONE_ROW;
        const double D = (double)cone_distance(r[0], r[1], r[2], all != 0);
        int len = 0;
        for (int c = 0; c < 3; ++c) len += snprintf(buf + len, sizeof(buf) - len, "%.16f\t", (double)(T)(r[c] * h0));
        for (int c = 0; c < 3; ++c) len += snprintf(buf + len, sizeof(buf) - len, "%.16f\t", (double)r[3 + c] * CONE_UNIT_V);
        if (all == 0) {
            len += snprintf(buf + len, sizeof(buf) - len, "%.16f\t%.16f\t%.16f\t%i\n", (double)(T)(r[6] * h0), D, out_list[z_index], index[k]);
        } else {
            int j = z_index;
            while (j++) {
                if (j >= n_limits) break;  // (the reference would read past its array here)
                if (limits[j] <= D) {
                    z_write = out_list[j];
                    break;
                }
            }
            len += snprintf(buf + len, sizeof(buf) - len, "%.16f\t%.16f\t%.16f\t%i\n", (double)(T)(r[6] * h0), D * (double)h0, z_write, index[k]);
        }
        out.append(buf, (size_t)len);
    }
    return out;
}

}  // namespace steps
