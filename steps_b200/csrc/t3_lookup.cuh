// t3_lookup.cuh -- the T^3 Ewald force correction D(d) read from the Ngrid^3 table by tricubic (Catmull-Rom) interpolation.
//
// What the reference defines (forces_cuda.cu:81-167, CPU twin ewald_space.cc:392-523): the table holds D at the centres of Ngrid^3
// cells of [-L/2, L/2)^3 (row-major x, y, z, component); a displacement d is mapped to the cell coordinate u = (d + L/2)/h - 1/2
// per axis (h = L/Ngrid), the four neighbours floor(u)-1 .. floor(u)+2 (periodic) are weighted with the Catmull-Rom cubic of the
// fraction t = u - floor(u), and D is the tensor-product sum over 4 x 4 x 4 neighbours.
//
// How it is evaluated here (B200): the 64 x 3 table values per pair are what bounds the T^3 kernels -- the L1 path delivers about
// 120 B/clk/SM whatever the access width or pattern (tools/ubench_loads.cu), and 8-byte loads reach only 72 B/clk.  So
//   * the kernels read ALIGNED ROW COPIES of the table: every (x, y) row is extended by three wrapped z-entries, so the 4 x 3 values
//     of the z-neighbours z0 .. z0+3 are contiguous for every z0 (no wrap arithmetic in z), and the table is stored once per
//     possible misalignment of 3 z0 reals against 16 bytes (2 copies in FP64, 4 in FP32), each shifted so that in the copy chosen by
//     z0 the 12 values start on a 16-byte boundary: a row of the stencil is six (three) 128-bit loads.  The copies keep the
//     footprint of the table itself (6.3 MB each at Ngrid = 63), which is what keeps the L1 hit rate of the gather high -- a layout
//     with one 128-byte line per (x, y, z0) window was measured and lost more to L1 misses than the wide loads won
//     (profiles/r2e_t3_zwin_sweep.txt);
//   * the contraction runs z first (one dot product of 4 per row and component, then one FMA per row and component with the
//     product of the x and y weights): 256 FMA-pipe instructions per pair instead of 272 for the weight-product-first order;
//   * weights in Horner form, cell coordinate by the reciprocal spacing, indices wrapped by compare-and-add.
// t3_correction_rowmajor() reads the caller's table as it is (any pair, exact-order kernels); t3_correction_zwin() reads the
// window copy (the action-reaction kernel).  Both are __host__ __device__: the CPU test tier runs them on the reference's table.
#pragma once
#include <cuda_runtime.h>

namespace steps {

// read-only table load: LDG.CI on the device; a plain load when the CPU test tier runs these functions on the host (tests/hostcheck)
template <typename T>
__host__ __device__ __forceinline__ T table_ld(const T *p) {
#ifdef __CUDA_ARCH__
    return __ldg(p);
#else
    return *p;
#endif
}

// aligned row copies: reals per 16 bytes = number of copies; row stride (reals) holds 3 (N + 3) values plus the largest shift
template <typename T>
__host__ __device__ constexpr int t3_copies() {
    return 16 / (int)sizeof(T);
}
template <typename T>
__host__ __device__ constexpr int t3_row_stride(int N) {
    return (3 * (N + 3) + t3_copies<T>() - 1 + t3_copies<T>() - 1) / t3_copies<T>() * t3_copies<T>();
}
// reals of one copy / of all copies
template <typename T>
__host__ __device__ constexpr size_t t3_copy_elems(int N) {
    return (size_t)N * N * t3_row_stride<T>(N);
}
template <typename T>
__host__ __device__ constexpr size_t t3_aligned_elems(int N) {
    return t3_copy_elems<T>(N) * t3_copies<T>();
}

// non-negative remainder (cell indices of displacements outside one period)
__host__ __device__ __forceinline__ int wrap_index(int i, int n) {
    const int r = i % n;
    return r < 0 ? r + n : r;
}

// Catmull-Rom weights of the neighbours -1, 0, +1, +2 for the fraction t in [0, 1), Horner form
template <typename T>
__host__ __device__ __forceinline__ void catmull_rom(T t, T (&w)[4]) {
    const T t2 = t * t;
    w[0] = (((T)(-0.5) * t + (T)1.0) * t - (T)0.5) * t;
    w[1] = ((T)1.5 * t - (T)2.5) * t2 + (T)1.0;
    w[2] = (((T)(-1.5) * t + (T)2.0) * t + (T)0.5) * t;
    w[3] = ((T)0.5 * t - (T)0.5) * t2;
}

// One axis: index of the FIRST of the four neighbours (wrapped into [0, n)) and their weights.  `u` is the cell coordinate.
template <typename T>
__host__ __device__ __forceinline__ int cubic_axis(T u, int n, T (&w)[4]) {
    const T cell = floor(u);
    catmull_rom<T>(u - cell, w);
    int first = (int)cell - 1;
    // a nearest-image displacement gives first in [-2, n-2]; anything else (a caller that did not wrap) takes the general remainder
    if (first < 0) first += n;
    if ((unsigned)first >= (unsigned)n) first = wrap_index((int)cell - 1, n);
    return first;
}

// the four wrapped indices first, first+1, .. of one axis
__host__ __device__ __forceinline__ void axis_indices(int first, int n, int (&idx)[4]) {
    int v = first;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        idx[k] = v;
        v = (v + 1 == n) ? 0 : v + 1;
    }
}

// geometry of a table: spacing as the reference forms it (L / Ngrid) and its reciprocal
struct T3Lookup {
    double L, halfL, h, inv_h;
    int N;
    const void *rowmajor;  // the caller's table [N][N][N][3]
    const void *zwin;      // aligned row copies (t3_aligned_fill), or nullptr
};

template <typename T>
__host__ __device__ __forceinline__ T3Lookup t3_lookup_make(double L, int N, const void *rowmajor, const void *zwin) {
    T3Lookup k;
    k.L = (double)(T)L;
    k.halfL = (double)((T)L * (T)0.5);
    k.h = (double)((T)L / (T)N);
    k.inv_h = (double)((T)N / (T)L);
    k.N = N;
    k.rowmajor = rowmajor;
    k.zwin = zwin;
    return k;
}

// D(d) from the caller's row-major table.  The cell coordinate is formed with the division the reference uses.
template <typename T>
__host__ __device__ __forceinline__ void t3_correction_rowmajor(const T3Lookup &k, T dx, T dy, T dz, T (&D)[3]) {
    const T halfL = (T)k.halfL, h = (T)k.h;
    const int N = k.N;
    T wx[4], wy[4], wz[4];
    int ix[4], iy[4], iz[4];
    axis_indices(cubic_axis<T>((dx + halfL) / h - (T)0.5, N, wx), N, ix);
    axis_indices(cubic_axis<T>((dy + halfL) / h - (T)0.5, N, wy), N, iy);
    axis_indices(cubic_axis<T>((dz + halfL) / h - (T)0.5, N, wz), N, iz);
    const T *__restrict__ tab = static_cast<const T *>(k.rowmajor);
    T s0 = 0, s1 = 0, s2 = 0;
#pragma unroll
    for (int a = 0; a < 4; ++a) {
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const T *__restrict__ row = tab + (size_t)(ix[a] * N + iy[b]) * (size_t)(3 * N);
            T p0 = 0, p1 = 0, p2 = 0;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const T *__restrict__ e = row + 3 * iz[c];
                p0 = fma(wz[c], table_ld(e), p0);
                p1 = fma(wz[c], table_ld(e + 1), p1);
                p2 = fma(wz[c], table_ld(e + 2), p2);
            }
            const T wab = wx[a] * wy[b];
            s0 = fma(wab, p0, s0);
            s1 = fma(wab, p1, s1);
            s2 = fma(wab, p2, s2);
        }
    }
    D[0] = s0; D[1] = s1; D[2] = s2;
}

// the 12 values of one z-window with 128-bit loads
__host__ __device__ __forceinline__ void zwin_load(const double *__restrict__ p, double (&e)[12]) {
#ifdef __CUDA_ARCH__
    const double2 *__restrict__ q = reinterpret_cast<const double2 *>(p);
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        const double2 v = __ldg(q + k);
        e[2 * k] = v.x;
        e[2 * k + 1] = v.y;
    }
#else
    for (int k = 0; k < 12; ++k) e[k] = p[k];
#endif
}
__host__ __device__ __forceinline__ void zwin_load(const float *__restrict__ p, float (&e)[12]) {
#ifdef __CUDA_ARCH__
    const float4 *__restrict__ q = reinterpret_cast<const float4 *>(p);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const float4 v = __ldg(q + k);
        e[4 * k] = v.x; e[4 * k + 1] = v.y; e[4 * k + 2] = v.z; e[4 * k + 3] = v.w;
    }
#else
    for (int k = 0; k < 12; ++k) e[k] = p[k];
#endif
}

// D(d) from the aligned row copies.  Cell coordinate by the reciprocal spacing (one rounding apart from the division: < 1e-14 of D).
template <typename T>
__host__ __device__ __forceinline__ void t3_correction_zwin(const T3Lookup &k, T dx, T dy, T dz, T (&D)[3]) {
    const T halfL = (T)k.halfL, inv_h = (T)k.inv_h;
    const int N = k.N;
    T wx[4], wy[4], wz[4];
    int ix[4], iy[4];
    axis_indices(cubic_axis<T>((dx + halfL) * inv_h - (T)0.5, N, wx), N, ix);
    axis_indices(cubic_axis<T>((dy + halfL) * inv_h - (T)0.5, N, wy), N, iy);
    const int z0 = cubic_axis<T>((dz + halfL) * inv_h - (T)0.5, N, wz);
    // the copy in which 3 z0 reals past a row start is a multiple of 16 bytes: shift = (-3 z0) mod copies
    constexpr int NC = t3_copies<T>();
    const int RS = t3_row_stride<T>(N);
    const int shift = (NC - (3 * z0) % NC) % NC;
    const T *__restrict__ win = static_cast<const T *>(k.zwin) + (size_t)shift * t3_copy_elems<T>(N) + (size_t)(shift + 3 * z0);
    T s0 = 0, s1 = 0, s2 = 0;
#pragma unroll
    for (int a = 0; a < 4; ++a) {
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            T e[12];
            zwin_load(win + (size_t)(ix[a] * N + iy[b]) * (size_t)RS, e);
            T p0 = wz[0] * e[0], p1 = wz[0] * e[1], p2 = wz[0] * e[2];
#pragma unroll
            for (int c = 1; c < 4; ++c) {
                p0 = fma(wz[c], e[3 * c], p0);
                p1 = fma(wz[c], e[3 * c + 1], p1);
                p2 = fma(wz[c], e[3 * c + 2], p2);
            }
            const T wab = wx[a] * wy[b];
            s0 = fma(wab, p0, s0);
            s1 = fma(wab, p1, s1);
            s2 = fma(wab, p2, s2);
        }
    }
    D[0] = s0; D[1] = s1; D[2] = s2;
}

// aligned row copies of a row-major table: copy c holds row (x, y) at reals [c, c + 3 (N + 3)) of its row slot, z-entries N .. N+2
// repeating 0 .. 2.  One call fills one (copy, row).
template <typename T>
__host__ __device__ __forceinline__ void t3_aligned_fill(const T *__restrict__ tab, int N, int copy, size_t row, T *__restrict__ out) {
    const int RS = t3_row_stride<T>(N);
    T *__restrict__ o = out + (size_t)copy * t3_copy_elems<T>(N) + row * (size_t)RS;
    for (int q = 0; q < copy; ++q) o[q] = (T)0;
    for (int z = 0; z < N + 3; ++z) {
        const int zs = z < N ? z : z - N;
        for (int c = 0; c < 3; ++c) o[copy + 3 * z + c] = tab[(row * (size_t)N + (size_t)zs) * 3 + c];
    }
    for (int q = copy + 3 * (N + 3); q < RS; ++q) o[q] = (T)0;
}

#ifdef __CUDACC__
template <typename T>
__global__ void t3_aligned_kernel(const T *__restrict__ tab, int N, T *__restrict__ out) {
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t rows = (size_t)N * N;
    if (t < rows * t3_copies<T>()) t3_aligned_fill<T>(tab, N, (int)(t / rows), t % rows, out);
}
#endif

}  // namespace steps
