// s1r2_lookup.cuh -- the S^1xR^2 Ewald force correction (D_rho, D_z) read from the [Nrho][Nz][2] table of the lookup build.
//
// What the reference defines (forces_cuda.cu:176-455, CPU twin ewald_space.cc:803-1048): rho = sqrt(dx^2 + dy^2) is mapped to the
// node coordinate ur = rho / drho (drho = rho_max / (Nrho - 1), nodes at 0, drho, .., clamped to the table), dz to the cell
// coordinate uz = (dz + Lz/2) / hz - 1/2 (Nz cells, periodic); EWALD_INTERPOLATION_ORDER picks nearest grid point (0), bilinear
// cloud-in-cell (2) or the quadratic triangular-shaped-cloud weights (4); D = (D_rho dx/rho, D_rho dy/rho, D_z).
//
// Here the three orders are ONE tensor-product routine: each axis yields a small stencil (1, 2 or 3 taps: indices and weights),
// the rho stencil clamps at the table edges the way each order of the reference does, the z stencil wraps, and the sum runs z outer,
// rho inner -- the order of the reference's own sums.  __host__ __device__: the CPU test tier runs it on the reference's tables.
#pragma once
#include "t3_lookup.cuh"

namespace steps {

template <typename T>
struct AxisStencil {
    int n;      // taps
    int idx[3];
    T w[3];
};

// quadratic (TSC) weights of the taps -1, 0, +1 for the offset s in [-1/2, 1/2] from the central node
template <typename T>
__host__ __device__ __forceinline__ void tsc_weights(T s, T (&w)[3]) {
    const T lo = (T)0.5 - s, hi = (T)0.5 + s;
    w[0] = (T)0.5 * lo * lo;
    w[1] = (T)0.75 - s * s;
    w[2] = (T)0.5 * hi * hi;
}

// radial axis: nodes 0 .. n-1 at spacing drho, no wrap; every order pins rho to [0, rho_max] first
template <typename T>
__host__ __device__ __forceinline__ AxisStencil<T> rho_stencil(int order, T rho, T rho_max, int n) {
    AxisStencil<T> st;
    const T drho = rho_max / (T)(n > 1 ? n - 1 : 1);
    T ur;
    if (rho < (T)0) ur = (T)0;
    else if (rho > rho_max) ur = (order == 0) ? ((drho > (T)0) ? rho_max / drho : (T)0) : (T)(n - 1);
    else ur = (order == 0 && !(drho > (T)0)) ? (T)0 : rho / drho;
    if (order == 0) {  // nearest node
        int i = (int)floor(ur + (T)0.5);
        i = i < 0 ? 0 : (i > n - 1 ? n - 1 : i);
        st.n = 1;
        st.idx[0] = i;
        st.w[0] = (T)1;
    } else if (order == 2) {  // the two nodes around ur; beyond the last interval the upper node carries everything
        int i0 = (int)floor(ur);
        T f = ur - (T)i0;
        if (i0 < 0) { i0 = 0; f = (T)0; }
        if (i0 > n - 2) { i0 = n - 2 > 0 ? n - 2 : 0; f = (T)1; }
        st.n = 2;
        st.idx[0] = i0; st.idx[1] = i0 + 1;
        st.w[0] = (T)1 - f; st.w[1] = f;
    } else {  // three nodes around the nearest one, indices pinned to the table
        const int j = (int)floor(ur + (T)0.5);
        tsc_weights<T>(ur - (T)j, st.w);
        st.n = 3;
        st.idx[0] = j - 1 < 0 ? 0 : j - 1;
        st.idx[1] = j < 0 ? 0 : (j > n - 1 ? n - 1 : j);
        st.idx[2] = j + 1 > n - 1 ? n - 1 : j + 1;
    }
    return st;
}

// periodic axis: n cells over [-Lz/2, Lz/2), value at the cell centres
template <typename T>
__host__ __device__ __forceinline__ AxisStencil<T> z_stencil(int order, T z, T Lz, int n) {
    AxisStencil<T> st;
    const T uz = (z + (T)0.5 * Lz) / (Lz / (T)n) - (T)0.5;
    if (order == 0) {
        st.n = 1;
        st.idx[0] = wrap_index((int)floor(uz + (T)0.5), n);
        st.w[0] = (T)1;
    } else if (order == 2) {
        const int i0 = (int)floor(uz);
        const T f = uz - (T)i0;
        st.n = 2;
        st.idx[0] = wrap_index(i0, n);
        st.idx[1] = wrap_index(st.idx[0] + 1, n);
        st.w[0] = (T)1 - f; st.w[1] = f;
    } else {
        const int j = (int)floor(uz + (T)0.5);
        tsc_weights<T>(uz - (T)j, st.w);
        st.n = 3;
        st.idx[0] = wrap_index(j - 1, n);
        st.idx[1] = wrap_index(j, n);
        st.idx[2] = wrap_index(j + 1, n);
    }
    return st;
}

// (D_rho, D_z) at (rho, z)
template <typename T>
__host__ __device__ __forceinline__ void s1r2_correction_rz(const T *__restrict__ tab, int order, int nrho, int nz, T rho_max, T Lz, T rho, T z,
                                                            T &Drho, T &Dz) {
    const AxisStencil<T> sr = rho_stencil<T>(order, rho, rho_max, nrho);
    const AxisStencil<T> sz = z_stencil<T>(order, z, Lz, nz);
    T d0 = 0, d1 = 0;
    // (fully unrolled with guards: the stencils stay in registers; the tap counts are uniform over a launch)
#pragma unroll
    for (int q = 0; q < 3; ++q) {
#pragma unroll
        for (int p = 0; p < 3; ++p) {
            if (q < sz.n && p < sr.n) {
                const T w = sr.w[p] * sz.w[q];
                const T *__restrict__ e = tab + ((size_t)sr.idx[p] * (size_t)nz + (size_t)sz.idx[q]) * 2u;
                d0 += w * table_ld(e);
                d1 += w * table_ld(e + 1);
            }
        }
    }
    Drho = d0;
    Dz = d1;
}

// Cartesian correction vector for the displacement (dx, dy, dz); on the axis (rho = 0) the radial part has no direction
template <typename T>
__host__ __device__ __forceinline__ void s1r2_correction(const T *__restrict__ tab, int order, int nrho, int nz, T rho_max, T Lz, T dx, T dy, T dz,
                                                         T (&D)[3]) {
    const T rho = sqrt(dx * dx + dy * dy);
    T Drho, Dz;
    s1r2_correction_rz<T>(tab, order, nrho, nz, rho_max, Lz, rho, dz, Drho, Dz);
    const T cx = (rho > 0) ? dx / rho : (T)0;
    const T cy = (rho > 0) ? dy / rho : (T)0;
    D[0] = Drho * cx;
    D[1] = Drho * cy;
    D[2] = Dz;
}

}  // namespace steps
