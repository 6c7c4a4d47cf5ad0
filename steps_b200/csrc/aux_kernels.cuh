// aux_kernels.cuh -- the O(N) kernels around the pair kernels: j-record packing, deterministic
// chunk reduction + background term, and the device-resident KDK integrator (step.cc:100-312).
#pragma once
#include "pair_generic.cuh"
#include "pair_r3_f32.cuh"
#include "pair_s1r2.cuh"
#include "pair_r3_sym.cuh"
#include "pair_r3_sym_f32.cuh"
#include "pair_s1r2_sym.cuh"
#include "pair_generic_sym.cuh"

namespace steps {

// max softening length per j-tile (softening is constant over a run: computed once per upload)
template <typename T>
__global__ void tile_smax_kernel(const T *__restrict__ s, int n, int tj, int n_tiles, T *__restrict__ smax_tile) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_tiles) return;
    T m = 0;
    const int j0 = t * tj;
    const int j1 = min(n, j0 + tj);
    for (int j = j0; j < j1; ++j) m = fmax(m, s[j]);
    smax_tile[t] = m;
}

// AoS positions + masses + softenings -> staged j-records (padded to whole tiles with massless,
// far-away records so that the tuned kernels need no bounds checks in the pair loop), plus one
// TileInfo64 per tile: bounding box and radial range of the tile's REAL particles, which lets a warp
// prove "every pair of my i-particles with this tile is farther apart than any s_i + s_j" and take the
// check-free pair loop.  One CTA of TJ threads per tile.
template <int TJ_>
__global__ void __launch_bounds__(TJ_) pack_kernel_f64(const double *__restrict__ x, const double *__restrict__ m, const double *__restrict__ s,
                                                        const double *__restrict__ smax_tile, JRec64 *__restrict__ out,
                                                        TileInfo64 *__restrict__ tinfo, int n, double zbox, int *__restrict__ z_outside, int planar) {
    const int t = blockIdx.x;
    const int j = t * TJ_ + threadIdx.x;
    // S^1xR^2: note any z outside [0, L) (the tuned image-sum kernel assumes |dz| < L, pair_s1r2.cuh)
    if (z_outside && j < n) {
        const double zz = x[3 * (size_t)j + 2];
        if (!(zz >= 0.0 && zz < zbox)) atomicOr(z_outside, 1);
    }
    JRec64 r;
    const double sm = smax_tile[t];
    double lo[3], hi[3], rlo, rhi;
    if (j < n) {
        r.x = x[3 * (size_t)j]; r.y = x[3 * (size_t)j + 1]; r.z = x[3 * (size_t)j + 2];
        r.m = m[j]; r.m15 = 1.5 * r.m; r.m1875 = 1.875 * r.m; r.s = s[j]; r.smax = sm;
        lo[0] = hi[0] = r.x; lo[1] = hi[1] = r.y; lo[2] = hi[2] = r.z;
        rlo = rhi = sqrt(r.x * r.x + r.y * r.y + r.z * r.z);
        if (planar) {  // S^1xR^2: bounds in the x,y plane only (z is periodic), cylindrical radius
            lo[2] = hi[2] = 0.0;
            rlo = rhi = sqrt(r.x * r.x + r.y * r.y);
        }
    } else {
        r.x = r.y = r.z = 1.0e20; r.m = r.m15 = r.m1875 = 0.0; r.s = 0.0; r.smax = sm;
        lo[0] = lo[1] = lo[2] = rlo = 1.0e300;
        hi[0] = hi[1] = hi[2] = -1.0e300;
        rhi = 0.0;
    }
    out[j] = r;
    // block reduction: 7 mins/maxes over TJ_ threads
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            lo[k] = fmin(lo[k], __shfl_xor_sync(0xffffffffu, lo[k], o));
            hi[k] = fmax(hi[k], __shfl_xor_sync(0xffffffffu, hi[k], o));
        }
        rlo = fmin(rlo, __shfl_xor_sync(0xffffffffu, rlo, o));
        rhi = fmax(rhi, __shfl_xor_sync(0xffffffffu, rhi, o));
    }
    __shared__ double red[TJ_ / 32][8];
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) {
        red[w][0] = lo[0]; red[w][1] = lo[1]; red[w][2] = lo[2];
        red[w][3] = hi[0]; red[w][4] = hi[1]; red[w][5] = hi[2];
        red[w][6] = rlo; red[w][7] = rhi;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        TileInfo64 ti;
        ti.lo[0] = red[0][0]; ti.lo[1] = red[0][1]; ti.lo[2] = red[0][2];
        ti.hi[0] = red[0][3]; ti.hi[1] = red[0][4]; ti.hi[2] = red[0][5];
        ti.rlo = red[0][6]; ti.rhi = red[0][7];
        for (int q = 1; q < TJ_ / 32; ++q) {
            for (int k = 0; k < 3; ++k) {
                ti.lo[k] = fmin(ti.lo[k], red[q][k]);
                ti.hi[k] = fmax(ti.hi[k], red[q][3 + k]);
            }
            ti.rlo = fmin(ti.rlo, red[q][6]);
            ti.rhi = fmax(ti.rhi, red[q][7]);
        }
        tinfo[t] = ti;
    }
}

template <int TJ_>
__global__ void __launch_bounds__(TJ_) pack_kernel_f32(const float *__restrict__ x, const float *__restrict__ m, const float *__restrict__ s,
                                                        const float *__restrict__ smax_tile, JRec32 *__restrict__ out,
                                                        TileInfo32 *__restrict__ tinfo, int n) {
    const int t = blockIdx.x;
    const int j = t * TJ_ + threadIdx.x;
    JRec32 r;
    const float sm = smax_tile[t];
    float lo[3], hi[3], rlo, rhi;
    if (j < n) {
        r.x = x[3 * (size_t)j]; r.y = x[3 * (size_t)j + 1]; r.z = x[3 * (size_t)j + 2];
        r.m = m[j]; r.s = s[j]; r.smax = sm;
        lo[0] = hi[0] = r.x; lo[1] = hi[1] = r.y; lo[2] = hi[2] = r.z;
        // bounds must be conservative: round the radius down for rlo and up for rhi
        const float r2 = r.x * r.x + r.y * r.y + r.z * r.z;
        rlo = __fsqrt_rd(r2) * 0.999999f;
        rhi = __fsqrt_ru(r2) * 1.000001f;
    } else {
        r.x = r.y = r.z = 1.0e15f; r.m = 0.f; r.s = 0.f; r.smax = sm;
        lo[0] = lo[1] = lo[2] = rlo = 3.0e38f;
        hi[0] = hi[1] = hi[2] = -3.0e38f;
        rhi = 0.f;
    }
    r.p0 = r.p1 = 0.f;
    out[j] = r;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            lo[k] = fminf(lo[k], __shfl_xor_sync(0xffffffffu, lo[k], o));
            hi[k] = fmaxf(hi[k], __shfl_xor_sync(0xffffffffu, hi[k], o));
        }
        rlo = fminf(rlo, __shfl_xor_sync(0xffffffffu, rlo, o));
        rhi = fmaxf(rhi, __shfl_xor_sync(0xffffffffu, rhi, o));
    }
    __shared__ float red[TJ_ / 32][8];
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) {
        red[w][0] = lo[0]; red[w][1] = lo[1]; red[w][2] = lo[2];
        red[w][3] = hi[0]; red[w][4] = hi[1]; red[w][5] = hi[2];
        red[w][6] = rlo; red[w][7] = rhi;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        TileInfo32 ti;
        ti.lo[0] = red[0][0]; ti.lo[1] = red[0][1]; ti.lo[2] = red[0][2];
        ti.hi[0] = red[0][3]; ti.hi[1] = red[0][4]; ti.hi[2] = red[0][5];
        ti.rlo = red[0][6]; ti.rhi = red[0][7];
        for (int q = 1; q < TJ_ / 32; ++q) {
            for (int k = 0; k < 3; ++k) {
                ti.lo[k] = fminf(ti.lo[k], red[q][k]);
                ti.hi[k] = fmaxf(ti.hi[k], red[q][3 + k]);
            }
            ti.rlo = fminf(ti.rlo, red[q][6]);
            ti.rhi = fmaxf(ti.rhi, red[q][7]);
        }
        tinfo[t] = ti;
    }
}

// linear interpolation of RADIAL_FORCE_TABLE, verbatim arithmetic of forces_cuda.cu:41-70
template <typename T>
__device__ __forceinline__ T cylindrical_force_correction(T r, T R, const T *__restrict__ table, int size) {
    const T step = R / (T)size;
    const int i = (int)floor(r / R * (size - 1));
    T corr = table[size - 1];
    if (i < size - 1) {
        const T X1 = step * i, Y1 = table[i], X2 = step * (i + 1), Y2 = table[i + 1];
        const T A = (Y2 - Y1) / (X2 - X1);
        const T B = Y1 - A * X1;
        corr = A * r + B;
    }
    return corr;
}

// F_i = sum over j-chunks (fixed order c = 0..n_chunks-1: deterministic) + background term.
// Writes AoS F at GLOBAL particle index (id_min + il).  fsym (optional): the j-side sums of the action-reaction
// kernel (pair_r3_sym.cuh), SoA [3][fsym_stride] by global particle index, accumulated with the sign of d = x_j - x_i
// as seen from the OTHER particle, hence subtracted here.
template <typename T>
__global__ void reduce_kernel(const T *__restrict__ fpart, int n_chunks, int fstride, int n_i, int id_min,
                              const T *__restrict__ x, T *__restrict__ F, const TopoParams tp, const T *__restrict__ fsym,
                              size_t fsym_stride, const unsigned long long *__restrict__ cmask = nullptr, int mask_words = 0,
                              int ib_size = 1) {
    const int il = blockIdx.x * blockDim.x + threadIdx.x;
    if (il >= n_i) return;
    T fx = 0, fy = 0, fz = 0;
    // action-reaction launch: only the chunks in which one of the i-block's tile ranges has tiles hold a partial sum for it (one bit
    // per chunk, built by the host from the rules; a block's ranges can leave whole chunks out in between)
    const unsigned long long *__restrict__ mk = cmask ? cmask + (size_t)(il / ib_size) * mask_words : nullptr;
    for (int c = 0; c < n_chunks; ++c) {
        if (mk && !((mk[c >> 6] >> (c & 63)) & 1ull)) continue;
        const T *__restrict__ p = fpart + (size_t)c * 3 * fstride;
        fx += p[il];
        fy += p[fstride + il];
        fz += p[2 * (size_t)fstride + il];
    }
    const size_t i = (size_t)id_min + il;
    if (fsym) {
        fx -= fsym[i];
        fy -= fsym[fsym_stride + i];
        fz -= fsym[2 * fsym_stride + i];
    }
    const T xi = x[3 * i], yi = x[3 * i + 1], zi = x[3 * i + 2];
    if (tp.bg_mode != 0) {
        const T B = (T)tp.bg_coeff;
        if (tp.topology == 0) {  // forces_cuda.cu:545-556
            fx += B * xi; fy += B * yi; fz += B * zi;
        } else if (tp.topology == 2 || tp.topology == 3) {
            // x,y only.  Radial table factor in the NOLOOKUP build (forces_cuda.cu:689-701, :747-753) and in
            // the quasi-periodic branch of the lookup build (:799-805); plain term for the lookup build
            // with IS_PERIODIC>=2 (:847-851).  Non-comoving: DE*x without table.
            T corr = (T)1;
            if (tp.bg_mode == 1 && (tp.topology == 3 || tp.is_periodic == 1)) {
                const T rxy = sqrt(xi * xi + yi * yi);
                corr = cylindrical_force_correction<T>(rxy, (T)tp.Rsim, static_cast<const T *>(tp.radial), tp.radial_size);
                fx += B * xi * corr;
                fy += B * yi * corr;
            } else {
                fx += B * xi;
                fy += B * yi;
            }
        }
    }
    F[3 * i] = fx; F[3 * i + 1] = fy; F[3 * i + 2] = fz;
}

// ---------------------------------------------------------------- KDK (step.cc:128-181, :254-269)
struct KdkScalars {
    double a3inv;   // (REAL)pow(a,-3.0)
    double twoH;    // 2.0*(REAL)Hubble_param
    double hhalf;   // (REAL)(h/2.0)
    double h;       // (REAL)h
    double L;
    double G;       // +1 (or -1 under GLASS_MAKING, global_variables.h:20-24)
    int topology;
};

template <typename T>
__device__ __forceinline__ T wrap_box(T xv, T L) {  // step.cc:151-180
    if (xv < 0) xv = xv + L;
    else if (xv >= L) xv = xv - L;
    return xv;
}

// first half kick + drift + wrap, for i in [lo, hi)
template <typename T>
__global__ void kick_drift_kernel(T *__restrict__ x, T *__restrict__ v, const T *__restrict__ F, int lo, int hi, KdkScalars k) {
    const int i = lo + blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= hi) return;
    const T L = (T)k.L;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const size_t q = 3 * (size_t)i + c;
        // reference expression is evaluated in double (G and 2.0 are double literals), then stored to REAL
        const T acc = (T)(k.G * (double)F[q] * (double)(T)k.a3inv - k.twoH * (double)v[q]);
        T vv = v[q] + acc * (T)k.hhalf;
        v[q] = vv;
        T xx = x[q] + vv * (T)k.h;
        if (k.topology == 1 || ((k.topology == 2 || k.topology == 3) && c == 2)) xx = wrap_box<T>(xx, L);
        x[q] = xx;
    }
}

__device__ __forceinline__ void atomic_max_nonneg(double *addr, double val) {
    // non-negative doubles order like their bit patterns
    atomicMax(reinterpret_cast<unsigned long long *>(addr), (unsigned long long)__double_as_longlong(val));
}

// second half kick (do_kick=1) + errmax = max_i |acc_i| / s_i   (step.cc:256-269; calculate_init_h :74-86
// with do_kick=0, which also wraps positions first, :42-71)
template <typename T>
__global__ void kick_errmax_kernel(T *__restrict__ x, T *__restrict__ v, const T *__restrict__ F, const T *__restrict__ soft,
                                   int lo, int hi, KdkScalars k, int do_kick, int do_wrap, double *__restrict__ errmax) {
    const int i = lo + blockIdx.x * blockDim.x + threadIdx.x;
    T err = 0;
    if (i < hi) {
        T acc2 = 0;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const size_t q = 3 * (size_t)i + c;
            if (do_wrap && (k.topology == 1 || ((k.topology == 2 || k.topology == 3) && c == 2))) {
                T xx = x[q];
                if (xx < 0) xx = xx + (T)k.L;
                if (xx >= (T)k.L) xx = xx - (T)k.L;
                x[q] = xx;
            }
            const T acc = (T)(k.G * (double)F[q] * (double)(T)k.a3inv - k.twoH * (double)v[q]);
            if (do_kick) v[q] = v[q] + acc * (T)k.hhalf;
            acc2 += acc * acc;
        }
        err = sqrt(acc2) / soft[i];
    }
    // block max
    double e = (double)err;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) e = fmax(e, __shfl_xor_sync(0xffffffffu, e, o));
    __shared__ double wmax[32];
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) wmax[w] = e;
    __syncthreads();
    if (w == 0) {
        e = (l < (blockDim.x >> 5)) ? wmax[l] : 0.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) e = fmax(e, __shfl_xor_sync(0xffffffffu, e, o));
        if (l == 0) atomic_max_nonneg(errmax, e);
    }
}

// FMA-pipe microbenchmark: `iters` dependent-chain-free FMAs per thread on 8 accumulators
template <typename T>
__global__ void fma_peak_kernel(T *out, int iters, T a, T b, int pattern) {
    if (pattern == 0) {  // c = fma(c, a, b): one varying register + two shared invariants (relies on the operand-reuse cache)
        T c0 = threadIdx.x, c1 = c0 + 1, c2 = c0 + 2, c3 = c0 + 3, c4 = c0 + 4, c5 = c0 + 5, c6 = c0 + 6, c7 = c0 + 7;
        for (int i = 0; i < iters; ++i) {
            c0 = fma(c0, a, b); c1 = fma(c1, a, b); c2 = fma(c2, a, b); c3 = fma(c3, a, b);
            c4 = fma(c4, a, b); c5 = fma(c5, a, b); c6 = fma(c6, a, b); c7 = fma(c7, a, b);
        }
        out[blockIdx.x * blockDim.x + threadIdx.x] = c0 + c1 + c2 + c3 + c4 + c5 + c6 + c7;
        return;
    }
    // c_k = fma(y_k, y_k, c_k): two distinct source registers per FMA.  On sm_100 an FP64 instruction occupies the
    // pipe for max(2, distinct non-reused 64-bit source registers) cycles per warp (tools/ubench2_fp64.cu), so this
    // form issues at the pipe's real rate (2.04 cyc) where c = fma(c, a, b) with shared invariants measured 2.2-2.27.
    T c0 = threadIdx.x, c1 = c0 + 1, c2 = c0 + 2, c3 = c0 + 3, c4 = c0 + 4, c5 = c0 + 5, c6 = c0 + 6, c7 = c0 + 7;
    const T t = (T)threadIdx.x * b;
    const T y0 = a + t, y1 = a + 2 * t, y2 = a + 3 * t, y3 = a + 4 * t, y4 = a + 5 * t, y5 = a + 6 * t, y6 = a + 7 * t, y7 = a + 8 * t;
    for (int i = 0; i < iters; ++i) {
        c0 = fma(y0, y0, c0); c1 = fma(y1, y1, c1); c2 = fma(y2, y2, c2); c3 = fma(y3, y3, c3);
        c4 = fma(y4, y4, c4); c5 = fma(y5, y5, c5); c6 = fma(y6, y6, c6); c7 = fma(y7, y7, c7);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = c0 + c1 + c2 + c3 + c4 + c5 + c6 + c7;
}

}  // namespace steps
