// pair_generic.cuh -- exact-branch pair kernels for every topology and both precisions.
//
// These follow the per-pair arithmetic of the reference CUDA kernels operation by operation
// (forces_cuda.cu:522-563 R^3, :567-645 T^3, :652-761 S^1xR^2 NOLOOKUP, :764-864 S^1xR^2 lookup,
// device helpers :41-70, :81-167, :188-455, :466-519) inside the B200 work decomposition of
// pair_r3.cuh: (i-block x j-chunk) CTAs, TMA-staged j tiles, register-blocked i-particles,
// deterministic chunk reduction.  The tuned R^3 kernels in pair_r3.cuh replace the R^3
// instantiations on the hot path; these remain the path for T^3 and S^1xR^2.
#pragma once
#include "pair_r3.cuh"

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

template <typename T> struct JRecOf;
template <> struct JRecOf<double> { using type = JRec64; };
template <> struct JRecOf<float> { using type = JRec32; };

// topology parameters shared by the pair kernels and the reduce/background kernel
struct TopoParams {
    int topology;       // STEPS_TOPO_*
    int is_periodic;
    int order;          // S^1xR^2 lookup interpolation order (0,2,4)
    int dim0, dim1;     // table dims
    int radial_size;
    int ewald_max;      // IS_PERIODIC+1                      (forces_cuda.cu:659)
    int bg_mode;        // 0 none, 1 mass_in_unit_sphere (comoving), 2 DE (non-comoving)
    double L, Rsim, ewald_cut, rho_max;
    double bg_coeff;    // mass_in_unit_sphere or DE
    const void *table;  // device
    const void *radial; // device
};

// ---------------------------------------------------------------- T^3 tricubic (forces_cuda.cu:87-167)
__host__ __device__ __forceinline__ int imodp(int i, int n) {
    int r = i % n;
    return (r < 0) ? (r + n) : r;
}

template <typename T>
__host__ __device__ __forceinline__ void map_to_centered_grid(T r, T L, int Ngrid, int &i0, T &fx) {
    const T grid_spacing = L / (T)Ngrid;
    const T u = (r + L * (T)0.5) / grid_spacing - (T)0.5;
    const T uf = floor(u);
    i0 = imodp((int)uf, Ngrid);
    fx = (T)(u - uf);
}

template <typename T>
__host__ __device__ __forceinline__ void cubic_weights(T t, T w[4]) {
    const T t2 = t * t;
    const T t3 = t * t2;
    w[0] = (T)(-0.5) * t3 + t2 - (T)0.5 * t;
    w[1] = (T)1.5 * t3 - (T)2.5 * t2 + (T)1.0;
    w[2] = (T)(-1.5) * t3 + (T)2.0 * t2 + (T)0.5 * t;
    w[3] = (T)0.5 * t3 - (T)0.5 * t2;
}

template <typename T>
__host__ __device__ __forceinline__ void t3_interpolate(int Ngrid, T L, const T *__restrict__ table, T dx, T dy, T dz, T D[3]) {
    int ix0, iy0, iz0;
    T fx, fy, fz;
    map_to_centered_grid(dx, L, Ngrid, ix0, fx);
    map_to_centered_grid(dy, L, Ngrid, iy0, fy);
    map_to_centered_grid(dz, L, Ngrid, iz0, fz);
    T wx[4], wy[4], wz[4];
    cubic_weights(fx, wx);
    cubic_weights(fy, wy);
    cubic_weights(fz, wz);
    int izs[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) izs[k] = imodp(iz0 - 1 + k, Ngrid);
    T s0 = 0, s1 = 0, s2 = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int ix = imodp(ix0 - 1 + i, Ngrid);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int iy = imodp(iy0 - 1 + j, Ngrid);
            const T w_xy = wx[i] * wy[j];
            const size_t row = (size_t)(ix * Ngrid + iy) * Ngrid;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const T w_xyz = w_xy * wz[k];
                const T *__restrict__ e = table + (row + izs[k]) * 3u;
                s0 += w_xyz * table_ld(e + 0);
                s1 += w_xyz * table_ld(e + 1);
                s2 += w_xyz * table_ld(e + 2);
            }
        }
    }
    D[0] = s0; D[1] = s1; D[2] = s2;
}

// ---------------------------------------------------------------- S^1xR^2 (rho,z) table (forces_cuda.cu:176-455)
template <typename T>
__host__ __device__ __forceinline__ void s1r2_ngp(const T *__restrict__ Tb, int Nrho, int Nz, T rho_max, T Lz, T rho, T z, T D[2]) {
    T rc = rho;
    if (rho < (T)0) rc = (T)0;
    if (rho > rho_max) rc = rho_max;
    const T half = (T)0.5 * Lz;
    const T drho = rho_max / (T)max(1, Nrho - 1);
    const T dz = Lz / (T)Nz;
    const T ur = (drho > (T)0) ? (rc / drho) : (T)0;
    const T uz = (z + half) / dz - (T)0.5;
    int ir = (int)floor(ur + (T)0.5);
    int iz = (int)floor(uz + (T)0.5);
    if (ir < 0) ir = 0;
    if (ir > Nrho - 1) ir = Nrho - 1;
    iz = imodp(iz, Nz);
    const size_t base = (size_t)ir * (size_t)Nz * 2u + (size_t)iz * 2u;
    D[0] = table_ld(Tb + base);
    D[1] = table_ld(Tb + base + 1);
}

template <typename T>
__host__ __device__ __forceinline__ void s1r2_cic(const T *__restrict__ Tb, int Nrho, int Nz, T rho_max, T Lz, T rho, T z, T D[2]) {
    T ur;
    const T drho = rho_max / (T)max(1, Nrho - 1);
    if (rho < (T)0) ur = 0;
    else if (rho > rho_max) ur = (T)(Nrho - 1);
    else ur = rho / drho;
    int ir0 = (int)floor(ur);
    const T half = (T)0.5 * Lz;
    const T dz = Lz / (T)Nz;
    T fr = ur - (T)ir0;
    if (ir0 < 0) { ir0 = 0; fr = 0; }
    if (ir0 > Nrho - 2) { ir0 = max(0, Nrho - 2); fr = (T)1; }
    const int ir1 = ir0 + 1;
    const T uz = (z + half) / dz - (T)0.5;
    int iz0 = (int)floor(uz);
    const T fz = uz - (T)iz0;
    iz0 = imodp(iz0, Nz);
    const int iz1 = imodp(iz0 + 1, Nz);
    T w00 = (T)1 - fr; w00 *= ((T)1 - fz);
    T w10 = fr; w10 *= ((T)1 - fz);
    T w01 = (T)1 - fr; w01 *= fz;
    const T w11 = fr * fz;
    auto get = [&](int ir, int iz, int c) -> T { return table_ld(Tb + ((size_t)ir * (size_t)Nz + (size_t)iz) * 2u + c); };
    D[0] = w00 * get(ir0, iz0, 0) + w10 * get(ir1, iz0, 0) + w01 * get(ir0, iz1, 0) + w11 * get(ir1, iz1, 0);
    D[1] = w00 * get(ir0, iz0, 1) + w10 * get(ir1, iz0, 1) + w01 * get(ir0, iz1, 1) + w11 * get(ir1, iz1, 1);
}

template <typename T>
__host__ __device__ __forceinline__ void s1r2_tsc(const T *__restrict__ Tb, int Nrho, int Nz, T rho_max, T Lz, T rho, T z, T D[2]) {
    T ur;
    const T drho = rho_max / (T)max(1, Nrho - 1);
    if (rho < (T)0) ur = (T)0;
    else if (rho > rho_max) ur = (T)(Nrho - 1);
    else ur = rho / drho;
    const T half = (T)0.5 * Lz;
    const T dz = Lz / (T)Nz;
    const T uz = (z + half) / dz - (T)0.5;
    const int jr = (int)floor(ur + (T)0.5);
    const int jz = (int)floor(uz + (T)0.5);
    const T sr = ur - (T)jr;
    const T sz = uz - (T)jz;
    const T wrm = (T)0.5 * ((T)0.5 - sr) * ((T)0.5 - sr);
    const T wrc = (T)0.75 - sr * sr;
    const T wrp = (T)0.5 * ((T)0.5 + sr) * ((T)0.5 + sr);
    const T wzv[3] = {(T)0.5 * ((T)0.5 - sz) * ((T)0.5 - sz), (T)0.75 - sz * sz, (T)0.5 * ((T)0.5 + sz) * ((T)0.5 + sz)};
    int ir0 = jr - 1; if (ir0 < 0) ir0 = 0;
    int ir1 = jr; if (ir1 < 0) ir1 = 0; if (ir1 > Nrho - 1) ir1 = Nrho - 1;
    int ir2 = jr + 1; if (ir2 > Nrho - 1) ir2 = Nrho - 1;
    const int izv[3] = {imodp(jz - 1, Nz), imodp(jz, Nz), imodp(jz + 1, Nz)};
    const size_t b0 = (size_t)ir0 * (size_t)Nz * 2u, b1 = (size_t)ir1 * (size_t)Nz * 2u, b2 = (size_t)ir2 * (size_t)Nz * 2u;
    T d0 = 0, d1 = 0;
#pragma unroll
    for (int q = 0; q < 3; ++q) {
        const T wz = wzv[q];
        const T wr0 = wrm * wz, wr1 = wrc * wz, wr2 = wrp * wz;
        const size_t c = (size_t)izv[q] * 2u;
        d0 += wr0 * table_ld(Tb + b0 + c) + wr1 * table_ld(Tb + b1 + c) + wr2 * table_ld(Tb + b2 + c);
        d1 += wr0 * table_ld(Tb + b0 + c + 1) + wr1 * table_ld(Tb + b1 + c + 1) + wr2 * table_ld(Tb + b2 + c + 1);
    }
    D[0] = d0; D[1] = d1;
}

template <typename T>
__host__ __device__ __forceinline__ void s1r2_interpolate(const TopoParams &tp, T dx, T dy, T dz, T D[3]) {
    const T *__restrict__ Tb = static_cast<const T *>(tp.table);
    const T rho = sqrt(dx * dx + dy * dy);
    T Drz[2];
    if (tp.order == 0) s1r2_ngp<T>(Tb, tp.dim0, tp.dim1, (T)tp.rho_max, (T)tp.L, rho, dz, Drz);
    else if (tp.order == 2) s1r2_cic<T>(Tb, tp.dim0, tp.dim1, (T)tp.rho_max, (T)tp.L, rho, dz, Drz);
    else s1r2_tsc<T>(Tb, tp.dim0, tp.dim1, (T)tp.rho_max, (T)tp.L, rho, dz, Drz);
    const T ex = (rho > 0) ? dx / rho : (T)0;
    const T ey = (rho > 0) ? dy / rho : (T)0;
    D[0] = Drz[0] * ex;
    D[1] = Drz[0] * ey;
    D[2] = Drz[1];
}

// ---------------------------------------------------------------- one (i,j) pair, exact reference arithmetic
template <typename T, int TOPO>
__host__ __device__ __forceinline__ void pair_exact(const TopoParams &tp, T xi, T yi, T zi, T si, T xj, T yj, T zj, T mj, T sj, T &ax,
                                           T &ay, T &az) {
    const T beta = si + sj;
    T dx = xj - xi, dy = yj - yi, dz = zj - zi;
    const T L = (T)tp.L;
    if (TOPO == 0) {
        const T r = sqrt(dx * dx + dy * dy + dz * dz);
        const T w = mj * softened_w<T>(r, beta);
        ax += w * dx; ay += w * dy; az += w * dz;
    } else if (TOPO == 1) {
        if (fabs(dx) > (T)0.5 * L) dx = dx - L * dx / fabs(dx);
        if (fabs(dy) > (T)0.5 * L) dy = dy - L * dy / fabs(dy);
        if (fabs(dz) > (T)0.5 * L) dz = dz - L * dz / fabs(dz);
        const T r = sqrt(dx * dx + dy * dy + dz * dz);
        if (tp.is_periodic == 1) {
            const T w = mj * softened_w<T>(r, beta);
            ax += w * dx; ay += w * dy; az += w * dz;
        } else {
            const T w = softened_w<T>(r, beta);
            T D[3];
            t3_interpolate<T>(tp.dim0, L, static_cast<const T *>(tp.table), dx, dy, dz, D);
            ax += mj * (w * dx - D[0]);
            ay += mj * (w * dy - D[1]);
            az += mj * (w * dz - D[2]);
        }
    } else if (TOPO == 3 && tp.is_periodic >= 2) {
        // direct real-space image sum along z (forces_cuda.cu:709-743; cut test as the CPU oracle, forces.cc:1274)
        T fx = 0, fy = 0, fz = 0;
        const T cut = (T)tp.ewald_cut * L;
        for (int m = -tp.ewald_max; m < tp.ewald_max + 1; ++m) {
            const T dzi = dz + ((T)m) * L;
            if (fabs(dzi) <= cut) {
                const T r = sqrt(dx * dx + dy * dy + dzi * dzi);
                const T w = mj * softened_w<T>(r, beta);
                fx += w * dx; fy += w * dy; fz += w * dzi;
            }
        }
        ax += fx; ay += fy; az += fz;
    } else {
        // nearest z image (forces_cuda.cu:674-677 / :785-788 / :826-829)
        if (dz > (T)0.5 * L) dz -= L;
        else if (dz < (T)(-0.5) * L) dz += L;
        const T r = sqrt(dx * dx + dy * dy + dz * dz);
        if (TOPO == 2 && tp.is_periodic >= 2) {
            const T w = softened_w<T>(r, beta);
            T D[3];
            s1r2_interpolate<T>(tp, dx, dy, dz, D);
            ax += mj * (w * dx - D[0]);
            ay += mj * (w * dy - D[1]);
            az += mj * (w * dz - D[2]);
        } else {
            const T w = mj * softened_w<T>(r, beta);
            ax += w * dx; ay += w * dy; az += w * dz;
        }
    }
}

template <typename T, int TOPO, int R, int THREADS, int TJ, int STAGES>
__global__ void __launch_bounds__(THREADS) force_generic_kernel(const R3LaunchArgs a, const TopoParams tp) {
    if (a.gate && *a.gate != a.gate_value) return;  // a tuned kernel handles this call (see pair_s1r2.cuh)
    using JRec = typename JRecOf<T>::type;
    constexpr int NWARPS = THREADS / 32;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    JRec *tiles = reinterpret_cast<JRec *>(smem_raw);
    uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw + (size_t)STAGES * TJ * sizeof(JRec));
    uint64_t *empty = full + STAGES;

    const int tid = threadIdx.x;
    const int jc = blockIdx.x / a.n_ib;
    const int ib = blockIdx.x - jc * a.n_ib;
    const int t0 = jc * a.tiles_per_chunk;
    const int t1 = min(t0 + a.tiles_per_chunk, a.n_tiles);
    const int nt = t1 - t0;
    const JRec *__restrict__ jrec = static_cast<const JRec *>(a.jrec);

    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], NWARPS);
        }
        fence_mbar_init();
    }
    __syncthreads();
    if (tid == 0) {
        const int npre = nt < STAGES ? nt : STAGES;
        for (int t = 0; t < npre; ++t) {
            mbar_arrive_expect_tx(&full[t], TJ * sizeof(JRec));
            tma_load_1d(tiles + (size_t)t * TJ, jrec + (size_t)(t0 + t) * TJ, TJ * sizeof(JRec), &full[t]);
        }
    }

    T xi[R], yi[R], zi[R], si[R], ax[R], ay[R], az[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
        int il = ib * (THREADS * R) + r * THREADS + tid;
        il = il < a.n_i ? il : a.n_i - 1;
        const JRec me = jrec[a.id_min + il];
        xi[r] = me.x; yi[r] = me.y; zi[r] = me.z; si[r] = me.s;
        ax[r] = ay[r] = az[r] = 0;
    }

    for (int t = 0; t < nt; ++t) {
        const int s = t % STAGES;
        const uint32_t ph = (uint32_t)(t / STAGES) & 1u;
        if (tid == 0 && t >= 1 && (t - 1 + STAGES) < nt) {
            const int sp = (t - 1) % STAGES;
            const uint32_t php = (uint32_t)((t - 1) / STAGES) & 1u;
            mbar_wait(&empty[sp], php);
            mbar_arrive_expect_tx(&full[sp], TJ * sizeof(JRec));
            tma_load_1d(tiles + (size_t)sp * TJ, jrec + (size_t)(t0 + t - 1 + STAGES) * TJ, TJ * sizeof(JRec), &full[sp]);
        }
        mbar_wait(&full[s], ph);
        const JRec *__restrict__ Tl = tiles + (size_t)s * TJ;
        // the last tile is padded with massless far-away records; skip them so that periodic wraps
        // and table lookups never see the padding coordinates
        const int jn = min(TJ, a.n_j - (t0 + t) * TJ);
        for (int jj = 0; jj < jn; ++jj) {
            const JRec q = Tl[jj];
#pragma unroll
            for (int r = 0; r < R; ++r)
                pair_exact<T, TOPO>(tp, xi[r], yi[r], zi[r], si[r], (T)q.x, (T)q.y, (T)q.z, (T)q.m, (T)q.s, ax[r], ay[r], az[r]);
        }
        __syncwarp();
        if ((tid & 31) == 0) mbar_arrive(&empty[s]);
    }

    T *__restrict__ fp = static_cast<T *>(a.fpart) + (size_t)jc * 3 * a.fstride;
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const int il = ib * (THREADS * R) + r * THREADS + tid;
        if (il < a.n_i) {
            fp[il] = ax[r];
            fp[a.fstride + il] = ay[r];
            fp[2 * (size_t)a.fstride + il] = az[r];
        }
    }
}

}  // namespace steps
