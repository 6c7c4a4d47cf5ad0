// pair_generic.cuh -- exact-branch pair kernels for every topology and both precisions.
//
// These keep the branch structure of the reference's per-pair arithmetic (forces_cuda.cu:522-563 R^3, :567-645 T^3, :652-761
// S^1xR^2 NOLOOKUP, :764-864 S^1xR^2 lookup; the softened kernel :466-519) -- every pair takes the exact path, nothing is
// deferred or approximated -- with the table lookups of t3_lookup.cuh / s1r2_lookup.cuh, inside the B200 work decomposition of
// pair_r3.cuh: (i-block x j-chunk) CTAs, TMA-staged j tiles, register-blocked i-particles,
// deterministic chunk reduction.  The tuned R^3 kernels in pair_r3.cuh replace the R^3
// instantiations on the hot path; these remain the path for T^3 and S^1xR^2.
#pragma once
#include "pair_r3.cuh"
#include "s1r2_lookup.cuh"

namespace steps {

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
    const void *table_zwin;  // device: z-window copy of the T^3 table (t3_lookup.cuh), or nullptr
};

// the table geometry of a T^3 engine (t3_lookup.cuh) / the correction vector of the S^1xR^2 lookup build (s1r2_lookup.cuh)
template <typename T>
__host__ __device__ __forceinline__ T3Lookup t3_lookup_of(const TopoParams &tp) {
    return t3_lookup_make<T>(tp.L, tp.dim0, tp.table, tp.table_zwin);
}

template <typename T>
__host__ __device__ __forceinline__ void s1r2_interpolate(const TopoParams &tp, T dx, T dy, T dz, T (&D)[3]) {
    s1r2_correction<T>(static_cast<const T *>(tp.table), tp.order, tp.dim0, tp.dim1, (T)tp.rho_max, (T)tp.L, dx, dy, dz, D);
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
            t3_correction_rowmajor<T>(t3_lookup_of<T>(tp), dx, dy, dz, D);
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
