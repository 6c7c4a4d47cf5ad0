// pair_generic_sym.cuh -- action-reaction evaluation for the table-lookup topologies: T^3 (nearest image + tricubic Ewald
// correction, forces_cuda.cu:567-645) and the S^1xR^2 lookup build (forces_cuda.cu:764-864), both precisions.
// OPT-IN (STEPS_B200_GEN_SYM=1 or steps_b200_engine_set_symmetric(e, 1)): written after round 1's GPU budget was spent; it has
// not run on a GPU yet (tests/test_gpu_generic_sym.py, gated by STEPS_B200_EXPERIMENTAL=1).
//
// The pair term of these topologies is  F_i += m_j t(d),  t(d) = w(|d|, s_i+s_j) d - D(d),  d = nearest image of x_j - x_i.
// w is even in d and the interpolated correction D is odd (the tables are built from inversion-symmetric image sums and the
// Catmull-Rom / TSC / CIC weights are mirror symmetric; in floating point D(-d) = -D(d) holds to the rounding of the cell
// coordinate, ~1e-15 of D), so one evaluation of t -- the 64-point gather is the cost of this path -- serves both particles:
//     F_i += m_j t,   F_j -= m_i t.
// Scheme, rules, passes and reductions are those of pair_r3_sym.cuh; the per-pair arithmetic is that of pair_exact()
// (pair_generic.cuh), operation by operation, with the mass factored out.
#pragma once
#include "pair_generic.cuh"
#include "pair_r3_sym.cuh"

namespace steps {

// unit-mass pair vector t of pair_exact<T, TOPO>: F_i += m_j t
template <typename T, int TOPO>
__host__ __device__ __forceinline__ void pair_exact_unit(const TopoParams &tp, T xi, T yi, T zi, T si, T xj, T yj, T zj, T sj, T &tx, T &ty, T &tz) {
    static_assert(TOPO == 1 || TOPO == 2, "table-lookup topologies");
    const T beta = si + sj;
    T dx = xj - xi, dy = yj - yi, dz = zj - zi;
    const T L = (T)tp.L;
    if (TOPO == 1) {
        if (fabs(dx) > (T)0.5 * L) dx = dx - L * dx / fabs(dx);
        if (fabs(dy) > (T)0.5 * L) dy = dy - L * dy / fabs(dy);
        if (fabs(dz) > (T)0.5 * L) dz = dz - L * dz / fabs(dz);
        const T r = sqrt(dx * dx + dy * dy + dz * dz);
        const T w = softened_w<T>(r, beta);
        if (tp.is_periodic == 1) {
            tx = w * dx; ty = w * dy; tz = w * dz;
        } else {
            T D[3];
            t3_correction_rowmajor<T>(t3_lookup_of<T>(tp), dx, dy, dz, D);
            tx = w * dx - D[0];
            ty = w * dy - D[1];
            tz = w * dz - D[2];
        }
    } else {
        if (dz > (T)0.5 * L) dz -= L;
        else if (dz < (T)(-0.5) * L) dz += L;
        const T r = sqrt(dx * dx + dy * dy + dz * dz);
        const T w = softened_w<T>(r, beta);
        if (tp.is_periodic >= 2) {
            T D[3];
            s1r2_interpolate<T>(tp, dx, dy, dz, D);
            tx = w * dx - D[0];
            ty = w * dy - D[1];
            tz = w * dz - D[2];
        } else {
            tx = w * dx; ty = w * dy; tz = w * dz;
        }
    }
}

// ---------------------------------------------------------------- T^3, IS_PERIODIC >= 2: lean arithmetic (FAST variants)
// Same quantities as pair_exact_unit<T, 1> with cheaper instruction sequences; every substitution moves a result by a few ulp at
// most (tolerance of the path: 1e-12):
//   * nearest image by d -= copysign(L, d) instead of d - L*d/|d| (a division whose quotient is +-1 up to one rounding of L*d);
//   * r^-3 = rsqrt(r2)^3 for r safely outside the softening length, the exact branches of force_softening otherwise;
//   * the correction from the z-window copy of the table with 128-bit loads (t3_correction_zwin, t3_lookup.cuh).
template <typename T>
__host__ __device__ __forceinline__ void pair_t3_fast_unit(const T3Lookup &k, T xi, T yi, T zi, T si, T xj, T yj, T zj, T sj, T &tx, T &ty, T &tz) {
    const T L = (T)k.L, halfL = (T)k.halfL;
    T dx = xj - xi, dy = yj - yi, dz = zj - zi;
    if (fabs(dx) > halfL) dx -= copysign(L, dx);
    if (fabs(dy) > halfL) dy -= copysign(L, dy);
    if (fabs(dz) > halfL) dz -= copysign(L, dz);
    const T r2 = dx * dx + dy * dy + dz * dz;
    const T beta = si + sj;
    T w;
    if (r2 > beta * beta * (T)1.0001) {
        const T y = rsqrt(r2);
        w = y * y * y;
    } else {
        w = softened_w<T>(sqrt(r2), beta);
    }
    T D[3];
    t3_correction_zwin<T>(k, dx, dy, dz, D);
    tx = fma(w, dx, -D[0]);
    ty = fma(w, dy, -D[1]);
    tz = fma(w, dz, -D[2]);
}

// tiles per window of the shared-memory j-side accumulator of this kernel: T^3 takes a short one (see the kernel), the S^1xR^2 lookup
// build, whose units are 20 x shorter and whose table is small, takes what fits (measured: 2.8e10 pairs/s with 2 tiles, 5e10 with 16)
__host__ __device__ constexpr int gen_sym_window(int topo, int minb, int base, int elem, int tj = 128) {
    return topo == 1 ? 2 : sym_window_tiles(minb, base, elem, tj);
}
// shared memory of this kernel besides the window of j-side accumulators
__host__ __device__ constexpr int sym_base_generic(int nwarps, int stages, int jrec_bytes, int elem, int tj = 128) {
    return stages * tj * jrec_bytes + 2 * nwarps * 3 * tj * elem + 2 * stages * 8;
}

template <typename T, int TOPO, int R, int THREADS, int TJ, int STAGES, int MINB, bool FAST>
__global__ void __launch_bounds__(THREADS, MINB) force_generic_sym_kernel(const SymLaunchArgs sa, const TopoParams tp) {
    using JRec = typename JRecOf<T>::type;
    constexpr int NWARPS = THREADS / 32;
    constexpr int IB = THREADS * R;
    // (a short window on purpose: a unit of this kernel takes ~0.5 ms, so the i-side round trips per window cost nothing, while every
    // KB of shared memory is taken from the L1 that serves the table gather -- measured: hit rate 85 % with a 16-tile window)
    constexpr int WB = gen_sym_window(TOPO, MINB, sym_base_generic(NWARPS, STAGES, (int)sizeof(JRec), (int)sizeof(T), TJ), (int)sizeof(T), TJ);
    static_assert(THREADS == TJ && TJ % 32 == 0 && IB % TJ == 0, "shape");
    const R3LaunchArgs &a = sa.a;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    JRec *tiles = reinterpret_cast<JRec *>(smem_raw);
    T *slots = reinterpret_cast<T *>(smem_raw + (size_t)STAGES * TJ * sizeof(JRec));  // [2][NWARPS][3][TJ]
    T *jacc = slots + 2 * NWARPS * 3 * TJ;  // [WB][3][TJ]: j-side sums of the current window (pair_r3_sym.cuh)
    uint64_t *full = reinterpret_cast<uint64_t *>(jacc + (size_t)WB * 3 * TJ);
    uint64_t *empty = full + STAGES;

    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const int2 od = sa.order[blockIdx.x];
    const int gs = od.x;  // superblock within the pass
    const int jc = od.y;  // j-chunk
    const int ib_lo = (sa.b0 + gs) * sa.sb;
    const int ib_hi = min(ib_lo + sa.sb, a.n_ib);
    const int c0 = jc * a.tiles_per_chunk;
    const int c1 = min(c0 + a.tiles_per_chunk, a.n_tiles);
    int TA = 0x7fffffff, TB = -1;  // tile range of the whole superblock inside this chunk
    for (int ib = ib_lo; ib < ib_hi; ++ib) {
        int ha, hb;
        sym_hull(sa.rules[ib], c0, c1, ha, hb);
        if (ha < hb) { TA = min(TA, ha); TB = max(TB, hb); }
    }
    if (TB <= TA) return;  // (the host's order table holds no such CTA)
    const JRec *__restrict__ jrec = static_cast<const JRec *>(a.jrec);

    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], NWARPS);
        }
        fence_mbar_init();
    }
    __syncthreads();
    int nsym = 0;
    int K0 = 0;  // tiles streamed so far by this CTA (pipeline stage / parity bookkeeping, pair_r3_sym.cuh)
    const uint64_t keep = l2_policy_keep();  // the i-side sums come back within milliseconds: ask L2 to hold them (ptx_helpers.cuh)
    static_assert(!FAST || TOPO == 1, "lean arithmetic exists for T^3 only");
    const T3Lookup fk = t3_lookup_of<T>(tp);
    // (the engine launches the FAST instantiation only for IS_PERIODIC >= 2: the nearest-image-only sum has no table)

    for (int w0 = TA; w0 < TB; w0 += WB) {
    const int w1 = min(w0 + WB, TB);
#pragma unroll 4
    for (int q = 0; q < WB * 3; ++q) jacc[q * TJ + tid] = 0;
    for (int ib = ib_lo; ib < ib_hi; ++ib) {
    const SymRule *__restrict__ rule = sa.rules + ib;
    int ha, hb;
    sym_hull(*rule, c0, c1, ha, hb);
    const int t0 = max(ha, w0), nt = min(hb, w1) - t0;
    if (nt <= 0) continue;
    const bool first = w0 <= ha;  // first window that reaches this block's tiles: the i-side sums start from zero
    if (tid == 0) {
        const int npre = nt < STAGES ? nt : STAGES;
        for (int t = 0; t < npre; ++t) {
            const int K = K0 + t;
            sym_wait_stage_free<STAGES>(empty, K);
            mbar_arrive_expect_tx(&full[K % STAGES], TJ * sizeof(JRec));
            tma_load_1d(tiles + (size_t)(K % STAGES) * TJ, jrec + (size_t)(t0 + t) * TJ, TJ * sizeof(JRec), &full[K % STAGES]);
        }
    }
    T *__restrict__ fp = static_cast<T *>(a.fpart) + (size_t)jc * 3 * a.fstride;

    T xi[R], yi[R], zi[R], si[R], mi[R], ax[R], ay[R], az[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const int il0 = ib * IB + r * THREADS + tid;
        const int il = il0 < a.n_i ? il0 : a.n_i - 1;
        const JRec me = jrec[a.id_min + il];
        xi[r] = me.x; yi[r] = me.y; zi[r] = me.z; si[r] = me.s;
        mi[r] = il0 < a.n_i ? (T)me.m : (T)0;  // a clamped duplicate must not act on the j side
        if (first || il0 >= a.n_i) {
            ax[r] = ay[r] = az[r] = 0;
        } else {  // sums of the earlier windows of this chunk
            ax[r] = ld_keep(fp + il0, keep);
            ay[r] = ld_keep(fp + a.fstride + il0, keep);
            az[r] = ld_keep(fp + 2 * (size_t)a.fstride + il0, keep);
        }
    }

    for (int t = 0; t < nt; ++t) {
        const int K = K0 + t;
        const int s = K % STAGES;
        const uint32_t ph = (uint32_t)(K / STAGES) & 1u;
        if (tid == 0 && t >= 1 && (t - 1 + STAGES) < nt) {
            const int Kn = K - 1 + STAGES;
            const int sp = Kn % STAGES;
            sym_wait_stage_free<STAGES>(empty, Kn);
            mbar_arrive_expect_tx(&full[sp], TJ * sizeof(JRec));
            tma_load_1d(tiles + (size_t)sp * TJ, jrec + (size_t)(t0 + t - 1 + STAGES) * TJ, TJ * sizeof(JRec), &full[sp]);
        }
        mbar_wait(&full[s], ph);
        const JRec *__restrict__ Tl = tiles + (size_t)s * TJ;
        const int cls = sym_tile_class(*rule, t0 + t);  // CTA-uniform
        // the last tile is padded with massless far-away records: they are never evaluated (periodic wraps and table
        // lookups must not see the padding coordinates)
        const int jn = min(TJ, a.n_j - (t0 + t) * TJ);
        if (cls == 1) {
            // ---- the i-block's own tiles: one-sided, exactly the loop of force_generic_kernel ----
            for (int jj = 0; jj < jn; ++jj) {
                const JRec q = Tl[jj];
#pragma unroll
                for (int r = 0; r < R; ++r)
                    pair_exact<T, TOPO>(tp, xi[r], yi[r], zi[r], si[r], (T)q.x, (T)q.y, (T)q.z, (T)q.m, (T)q.s, ax[r], ay[r], az[r]);
            }
        } else if (cls == 2) {
            // ---- symmetric tile: the 32 records of a group visit the lanes systolically (pair_r3_sym.cuh) ----
            T *__restrict__ slot = slots + ((size_t)(nsym & 1) * NWARPS + warp) * 3 * TJ;
            for (int g0 = 0; g0 < TJ; g0 += 32) {
                T vx = 0, vy = 0, vz = 0;
#pragma unroll 1
                for (int s2 = 0; s2 < 32; ++s2) {
                    const int jidx = g0 + ((lane + s2) & 31);
                    if (jidx < jn) {
                        const JRec q = Tl[jidx];
#pragma unroll
                        for (int r = 0; r < R; ++r) {
                            T tx, ty, tz;
                            if (FAST)
                                pair_t3_fast_unit<T>(fk, xi[r], yi[r], zi[r], si[r], (T)q.x, (T)q.y, (T)q.z, (T)q.s, tx, ty, tz);
                            else
                                pair_exact_unit<T, TOPO>(tp, xi[r], yi[r], zi[r], si[r], (T)q.x, (T)q.y, (T)q.z, (T)q.s, tx, ty, tz);
                            const T mj = (T)q.m;
                            ax[r] += mj * tx;
                            ay[r] += mj * ty;
                            az[r] += mj * tz;
                            vx += mi[r] * tx;
                            vy += mi[r] * ty;
                            vz += mi[r] * tz;
                        }
                    }
                    __syncwarp();
                    // the accumulator follows its record: lane l takes over the record lane l+1 just worked on
                    vx = __shfl_sync(0xffffffffu, vx, (lane + 1) & 31);
                    vy = __shfl_sync(0xffffffffu, vy, (lane + 1) & 31);
                    vz = __shfl_sync(0xffffffffu, vz, (lane + 1) & 31);
                }
                slot[g0 + lane] = vx;
                slot[TJ + g0 + lane] = vy;
                slot[2 * TJ + g0 + lane] = vz;
            }
            __syncthreads();
            {
                // the warps' sums in warp order onto the window's accumulator (blocks of the superblock arrive in block order)
                const T *__restrict__ sb = slots + (size_t)(nsym & 1) * NWARPS * 3 * TJ;
                T *__restrict__ ja = jacc + (size_t)(t0 + t - w0) * 3 * TJ + tid;
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    T v = 0;
#pragma unroll
                    for (int w = 0; w < NWARPS; ++w) v += sb[((size_t)w * 3 + c) * TJ + tid];
                    ja[c * TJ] += v;
                }
            }
            ++nsym;
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[s]);
    }
    K0 += nt;

#pragma unroll
    for (int r = 0; r < R; ++r) {
        const int il = ib * IB + r * THREADS + tid;
        if (il < a.n_i) {
            st_keep(fp + il, ax[r], keep);
            st_keep(fp + a.fstride + il, ay[r], keep);
            st_keep(fp + 2 * (size_t)a.fstride + il, az[r], keep);
        }
    }
    }  // i-blocks of the superblock
    // the window's j-side sums: one row segment per (superblock, tile); streamed (read once, by the row reduction)
    {
        T *__restrict__ gp = static_cast<T *>(sa.gpart) + (size_t)gs * 3 * sa.n_pad + (size_t)w0 * TJ + tid;
        for (int tl = 0; tl < w1 - w0; ++tl) {
#pragma unroll
            for (int c = 0; c < 3; ++c) __stcs(gp + (size_t)c * sa.n_pad + (size_t)tl * TJ, jacc[(tl * 3 + c) * TJ + tid]);
        }
    }
    }  // windows
}

}  // namespace steps
