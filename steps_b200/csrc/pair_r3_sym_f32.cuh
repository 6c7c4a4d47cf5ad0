// pair_r3_sym_f32.cuh -- action-reaction R^3 pair kernel of the single-precision build (BASELINE.json configs[4]).
//
// The scheme of pair_r3_sym.cuh (one evaluation of g = w(r_ij) and d = x_j - x_i serves both particles; records visit the
// lanes of a warp systolically with their accumulators rotating by shuffle; one partial row per (i-block, j-tile);
// SymRule tables built by the host) with the FP32 pair arithmetic of pair_r3_f32.cuh:
//     3 FADD (d), FMUL + 2 FFMA (r2), y = MUFU.RSQ(r2), t = y*y, e = fma(-r2,t,1), c = t*y, q = fma(1.5,e,1), g = c*q,
//     wi = g*m_j, wj = g*m_i, 3 FFMA into F_i, 3 FFMA into the accumulator of j
// = 19 FP32-pipe instructions per unordered pair (9.5 per interaction) against 14 per interaction one-sided.
// Records are 32 bytes {x,y,z,m | s,smax,-,-}; the visiting copy of a tile is one float4 (x,y,z,m) per record, each group of
// 32 stored twice in a row of 64 (no wrap, conflict-free 128-bit loads).
#pragma once
#include "pair_r3_f32.cuh"
#include "pair_r3_sym.cuh"

namespace steps {

__device__ __noinline__ float sym_exact_w_f32(float r2, float beta) { return softened_w<float>(sqrtf(r2), beta); }

#define STEPS_PAIR_SYM_CORE_F32(Q, r, YB_EXPR)                   \
    const float dx = (Q).x - xi[r];                              \
    const float dy = (Q).y - yi[r];                              \
    const float dz = (Q).z - zi[r];                              \
    float r2 = dx * dx;                                          \
    r2 = fmaf(dy, dy, r2);                                       \
    r2 = fmaf(dz, dz, r2);                                       \
    int yb = __float_as_int(rsqrt_seed(r2));                     \
    YB_EXPR;                                                     \
    const float y = __int_as_float(yb);                          \
    const float tt = y * y;                                      \
    const float e = fmaf(-r2, tt, 1.0f);                         \
    const float c = tt * y;                                      \
    const float q = fmaf(1.5f, e, 1.0f);                         \
    const float g = c * q;

template <int R, int TJ, int THREADS, bool CHECKED, int UNR>
__device__ __forceinline__ void sym_tile_f32(const JRec32 *__restrict__ T, const float4 *__restrict__ stage, int lane, int tid,
                                             const float (&xi)[R], const float (&yi)[R], const float (&zi)[R], const float (&mi)[R],
                                             float (&ax)[R], float (&ay)[R], float (&az)[R], const int (&thr)[R], float *__restrict__ slot,
                                             const JRec32 *__restrict__ jrec, int id_min, int n_i, int ib) {
    constexpr int IB = THREADS * R;
    for (int g0 = 0; g0 < TJ; g0 += 32) {
        float vx = 0.f, vy = 0.f, vz = 0.f;
        const JRec32 *__restrict__ G = T + g0;
        const float4 *__restrict__ S = stage + (g0 * 2) + lane;  // row of 64 per group: element l+s holds record (l+s) mod 32
        float4 nq = S[0];
#pragma unroll UNR
        for (int s2 = 0; s2 < 32; ++s2) {
            const float4 q4 = nq;
            nq = S[s2 + 1];  // record of the next step (prefetched)
            int ymin = 0x7fffffff;
#pragma unroll
            for (int r = 0; r < R; ++r) {
                STEPS_PAIR_SYM_CORE_F32(q4, r, if (CHECKED) { yb = (__float_as_int(r2) <= thr[r]) ? 0 : yb; ymin = min(ymin, yb); })
                const float wi = g * q4.w;
                const float wj = g * mi[r];
                ax[r] = fmaf(wi, dx, ax[r]);
                ay[r] = fmaf(wi, dy, ay[r]);
                az[r] = fmaf(wi, dz, az[r]);
                vx = fmaf(wj, dx, vx);
                vy = fmaf(wj, dy, vy);
                vz = fmaf(wj, dz, vz);
            }
            if (CHECKED) {
                if (ymin == 0) {
                    const float sj = G[(lane + s2) & 31].s;
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        const float dx = q4.x - xi[r], dy = q4.y - yi[r], dz = q4.z - zi[r];
                        float r2 = dx * dx;
                        r2 = fmaf(dy, dy, r2);
                        r2 = fmaf(dz, dz, r2);
                        if (__float_as_int(r2) <= thr[r]) {
                            int il = ib * IB + r * THREADS + tid;
                            il = il < n_i ? il : n_i - 1;
                            const float w = sym_exact_w_f32(r2, jrec[id_min + il].s + sj);
                            const float wi = w * q4.w, wj = w * mi[r];
                            ax[r] = fmaf(wi, dx, ax[r]);
                            ay[r] = fmaf(wi, dy, ay[r]);
                            az[r] = fmaf(wi, dz, az[r]);
                            vx = fmaf(wj, dx, vx);
                            vy = fmaf(wj, dy, vy);
                            vz = fmaf(wj, dz, vz);
                        }
                    }
                }
                __syncwarp();
            }
            vx = __shfl_sync(0xffffffffu, vx, (lane + 1) & 31);
            vy = __shfl_sync(0xffffffffu, vy, (lane + 1) & 31);
            vz = __shfl_sync(0xffffffffu, vz, (lane + 1) & 31);
        }
        slot[g0 + lane] = vx;
        slot[TJ + g0 + lane] = vy;
        slot[2 * TJ + g0 + lane] = vz;
    }
}

// shared memory of the FP32 action-reaction kernel besides the window of j-side accumulators (pair_r3_sym.cuh: sym_window_tiles)
__host__ __device__ constexpr int sym_base_f32(int nwarps, int stages, int tj = 128) {
    return stages * (tj * 32 + 32) + nwarps * 48 + 2 * tj * 16 + 2 * nwarps * 3 * tj * 4 + 2 * stages * 8;
}

template <int R, int THREADS, int TJ, int STAGES, int MINB, int UNR>
__global__ void __launch_bounds__(THREADS, MINB) force_r3_f32_sym_kernel(const SymLaunchArgs sa) {
    constexpr int NWARPS = THREADS / 32;
    constexpr int JB = 16;
    constexpr int IB = THREADS * R;
    constexpr int WB = sym_window_tiles(MINB, sym_base_f32(NWARPS, STAGES, TJ), 4, TJ);
    static_assert(sizeof(JRec32) == 32 && sizeof(TileInfo32) == 32 && sizeof(WarpBounds32) == 48, "sym_base_f32");
    static_assert(THREADS == TJ && TJ % 32 == 0 && IB % TJ == 0, "shape");
    const R3LaunchArgs &a = sa.a;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    JRec32 *tiles = reinterpret_cast<JRec32 *>(smem_raw);
    TileInfo32 *tinfo_s = reinterpret_cast<TileInfo32 *>(smem_raw + (size_t)STAGES * TJ * sizeof(JRec32));
    WarpBounds32 *wb_s = reinterpret_cast<WarpBounds32 *>(tinfo_s + STAGES);
    float4 *stage = reinterpret_cast<float4 *>(wb_s + NWARPS);                 // [TJ/32][64]
    float *slots = reinterpret_cast<float *>(stage + 2 * TJ);                   // [2][NWARPS][3][TJ]
    float *jacc = slots + 2 * NWARPS * 3 * TJ;                                  // [WB][3][TJ]: j-side sums of the current window (pair_r3_sym.cuh)
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
    const JRec32 *__restrict__ jrec = static_cast<const JRec32 *>(a.jrec);
    const TileInfo32 *__restrict__ tinfo = static_cast<const TileInfo32 *>(a.tinfo);
    constexpr uint32_t TILE_TX = TJ * sizeof(JRec32) + sizeof(TileInfo32);

    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], NWARPS);
        }
        fence_mbar_init();
    }
    __syncthreads();
    const WarpBounds32 *__restrict__ wb = wb_s + warp;
    int nsym = 0;
    int K0 = 0;  // tiles streamed so far by this CTA (pipeline stage / parity bookkeeping, pair_r3_sym.cuh)
    const uint64_t keep = l2_policy_keep();  // the i-side sums come back within milliseconds: ask L2 to hold them (ptx_helpers.cuh)

    for (int w0 = TA; w0 < TB; w0 += WB) {
    const int w1 = min(w0 + WB, TB);
#pragma unroll 4
    for (int k = 0; k < WB * 3; ++k) jacc[k * TJ + tid] = 0.f;
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
            mbar_arrive_expect_tx(&full[K % STAGES], TILE_TX);
            tma_load_1d(tiles + (size_t)(K % STAGES) * TJ, jrec + (size_t)(t0 + t) * TJ, TJ * sizeof(JRec32), &full[K % STAGES]);
            tma_load_1d(tinfo_s + (K % STAGES), tinfo + (t0 + t), sizeof(TileInfo32), &full[K % STAGES]);
        }
    }
    float *__restrict__ fp = static_cast<float *>(a.fpart) + (size_t)jc * 3 * a.fstride;

    float xi[R], yi[R], zi[R], mi[R], ax[R], ay[R], az[R];
    {
        float lo[3] = {3e38f, 3e38f, 3e38f}, hi[3] = {-3e38f, -3e38f, -3e38f}, rlo = 3e38f, rhi = 0.f, smx = 0.f;
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int il0 = ib * IB + r * THREADS + tid;
            const int il = il0 < a.n_i ? il0 : a.n_i - 1;
            const JRec32 me = jrec[a.id_min + il];
            xi[r] = me.x; yi[r] = me.y; zi[r] = me.z;
            mi[r] = il0 < a.n_i ? me.m : 0.f;  // a clamped duplicate must not act on the j side
            if (first || il0 >= a.n_i) {
                ax[r] = ay[r] = az[r] = 0.f;
            } else {  // sums of the earlier windows of this chunk
                ax[r] = ld_keep(fp + il0, keep);
                ay[r] = ld_keep(fp + a.fstride + il0, keep);
                az[r] = ld_keep(fp + 2 * (size_t)a.fstride + il0, keep);
            }
            lo[0] = fminf(lo[0], me.x); hi[0] = fmaxf(hi[0], me.x);
            lo[1] = fminf(lo[1], me.y); hi[1] = fmaxf(hi[1], me.y);
            lo[2] = fminf(lo[2], me.z); hi[2] = fmaxf(hi[2], me.z);
            const float rr = sqrtf(me.x * me.x + me.y * me.y + me.z * me.z);
            rlo = fminf(rlo, rr); rhi = fmaxf(rhi, rr);
            smx = fmaxf(smx, me.s);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                lo[k] = fminf(lo[k], __shfl_xor_sync(0xffffffffu, lo[k], o));
                hi[k] = fmaxf(hi[k], __shfl_xor_sync(0xffffffffu, hi[k], o));
            }
            rlo = fminf(rlo, __shfl_xor_sync(0xffffffffu, rlo, o));
            rhi = fmaxf(rhi, __shfl_xor_sync(0xffffffffu, rhi, o));
            smx = fmaxf(smx, __shfl_xor_sync(0xffffffffu, smx, o));
        }
        if (lane == 0) {
            WarpBounds32 &wbw = wb_s[warp];
            wbw.lo[0] = lo[0]; wbw.lo[1] = lo[1]; wbw.lo[2] = lo[2];
            wbw.hi[0] = hi[0]; wbw.hi[1] = hi[1]; wbw.hi[2] = hi[2];
            wbw.rlo = rlo; wbw.rhi = rhi; wbw.smax = smx;
        }
        __syncwarp();
    }

    for (int t = 0; t < nt; ++t) {
        const int K = K0 + t;
        const int s = K % STAGES;
        const uint32_t ph = (uint32_t)(K / STAGES) & 1u;
        if (tid == 0 && t >= 1 && (t - 1 + STAGES) < nt) {
            const int Kn = K - 1 + STAGES;
            const int sp = Kn % STAGES;
            sym_wait_stage_free<STAGES>(empty, Kn);
            mbar_arrive_expect_tx(&full[sp], TILE_TX);
            tma_load_1d(tiles + (size_t)sp * TJ, jrec + (size_t)(t0 + t - 1 + STAGES) * TJ, TJ * sizeof(JRec32), &full[sp]);
            tma_load_1d(tinfo_s + sp, tinfo + (t0 + t - 1 + STAGES), sizeof(TileInfo32), &full[sp]);
        }
        mbar_wait(&full[s], ph);
        const JRec32 *__restrict__ T = tiles + (size_t)s * TJ;
        const int cls = sym_tile_class(*rule, t0 + t);  // CTA-uniform
        if (cls != 0) {
            const float smax = T[0].smax;
            bool far;
            {
                const TileInfo32 *__restrict__ ti = tinfo_s + s;
                float gap2 = 0.f;
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const float gp = fmaxf(fmaxf(wb->lo[k] - ti->hi[k], ti->lo[k] - wb->hi[k]), 0.f);
                    gap2 = fmaf(gp, gp, gap2);
                }
                const float rg = fmaxf(wb->rlo - ti->rhi, ti->rlo - wb->rhi);
                const float b = (wb->smax + smax) * 1.001f;
                far = (gap2 > b * b) || (rg > b);
            }
            int thr[R];
#pragma unroll
            for (int r = 0; r < R; ++r) thr[r] = -1;  // far tiles: nothing is ever flagged (bits of r2 >= 0)
            if (!far) {
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    int il = ib * IB + r * THREADS + tid;
                    il = il < a.n_i ? il : a.n_i - 1;
                    const float b = jrec[a.id_min + il].s + smax;
                    thr[r] = __float_as_int(b * b) + 4;  // conservative, as pair_r3_f32.cuh
                }
            }
            if (cls == 1) {
                // the i-block's own tiles: one-sided, checked loop
                for (int j0 = 0; j0 < TJ; j0 += JB) {
                    int ymin = 0x7fffffff;
#pragma unroll 1
                    for (int jj = 0; jj < JB; ++jj) {
                        const float4 q4 = *reinterpret_cast<const float4 *>(&T[j0 + jj].x);
#pragma unroll
                        for (int r = 0; r < R; ++r) {
                            STEPS_PAIR_SYM_CORE_F32(q4, r, yb = (__float_as_int(r2) <= thr[r]) ? 0 : yb; ymin = min(ymin, yb))
                            const float wi = g * q4.w;
                            ax[r] = fmaf(wi, dx, ax[r]);
                            ay[r] = fmaf(wi, dy, ay[r]);
                            az[r] = fmaf(wi, dz, az[r]);
                        }
                    }
                    if (ymin == 0) {
#pragma unroll
                        for (int r = 0; r < R; ++r) {
                            int il = ib * IB + r * THREADS + tid;
                            il = il < a.n_i ? il : a.n_i - 1;
                            const float3 f = near_pairs_f32(T + j0, JB, xi[r], yi[r], zi[r], jrec[a.id_min + il].s, thr[r]);
                            ax[r] += f.x; ay[r] += f.y; az[r] += f.z;
                        }
                    }
                }
            } else {
                float *__restrict__ slot = slots + ((size_t)(nsym & 1) * NWARPS + warp) * 3 * TJ;
                {
                    const float4 q4 = *reinterpret_cast<const float4 *>(&T[tid].x);
                    float4 *__restrict__ sq = stage + (tid >> 5) * 64 + (tid & 31);
                    sq[0] = q4;
                    sq[32] = q4;
                }
                __syncthreads();
                if (far)
                    sym_tile_f32<R, TJ, THREADS, false, UNR>(T, stage, lane, tid, xi, yi, zi, mi, ax, ay, az, thr, slot, jrec, a.id_min, a.n_i, ib);
                else
                    sym_tile_f32<R, TJ, THREADS, true, UNR>(T, stage, lane, tid, xi, yi, zi, mi, ax, ay, az, thr, slot, jrec, a.id_min, a.n_i, ib);
                __syncthreads();
                {
                    // the warps' sums in warp order onto the window's accumulator (blocks of the superblock arrive in block order)
                    const float *__restrict__ sb = slots + (size_t)(nsym & 1) * NWARPS * 3 * TJ;
                    float *__restrict__ ja = jacc + (size_t)(t0 + t - w0) * 3 * TJ + tid;
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        float v = 0.f;
#pragma unroll
                        for (int w = 0; w < NWARPS; ++w) v += sb[((size_t)w * 3 + c) * TJ + tid];
                        ja[c * TJ] += v;
                    }
                }
                ++nsym;
            }
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
        float *__restrict__ gp = static_cast<float *>(sa.gpart) + (size_t)gs * 3 * sa.n_pad + (size_t)w0 * TJ + tid;
        for (int tl = 0; tl < w1 - w0; ++tl) {
#pragma unroll
            for (int c = 0; c < 3; ++c) __stcs(gp + (size_t)c * sa.n_pad + (size_t)tl * TJ, jacc[(tl * 3 + c) * TJ + tid]);
        }
    }
    }  // windows
}

}  // namespace steps
