// pair_s1r2_sym.cuh -- action-reaction pair kernel of the S^1xR^2 slab, FP64 NOLOOKUP build, IS_PERIODIC 2..4
// (BASELINE.json configs[3]).  OPT-IN (STEPS_B200_S1R2_SYM=1) until the whole of tests/test_gpu_s1r2_sym.py has run on a GPU: 12 of its 16
// tests passed on a B200 in the last seconds of round 1's GPU budget (max |dF|/|F| 1.2e-13 against the reference, IS_PERIODIC 2..4), and
// it measured 3.58e11 pairs/s against 2.1e11 for the one-sided kernel at N = 200k (profiles/r1ad_*, r1ae_*).
//
// The image sum of forces_periodic_z (forces.cc:1262-1290, forces_cuda.cu:709-743)
//     F_i += m_j sum_{m=-M..M, |dz_m| <= cut} w(r_m, s_i+s_j) (dx, dy, dz_m),   dz_m = z_j - z_i + m L
// is antisymmetric under i <-> j: the set {|dz_m|} is the same for (j, i) (m -> -m), so with the unit-mass sums
//     W = sum_m w_m,   Z = sum_m w_m dz_m
// of ONE evaluation   F_i += m_j (W dx, W dy, Z)   and   F_j -= m_i (W dx, W dy, Z).
// The image slots (the expensive part: 2M slots x 10 FP64 instructions) are therefore evaluated once per unordered pair;
// the one-sided kernel of pair_s1r2.cuh spends 2M x 10 + 9 FP64 instructions per directed pair, this one
// (2M x 10 + 5 + 8) / 2 -- 38 against 69 at IS_PERIODIC = 2.
// Scheme, rules, passes and reductions are those of pair_r3_sym.cuh (systolic visiting records with rotating
// accumulators, one partial row per (i-block, j-tile), SymRule tables from the host); slot masking, the planar far/near
// classification and the exact slow path for the three nearest slots are those of pair_s1r2.cuh.  Like that kernel it
// requires every z inside [0, L): the engine reads the pack kernel's flag and takes the one-sided path otherwise.
#pragma once
#include "pair_r3_sym.cuh"
#include "pair_s1r2.cuh"

namespace steps {

// all image slots of one (i, j) pair: leaves dx, dy, dz, Wsum, Zsum in scope.  Same slot list as STEPS_S1R2_PAIR.
#define STEPS_S1R2_PAIR_CORE(NEARCHK)                                                                                    \
    const double dx = xy.x - xi[r];                                                                                      \
    const double dy = xy.y - yi[r];                                                                                      \
    const double dz = zm.x - zi[r];                                                                                      \
    double dead = dx * dx;                                                                                               \
    const double dxy2 = fma(dy, dy, dead);                                                                               \
    double Wsum, Zsum;                                                                                                   \
    STEPS_S1R2_SLOT(true, dz, { if (NEARCHK) { yh = (__double2hiint(r2) <= thr[r]) ? 0 : yh; ymin = min(ymin, yh); } })           \
    STEPS_S1R2_SLOT(false, dz - L, { if (NEARCHK) { yh = (__double2hiint(r2) <= thr[r]) ? 0 : yh; ymin = min(ymin, yh); } })      \
    STEPS_S1R2_SLOT(false, dz + L, { if (NEARCHK) { yh = (__double2hiint(r2) <= thr[r]) ? 0 : yh; ymin = min(ymin, yh); } })      \
    _Pragma("unroll") for (int m = 2; m <= M - 2; ++m) {                                                                 \
        STEPS_S1R2_SLOT(false, dz + ((double)m) * L, {})                                                                 \
        STEPS_S1R2_SLOT(false, dz - ((double)m) * L, {})                                                                 \
    }                                                                                                                    \
    STEPS_S1R2_SLOT(false, dz + LM1, { yh = le_pos(dzi_, cut_bits) ? yh : 0; })                                          \
    STEPS_S1R2_SLOT(false, dz - LM1, { yh = le_neg(dzi_, cutneg_bits) ? yh : 0; })                                       \
    STEPS_S1R2_SLOT(false, dz - copysign(LM, dz), { yh = le_pos(fabs(dzi_), cut_bits) ? yh : 0; })

struct S1R2SymRegs {
    double L, LM1, LM;
    long long cut_bits;
    unsigned long long cutneg_bits;
};

// One symmetric (i-warp x j-tile) block; see sym_tile of pair_r3_sym.cuh.
template <int R, int TJ, int THREADS, bool CHECKED, int UNR, int M>
__device__ __forceinline__ void sym_tile_s1r2(const JRec64 *__restrict__ T, const double *__restrict__ soa, int lane, int tid,
                                              const double (&xi)[R], const double (&yi)[R], const double (&zi)[R], const double (&mi)[R],
                                              double (&ax)[R], double (&ay)[R], double (&az)[R], const int (&thr)[R],
                                              double *__restrict__ slot, const JRec64 *__restrict__ jrec, int id_min, int n_i, int ib,
                                              const S1R2SymRegs &k) {
    constexpr int IB = THREADS * R;
    const double L = k.L, LM1 = k.LM1, LM = k.LM;
    const long long cut_bits = k.cut_bits;
    const unsigned long long cutneg_bits = k.cutneg_bits;
    for (int g0 = 0; g0 < TJ; g0 += 32) {
        double vx = 0.0, vy = 0.0, vz = 0.0;
        const JRec64 *__restrict__ G = T + g0;
        const double2 *__restrict__ SXY = reinterpret_cast<const double2 *>(soa) + (g0 * 2) + lane;
        const double2 *__restrict__ SZM = SXY + 2 * TJ;
        double2 nxy = SXY[0], nzm = SZM[0];
#pragma unroll UNR
        for (int s2 = 0; s2 < 32; ++s2) {
            const double2 xy = nxy, zm = nzm;
            nxy = SXY[s2 + 1];
            nzm = SZM[s2 + 1];
            int ymin = 0x7fffffff;
#pragma unroll
            for (int r = 0; r < R; ++r) {
                STEPS_S1R2_PAIR_CORE(CHECKED)
                const double mW = zm.y * Wsum;
                const double nW = mi[r] * Wsum;
                ax[r] = fma(mW, dx, ax[r]);
                ay[r] = fma(mW, dy, ay[r]);
                az[r] = fma(zm.y, Zsum, az[r]);
                vx = fma(nW, dx, vx);
                vy = fma(nW, dy, vy);
                vz = fma(mi[r], Zsum, vz);
            }
            if (CHECKED) {
                if (ymin == 0) {
                    // rare: a nearest-image evaluation (m = -1, 0, +1) of this step lies inside the softening radius: exact branches
                    const double sj = G[(lane + s2) & 31].s;
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        const double dx = xy.x - xi[r], dy = xy.y - yi[r], dz = zm.x - zi[r];
                        const double dxy2 = fma(dy, dy, dx * dx);
#pragma unroll
                        for (int m = -1; m <= 1; ++m) {
                            const double dzi = dz + ((double)m) * L;
                            const double r2 = fma(dzi, dzi, dxy2);
                            if (__double2hiint(r2) <= thr[r]) {
                                int il = ib * IB + r * THREADS + tid;
                                il = il < n_i ? il : n_i - 1;
                                const double w = sym_exact_w(r2, jrec[id_min + il].s + sj);
                                const double wi = w * zm.y, wj = w * mi[r];
                                ax[r] = fma(wi, dx, ax[r]);
                                ay[r] = fma(wi, dy, ay[r]);
                                az[r] = fma(wi, dzi, az[r]);
                                vx = fma(wj, dx, vx);
                                vy = fma(wj, dy, vy);
                                vz = fma(wj, dzi, vz);
                            }
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

template <int R, int THREADS, int TJ, int STAGES, int MINB, int UNR, int M>
__global__ void __launch_bounds__(THREADS, MINB) force_s1r2nl_f64_sym_kernel(const SymLaunchArgs sa, const S1R2Consts kc) {
    static_assert(M >= 3, "IS_PERIODIC >= 2");
    constexpr int NWARPS = THREADS / 32;
    constexpr int JB = 16;
    constexpr int IB = THREADS * R;
    constexpr int WB = sym_window_tiles(MINB, sym_base_f64(NWARPS, STAGES, TJ), 8, TJ);
    static_assert(THREADS == TJ && TJ % 32 == 0 && IB % TJ == 0 && TJ % JB == 0, "shape");
    const R3LaunchArgs &a = sa.a;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    JRec64 *tiles = reinterpret_cast<JRec64 *>(smem_raw);
    TileInfo64 *tinfo_s = reinterpret_cast<TileInfo64 *>(smem_raw + (size_t)STAGES * TJ * sizeof(JRec64));
    WarpBounds64 *wb_s = reinterpret_cast<WarpBounds64 *>(tinfo_s + STAGES);
    double *slots = reinterpret_cast<double *>(wb_s + NWARPS);  // [2][NWARPS][3][TJ]
    double *soa = slots + 2 * NWARPS * 3 * TJ;                  // staged (x,y) / (z,m) copy of the current symmetric tile
    double *jacc = soa + 8 * TJ;                                // [WB][3][TJ]: j-side sums of the current window (pair_r3_sym.cuh)
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
    const JRec64 *__restrict__ jrec = static_cast<const JRec64 *>(a.jrec);
    const TileInfo64 *__restrict__ tinfo = static_cast<const TileInfo64 *>(a.tinfo);
    constexpr uint32_t TILE_TX = TJ * sizeof(JRec64) + sizeof(TileInfo64);

    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], NWARPS);
        }
        fence_mbar_init();
    }
    __syncthreads();
    const WarpBounds64 *__restrict__ wb = wb_s + warp;
    S1R2SymRegs k;
    k.L = kc.L;
    k.LM1 = ((double)(M - 1)) * kc.L;  // (T)m * L as the reference forms it
    k.LM = ((double)M) * kc.L;
    asm volatile("mov.b64 %0, %0;" : "+d"(k.LM));  // as pair_s1r2.cuh: keep M*L in a vector register for the sign select
    k.cut_bits = __double_as_longlong(kc.cut);
    k.cutneg_bits = (unsigned long long)k.cut_bits | 0x8000000000000000ull;
    const double L = k.L, LM1 = k.LM1, LM = k.LM;
    const long long cut_bits = k.cut_bits;
    const unsigned long long cutneg_bits = k.cutneg_bits;
    int nsym = 0;
    int K0 = 0;  // tiles streamed so far by this CTA (pipeline stage / parity bookkeeping, pair_r3_sym.cuh)
    const uint64_t keep = l2_policy_keep();  // the i-side sums come back within milliseconds: ask L2 to hold them (ptx_helpers.cuh)

    for (int w0 = TA; w0 < TB; w0 += WB) {
    const int w1 = min(w0 + WB, TB);
#pragma unroll 4
    for (int q = 0; q < WB * 3; ++q) jacc[q * TJ + tid] = 0.0;
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
            tma_load_1d(tiles + (size_t)(K % STAGES) * TJ, jrec + (size_t)(t0 + t) * TJ, TJ * sizeof(JRec64), &full[K % STAGES]);
            tma_load_1d(tinfo_s + (K % STAGES), tinfo + (t0 + t), sizeof(TileInfo64), &full[K % STAGES]);
        }
    }
    double *__restrict__ fp = static_cast<double *>(a.fpart) + (size_t)jc * 3 * a.fstride;

    double xi[R], yi[R], zi[R], mi[R], ax[R], ay[R], az[R];
    {
        // planar bounds (x,y box and cylindrical radius): the tile bounds of this topology are planar too (pack kernel)
        double lo[2] = {1e300, 1e300}, hi[2] = {-1e300, -1e300}, rlo = 1e300, rhi = 0.0, smx = 0.0;
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int il0 = ib * IB + r * THREADS + tid;
            const int il = il0 < a.n_i ? il0 : a.n_i - 1;
            const JRec64 me = jrec[a.id_min + il];
            xi[r] = me.x; yi[r] = me.y; zi[r] = me.z;
            mi[r] = il0 < a.n_i ? me.m : 0.0;  // a clamped duplicate must not act on the j side
            if (first || il0 >= a.n_i) {
                ax[r] = ay[r] = az[r] = 0.0;
            } else {  // sums of the earlier windows of this chunk
                ax[r] = ld_keep(fp + il0, keep);
                ay[r] = ld_keep(fp + a.fstride + il0, keep);
                az[r] = ld_keep(fp + 2 * (size_t)a.fstride + il0, keep);
            }
            lo[0] = fmin(lo[0], me.x); hi[0] = fmax(hi[0], me.x);
            lo[1] = fmin(lo[1], me.y); hi[1] = fmax(hi[1], me.y);
            const double rr = sqrt(me.x * me.x + me.y * me.y);
            rlo = fmin(rlo, rr); rhi = fmax(rhi, rr);
            smx = fmax(smx, me.s);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                lo[q] = fmin(lo[q], __shfl_xor_sync(0xffffffffu, lo[q], o));
                hi[q] = fmax(hi[q], __shfl_xor_sync(0xffffffffu, hi[q], o));
            }
            rlo = fmin(rlo, __shfl_xor_sync(0xffffffffu, rlo, o));
            rhi = fmax(rhi, __shfl_xor_sync(0xffffffffu, rhi, o));
            smx = fmax(smx, __shfl_xor_sync(0xffffffffu, smx, o));
        }
        if (lane == 0) {
            WarpBounds64 &wbw = wb_s[warp];
            wbw.lo[0] = lo[0]; wbw.lo[1] = lo[1]; wbw.lo[2] = 0.0;
            wbw.hi[0] = hi[0]; wbw.hi[1] = hi[1]; wbw.hi[2] = 0.0;
            wbw.rlo = rlo; wbw.rhi = rhi; wbw.smax = smx * 1.0000001; wbw.pad = 0.0;
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
            tma_load_1d(tiles + (size_t)sp * TJ, jrec + (size_t)(t0 + t - 1 + STAGES) * TJ, TJ * sizeof(JRec64), &full[sp]);
            tma_load_1d(tinfo_s + sp, tinfo + (t0 + t - 1 + STAGES), sizeof(TileInfo64), &full[sp]);
        }
        mbar_wait(&full[s], ph);
        const JRec64 *__restrict__ T = tiles + (size_t)s * TJ;
        const int cls = sym_tile_class(*rule, t0 + t);  // CTA-uniform
        if (cls != 0) {
            const double smax = T[0].smax;
            bool far;
            {
                const TileInfo64 *__restrict__ ti = tinfo_s + s;
                double gap2 = 0.0;
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    const double gp = fmax(fmax(wb->lo[q] - ti->hi[q], ti->lo[q] - wb->hi[q]), 0.0);
                    gap2 = fma(gp, gp, gap2);
                }
                const double rg = fmax(wb->rlo - ti->rhi, ti->rlo - wb->rhi);
                const double b = (wb->smax + smax) * 1.000001;
                far = (gap2 > b * b) || (rg > b);
            }
            int thr[R];
#pragma unroll
            for (int r = 0; r < R; ++r) thr[r] = -1;  // far tiles: nothing is ever flagged
            if (!far) {
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    int il = ib * IB + r * THREADS + tid;
                    il = il < a.n_i ? il : a.n_i - 1;
                    const double b = jrec[a.id_min + il].s + smax;
                    thr[r] = __double2hiint(b * b) + 1;  // conservative: r2 < b*b => hi(r2) <= hi(b*b)
                }
            }
            if (cls == 1) {
                // ---- the i-block's own tiles: one-sided evaluation (the loop of pair_s1r2.cuh) ----
                if (far) {
                    int ymin = 0;
                    (void)ymin;
#pragma unroll 1
                    for (int jj = 0; jj < TJ; ++jj) {
                        const double2 xy = *reinterpret_cast<const double2 *>(&T[jj].x);
                        const double2 zm = *reinterpret_cast<const double2 *>(&T[jj].z);
#pragma unroll
                        for (int r = 0; r < R; ++r) STEPS_S1R2_PAIR(false)
                    }
                } else {
                    for (int j0 = 0; j0 < TJ; j0 += JB) {
                        int ymin = 0x7fffffff;
#pragma unroll 1
                        for (int jj = 0; jj < JB; ++jj) {
                            const double2 xy = *reinterpret_cast<const double2 *>(&T[j0 + jj].x);
                            const double2 zm = *reinterpret_cast<const double2 *>(&T[j0 + jj].z);
#pragma unroll
                            for (int r = 0; r < R; ++r) STEPS_S1R2_PAIR(true)
                        }
                        if (ymin == 0) {
#pragma unroll
                            for (int r = 0; r < R; ++r) {
                                int il = ib * IB + r * THREADS + tid;
                                il = il < a.n_i ? il : a.n_i - 1;
                                const double3 f = near_pairs_s1r2_f64(T + j0, JB, xi[r], yi[r], zi[r], jrec[a.id_min + il].s, thr[r], L);
                                ax[r] += f.x; ay[r] += f.y; az[r] += f.z;
                            }
                        }
                    }
                }
            } else {
                // ---- symmetric tile: systolic visit of 32 records per group ----
                double *__restrict__ slot = slots + ((size_t)(nsym & 1) * NWARPS + warp) * 3 * TJ;
                {
                    const double2 xy = *reinterpret_cast<const double2 *>(&T[tid].x);
                    const double2 zm = *reinterpret_cast<const double2 *>(&T[tid].z);
                    double2 *__restrict__ sxy = reinterpret_cast<double2 *>(soa) + (tid >> 5) * 64 + (tid & 31);
                    double2 *__restrict__ szm = sxy + 2 * TJ;
                    sxy[0] = xy; sxy[32] = xy;
                    szm[0] = zm; szm[32] = zm;
                }
                __syncthreads();
                if (far)
                    sym_tile_s1r2<R, TJ, THREADS, false, UNR, M>(T, soa, lane, tid, xi, yi, zi, mi, ax, ay, az, thr, slot, jrec, a.id_min, a.n_i, ib, k);
                else
                    sym_tile_s1r2<R, TJ, THREADS, true, UNR, M>(T, soa, lane, tid, xi, yi, zi, mi, ax, ay, az, thr, slot, jrec, a.id_min, a.n_i, ib, k);
                __syncthreads();
                {
                    // the warps' sums in warp order onto the window's accumulator (blocks of the superblock arrive in block order)
                    const double *__restrict__ sb = slots + (size_t)(nsym & 1) * NWARPS * 3 * TJ;
                    double *__restrict__ ja = jacc + (size_t)(t0 + t - w0) * 3 * TJ + tid;
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        double v = 0.0;
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
        double *__restrict__ gp = static_cast<double *>(sa.gpart) + (size_t)gs * 3 * sa.n_pad + (size_t)w0 * TJ + tid;
        for (int tl = 0; tl < w1 - w0; ++tl) {
#pragma unroll
            for (int c = 0; c < 3; ++c) __stcs(gp + (size_t)c * sa.n_pad + (size_t)tl * TJ, jacc[(tl * 3 + c) * TJ + tid]);
        }
    }
    }  // windows
}

}  // namespace steps
