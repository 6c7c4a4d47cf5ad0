// pair_r3_sym.cuh -- action-reaction (Newton's third law) R^3 FP64 pair kernel.
//
// Same force law as pair_r3.cuh / forces() (forces.cc:510-577): F_i = sum_j m_j w(r_ij, s_i+s_j) (x_j - x_i).
// w and d = x_j - x_i are symmetric / antisymmetric in (i, j), so one evaluation of g = w(r_ij) and d serves
// BOTH particles:   F_i += m_j g d    and    F_j -= m_i g d.
// The one-sided kernel spends 15 FP64-pipe instructions per DIRECTED pair (the pipe that bounds this path); this
// kernel spends 20 per UNORDERED pair = 10 per directed pair:
//     3 DADD (d), DMUL + 2 DFMA (r2), MUFU seed, t = y0^2, e = fma(-r2,t,1), c = t y0,
//     q = fma(fma(e,1.875,1.5),e,1), g = c q, wi = g m_j, wj = g m_i, 3 DFMA into F_i, 3 DFMA into the j accumulator.
//
// How the j-side sum is formed without atomics (deterministic):
//   * a warp holds 32*R i-particles in registers (R per lane) and walks a j-tile in groups of 32 records;
//   * inside a group the 32 records VISIT the lanes systolically: at step s lane l works on record (l+s) mod 32 and
//     owns that record's accumulator (3 doubles), which rotates one lane per step by warp shuffle; after 32 steps each
//     record has met all 32*R i-particles of the warp and its accumulator is back in lane l = record index;
//   * the warps of a CTA deposit their accumulators in shared memory and the CTA adds them in warp order into a j-side
//     accumulator that STAYS in shared memory for a window of WB tiles while the CTA works through the SB i-blocks of its
//     superblock one after the other (the i-side accumulators of a block travel to the partial-sum buffer and back between
//     windows: 18 KB per block, L2 traffic).  One partial row per (SUPERBLOCK, j-tile) reaches global memory -- SB times fewer
//     rows than one per (i-block, j-tile) -- and a reduce kernel adds the rows in superblock order.  Every sum has a fixed
//     order (warps, blocks of the superblock, superblocks): bitwise deterministic.
// Which (i-block, j-tile) combinations are evaluated is given by one SymRule per i-block: the tiles holding the
// i-block's own particles are evaluated one-sidedly (both directions appear there, self pair included), tiles in
// the rule's symmetric ranges are evaluated once for both sides, all other tiles are skipped (their pairs are
// evaluated by the CTA of the other block).  The host builds the rules so that every unordered pair of the whole
// job is covered exactly once (engine.cu: build_sym_rules; multi-GPU: ring assignment of block pairs).
#pragma once
#include "pair_r3.cuh"

namespace steps {

constexpr int SYM_MAX_RANGES = 5;
struct __align__(16) SymRule {
    int diag_lo, diag_hi;  // tiles [diag_lo, diag_hi): the i-block's own particles (one-sided evaluation)
    int n_sym;             // number of symmetric tile ranges
    int sym_lo[SYM_MAX_RANGES], sym_hi[SYM_MAX_RANGES];
    int pad[3];
};
static_assert(sizeof(SymRule) == 64, "SymRule layout");

// 0 = skip, 1 = one-sided (diagonal), 2 = symmetric
__host__ __device__ __forceinline__ int sym_tile_class(const SymRule &r, int t) {
    if (t >= r.diag_lo && t < r.diag_hi) return 1;
    for (int k = 0; k < r.n_sym; ++k)
        if (t >= r.sym_lo[k] && t < r.sym_hi[k]) return 2;
    return 0;
}

struct SymLaunchArgs {
    R3LaunchArgs a;        // jrec/tinfo/fpart/id_min/n_i/tiles_per_chunk/n_tiles/n_j/fstride as for the one-sided kernel;
                           // a.n_ib = number of local i-blocks of the call
    const SymRule *rules;  // one per local i-block (index 0 = the block starting at id_min)
    void *gpart;           // j-side partial rows: [superblock of the pass][3][n_pad], REAL of the build
    int b0;                // first local superblock of this pass
    int n_pad;             // n_tiles * TJ
    int sb;                // i-blocks per superblock
    const int2 *order;     // one entry per CTA of this pass: (superblock within the pass, j-chunk), heaviest CTAs first
};

// tile range [ta, tb) a rule touches inside [c0, c1): hull of the own tiles and the symmetric ranges (empty: tb <= ta)
__host__ __device__ __forceinline__ void sym_hull(const SymRule &r, int c0, int c1, int &ta, int &tb) {
    ta = 0x7fffffff;
    tb = -1;
    {
        const int lo = r.diag_lo > c0 ? r.diag_lo : c0, hi = r.diag_hi < c1 ? r.diag_hi : c1;
        if (lo < hi) { ta = lo < ta ? lo : ta; tb = hi > tb ? hi : tb; }
    }
    for (int k = 0; k < r.n_sym; ++k) {
        const int lo = r.sym_lo[k] > c0 ? r.sym_lo[k] : c0, hi = r.sym_hi[k] < c1 ? r.sym_hi[k] : c1;
        if (lo < hi) { ta = lo < ta ? lo : ta; tb = hi > tb ? hi : tb; }
    }
}

// Pipeline bookkeeping shared by the action-reaction kernels: tiles are numbered by a counter K that runs over the whole life
// of the CTA (all segments), stage = K % STAGES, parity of the full barrier = (K / STAGES) & 1.  The producer may refill the
// stage of tile K once the consumers have released tile K - STAGES.
template <int STAGES>
__device__ __forceinline__ void sym_wait_stage_free(uint64_t *empty, int K) {
    if (K >= STAGES) mbar_wait(&empty[K % STAGES], (uint32_t)(K / STAGES - 1) & 1u);
}

// exact softened kernel for one flagged pair (rare; out of line)
__device__ __noinline__ double sym_exact_w(double r2, double beta) { return softened_w<double>(sqrt(r2), beta); }

#define STEPS_PAIR_SYM_CORE(XJ, YJ, ZJ, r, YH_EXPR)                                        \
    const double dx = (XJ) - xi[r];                                                        \
    const double dy = (YJ) - yi[r];                                                        \
    const double dz = (ZJ) - zi[r];                                                        \
    const double dx2 = dx * dx;                                                            \
    double r2 = fma(dy, dy, dx2);                                                          \
    r2 = fma(dz, dz, r2);                                                                  \
    int yh = __double2hiint(rsqrt_seed(r2));                                               \
    YH_EXPR;                                                                               \
    const double y0 = __hiloint2double(yh, __double2loint(dx2));                           \
    const double tt = y0 * y0;                                                             \
    const double e = fma(-r2, tt, 1.0);                                                    \
    const double c = tt * y0;                                                              \
    double q = fma(e, 1.875, 1.5);                                                         \
    q = fma(q, e, 1.0);                                                                    \
    const double g = c * q;

// One symmetric (i-warp x j-tile) block.  CHECKED = false: every pair is provably outside the softening radius (no
// per-pair test at all); true: pairs with hi(r2) <= thr[r] are masked out of the fast arithmetic (seed := 0 => g = 0
// exactly) and re-evaluated with the reference's exact branches before the accumulators rotate.
template <int R, int TJ, int THREADS, bool CHECKED, int UNR>
__device__ __forceinline__ void sym_tile(const JRec64 *__restrict__ T, const double *__restrict__ soa, int lane, int tid,
                                         const double (&xi)[R], const double (&yi)[R],
                                         const double (&zi)[R], const double (&mi)[R], double (&ax)[R], double (&ay)[R], double (&az)[R],
                                         const int (&thr)[R], double *__restrict__ slot, const JRec64 *__restrict__ jrec, int id_min,
                                         int n_i, int ib) {
    constexpr int IB = THREADS * R;
    for (int g0 = 0; g0 < TJ; g0 += 32) {
        double vx = 0.0, vy = 0.0, vz = 0.0;
        const JRec64 *__restrict__ G = T + g0;
        // per-lane records come from the staged copy of the tile: (x,y) pairs and (z,m) pairs as 16-byte elements, each group
        // of 32 records stored TWICE in a row of 64, so that lane l finds record (l+s) mod 32 at element l+s -- no wrap, no
        // index arithmetic (immediate offsets once the step loop is unrolled) and conflict-free 128-bit loads, where the
        // 64-byte AoS records would collide 16-way
        const double2 *__restrict__ SXY = reinterpret_cast<const double2 *>(soa) + (g0 * 2) + lane;
        const double2 *__restrict__ SZM = SXY + 2 * TJ;
        double2 nxy = SXY[0], nzm = SZM[0];
#pragma unroll UNR
        for (int s2 = 0; s2 < 32; ++s2) {
            const double2 xy = nxy, zm = nzm;
            nxy = SXY[s2 + 1];  // record of the next step (prefetched)
            nzm = SZM[s2 + 1];
            int ymin = 0x7fffffff;
#pragma unroll
            for (int r = 0; r < R; ++r) {
                STEPS_PAIR_SYM_CORE(xy.x, xy.y, zm.x, r, if (CHECKED) { yh = (__double2hiint(r2) <= thr[r]) ? 0 : yh; ymin = min(ymin, yh); })
                const double wi = g * zm.y;
                const double wj = g * mi[r];
                ax[r] = fma(wi, dx, ax[r]);
                ay[r] = fma(wi, dy, ay[r]);
                az[r] = fma(wi, dz, az[r]);
                vx = fma(wj, dx, vx);
                vy = fma(wj, dy, vy);
                vz = fma(wj, dz, vz);
            }
            if (CHECKED) {
                if (ymin == 0) {
                    // rare: some pair of this step lies inside the softening radius (or coincides): exact branches
                    const double sj = G[(lane + s2) & 31].s;
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        const double dx = xy.x - xi[r], dy = xy.y - yi[r], dz = zm.x - zi[r];
                        double r2 = dx * dx;
                        r2 = fma(dy, dy, r2);
                        r2 = fma(dz, dz, r2);
                        if (__double2hiint(r2) <= thr[r]) {
                            int il = ib * IB + r * THREADS + tid;
                            il = il < n_i ? il : n_i - 1;
                            const double w = sym_exact_w(r2, jrec[id_min + il].s + sj);
                            const double wi = w * zm.y, wj = w * mi[r];
                            ax[r] = fma(wi, dx, ax[r]);
                            ay[r] = fma(wi, dy, ay[r]);
                            az[r] = fma(wi, dz, az[r]);
                            vx = fma(wj, dx, vx);
                            vy = fma(wj, dy, vy);
                            vz = fma(wj, dz, vz);
                        }
                    }
                }
                __syncwarp();
            }
            // the accumulator follows its record: lane l takes over the record lane l+1 just worked on
            vx = __shfl_sync(0xffffffffu, vx, (lane + 1) & 31);
            vy = __shfl_sync(0xffffffffu, vy, (lane + 1) & 31);
            vz = __shfl_sync(0xffffffffu, vz, (lane + 1) & 31);
        }
        // after 32 rotations lane l holds the accumulator of record g0 + l
        slot[g0 + lane] = vx;
        slot[TJ + g0 + lane] = vy;
        slot[2 * TJ + g0 + lane] = vz;
    }
}

// Tiles of one window of the j-side accumulator that lives in shared memory (see the header comment): as many as fit beside the
// kernel's other buffers (`base` bytes) when `minb` CTAs share the 227 KB of an SM, 16 at most; elem = sizeof(REAL).
// Host and device evaluate the same function, so the launch's shared-memory size and the kernel's carve-up agree.
__host__ __device__ constexpr int sym_window_tiles(int minb, int base, int elem, int tj = 128) {
    const int room = 227 * 1024 / minb - 1024 - base;
    const int w = room / (3 * tj * elem);
    return w >= 16 ? 16 : w >= 8 ? 8 : w >= 4 ? 4 : 2;
}
// shared memory of the FP64 action-reaction kernels besides the window: staged tiles + tile bounds | warp bounds | 2 x per-warp
// accumulator slots | visiting copy of a tile | barriers
__host__ __device__ constexpr int sym_base_f64(int nwarps, int stages, int tj = 128) {
    return stages * (tj * 64 + 64) + nwarps * 80 + 2 * nwarps * 3 * tj * 8 + 8 * tj * 8 + 2 * stages * 8;
}

template <int R, int THREADS, int TJ, int STAGES, int MINB, int UNR>
__global__ void __launch_bounds__(THREADS, MINB) force_r3_f64_sym_kernel(const SymLaunchArgs sa) {
    constexpr int NWARPS = THREADS / 32;
    constexpr int JB = 16;
    constexpr int IB = THREADS * R;
    constexpr int WB = sym_window_tiles(MINB, sym_base_f64(NWARPS, STAGES, TJ), 8, TJ);
    static_assert(sizeof(JRec64) == 64 && sizeof(TileInfo64) == 64 && sizeof(WarpBounds64) == 80, "sym_base_f64");
    static_assert(THREADS == TJ && TJ % 32 == 0 && IB % TJ == 0, "shape");
    const R3LaunchArgs &a = sa.a;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    JRec64 *tiles = reinterpret_cast<JRec64 *>(smem_raw);
    TileInfo64 *tinfo_s = reinterpret_cast<TileInfo64 *>(smem_raw + (size_t)STAGES * TJ * sizeof(JRec64));
    WarpBounds64 *wb_s = reinterpret_cast<WarpBounds64 *>(tinfo_s + STAGES);
    double *slots = reinterpret_cast<double *>(wb_s + NWARPS);  // [2][NWARPS][3][TJ]
    double *soa = slots + 2 * NWARPS * 3 * TJ;  // double2 [2][TJ/32][64]: (x,y) and (z,m) of the current symmetric tile, see sym_tile
    double *jacc = soa + 8 * TJ;                // [WB][3][TJ]: j-side sums of the current window; column tid belongs to thread tid
    uint64_t *full = reinterpret_cast<uint64_t *>(jacc + (size_t)WB * 3 * TJ);
    uint64_t *empty = full + STAGES;

    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const int2 od = sa.order[blockIdx.x];
    const int gs = od.x;             // superblock within the pass
    const int jc = od.y;             // j-chunk
    const int ib_lo = (sa.b0 + gs) * sa.sb;
    const int ib_hi = min(ib_lo + sa.sb, a.n_ib);
    const int c0 = jc * a.tiles_per_chunk;
    const int c1 = min(c0 + a.tiles_per_chunk, a.n_tiles);
    // tile range of the whole superblock inside this chunk
    int TA = 0x7fffffff, TB = -1;
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
    int nsym = 0;  // symmetric tiles processed so far (selects the slot buffer)
    int K0 = 0;    // tiles streamed so far by this CTA (pipeline stage / parity bookkeeping)
    const uint64_t keep = l2_policy_keep();  // the i-side sums come back within milliseconds: ask L2 to hold them

    for (int w0 = TA; w0 < TB; w0 += WB) {
        const int w1 = min(w0 + WB, TB);
#pragma unroll 4
        for (int k = 0; k < WB * 3; ++k) jacc[k * TJ + tid] = 0.0;
        for (int ib = ib_lo; ib < ib_hi; ++ib) {
            const SymRule *__restrict__ rule = sa.rules + ib;  // read through L1 when needed: a register copy would cost 13 registers
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

            // (the softening length of the i-particles is needed in near tiles only: re-read from the record there, no registers)
            double xi[R], yi[R], zi[R], mi[R], ax[R], ay[R], az[R];
            {
                double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300}, rlo = 1e300, rhi = 0.0;
                double smx = 0.0;
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
                    lo[2] = fmin(lo[2], me.z); hi[2] = fmax(hi[2], me.z);
                    const double rr = sqrt(me.x * me.x + me.y * me.y + me.z * me.z);
                    rlo = fmin(rlo, rr); rhi = fmax(rhi, rr);
                    smx = fmax(smx, me.s);
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
                    for (int k = 0; k < 3; ++k) {
                        lo[k] = fmin(lo[k], __shfl_xor_sync(0xffffffffu, lo[k], o));
                        hi[k] = fmax(hi[k], __shfl_xor_sync(0xffffffffu, hi[k], o));
                    }
                    rlo = fmin(rlo, __shfl_xor_sync(0xffffffffu, rlo, o));
                    rhi = fmax(rhi, __shfl_xor_sync(0xffffffffu, rhi, o));
                    smx = fmax(smx, __shfl_xor_sync(0xffffffffu, smx, o));
                }
                if (lane == 0) {
                    WarpBounds64 &wbw = wb_s[warp];
                    wbw.lo[0] = lo[0]; wbw.lo[1] = lo[1]; wbw.lo[2] = lo[2];
                    wbw.hi[0] = hi[0]; wbw.hi[1] = hi[1]; wbw.hi[2] = hi[2];
                    wbw.rlo = rlo; wbw.rhi = rhi; wbw.smax = smx; wbw.pad = 0.0;
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
                        for (int k = 0; k < 3; ++k) {
                            const double gp = fmax(fmax(wb->lo[k] - ti->hi[k], ti->lo[k] - wb->hi[k]), 0.0);
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
                            // conservative: r2 < b*b  =>  hi(r2) <= hi(b*b); one extra ulp of the high word for the rounding of b*b
                            thr[r] = __double2hiint(b * b) + 1;
                        }
                    }
                    if (cls == 1) {
                        // ---- the i-block's own tiles: one-sided evaluation, checked loop (pair_r3.cuh near branch) ----
                        for (int j0 = 0; j0 < TJ; j0 += JB) {
                            int ymin = 0x7fffffff;
#pragma unroll 1
                            for (int jj = 0; jj < JB; ++jj) {
                                const double2 xy = *reinterpret_cast<const double2 *>(&T[j0 + jj].x);
                                const double2 zm = *reinterpret_cast<const double2 *>(&T[j0 + jj].z);
#pragma unroll
                                for (int r = 0; r < R; ++r) {
                                    STEPS_PAIR_SYM_CORE(xy.x, xy.y, zm.x, r, yh = (__double2hiint(r2) <= thr[r]) ? 0 : yh; ymin = min(ymin, yh))
                                    const double wi = g * zm.y;
                                    ax[r] = fma(wi, dx, ax[r]);
                                    ay[r] = fma(wi, dy, ay[r]);
                                    az[r] = fma(wi, dz, az[r]);
                                }
                            }
                            if (ymin == 0) {
#pragma unroll
                                for (int r = 0; r < R; ++r) {
                                    int il = ib * IB + r * THREADS + tid;
                                    il = il < a.n_i ? il : a.n_i - 1;
                                    const double3 f = near_pairs_f64(T + j0, JB, xi[r], yi[r], zi[r], jrec[a.id_min + il].s, thr[r]);
                                    ax[r] += f.x; ay[r] += f.y; az[r] += f.z;
                                }
                            }
                        }
                    } else {
                        // ---- symmetric tile: systolic visit of 32 records per group ----
                        double *__restrict__ slot = slots + ((size_t)(nsym & 1) * NWARPS + warp) * 3 * TJ;
                        // staged copy of the tile's hot fields (layout: sym_tile).  Every warp left the previous symmetric tile's loop before the
                        // __syncthreads that preceded its row combine, so the buffer is free to overwrite here.
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
                            sym_tile<R, TJ, THREADS, false, UNR>(T, soa, lane, tid, xi, yi, zi, mi, ax, ay, az, thr, slot, jrec, a.id_min, a.n_i, ib);
                        else
                            sym_tile<R, TJ, THREADS, true, UNR>(T, soa, lane, tid, xi, yi, zi, mi, ax, ay, az, thr, slot, jrec, a.id_min, a.n_i, ib);
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
        }
        // the window's j-side sums: one row segment per (superblock, tile); streamed (read once, by the row reduction)
        {
            double *__restrict__ gp = static_cast<double *>(sa.gpart) + (size_t)gs * 3 * sa.n_pad + (size_t)w0 * TJ + tid;
            for (int tl = 0; tl < w1 - w0; ++tl) {
#pragma unroll
                for (int c = 0; c < 3; ++c) __stcs(gp + (size_t)c * sa.n_pad + (size_t)tl * TJ, jacc[(tl * 3 + c) * TJ + tid]);
            }
        }
    }
}

// j-side reduction of one pass: fsym[c][j] += sum over the pass's superblocks (in order) of the rows that hold a symmetric
// contribution for j's tile (a superblock's row holds one iff one of its i-blocks has the tile in a symmetric range).
// One thread per j.
template <typename T>
__global__ void reduce_sym_kernel(const T *__restrict__ gpart, const SymRule *__restrict__ rules, int b0, int nsb, int sb, int n_ib, int n_pad, int tj,
                                  T *__restrict__ fsym) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_pad) return;
    const int t = j / tj;
    T sx = 0, sy = 0, sz = 0;
    for (int g = 0; g < nsb; ++g) {
        const int ib_lo = (b0 + g) * sb, ib_hi = min(ib_lo + sb, n_ib);
        bool any = false;
        for (int ib = ib_lo; ib < ib_hi && !any; ++ib) any = sym_tile_class(rules[ib], t) == 2;
        if (any) {
            const T *__restrict__ row = gpart + (size_t)g * 3 * n_pad + j;
            sx += __ldcs(row);
            sy += __ldcs(row + (size_t)n_pad);
            sz += __ldcs(row + 2 * (size_t)n_pad);
        }
    }
    fsym[j] += sx;
    fsym[(size_t)n_pad + j] += sy;
    fsym[2 * (size_t)n_pad + j] += sz;
}

}  // namespace steps
