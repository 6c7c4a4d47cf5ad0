// pair_s1r2.cuh -- tuned FP64 pair kernel of the S^1xR^2 slab, NOLOOKUP build, IS_PERIODIC >= 2
// (BASELINE.json configs[3]: N = 4M, periodic image sum along z, open transverse boundaries).
//
// Replaces the multi-image branch of ForceKernel_periodic_z (reference forces_cuda.cu:709-743; CPU forces.cc:1262-1290):
//     for m = -M .. M (M = IS_PERIODIC+1):  dz_m = dz + m*L;  if |dz_m| <= (M-0.4)*L:  F_i += m_j w(r_m, s_i+s_j) (dx, dy, dz_m)
// Same decomposition as pair_r3.cuh ((i-block x j-chunk) CTAs, TMA-staged j tiles, register-blocked i-particles,
// deterministic chunk reduction) and the same 15-instruction far-field pair math per image.  What is specific here:
//   * with z in [0, L) (the integrator wraps it, step.cc:151-180) |dz| < L, so images |m| <= M-2 are always inside the cut,
//     m = +-(M-1) are inside for part of the dz range, and at most ONE of m = +-M is (the one opposite in sign to dz).
//     The kernel therefore evaluates 2M image SLOTS instead of 2M+1 images: 2M-3 unconditional, two masked by the
//     reference's exact predicate |dz_m| <= cut, and one sign-selected slot for +-M, also masked by the exact predicate.
//     Masking = the slot's rsqrt seed is replaced by a denormal, which makes its weight exactly 0 (no branch, no inf*0).
//     The engine only takes this kernel when every z is inside [0, L) (checked on the device by the pack kernel);
//     otherwise the exact-branch kernel of pair_generic.cuh runs.
//   * dx^2+dy^2 is shared by the slots; the x,y accumulation uses the slot sum W = sum_m w_m (2 DFMA per pair instead of
//     2 per image); z accumulates per slot.
//   * only the slots m = 0, +-1 can come within a softening length (s_i+s_j << L): they carry the integer near-test of
//     pair_r3.cuh and flagged (pair, image) evaluations are redone with the reference's exact branches in a rare slow path.
#pragma once
#include "pair_generic.cuh"

namespace steps {

struct S1R2Consts {
    double L;      // period along z
    double cut;    // (T)ewald_cut * L, the reference's image cut (forces.cc:1274: <=)
    int M;         // ewald_max = IS_PERIODIC + 1
};

// exact re-evaluation of the flagged nearest-image evaluations of one sub-block (slots m = -1, 0, +1)
__device__ __noinline__ double3 near_pairs_s1r2_f64(const JRec64 *__restrict__ T, int nj, double xi, double yi, double zi, double si, int thr,
                                                     double L) {
    double fx = 0.0, fy = 0.0, fz = 0.0;
    for (int jj = 0; jj < nj; ++jj) {
        const JRec64 q = T[jj];
        const double dx = q.x - xi;
        const double dy = q.y - yi;
        const double dz = q.z - zi;
        const double dxy2 = fma(dy, dy, dx * dx);
#pragma unroll
        for (int m = -1; m <= 1; ++m) {
            const double dzi = dz + ((double)m) * L;
            const double r2 = fma(dzi, dzi, dxy2);
            if (__double2hiint(r2) <= thr) {
                const double w = q.m * softened_w<double>(sqrt(r2), si + q.s);
                fx = fma(w, dx, fx);
                fy = fma(w, dy, fy);
                fz = fma(w, dzi, fz);
            }
        }
    }
    return make_double3(fx, fy, fz);
}

// one image slot: unit-mass weight w' = r_m^-3 of the image at z-offset DZI, accumulated into Wsum = sum w' and Zsum = sum w' dz_m
// (the mass multiplies once per pair).  The series coefficients are immediates, so the slot reads no j-registers.
// YH_EXPR turns the raw seed high word into the masked one.
#define STEPS_S1R2_SLOT(FIRST, DZI, YH_EXPR)                                           \
    {                                                                                  \
        const double dzi_ = (DZI);                                                     \
        const double r2 = fma(dzi_, dzi_, dxy2);                                       \
        int yh = __double2hiint(rsqrt_seed(r2));                                       \
        YH_EXPR;                                                                       \
        const double y0 = __hiloint2double(yh, __double2loint(dead));                  \
        const double tt = y0 * y0;                                                     \
        const double e = fma(-r2, tt, 1.0);                                            \
        const double c = tt * y0;                                                      \
        double p = fma(e, 1.875, 1.5);                                                 \
        p = fma(p, e, 1.0);                                                            \
        const double w = c * p;                                                        \
        if (FIRST) {                                                                   \
            Wsum = w;                                                                  \
            Zsum = w * dzi_;                                                           \
        } else {                                                                       \
            Wsum += w;                                                                 \
            Zsum = fma(w, dzi_, Zsum);                                                 \
        }                                                                              \
        dead = tt;                                                                     \
    }

// |v| <= cut for doubles of known sign, on the integer pipe: positive doubles order like their bit patterns
__device__ __forceinline__ bool le_pos(double v, long long cut_bits) { return __double_as_longlong(v) <= cut_bits; }
__device__ __forceinline__ bool le_neg(double v, unsigned long long cutneg_bits) { return (unsigned long long)__double_as_longlong(v) <= cutneg_bits; }

// all image slots of one (i, j) pair.  NEARCHK: the nearest-image slots carry the integer softening test (near tiles);
// far tiles (every pair provably farther apart in the x,y plane than any s_i + s_j) skip it.
#define STEPS_S1R2_PAIR(NEARCHK)                                                                                         \
    {                                                                                                                    \
        const double dx = xy.x - xi[r];                                                                                  \
        const double dy = xy.y - yi[r];                                                                                  \
        const double dz = zm.x - zi[r];                                                                                  \
        double dead = dx * dx; /* the seed's low word comes from a dead value (any low word will do, see pair_r3.cuh) */ \
        const double dxy2 = fma(dy, dy, dead);                                                                           \
        double Wsum, Zsum;                                                                                               \
        /* nearest-image candidates (m = 0, -1, +1): always inside the cut, may be inside a softening length */          \
        STEPS_S1R2_SLOT(true, dz, { if (NEARCHK) { yh = (__double2hiint(r2) <= thr[r]) ? 0 : yh; ymin = min(ymin, yh); } })       \
        STEPS_S1R2_SLOT(false, dz - L, { if (NEARCHK) { yh = (__double2hiint(r2) <= thr[r]) ? 0 : yh; ymin = min(ymin, yh); } })  \
        STEPS_S1R2_SLOT(false, dz + L, { if (NEARCHK) { yh = (__double2hiint(r2) <= thr[r]) ? 0 : yh; ymin = min(ymin, yh); } })  \
        /* always inside the cut, never softened */                                                                      \
        _Pragma("unroll") for (int m = 2; m <= M - 2; ++m) {                                                             \
            STEPS_S1R2_SLOT(false, dz + ((double)m) * L, {})                                                             \
            STEPS_S1R2_SLOT(false, dz - ((double)m) * L, {})                                                             \
        }                                                                                                                \
        /* m = +-(M-1): the reference's exact predicate |dz_m| <= cut; the sign of dz_m is known because |dz| < L */     \
        STEPS_S1R2_SLOT(false, dz + LM1, { yh = le_pos(dzi_, cut_bits) ? yh : 0; })                                      \
        STEPS_S1R2_SLOT(false, dz - LM1, { yh = le_neg(dzi_, cutneg_bits) ? yh : 0; })                                   \
        /* m = +-M: only the image opposite in sign to dz can be inside the cut */                                       \
        STEPS_S1R2_SLOT(false, dz - copysign(LM, dz), { yh = le_pos(fabs(dzi_), cut_bits) ? yh : 0; })                   \
        const double mW = zm.y * Wsum;                                                                                   \
        ax[r] = fma(mW, dx, ax[r]);                                                                                      \
        ay[r] = fma(mW, dy, ay[r]);                                                                                      \
        az[r] = fma(zm.y, Zsum, az[r]);                                                                                  \
    }

template <int R, int THREADS, int TJ, int STAGES, int MINB, int M>
__global__ void __launch_bounds__(THREADS, MINB) force_s1r2nl_f64_kernel(const R3LaunchArgs a, const S1R2Consts k) {
    static_assert(M >= 3, "IS_PERIODIC >= 2");
    if (a.gate && *a.gate != a.gate_value) return;  // some z outside [0, L): the exact-branch kernel handles this call
    constexpr int NWARPS = THREADS / 32;
    constexpr int JB = 16;
    static_assert(TJ % JB == 0, "tile must be a whole number of sub-blocks");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    JRec64 *tiles = reinterpret_cast<JRec64 *>(smem_raw);
    TileInfo64 *tinfo_s = reinterpret_cast<TileInfo64 *>(smem_raw + (size_t)STAGES * TJ * sizeof(JRec64));
    WarpBounds64 *wb_s = reinterpret_cast<WarpBounds64 *>(tinfo_s + STAGES);
    uint64_t *full = reinterpret_cast<uint64_t *>(wb_s + NWARPS);
    uint64_t *empty = full + STAGES;

    const int tid = threadIdx.x;
    const int jc = blockIdx.x / a.n_ib;
    const int ib = blockIdx.x - jc * a.n_ib;
    const int t0 = jc * a.tiles_per_chunk;
    const int t1 = min(t0 + a.tiles_per_chunk, a.n_tiles);
    const int nt = t1 - t0;
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
    if (tid == 0) {
        const int npre = nt < STAGES ? nt : STAGES;
        for (int t = 0; t < npre; ++t) {
            mbar_arrive_expect_tx(&full[t], TILE_TX);
            tma_load_1d(tiles + (size_t)t * TJ, jrec + (size_t)(t0 + t) * TJ, TJ * sizeof(JRec64), &full[t]);
            tma_load_1d(tinfo_s + t, tinfo + (t0 + t), sizeof(TileInfo64), &full[t]);
        }
    }

    double xi[R], yi[R], zi[R], ax[R], ay[R], az[R];
    float si_up[R];
    {
        // bounds of this warp's i-particles in the x,y plane (box and cylindrical radius): the tile bounds written by the
        // pack kernel are planar too for this topology, so the classification below never looks at z
        double lo[2] = {1e300, 1e300}, hi[2] = {-1e300, -1e300}, rlo = 1e300, rhi = 0.0;
        float smx = 0.f;
#pragma unroll
        for (int r = 0; r < R; ++r) {
            int il = ib * (THREADS * R) + r * THREADS + tid;
            il = il < a.n_i ? il : a.n_i - 1;
            const JRec64 me = jrec[a.id_min + il];
            xi[r] = me.x; yi[r] = me.y; zi[r] = me.z;
            si_up[r] = __double2float_ru(me.s);
            ax[r] = ay[r] = az[r] = 0.0;
            lo[0] = fmin(lo[0], me.x); hi[0] = fmax(hi[0], me.x);
            lo[1] = fmin(lo[1], me.y); hi[1] = fmax(hi[1], me.y);
            const double rr = sqrt(me.x * me.x + me.y * me.y);
            rlo = fmin(rlo, rr); rhi = fmax(rhi, rr);
            smx = fmaxf(smx, si_up[r]);
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
            smx = fmaxf(smx, __shfl_xor_sync(0xffffffffu, smx, o));
        }
        if ((tid & 31) == 0) {
            WarpBounds64 &wb = wb_s[tid >> 5];
            wb.lo[0] = lo[0]; wb.lo[1] = lo[1]; wb.lo[2] = 0.0;
            wb.hi[0] = hi[0]; wb.hi[1] = hi[1]; wb.hi[2] = 0.0;
            wb.rlo = rlo; wb.rhi = rhi; wb.smax = (double)smx; wb.pad = 0.0;
        }
        __syncwarp();
    }
    const WarpBounds64 *__restrict__ wb = wb_s + (tid >> 5);
    const double L = k.L;
    const double LM1 = ((double)(M - 1)) * L;  // (T)m * L as the reference forms it
    double LM = ((double)M) * L;
    asm volatile("mov.b64 %0, %0;" : "+d"(LM));  // keep M*L in a vector register: the sign-select below is then one LOP3 on its high word
    const long long cut_bits = __double_as_longlong(k.cut);
    const unsigned long long cutneg_bits = (unsigned long long)cut_bits | 0x8000000000000000ull;

    for (int t = 0; t < nt; ++t) {
        const int s = t % STAGES;
        const uint32_t ph = (uint32_t)(t / STAGES) & 1u;
        if (tid == 0 && t >= 1 && (t - 1 + STAGES) < nt) {
            const int sp = (t - 1) % STAGES;
            const uint32_t php = (uint32_t)((t - 1) / STAGES) & 1u;
            mbar_wait(&empty[sp], php);
            mbar_arrive_expect_tx(&full[sp], TILE_TX);
            tma_load_1d(tiles + (size_t)sp * TJ, jrec + (size_t)(t0 + t - 1 + STAGES) * TJ, TJ * sizeof(JRec64), &full[sp]);
            tma_load_1d(tinfo_s + sp, tinfo + (t0 + t - 1 + STAGES), sizeof(TileInfo64), &full[sp]);
        }
        mbar_wait(&full[s], ph);
        const JRec64 *__restrict__ T = tiles + (size_t)s * TJ;
        const double smax = T[0].smax;
        bool far;
        {
            const TileInfo64 *__restrict__ ti = tinfo_s + s;
            double gap2 = 0.0;
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const double g = fmax(fmax(wb->lo[q] - ti->hi[q], ti->lo[q] - wb->hi[q]), 0.0);
                gap2 = fma(g, g, gap2);
            }
            const double rg = fmax(wb->rlo - ti->rhi, ti->rlo - wb->rhi);
            const double b = (wb->smax + smax) * 1.000001;
            far = (gap2 > b * b) || (rg > b);
        }
        if (far) {
            int thr[1] = {0};
            int ymin = 0;
            (void)thr; (void)ymin;
#pragma unroll 1
            for (int jj = 0; jj < TJ; ++jj) {
                const double2 xy = *reinterpret_cast<const double2 *>(&T[jj].x);
                const double2 zm = *reinterpret_cast<const double2 *>(&T[jj].z);
#pragma unroll
                for (int r = 0; r < R; ++r) STEPS_S1R2_PAIR(false)
            }
        } else {
            int thr[R];
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const double b = (double)si_up[r] + smax;
                thr[r] = __double2hiint(b * b) + 1;  // conservative: r2 < b*b => hi(r2) <= hi(b*b)
            }
            for (int j0 = 0; j0 < TJ; j0 += JB) {
                int ymin = 0x7fffffff;  // becomes 0 iff a nearest-image evaluation of this sub-block was flagged
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
                        int il = ib * (THREADS * R) + r * THREADS + tid;
                        il = il < a.n_i ? il : a.n_i - 1;
                        const double3 f = near_pairs_s1r2_f64(T + j0, JB, xi[r], yi[r], zi[r], jrec[a.id_min + il].s, thr[r], L);
                        ax[r] += f.x; ay[r] += f.y; az[r] += f.z;
                    }
                }
            }
        }
        __syncwarp();
        if ((tid & 31) == 0) mbar_arrive(&empty[s]);
    }

    double *__restrict__ fp = static_cast<double *>(a.fpart) + (size_t)jc * 3 * a.fstride;
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
