// pair_r3.cuh -- the R^3 (non-periodic, compactified) pair kernels, FP64 and FP32.
//
// Replaces ForceKernel (reference forces_cuda.cu:522-563) / the j-loop of forces() (forces.cc:535-556):
//     F_i = sum_{j=0}^{N-1} m_j * w(r_ij, s_i + s_j) * (x_j - x_i)
// with w the cubic-spline softened 1/r^3 of force_softening (forces.cc:52-87, forces_cuda.cu:466-519).
//
// Design (B200 / sm_100a):
//   * work unit = (i-block of THREADS*R particles) x (j-chunk of `tiles_per_chunk` tiles); one CTA per unit.
//     Splitting j gives units fine enough to fill 148 SMs evenly at any N (tail < 1/units-per-slot) and
//     partial sums are combined afterwards in fixed chunk order (deterministic, see reduce kernel).
//   * j-tiles (TJ records of 64 B / 32 B) are staged into shared memory by 1-D TMA bulk copies
//     (cp.async.bulk + mbarrier complete_tx), STAGES deep; one elected thread is the producer.
//   * each thread keeps R i-particles in registers; a warp walks the tile with broadcast 128-bit
//     shared loads, so one LDS serves 32*R pairs.
//   * far-field pair math runs 15 FP64-pipe instructions (the pipe that bounds this kernel):
//       3 DADD (d = xj - xi), DMUL+2 DFMA (r2), y0 = MUFU.RSQ64H(r2) [SFU, free],
//       t = y0*y0, e = fma(-r2,t,1), c = t*y0, p = fma(fma(1.875m,e,1.5m),e,m), w = c*p   (6)
//       3 DFMA accumulate.            m*r^-3 = m*y0^3*(1-e)^(-3/2), |e|<~2^-21, series error ~2.2*e^3 < 1e-18.
//     (m, 1.5m, 1.875m) are pre-staged per j so the mass multiply costs nothing.
//   * the softened branch (r < s_i+s_j) and the self pair (r=0) are detected WITHOUT touching the FP64
//     pipe: an integer compare of the high word of r2 against a conservative per-(i, tile) threshold
//     (s_i + max_{j in tile} s_j)^2.  Flagged pairs are masked out of the fast accumulation and
//     re-evaluated exactly (reference branch structure) in a rare deferred pass per sub-block.
#pragma once
#include "ptx_helpers.cuh"

namespace steps {

// One j-particle as staged for the FP64 kernels: 64 B, 16 B-aligned halves for LDS.128.
struct __align__(16) JRec64 {
    double x, y;        // LDS.128 #1
    double z, m;        // LDS.128 #2
    double m15, m1875;  // LDS.128 #3   1.5*m, 1.875*m
    double s, smax;     // slow path only: softening length, max softening over this record's tile
};
static_assert(sizeof(JRec64) == 64, "JRec64 layout");

// Bounds of the real particles of one j-tile (written by the pack kernel every step): axis-aligned box and
// range of |x| about the origin (the compactified R^3 load is centred on the origin and ordered by radius).
struct __align__(16) TileInfo64 {
    double lo[3], hi[3];
    double rlo, rhi;
};
static_assert(sizeof(TileInfo64) == 64, "TileInfo64 layout");

// FP32: 32 B.
struct __align__(16) JRec32 {
    float x, y, z, m;     // LDS.128 #1
    float s, smax, p0, p1;  // slow path only
};
static_assert(sizeof(JRec32) == 32, "JRec32 layout");

// Exact softened kernel, branch structure and operation order of force_softening_cuda
// (forces_cuda.cu:466-519).  Returns w (without the mass factor).
template <typename T>
__host__ __device__ __forceinline__ T softened_w(T r, T beta) {
    const T half_beta = beta * (T)0.5;
    const T r2 = r * r;
    const T r3 = r2 * r;
    T w;
    if (r >= beta) {
        w = (T)1.0 / r3;
    } else if (r > half_beta) {
        const T b2 = beta * beta, b3 = b2 * beta, b4 = b2 * b2, b5 = b4 * beta, b6 = b3 * b3;
        const T C0 = (T)(-32.0) / ((T)3.0 * b6);
        const T C1 = (T)(38.4) / b5;
        const T C2 = (T)(-48.0) / b4;
        const T C3 = (T)(64.0) / ((T)3.0 * b3);
        const T C4 = (T)(-1.0 / 15.0);
        w = C0 * r3 + C1 * r2 + C2 * r + C3 + C4 * ((T)1.0 / r3);
    } else {
        const T b2 = beta * beta, b3 = b2 * beta, b4 = b2 * b2, b5 = b4 * beta, b6 = b3 * b3;
        const T C0 = (T)(32.0) / b6;
        const T C1 = (T)(-38.4) / b5;
        const T C2 = (T)(32.0 / 3.0) / b3;
        w = C0 * r3 + C1 * r2 + C2;
    }
    return w;
}

struct R3LaunchArgs {
    const void *tinfo;  // TileInfo64* per j-tile (tuned FP64 kernel only)
    const void *jrec;   // JRec64* / JRec32*, padded to a whole number of tiles
    void *fpart;        // partial sums, [n_chunks][3][fstride]
    int id_min;         // first i of the call
    int n_i;            // number of i-particles in the call
    int n_ib;           // number of i-blocks
    int tiles_per_chunk;
    int n_tiles;        // total j tiles
    int n_j;            // number of real j-particles (N)
    int fstride;        // >= n_i
    const int *gate;    // optional device flag: the kernel returns at once unless *gate == gate_value (nullptr: always run)
    int gate_value;
};

// Slow path of the tuned kernels: the pairs of one sub-block that the integer test flagged (r2 below the
// conservative threshold), evaluated with the reference's exact branch structure.  Out of line on purpose: it
// runs for a tiny fraction of the sub-blocks and must not cost the hot loop registers or instruction cache.
__device__ __noinline__ double3 near_pairs_f64(const JRec64 *__restrict__ T, int nj, double xi, double yi, double zi, double si, int thr) {
    double fx = 0.0, fy = 0.0, fz = 0.0;
    for (int jj = 0; jj < nj; ++jj) {
        const JRec64 q = T[jj];
        const double dx = q.x - xi;
        const double dy = q.y - yi;
        const double dz = q.z - zi;
        double r2 = dx * dx;
        r2 = fma(dy, dy, r2);
        r2 = fma(dz, dz, r2);
        if (__double2hiint(r2) <= thr) {
            const double w = q.m * softened_w<double>(sqrt(r2), si + q.s);
            fx = fma(w, dx, fx);
            fy = fma(w, dy, fy);
            fz = fma(w, dz, fz);
        }
    }
    return make_double3(fx, fy, fz);
}

// bounds of the i-particles one warp holds in registers (kept in shared memory; read once per tile)
struct __align__(16) WarpBounds64 {
    double lo[3], hi[3];
    double rlo, rhi;
    double smax, pad;
};

// far-field pair: 15 FP64-pipe instructions + one MUFU, nothing else (see the header comment)
#define STEPS_PAIR_FAR_F64(XJ, YJ, ZJ, MJ, M15J, M1875J, r)                                       \
    {                                                                                             \
        const double dx = (XJ) - xi[r];                                                           \
        const double dy = (YJ) - yi[r];                                                           \
        const double dz = (ZJ) - zi[r];                                                           \
        const double dx2 = dx * dx;                                                               \
        double r2 = fma(dy, dy, dx2);                                                             \
        r2 = fma(dz, dz, r2);                                                                     \
        const double y0 = __hiloint2double(__double2hiint(rsqrt_seed(r2)), __double2loint(dx2));  \
        const double tt = y0 * y0;                                                                \
        const double e = fma(-r2, tt, 1.0);                                                       \
        const double c = tt * y0;                                                                 \
        double p = fma((M1875J), e, (M15J));                                                      \
        p = fma(p, e, (MJ));                                                                      \
        const double w = c * p;                                                                   \
        ax[r] = fma(w, dx, ax[r]);                                                                \
        ay[r] = fma(w, dy, ay[r]);                                                                \
        az[r] = fma(w, dz, az[r]);                                                                \
    }

template <int R, int THREADS, int TJ, int STAGES, int MINB, int UNROLL>
__global__ void __launch_bounds__(THREADS, MINB) force_r3_f64_kernel(const R3LaunchArgs a) {
    constexpr int NWARPS = THREADS / 32;
    constexpr int JB = 16;  // sub-block between slow-path checks (near tiles)
    static_assert(TJ % JB == 0 && JB % UNROLL == 0, "tile must be a whole number of sub-blocks");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    JRec64 *tiles = reinterpret_cast<JRec64 *>(smem_raw);
    TileInfo64 *tinfo_s = reinterpret_cast<TileInfo64 *>(smem_raw + (size_t)STAGES * TJ * sizeof(JRec64));
    WarpBounds64 *wb_s = reinterpret_cast<WarpBounds64 *>(tinfo_s + STAGES);
    uint64_t *full = reinterpret_cast<uint64_t *>(wb_s + NWARPS);
    uint64_t *empty = full + STAGES;

    const int tid = threadIdx.x;
    const int jc = blockIdx.x / a.n_ib;  // chunk-major: concurrently running CTAs stream the same j-chunk from L2
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

    // register-blocked i-particles.  The softening length is only needed (a) for the near thresholds,
    // where a float rounded UP is enough (conservative), and (b) exactly in the rare slow path, which re-reads it.
    double xi[R], yi[R], zi[R], ax[R], ay[R], az[R];
    float si_up[R];
    {
        double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300}, rlo = 1e300, rhi = 0.0;
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
            lo[2] = fmin(lo[2], me.z); hi[2] = fmax(hi[2], me.z);
            const double rr = sqrt(me.x * me.x + me.y * me.y + me.z * me.z);
            rlo = fmin(rlo, rr); rhi = fmax(rhi, rr);
            smx = fmaxf(smx, si_up[r]);
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
            smx = fmaxf(smx, __shfl_xor_sync(0xffffffffu, smx, o));
        }
        if ((tid & 31) == 0) {
            WarpBounds64 &wb = wb_s[tid >> 5];
            wb.lo[0] = lo[0]; wb.lo[1] = lo[1]; wb.lo[2] = lo[2];
            wb.hi[0] = hi[0]; wb.hi[1] = hi[1]; wb.hi[2] = hi[2];
            wb.rlo = rlo; wb.rhi = rhi; wb.smax = (double)smx; wb.pad = 0.0;
        }
        __syncwarp();
    }
    const WarpBounds64 *__restrict__ wb = wb_s + (tid >> 5);

    for (int t = 0; t < nt; ++t) {
        const int s = t % STAGES;
        const uint32_t ph = (uint32_t)(t / STAGES) & 1u;
        // producer: refill the stage consumed one iteration ago (all warps have very likely left it)
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
        // warp-uniform classification: is every (i of this warp, j of this tile) pair provably farther apart than
        // any s_i + s_j?  Box distance or radial gap (reverse triangle inequality), with a 1e-6 safety margin.
        bool far;
        {
            const TileInfo64 *__restrict__ ti = tinfo_s + s;
            double gap2 = 0.0;
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const double g = fmax(fmax(wb->lo[k] - ti->hi[k], ti->lo[k] - wb->hi[k]), 0.0);
                gap2 = fma(g, g, gap2);
            }
            const double rg = fmax(wb->rlo - ti->rhi, ti->rlo - wb->rhi);
            const double b = (wb->smax + smax) * 1.000001;
            far = (gap2 > b * b) || (rg > b);
        }
        if (far) {
#pragma unroll UNROLL
            for (int jj = 0; jj < TJ; ++jj) {
                const double2 xy = *reinterpret_cast<const double2 *>(&T[jj].x);
                const double2 zm = *reinterpret_cast<const double2 *>(&T[jj].z);
                const double2 mm = *reinterpret_cast<const double2 *>(&T[jj].m15);
#pragma unroll
                for (int r = 0; r < R; ++r) STEPS_PAIR_FAR_F64(xy.x, xy.y, zm.x, zm.y, mm.x, mm.y, r)
            }
        } else {
            int thr[R];
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const double b = (double)si_up[r] + smax;
                // conservative: r2 < b*b  =>  hi(r2) <= hi(b*b); one extra ulp of the high word for rounding of b*b
                thr[r] = __double2hiint(b * b) + 1;
            }
            for (int j0 = 0; j0 < TJ; j0 += JB) {
                int ymin = 0x7fffffff;  // becomes 0 iff some pair of this sub-block was flagged
#pragma unroll UNROLL
                for (int jj = 0; jj < JB; ++jj) {
                    const double2 xy = *reinterpret_cast<const double2 *>(&T[j0 + jj].x);
                    const double2 zm = *reinterpret_cast<const double2 *>(&T[j0 + jj].z);
                    const double2 mm = *reinterpret_cast<const double2 *>(&T[j0 + jj].m15);
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        const double dx = xy.x - xi[r];
                        const double dy = xy.y - yi[r];
                        const double dz = zm.x - zi[r];
                        const double dx2 = dx * dx;
                        double r2 = fma(dy, dy, dx2);
                        r2 = fma(dz, dz, r2);
                        // seed with the near-mask folded in: hi(y0) = 0 for flagged pairs => y0 denormal => t = c = w = 0
                        // exactly (and no inf/NaN from r2 = 0); ISETP + SEL + VIMNMX, nothing on the FP64 pipe.  The low
                        // word of the seed comes from a dead value instead of being zeroed (see STEPS_PAIR_FAR_F64).
                        int yh = __double2hiint(rsqrt_seed(r2));
                        yh = (__double2hiint(r2) <= thr[r]) ? 0 : yh;
                        ymin = min(ymin, yh);
                        const double y0 = __hiloint2double(yh, __double2loint(dx2));
                        const double tt = y0 * y0;
                        const double e = fma(-r2, tt, 1.0);
                        const double c = tt * y0;
                        double p = fma(mm.y, e, mm.x);
                        p = fma(p, e, zm.y);
                        const double w = c * p;
                        ax[r] = fma(w, dx, ax[r]);
                        ay[r] = fma(w, dy, ay[r]);
                        az[r] = fma(w, dz, az[r]);
                    }
                }
                if (ymin == 0) {
                    // rare: re-evaluate the flagged pairs of this sub-block with the reference's exact branches
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        int il = ib * (THREADS * R) + r * THREADS + tid;
                        il = il < a.n_i ? il : a.n_i - 1;
                        const double3 f = near_pairs_f64(T + j0, JB, xi[r], yi[r], zi[r], jrec[a.id_min + il].s, thr[r]);
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
