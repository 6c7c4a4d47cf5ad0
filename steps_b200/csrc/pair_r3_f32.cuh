// pair_r3_f32.cuh -- the tuned R^3 pair kernel of the single-precision build (-DUSE_SINGLE_PRECISION,
// reference global_variables.h:26-32; BASELINE.json configs[4]: N = 16M FP32).
//
// Same decomposition as the FP64 kernel in pair_r3.cuh: (i-block x j-chunk) CTAs, TMA-staged j tiles,
// register-blocked i-particles, warp-uniform far/near tile classification, deterministic chunk reduction.
// Far-field pair math, 14 FP32-pipe instructions + one MUFU.RSQ:
//     3 FADD (d = xj - xi), FMUL + 2 FFMA (r2), y = MUFU.RSQ(r2),
//     t = y*y, e = fma(-r2, t, 1), c = t*y, p = fma(1.5m, e, m), w = c*p, 3 FFMA accumulate.
// The first-order correction e removes the ~2 ulp error of the hardware seed (cubed: ~6 ulp), leaving
// ~1 ulp per pair -- the level of the reference's own sqrtf/powf (forces.cc:52-87 in float).
#pragma once
#include "pair_r3.cuh"

namespace steps {

struct __align__(16) TileInfo32 {
    float lo[3], hi[3];
    float rlo, rhi;
};
static_assert(sizeof(TileInfo32) == 32, "TileInfo32 layout");

struct __align__(16) WarpBounds32 {
    float lo[3], hi[3];
    float rlo, rhi;
    float smax, pad[3];
};

__device__ __noinline__ float3 near_pairs_f32(const JRec32 *__restrict__ T, int nj, float xi, float yi, float zi, float si, int thr) {
    float fx = 0.f, fy = 0.f, fz = 0.f;
    for (int jj = 0; jj < nj; ++jj) {
        const JRec32 q = T[jj];
        const float dx = q.x - xi;
        const float dy = q.y - yi;
        const float dz = q.z - zi;
        float r2 = dx * dx;
        r2 = fmaf(dy, dy, r2);
        r2 = fmaf(dz, dz, r2);
        if (__float_as_int(r2) <= thr) {
            const float w = q.m * softened_w<float>(sqrtf(r2), si + q.s);
            fx = fmaf(w, dx, fx);
            fy = fmaf(w, dy, fy);
            fz = fmaf(w, dz, fz);
        }
    }
    return make_float3(fx, fy, fz);
}

template <int R, int THREADS, int TJ, int STAGES, int MINB, int UNROLL>
__global__ void __launch_bounds__(THREADS, MINB) force_r3_f32_kernel(const R3LaunchArgs a) {
    constexpr int NWARPS = THREADS / 32;
    constexpr int JB = 16;
    static_assert(TJ % JB == 0 && JB % UNROLL == 0, "tile must be a whole number of sub-blocks");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    JRec32 *tiles = reinterpret_cast<JRec32 *>(smem_raw);
    TileInfo32 *tinfo_s = reinterpret_cast<TileInfo32 *>(smem_raw + (size_t)STAGES * TJ * sizeof(JRec32));
    WarpBounds32 *wb_s = reinterpret_cast<WarpBounds32 *>(tinfo_s + STAGES);
    uint64_t *full = reinterpret_cast<uint64_t *>(wb_s + NWARPS);
    uint64_t *empty = full + STAGES;

    const int tid = threadIdx.x;
    const int jc = blockIdx.x / a.n_ib;
    const int ib = blockIdx.x - jc * a.n_ib;
    const int t0 = jc * a.tiles_per_chunk;
    const int t1 = min(t0 + a.tiles_per_chunk, a.n_tiles);
    const int nt = t1 - t0;
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
    if (tid == 0) {
        const int npre = nt < STAGES ? nt : STAGES;
        for (int t = 0; t < npre; ++t) {
            mbar_arrive_expect_tx(&full[t], TILE_TX);
            tma_load_1d(tiles + (size_t)t * TJ, jrec + (size_t)(t0 + t) * TJ, TJ * sizeof(JRec32), &full[t]);
            tma_load_1d(tinfo_s + t, tinfo + (t0 + t), sizeof(TileInfo32), &full[t]);
        }
    }

    float xi[R], yi[R], zi[R], ax[R], ay[R], az[R];
    {
        float lo[3] = {3e38f, 3e38f, 3e38f}, hi[3] = {-3e38f, -3e38f, -3e38f}, rlo = 3e38f, rhi = 0.f, smx = 0.f;
#pragma unroll
        for (int r = 0; r < R; ++r) {
            int il = ib * (THREADS * R) + r * THREADS + tid;
            il = il < a.n_i ? il : a.n_i - 1;
            const JRec32 me = jrec[a.id_min + il];
            xi[r] = me.x; yi[r] = me.y; zi[r] = me.z;
            ax[r] = ay[r] = az[r] = 0.f;
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
        if ((tid & 31) == 0) {
            WarpBounds32 &wb = wb_s[tid >> 5];
            wb.lo[0] = lo[0]; wb.lo[1] = lo[1]; wb.lo[2] = lo[2];
            wb.hi[0] = hi[0]; wb.hi[1] = hi[1]; wb.hi[2] = hi[2];
            wb.rlo = rlo; wb.rhi = rhi; wb.smax = smx;
        }
        __syncwarp();
    }
    const WarpBounds32 *__restrict__ wb = wb_s + (tid >> 5);

    for (int t = 0; t < nt; ++t) {
        const int s = t % STAGES;
        const uint32_t ph = (uint32_t)(t / STAGES) & 1u;
        if (tid == 0 && t >= 1 && (t - 1 + STAGES) < nt) {
            const int sp = (t - 1) % STAGES;
            const uint32_t php = (uint32_t)((t - 1) / STAGES) & 1u;
            mbar_wait(&empty[sp], php);
            mbar_arrive_expect_tx(&full[sp], TILE_TX);
            tma_load_1d(tiles + (size_t)sp * TJ, jrec + (size_t)(t0 + t - 1 + STAGES) * TJ, TJ * sizeof(JRec32), &full[sp]);
            tma_load_1d(tinfo_s + sp, tinfo + (t0 + t - 1 + STAGES), sizeof(TileInfo32), &full[sp]);
        }
        mbar_wait(&full[s], ph);
        const JRec32 *__restrict__ T = tiles + (size_t)s * TJ;
        const float smax = T[0].smax;
        bool far;
        {
            const TileInfo32 *__restrict__ ti = tinfo_s + s;
            float gap2 = 0.f;
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const float g = fmaxf(fmaxf(wb->lo[k] - ti->hi[k], ti->lo[k] - wb->hi[k]), 0.f);
                gap2 = fmaf(g, g, gap2);
            }
            const float rg = fmaxf(wb->rlo - ti->rhi, ti->rlo - wb->rhi);
            const float b = (wb->smax + smax) * 1.001f;  // FP32: coordinates carry ~1e-7 relative rounding; keep a wide margin
            far = (gap2 > b * b) || (rg > b);
        }
        if (far) {
#pragma unroll UNROLL
            for (int jj = 0; jj < TJ; ++jj) {
                const float4 q = *reinterpret_cast<const float4 *>(&T[jj].x);
                const float m15 = 1.5f * q.w;
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    const float dx = q.x - xi[r];
                    const float dy = q.y - yi[r];
                    const float dz = q.z - zi[r];
                    float r2 = dx * dx;
                    r2 = fmaf(dy, dy, r2);
                    r2 = fmaf(dz, dz, r2);
                    const float y = rsqrt_seed(r2);
                    const float tt = y * y;
                    const float e = fmaf(-r2, tt, 1.0f);
                    const float c = tt * y;
                    const float p = fmaf(m15, e, q.w);
                    const float w = c * p;
                    ax[r] = fmaf(w, dx, ax[r]);
                    ay[r] = fmaf(w, dy, ay[r]);
                    az[r] = fmaf(w, dz, az[r]);
                }
            }
        } else {
            int thr[R];
#pragma unroll
            for (int r = 0; r < R; ++r) {
                int il = ib * (THREADS * R) + r * THREADS + tid;
                il = il < a.n_i ? il : a.n_i - 1;
                const float b = jrec[a.id_min + il].s + smax;
                // conservative: r2 < b*b => bits(r2) <= bits(b*b); a few ulps up for the rounding of b and b*b
                thr[r] = __float_as_int(b * b) + 4;
            }
            for (int j0 = 0; j0 < TJ; j0 += JB) {
                int ymin = 0x7fffffff;
#pragma unroll UNROLL
                for (int jj = 0; jj < JB; ++jj) {
                    const float4 q = *reinterpret_cast<const float4 *>(&T[j0 + jj].x);
                    const float m15 = 1.5f * q.w;
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        const float dx = q.x - xi[r];
                        const float dy = q.y - yi[r];
                        const float dz = q.z - zi[r];
                        float r2 = dx * dx;
                        r2 = fmaf(dy, dy, r2);
                        r2 = fmaf(dz, dz, r2);
                        // near mask folded into the seed: y = 0 for flagged pairs => w = 0 exactly, no inf/NaN from r2 = 0
                        int yb = __float_as_int(rsqrt_seed(r2));
                        yb = (__float_as_int(r2) <= thr[r]) ? 0 : yb;
                        ymin = min(ymin, yb);
                        const float y = __int_as_float(yb);
                        const float tt = y * y;
                        const float e = fmaf(-r2, tt, 1.0f);
                        const float c = tt * y;
                        const float p = fmaf(m15, e, q.w);
                        const float w = c * p;
                        ax[r] = fmaf(w, dx, ax[r]);
                        ay[r] = fmaf(w, dy, ay[r]);
                        az[r] = fmaf(w, dz, az[r]);
                    }
                }
                if (ymin == 0) {
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        int il = ib * (THREADS * R) + r * THREADS + tid;
                        il = il < a.n_i ? il : a.n_i - 1;
                        const float3 f = near_pairs_f32(T + j0, JB, xi[r], yi[r], zi[r], jrec[a.id_min + il].s, thr[r]);
                        ax[r] += f.x; ay[r] += f.y; az[r] += f.z;
                    }
                }
            }
        }
        __syncwarp();
        if ((tid & 31) == 0) mbar_arrive(&empty[s]);
    }

    float *__restrict__ fp = static_cast<float *>(a.fpart) + (size_t)jc * 3 * a.fstride;
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
