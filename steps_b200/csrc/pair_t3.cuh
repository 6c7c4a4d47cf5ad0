// pair_t3.cuh -- tuned T^3 (fully periodic) FP64 pair kernel for IS_PERIODIC >= 2 (Ewald lookup table).
//
// Same algorithm as forces_periodic() / the reference CUDA kernel (forces.cc:776-876, forces_cuda.cu:567-645):
//     F_i = sum_j m_j ( w(r_ij, s_i+s_j) d_ij - D(d_ij) ),   d_ij = nearest periodic image of x_j - x_i,
// D = tricubic (Catmull-Rom) interpolation of T3_EWALD_FORCE_TABLE at d_ij (forces_cuda.cu:81-167: 64 wrapped
// neighbours starting at i0-1 of u = (d + L/2)/(L/Ngrid) - 1/2).  What changes is how it is evaluated:
//   * nearest image by compare + subtract of +-L (the reference's d - L*d/|d| is the same value to an ulp of L);
//   * w = r^-3 from the MUFU.RSQ64H seed and the e-series of pair_r3.cuh (6 FP64 instructions) instead of sqrt + division;
//     pairs that may lie inside the softening radius (integer test of hi(r2) against a per-(i, tile) threshold) are masked
//     out of it and re-evaluated with the reference's exact branches per 16-record sub-block (rare);
//   * the table is re-laid once per upload as (Ngrid+4)^3 cells of 32 bytes {Dx, Dy, Dz, 0} with a periodic halo of 2:
//     no modulo in the pair loop (the reference wraps 24 indices per pair with integer %), and a cell is one aligned
//     32-byte sector (a 128-bit + a 64-bit load) instead of 24 bytes straddling sectors; the 4 cells of a z-row are contiguous;
//   * the 4x4x4 sum is contracted z-first (per (x,y) row: 3 x (DMUL + 3 DFMA), then one DFMA per component with
//     wx*wy): 256 FP64 instructions instead of 272 and two live weights instead of one product per cell.
// The kernel is load-bound (128 scattered 128-bit table reads per pair through L1/L2), not FP64-bound: see DESIGN.md.
// Preconditions checked by the engine: FP64, IS_PERIODIC >= 2, every coordinate inside [0, L) (device flag written by
// the pack kernel; otherwise the exact-branch kernel of pair_generic.cuh runs in the same launch shape).
#pragma once
#include "pair_generic.cuh"

namespace steps {

struct T3Consts {
    double L, halfL;
    double h, inv_h;       // grid spacing L / Ngrid and its reciprocal
    int P;                 // padded table dimension, Ngrid + 4
    int ngrid;
    const double4 *tab;    // [P][P][P] cells {Dx, Dy, Dz, 0}; cell (a,b,c) holds table entry ((a-2) mod Ngrid, ...)
};

// periodic halo copy of the reference-layout table [Ngrid^3][3] into the padded cell layout
__global__ void t3_pad_table_kernel(const double *__restrict__ table, int ngrid, int P, double4 *__restrict__ out) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t total = (size_t)P * P * P;
    if (idx >= total) return;
    const int c = (int)(idx % P);
    const int b = (int)((idx / P) % P);
    const int a = (int)(idx / ((size_t)P * P));
    const int ix = imodp(a - 2, ngrid), iy = imodp(b - 2, ngrid), iz = imodp(c - 2, ngrid);
    const double *__restrict__ e = table + (((size_t)ix * ngrid + iy) * ngrid + iz) * 3u;
    out[idx] = make_double4(e[0], e[1], e[2], 0.0);
}

__device__ __forceinline__ double t3_wrap(double d, double L, double halfL) {
    // reference: if (fabs(d) > 0.5*L) d = d - L*d/fabs(d)
    return (fabs(d) > halfL) ? d - copysign(L, d) : d;
}

// one 32-byte cell through the read-only path: (Dx, Dy) as one 128-bit load, Dz as a 64-bit load of the same sector
__device__ __forceinline__ double3 t3_cell(const double4 *__restrict__ c) {
    const double2 xy = __ldg(reinterpret_cast<const double2 *>(c));
    const double z = __ldg(reinterpret_cast<const double *>(c) + 2);
    return make_double3(xy.x, xy.y, z);
}

// slow path: the pairs of one sub-block that the integer test flagged, exact softened kernel (table part already added)
__device__ __noinline__ double3 near_pairs_t3(const JRec64 *__restrict__ T, int nj, double xi, double yi, double zi, double si, int thr,
                                              double L, double halfL) {
    double fx = 0.0, fy = 0.0, fz = 0.0;
    for (int jj = 0; jj < nj; ++jj) {
        const JRec64 q = T[jj];
        const double dx = t3_wrap(q.x - xi, L, halfL);
        const double dy = t3_wrap(q.y - yi, L, halfL);
        const double dz = t3_wrap(q.z - zi, L, halfL);
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

// grid coordinate of one component: padded index of the first stencil cell (i0 - 1 + 2) and the cubic weights
__device__ __forceinline__ int t3_axis(double d, const T3Consts &k, double (&w)[4]) {
    // u = (d + L/2) / h - 1/2 with the quotient of the reference's division (forces_cuda.cu:90-93): reciprocal multiply plus
    // one FMA residual step gives the correctly rounded quotient (a bare reciprocal multiply is off by an ulp of u ~ 1e-14 in the
    // fractional coordinate, which the strongly cancelling lattice forces of this topology would show at the 1e-12 level)
    const double num = d + k.halfL;
    double qd = num * k.inv_h;
    qd = fma(fma(-qd, k.h, num), k.inv_h, qd);
    const double u = qd - 0.5;
    int i0 = __double2int_rd(u);
    i0 = max(-1, min(i0, k.ngrid - 1));  // d in [-L/2, L/2] gives exactly this range; the clamp only guards bad input
    const double t = u - (double)i0;
    const double t2 = t * t;
    const double t3 = t * t2;
    w[0] = -0.5 * t3 + t2 - 0.5 * t;          // forces_cuda.cu:104-113 (get_cubic_weights)
    w[1] = 1.5 * t3 - 2.5 * t2 + 1.0;
    w[2] = -1.5 * t3 + 2.0 * t2 + 0.5 * t;
    w[3] = 0.5 * t3 - 0.5 * t2;
    return i0 + 1;
}

template <int R, int THREADS, int TJ, int STAGES, int MINB>
__global__ void __launch_bounds__(THREADS, MINB) force_t3_f64_kernel(const R3LaunchArgs a, const T3Consts k) {
    if (a.gate && *a.gate != a.gate_value) return;  // some coordinate lies outside [0, L): the exact-branch kernel handles this call
    constexpr int NWARPS = THREADS / 32;
    constexpr int JB = 16;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    JRec64 *tiles = reinterpret_cast<JRec64 *>(smem_raw);
    uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw + (size_t)STAGES * TJ * sizeof(JRec64));
    uint64_t *empty = full + STAGES;

    const int tid = threadIdx.x;
    const int jc = blockIdx.x / a.n_ib;
    const int ib = blockIdx.x - jc * a.n_ib;
    const int t0 = jc * a.tiles_per_chunk;
    const int t1 = min(t0 + a.tiles_per_chunk, a.n_tiles);
    const int nt = t1 - t0;
    const JRec64 *__restrict__ jrec = static_cast<const JRec64 *>(a.jrec);

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
            mbar_arrive_expect_tx(&full[t], TJ * sizeof(JRec64));
            tma_load_1d(tiles + (size_t)t * TJ, jrec + (size_t)(t0 + t) * TJ, TJ * sizeof(JRec64), &full[t]);
        }
    }

    double xi[R], yi[R], zi[R], si[R], ax[R], ay[R], az[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
        int il = ib * (THREADS * R) + r * THREADS + tid;
        il = il < a.n_i ? il : a.n_i - 1;
        const JRec64 me = jrec[a.id_min + il];
        xi[r] = me.x; yi[r] = me.y; zi[r] = me.z; si[r] = me.s;
        ax[r] = ay[r] = az[r] = 0.0;
    }
    const size_t P = (size_t)k.P;

    for (int t = 0; t < nt; ++t) {
        const int s = t % STAGES;
        const uint32_t ph = (uint32_t)(t / STAGES) & 1u;
        if (tid == 0 && t >= 1 && (t - 1 + STAGES) < nt) {
            const int sp = (t - 1) % STAGES;
            const uint32_t php = (uint32_t)((t - 1) / STAGES) & 1u;
            mbar_wait(&empty[sp], php);
            mbar_arrive_expect_tx(&full[sp], TJ * sizeof(JRec64));
            tma_load_1d(tiles + (size_t)sp * TJ, jrec + (size_t)(t0 + t - 1 + STAGES) * TJ, TJ * sizeof(JRec64), &full[sp]);
        }
        mbar_wait(&full[s], ph);
        const JRec64 *__restrict__ Tl = tiles + (size_t)s * TJ;
        const double smax = Tl[0].smax;
        // the last tile is padded with massless far-away records: never feed those to the wrap / the table
        const int jn = min(TJ, a.n_j - (t0 + t) * TJ);
        int thr[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const double b = si[r] + smax;
            thr[r] = __double2hiint(b * b) + 1;  // conservative: r2 < b*b => hi(r2) <= hi(b*b) (+1 for the rounding of b*b)
        }
        for (int j0 = 0; j0 < jn; j0 += JB) {
            const int nb = min(JB, jn - j0);
            int flagged = 0;
            for (int jj = 0; jj < nb; ++jj) {
                const double2 xy = *reinterpret_cast<const double2 *>(&Tl[j0 + jj].x);
                const double2 zm = *reinterpret_cast<const double2 *>(&Tl[j0 + jj].z);
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    const double dx = t3_wrap(xy.x - xi[r], k.L, k.halfL);
                    const double dy = t3_wrap(xy.y - yi[r], k.L, k.halfL);
                    const double dz = t3_wrap(zm.x - zi[r], k.L, k.halfL);
                    double r2 = dx * dx;
                    r2 = fma(dy, dy, r2);
                    r2 = fma(dz, dz, r2);
                    int yh = __double2hiint(rsqrt_seed(r2));
                    const bool near = __double2hiint(r2) <= thr[r];
                    yh = near ? 0 : yh;  // masked: y0 = 0 => w = 0 exactly (and no inf/NaN from r2 = 0)
                    flagged |= near ? 1 : 0;
                    const double y0 = __hiloint2double(yh, 0);
                    const double tt = y0 * y0;
                    const double e = fma(-r2, tt, 1.0);
                    const double c = tt * y0;
                    double q = fma(e, 1.875, 1.5);
                    q = fma(q, e, 1.0);
                    const double w = c * q;
                    // ---- D(d): tricubic interpolation of the padded table, z-first contraction
                    double wx[4], wy[4], wz[4];
                    const int ix = t3_axis(dx, k, wx);
                    const int iy = t3_axis(dy, k, wy);
                    const int iz = t3_axis(dz, k, wz);
                    const double4 *__restrict__ base = k.tab + ((size_t)ix * P + iy) * P + iz;
                    double sx = 0.0, sy = 0.0, sz = 0.0;
#pragma unroll
                    for (int p = 0; p < 4; ++p) {
#pragma unroll
                        for (int qy = 0; qy < 4; ++qy) {
                            const double4 *__restrict__ row = base + ((size_t)p * P + qy) * P;
                            const double3 c0 = t3_cell(row), c1 = t3_cell(row + 1), c2 = t3_cell(row + 2), c3 = t3_cell(row + 3);
                            double rx = wz[0] * c0.x, ry = wz[0] * c0.y, rz = wz[0] * c0.z;
                            rx = fma(wz[1], c1.x, rx); ry = fma(wz[1], c1.y, ry); rz = fma(wz[1], c1.z, rz);
                            rx = fma(wz[2], c2.x, rx); ry = fma(wz[2], c2.y, ry); rz = fma(wz[2], c2.z, rz);
                            rx = fma(wz[3], c3.x, rx); ry = fma(wz[3], c3.y, ry); rz = fma(wz[3], c3.z, rz);
                            const double wxy = wx[p] * wy[qy];
                            sx = fma(wxy, rx, sx); sy = fma(wxy, ry, sy); sz = fma(wxy, rz, sz);
                        }
                    }
                    // F += m_j (w d - D)      (forces_cuda.cu:632-634)
                    ax[r] = fma(zm.y, fma(w, dx, -sx), ax[r]);
                    ay[r] = fma(zm.y, fma(w, dy, -sy), ay[r]);
                    az[r] = fma(zm.y, fma(w, dz, -sz), az[r]);
                }
            }
            if (flagged) {
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    const double3 f = near_pairs_t3(Tl + j0, nb, xi[r], yi[r], zi[r], si[r], thr[r], k.L, k.halfL);
                    ax[r] += f.x; ay[r] += f.y; az[r] += f.z;
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
