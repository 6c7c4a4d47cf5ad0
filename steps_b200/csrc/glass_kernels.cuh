// glass_kernels.cuh -- the KDK kernels of a GLASS_MAKING build: the same kick / drift / kick as
// kick_drift_kernel and kick_errmax_kernel (aux_kernels.cuh; the caller passes G = -1, global_variables.h:19-23) plus the
// diagnostics the reference's step() accumulates in that mode (step.cc:107-121, :143-148, :270-303):
//   first half:   disp_i = |v_i h| after the kick           -> sum, max      (dmean, dmax)
//   second half:  |F_i|, |ACCELERATION_i|, |v_i| after the kick -> sum, max each (F_mean, Fmax, A_mean, A_max, V_mean, V_max)
// Sums are formed deterministically: one (sum, max) pair per quantity per block in a partial buffer, reduced by one block in
// fixed order.  OPT-IN mode of the engine (steps_b200_engine_set_glass_making); written after round 1's GPU budget was
// spent, not yet run on a GPU (tests/test_gpu_glass.py, gated by STEPS_B200_EXPERIMENTAL=1).
#pragma once
#include "aux_kernels.cuh"

namespace steps {

constexpr int GLASS_NQ = 4;  // quantities: 0 displacement, 1 force, 2 acceleration, 3 velocity; slot 2q = sum, 2q+1 = max

// block-wide (sum, max) of one non-negative value per thread -> out[0], out[1] (thread 0 writes)
__device__ __forceinline__ void glass_block_sum_max(double val, double *__restrict__ out, double *__restrict__ sh /* [2][32] */) {
    double s = val, m = val;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s += __shfl_xor_sync(0xffffffffu, s, o);
        m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
    }
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    __syncthreads();  // sh may still be read by the previous call
    if (l == 0) { sh[w] = s; sh[32 + w] = m; }
    __syncthreads();
    if (w == 0) {
        const int nw = blockDim.x >> 5;
        s = l < nw ? sh[l] : 0.0;
        m = l < nw ? sh[32 + l] : 0.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            s += __shfl_xor_sync(0xffffffffu, s, o);
            m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
        }
        if (l == 0) { out[0] = s; out[1] = m; }
    }
}

// first half kick + drift + wrap (kick_drift_kernel) + displacement diagnostics; part = [gridDim.x][2 * GLASS_NQ]
template <typename T>
__global__ void glass_kick_drift_kernel(T *__restrict__ x, T *__restrict__ v, const T *__restrict__ F, int lo, int hi, KdkScalars k,
                                        double *__restrict__ part) {
    __shared__ double sh[64];
    const int i = lo + blockIdx.x * blockDim.x + threadIdx.x;
    T disp = 0;
    if (i < hi) {
        const T L = (T)k.L;
        T d2 = 0;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const size_t q = 3 * (size_t)i + c;
            const T acc = (T)(k.G * (double)F[q] * (double)(T)k.a3inv - k.twoH * (double)v[q]);
            T vv = v[q] + acc * (T)k.hhalf;
            v[q] = vv;
            T xx = x[q] + vv * (T)k.h;
            if (k.topology == 1 || ((k.topology == 2 || k.topology == 3) && c == 2)) xx = wrap_box<T>(xx, L);
            x[q] = xx;
            const T dv = vv * (T)k.h;  // step.cc:144: pow(v*(REAL)h, 2)
            d2 += dv * dv;
        }
        disp = sqrt(d2);
    }
    glass_block_sum_max((double)disp, part + (size_t)blockIdx.x * 2 * GLASS_NQ, sh);
}

// second half kick + errmax (kick_errmax_kernel, do_kick = 1, no wrap) + force / acceleration / velocity diagnostics
template <typename T>
__global__ void glass_kick_errmax_kernel(T *__restrict__ v, const T *__restrict__ F, const T *__restrict__ soft, int lo, int hi, KdkScalars k,
                                         double *__restrict__ errmax, double *__restrict__ part) {
    __shared__ double sh[64];
    const int i = lo + blockIdx.x * blockDim.x + threadIdx.x;
    T err = 0, fabs_ = 0, aabs = 0, vabs = 0;
    if (i < hi) {
        T acc2 = 0, f2 = 0, v2 = 0;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const size_t q = 3 * (size_t)i + c;
            const T f = F[q];
            const T acc = (T)(k.G * (double)f * (double)(T)k.a3inv - k.twoH * (double)v[q]);
            const T vv = v[q] + acc * (T)k.hhalf;
            v[q] = vv;
            acc2 += acc * acc;
            f2 += f * f;
            v2 += vv * vv;
        }
        aabs = sqrt(acc2);
        err = aabs / soft[i];
        fabs_ = sqrt(f2);
        vabs = sqrt(v2);
    }
    double *__restrict__ out = part + (size_t)blockIdx.x * 2 * GLASS_NQ;
    glass_block_sum_max((double)fabs_, out + 2, sh);
    glass_block_sum_max((double)aabs, out + 4, sh);
    glass_block_sum_max((double)vabs, out + 6, sh);
    // errmax exactly as kick_errmax_kernel: block max -> atomicMax on the ordered bits
    double e = (double)err;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) e = fmax(e, __shfl_xor_sync(0xffffffffu, e, o));
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    __syncthreads();
    if (l == 0) sh[w] = e;
    __syncthreads();
    if (w == 0) {
        e = (l < (blockDim.x >> 5)) ? sh[l] : 0.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) e = fmax(e, __shfl_xor_sync(0xffffffffu, e, o));
        if (l == 0) atomic_max_nonneg(errmax, e);
    }
}

// partials [nblocks][2*GLASS_NQ] -> out[q] = sum, out[GLASS_NQ + q] = max for the quantities [q0, q1): one block, fixed order
// (thread t adds blocks t, t+256, ...).  Sums and maxima sit in separate halves so that two all-reduces serve a multi-GPU job.
__global__ void glass_finish_kernel(const double *__restrict__ part, int nblocks, int q0, int q1, double *__restrict__ out) {
    __shared__ double sh[64];
    for (int q = q0; q < q1; ++q) {
        double s = 0.0, m = 0.0;
        for (int b = threadIdx.x; b < nblocks; b += blockDim.x) {
            s += part[(size_t)b * 2 * GLASS_NQ + 2 * q];
            m = fmax(m, part[(size_t)b * 2 * GLASS_NQ + 2 * q + 1]);
        }
        // (sum, max) are reduced separately: glass_block_sum_max takes one value, so run it twice
        double tmp[2] = {0.0, 0.0};
        glass_block_sum_max(s, tmp, sh);
        const double ssum = tmp[0];
        glass_block_sum_max(m, tmp, sh);
        if (threadIdx.x == 0) {
            out[q] = ssum;
            out[GLASS_NQ + q] = tmp[1];
        }
    }
}

}  // namespace steps
