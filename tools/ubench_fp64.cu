// ubench_fp64.cu -- FP64-pipe microbenchmarks for sm_100a (development aid; results go to profiles/).
// Measures how many FP64-pipe warp-instructions per clock per SM the B200 sustains for instruction mixes
// that bracket the pair kernel: pure DFMA, DADD/DMUL/DFMA blend, + MUFU.RSQ64H, + ALU ops, + LDS.128.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/ubench_fp64 tools/ubench_fp64.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ double rsq64h(double x) {
    double y;
    asm volatile("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    return y;
}

// MODE 0: 16 DFMA / iter
// MODE 1: 15 DFMA + 1 MUFU.RSQ64H
// MODE 2: 15 DFMA + 1 MUFU.RSQ64H + ISETP/SEL/VIMNMX (3 ALU)
// MODE 3: blend 8 DFMA + 5 DMUL + 3 DADD (the pair kernel's FP64 mix, 16 instr)
// MODE 4: 16 DADD
// MODE 5: 16 DMUL
// MODE 6: 15 DFMA + 1 MUFU + 3 ALU + 1 LDS.128 per 1.33 iter-equivalent (3 LDS.128 per 4 "pairs")
// MODE 7: 14 DFMA + 2 MUFU.RSQ64H
// MODE 8: 8 DFMA + 8 FFMA (does the FP32 pipe dual-issue beside FP64?)
template <int MODE>
__global__ void __launch_bounds__(256) k(double *out, int iters, double a, double b, int thr) {
    __shared__ double4 sm[64];
    if (threadIdx.x < 64) sm[threadIdx.x] = make_double4(a, b, a, b);
    __syncthreads();
    double c0 = threadIdx.x + 1.0, c1 = c0 + 1, c2 = c0 + 2, c3 = c0 + 3, c4 = c0 + 4, c5 = c0 + 5, c6 = c0 + 6, c7 = c0 + 7;
    float f0 = threadIdx.x, f1 = f0 + 1, f2 = f0 + 2, f3 = f0 + 3;
    int ymin = 0x7fffffff;
    double seed = c0;
    for (int i = 0; i < iters; ++i) {
        if (MODE == 0) {
            c0 = fma(c0, a, b); c1 = fma(c1, a, b); c2 = fma(c2, a, b); c3 = fma(c3, a, b);
            c4 = fma(c4, a, b); c5 = fma(c5, a, b); c6 = fma(c6, a, b); c7 = fma(c7, a, b);
            c0 = fma(c0, a, b); c1 = fma(c1, a, b); c2 = fma(c2, a, b); c3 = fma(c3, a, b);
            c4 = fma(c4, a, b); c5 = fma(c5, a, b); c6 = fma(c6, a, b); c7 = fma(c7, a, b);
        } else if (MODE == 1 || MODE == 2 || MODE == 6 || MODE == 7) {
            double y = rsq64h(seed);
            int yh = __double2hiint(y);
            if (MODE == 2 || MODE == 6) {
                yh = (__double2hiint(c7) <= thr) ? 0 : yh;
                ymin = min(ymin, yh);
            }
            if (MODE == 6) {
                const double4 q = sm[(i + threadIdx.x / 32) & 63];
                c6 += q.x * 1e-300;  // folded below into the DFMA count (1 of the 15)
                seed = __hiloint2double(yh, __double2loint(q.w));
            } else {
                seed = __hiloint2double(yh ^ 0x00100000, 0);
            }
            if (MODE == 7) {
                double y2 = rsq64h(c6);
                c6 = __hiloint2double(__double2hiint(y2) | 0x3ff00000, 0);
            }
            c0 = fma(c0, a, b); c1 = fma(c1, a, b); c2 = fma(c2, a, b); c3 = fma(c3, a, b);
            c4 = fma(c4, a, b); c5 = fma(c5, a, b); if (MODE != 6 && MODE != 7) c6 = fma(c6, a, b); c7 = fma(c7, a, b);
            c0 = fma(c0, a, b); c1 = fma(c1, a, b); c2 = fma(c2, a, b); c3 = fma(c3, a, b);
            c4 = fma(c4, a, b); c5 = fma(c5, a, b); if (MODE != 7) c6 = fma(c6, a, b);
        } else if (MODE == 3) {
            c0 = fma(c0, a, b); c1 = c1 * a; c2 = c2 + b; c3 = fma(c3, a, b);
            c4 = c4 * a; c5 = fma(c5, a, b); c6 = c6 + b; c7 = fma(c7, a, b);
            c0 = c0 * a; c1 = fma(c1, a, b); c2 = fma(c2, a, b); c3 = c3 * a;
            c4 = fma(c4, a, b); c5 = c5 + b; c6 = c6 * a; c7 = fma(c7, a, b);
        } else if (MODE == 4) {
            c0 += b; c1 += b; c2 += b; c3 += b; c4 += b; c5 += b; c6 += b; c7 += b;
            c0 += a; c1 += a; c2 += a; c3 += a; c4 += a; c5 += a; c6 += a; c7 += a;
        } else if (MODE == 5) {
            c0 *= b; c1 *= b; c2 *= b; c3 *= b; c4 *= b; c5 *= b; c6 *= b; c7 *= b;
            c0 *= a; c1 *= a; c2 *= a; c3 *= a; c4 *= a; c5 *= a; c6 *= a; c7 *= a;
        } else if (MODE == 8) {
            c0 = fma(c0, a, b); c1 = fma(c1, a, b); c2 = fma(c2, a, b); c3 = fma(c3, a, b);
            c4 = fma(c4, a, b); c5 = fma(c5, a, b); c6 = fma(c6, a, b); c7 = fma(c7, a, b);
            f0 = fmaf(f0, 1.0001f, 0.5f); f1 = fmaf(f1, 1.0001f, 0.5f); f2 = fmaf(f2, 1.0001f, 0.5f); f3 = fmaf(f3, 1.0001f, 0.5f);
            f0 = fmaf(f0, 0.9999f, 0.5f); f1 = fmaf(f1, 0.9999f, 0.5f); f2 = fmaf(f2, 0.9999f, 0.5f); f3 = fmaf(f3, 0.9999f, 0.5f);
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = c0 + c1 + c2 + c3 + c4 + c5 + c6 + c7 + seed + ymin + f0 + f1 + f2 + f3;
}

template <int MODE>
void run(const char *name, int fp64_per_iter, int blocks_per_sm, int sms, double *d) {
    const int iters = 1 << 14;
    const int blocks = sms * blocks_per_sm;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        CK(cudaEventRecord(e0));
        k<MODE><<<blocks, 256>>>(d, iters, 1.0000001, 1e-9, 0x3ff00000);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0 && ms < best) best = ms;
    }
    CK(cudaGetLastError());
    const double warp_instr = (double)fp64_per_iter * iters * blocks * 8.0;  // 8 warps per block
    const double cycles = best * 1e-3 * 1.965e9;
    printf("%-44s blocks/SM=%d  %8.3f ms  FP64 warp-instr/clk/SM = %.3f  (lanes/clk/SM = %.1f)  cyc per 15-FP64 'pair' per SMSP = %.2f\n", name, blocks_per_sm,
           best, warp_instr / cycles / sms, 32.0 * warp_instr / cycles / sms, 15.0 * 4.0 / (warp_instr / cycles / sms));
}

int main() {
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    printf("%s, %d SMs, nominal clock %.0f MHz\n", prop.name, prop.multiProcessorCount, prop.clockRate / 1e3);
    double *d;
    CK(cudaMalloc(&d, (size_t)prop.multiProcessorCount * 8 * 256 * 8));
    const int sms = prop.multiProcessorCount;
    for (int bps : {2, 4, 8}) {
        run<0>("16 DFMA", 16, bps, sms, d);
        run<3>("8 DFMA + 5 DMUL + 3 DADD", 16, bps, sms, d);
        run<4>("16 DADD", 16, bps, sms, d);
        run<5>("16 DMUL", 16, bps, sms, d);
        run<1>("15 DFMA + 1 MUFU.RSQ64H", 15, bps, sms, d);
        run<7>("14 DFMA + 2 MUFU.RSQ64H", 14, bps, sms, d);
        run<2>("15 DFMA + 1 MUFU.RSQ64H + 3 ALU", 15, bps, sms, d);
        run<6>("15 DFMA + 1 MUFU + 3 ALU + 1 LDS.128", 15, bps, sms, d);
        run<8>("8 DFMA + 8 FFMA", 8, bps, sms, d);
    }
    return 0;
}
