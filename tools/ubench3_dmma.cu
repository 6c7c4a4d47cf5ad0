// ubench3_dmma.cu -- does the FP64 tensor instruction (DMMA, mma.sync.m8n8k4.f64) issue beside the FP64 CUDA-core pipe on
// sm_100a, and at what rate?  (development aid; results go to profiles/)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/ubench3_dmma tools/ubench3_dmma.cu
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)
#define DFMA(d, a, b, c) asm volatile("fma.rn.f64 %0, %1, %2, %3;" : "=d"(d) : "d"(a), "d"(b), "d"(c))
#define DMMA(c0, c1, a, b) asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b))

// per loop iteration: NF DFMA (x = fma(y,y,x), 2 distinct registers: 2.04 cyc each alone) + ND DMMA on independent accumulators
template <int NF, int ND>
__global__ void __launch_bounds__(128) k(double *out, int iters, double a, double b) {
    double x[8], y[8], c0[8], c1[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        x[q] = threadIdx.x + q + 1.0;
        y[q] = a + 1e-9 * q + 1e-12 * threadIdx.x;
        c0[q] = 0.0; c1[q] = 0.0;
    }
    double av = a + 1e-10 * threadIdx.x, bv = b + 1e-11 * threadIdx.x;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int q = 0; q < 16; ++q) {
            if (q < NF) DFMA(x[q & 7], y[q & 7], y[q & 7], x[q & 7]);
            if (q < ND) DMMA(c0[q & 7], c1[q & 7], av, bv);
        }
    }
    double s = 0;
#pragma unroll
    for (int q = 0; q < 8; ++q) s += x[q] + y[q] + c0[q] + c1[q];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NF, int ND>
void run(const char *name, int blocks_per_sm, int sms, double *d) {
    const int iters = 1 << 12;
    const int blocks = sms * blocks_per_sm;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        CK(cudaEventRecord(e0));
        k<NF, ND><<<blocks, 128>>>(d, iters, 1.0000001, 1e-9);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0 && ms < best) best = ms;
    }
    CK(cudaGetLastError());
    const double cyc = best * 1e-3 * 1.965e9 / ((double)iters * blocks_per_sm);
    const double tf = ((double)NF * 2 * 32 + (double)ND * 512) * iters * blocks * 4 / (best * 1e-3) / 1e12;
    printf("%-28s warps/SMSP=%d  %8.3f ms  cycles per iteration per SMSP = %7.2f   total %.2f TFLOP/s (DFMA 64 + DMMA 512 flop per warp instr)\n", name,
           blocks_per_sm, best, cyc, tf);
}

int main() {
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    printf("%s, %d SMs\n", prop.name, prop.multiProcessorCount);
    const int sms = prop.multiProcessorCount;
    double *d;
    CK(cudaMalloc(&d, (size_t)(1 << 22) * 8));
    for (int bps : {1, 2, 4}) {
        run<16, 0>("16 DFMA", bps, sms, d);
        run<0, 1>("1 DMMA.884", bps, sms, d);
        run<0, 4>("4 DMMA.884", bps, sms, d);
        run<0, 8>("8 DMMA.884", bps, sms, d);
        run<16, 1>("16 DFMA + 1 DMMA", bps, sms, d);
        run<16, 2>("16 DFMA + 2 DMMA", bps, sms, d);
        run<12, 1>("12 DFMA + 1 DMMA", bps, sms, d);
        run<12, 2>("12 DFMA + 2 DMMA", bps, sms, d);
        run<8, 8>("8 DFMA + 8 DMMA", bps, sms, d);
    }
    return 0;
}
