// ubench2_fp64.cu -- second FP64 issue-cost microbenchmark for sm_100a (development aid; results go to profiles/).
// Questions it answers for the R^3 FP64 pair kernel (pair_r3.cuh):
//   (a) what does a DFMA cost with 1, 2, 3 distinct register operands (register-file bank pressure)?
//   (b) what does one MUFU.RSQ64H / MUFU.RSQ (f32) / MUFU.RCP64H cost beside a stream of FP64 instructions?
//   (c) what does a warp-uniform LDS.32/.64/.128 cost beside a stream of FP64 instructions?
//   (d) the distribution of e = 1 - r2*y0^2 for the MUFU.RSQ64H seed y0 (bounds the series truncation of the pair math)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/ubench2_fp64 tools/ubench2_fp64.cu
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

#define DFMA(d, a, b, c) asm volatile("fma.rn.f64 %0, %1, %2, %3;" : "=d"(d) : "d"(a), "d"(b), "d"(c))
#define DMUL(d, a, b) asm volatile("mul.rn.f64 %0, %1, %2;" : "=d"(d) : "d"(a), "d"(b))
#define DADD(d, a, b) asm volatile("add.rn.f64 %0, %1, %2;" : "=d"(d) : "d"(a), "d"(b))

// 8 independent chains x[k]; y[k], z[k] are per-chain loop-invariant registers.
template <int MODE>
__global__ void __launch_bounds__(128) k(double *out, int iters, double a, double b) {
    __shared__ double4 sm[64];
    if (threadIdx.x < 64) sm[threadIdx.x] = make_double4(a, b, a, b);
    __syncthreads();
    double x[8], y[8], z[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        x[q] = threadIdx.x + q + 1.0;
        y[q] = a + 1e-9 * q + 1e-12 * threadIdx.x;
        z[q] = b + 1e-9 * q + 1e-12 * threadIdx.x;
    }
    double sink = 0.0;
    float fsink = 1.0f + threadIdx.x;
    unsigned usink = 0;
    const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(sm);
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int rep = 0; rep < 2; ++rep) {
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                if (MODE == 0) DFMA(x[q], x[q], a, b);               // 1 varying register + 2 shared invariants
                else if (MODE == 1) DFMA(x[q], y[q], z[q], x[q]);    // 3 distinct registers
                else if (MODE == 2) DFMA(x[q], y[q], y[q], x[q]);    // 2 distinct registers
                else if (MODE == 3) DMUL(x[q], x[q], y[q]);          // 2 distinct
                else if (MODE == 4) DADD(x[q], x[q], y[q]);          // 2 distinct
                else if (MODE == 5) { if (q & 1) DFMA(x[q], y[q], z[q], x[q]); else DMUL(x[q], x[q], y[q]); }   // alternate DFMA(3)/DMUL
                else if (MODE == 6) { if (q & 1) DFMA(x[q], y[q], z[q], x[q]); else DADD(x[q], x[q], y[q]); }   // alternate DFMA(3)/DADD
                else if (MODE == 7) DFMA(x[q], x[q], y[q], z[q]);    // 3 distinct, accumulator in multiplicand position
                else DFMA(x[q], x[q], a, b);                          // MODE >= 10: 16 DFMA baseline + 1 extra op below
            }
        }
        // one extra instruction per 16 FP64 instructions
        if (MODE == 10) { double t; asm volatile("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(t) : "d"(x[i & 7])); sink += 0; asm volatile("" :: "d"(t)); }
        if (MODE == 11) { float t; asm volatile("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fsink)); asm volatile("" :: "f"(t)); }
        if (MODE == 12) { double t; asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(t) : "d"(x[i & 7])); asm volatile("" :: "d"(t)); }
        if (MODE == 13) { unsigned t; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(t) : "r"(sbase + ((i & 63) << 5))); asm volatile("" :: "r"(t)); }
        if (MODE == 14) { double t; asm volatile("ld.shared.f64 %0, [%1];" : "=d"(t) : "r"(sbase + ((i & 63) << 5))); asm volatile("" :: "d"(t)); }
        if (MODE == 15) { double t, u; asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(t), "=d"(u) : "r"(sbase + ((i & 63) << 5))); asm volatile("" :: "d"(t), "d"(u)); }
        if (MODE == 16) { unsigned t; asm volatile("add.u32 %0, %1, %2;" : "=r"(t) : "r"(usink), "r"(i)); usink = t; }
        if (MODE == 17) { float t; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(t) : "f"(fsink), "f"(1.0001f), "f"(0.5f)); fsink = t; }
        if (MODE == 18) {  // 3 LDS.128 (one j-record) per 16 FP64
            double t, u;
            asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(t), "=d"(u) : "r"(sbase + ((i & 31) << 6))); asm volatile("" :: "d"(t), "d"(u));
            asm volatile("ld.shared.v2.f64 {%0, %1}, [%2+16];" : "=d"(t), "=d"(u) : "r"(sbase + ((i & 31) << 6))); asm volatile("" :: "d"(t), "d"(u));
            asm volatile("ld.shared.v2.f64 {%0, %1}, [%2+32];" : "=d"(t), "=d"(u) : "r"(sbase + ((i & 31) << 6))); asm volatile("" :: "d"(t), "d"(u));
        }
    }
    double s = sink + fsink + usink;
#pragma unroll
    for (int q = 0; q < 8; ++q) s += x[q] + y[q] + z[q];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char *name, int blocks_per_sm, int sms, double *d) {
    const int iters = 1 << 13;
    const int blocks = sms * blocks_per_sm;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        CK(cudaEventRecord(e0));
        k<MODE><<<blocks, 128>>>(d, iters, 1.0000001, 1e-9);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0 && ms < best) best = ms;
    }
    CK(cudaGetLastError());
    // cycles per SMSP per group of 16 FP64 instructions (+ the extra op)
    const double warps_per_smsp = blocks_per_sm * 4 / 4.0;
    const double cyc = best * 1e-3 * 1.965e9 / ((double)iters * warps_per_smsp);
    printf("%-52s warps/SMSP=%g  %8.3f ms  cycles per 16-FP64 group per SMSP = %6.2f  (per FP64 instr %.3f)\n", name, warps_per_smsp, best, cyc, cyc / 16.0);
}

__global__ void seed_stats(double *e_out, int n, int garbage_lo) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    // r2 sweeps [1, 4) (one full period of the mantissa/exponent-parity pattern), scrambled low bits
    unsigned long long h = (unsigned long long)i * 0x9E3779B97F4A7C15ull;
    const double u = (double)(h >> 11) * (1.0 / 9007199254740992.0);
    const double r2 = 1.0 + 3.0 * u;
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(r2));
    if (garbage_lo) y = __hiloint2double(__double2hiint(y), (int)(h & 0xffffffffu));
    const double tt = y * y;
    e_out[i] = fma(-r2, tt, 1.0);
}

int main() {
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    printf("%s, %d SMs\n", prop.name, prop.multiProcessorCount);
    const int sms = prop.multiProcessorCount;
    double *d;
    CK(cudaMalloc(&d, (size_t)(1 << 24) * 8));
    for (int bps : {2, 4}) {
        run<0>("16 DFMA x=fma(x,a,b)        [1 varying reg]", bps, sms, d);
        run<2>("16 DFMA x=fma(y,y,x)        [2 distinct]", bps, sms, d);
        run<1>("16 DFMA x=fma(y,z,x)        [3 distinct]", bps, sms, d);
        run<7>("16 DFMA x=fma(x,y,z)        [3 distinct]", bps, sms, d);
        run<3>("16 DMUL x=x*y", bps, sms, d);
        run<4>("16 DADD x=x+y", bps, sms, d);
        run<5>("8 DFMA(3 distinct) + 8 DMUL alternating", bps, sms, d);
        run<6>("8 DFMA(3 distinct) + 8 DADD alternating", bps, sms, d);
        run<10>("16 DFMA + 1 MUFU.RSQ64H", bps, sms, d);
        run<11>("16 DFMA + 1 MUFU.RSQ (f32)", bps, sms, d);
        run<12>("16 DFMA + 1 MUFU.RCP64H", bps, sms, d);
        run<13>("16 DFMA + 1 LDS.32 (uniform addr)", bps, sms, d);
        run<14>("16 DFMA + 1 LDS.64 (uniform addr)", bps, sms, d);
        run<15>("16 DFMA + 1 LDS.128 (uniform addr)", bps, sms, d);
        run<18>("16 DFMA + 3 LDS.128 (uniform addr)", bps, sms, d);
        run<16>("16 DFMA + 1 IADD", bps, sms, d);
        run<17>("16 DFMA + 1 FFMA", bps, sms, d);
    }
    const int n = 1 << 22;
    std::vector<double> h(n);
    for (int g = 0; g < 2; ++g) {
        seed_stats<<<n / 256, 256>>>(d, n, g);
        CK(cudaMemcpy(h.data(), d, (size_t)n * 8, cudaMemcpyDeviceToHost));
        double lo = 1e300, hi = -1e300, s = 0, s2 = 0;
        for (double v : h) { lo = fmin(lo, v); hi = fmax(hi, v); s += v; s2 += v * v; }
        printf("MUFU.RSQ64H seed, %s low word: e = 1 - r2*y0^2 over r2 in [1,4): min %.4e  max %.4e  mean %.4e  rms %.4e   1.875*max(e^2) = %.3e  2.1875*max|e|^3 = %.3e\n",
               g ? "garbage" : "zero", lo, hi, s / n, sqrt(s2 / n), 1.875 * fmax(lo * lo, hi * hi), 2.1875 * pow(fmax(-lo, hi), 3));
    }
    return 0;
}
