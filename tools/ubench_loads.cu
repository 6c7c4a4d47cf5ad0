// ubench_loads.cu -- load-path microbenchmarks for sm_100a (development aid; results go to profiles/).
// Question behind it (the T^3 gather): what does a warp-wide load cost when all 32 lanes read the SAME address
// (the lattice-ordered T^3 case: all lanes of a systolic step fall into one table cell), as a function of the access width and
// of the path (shared memory / L1-resident global)?  If the register write-back (128 B/clk/SM) bounds broadcast loads too, the
// 1536 table bytes a tricubic evaluation consumes cost 12 clk per evaluation per SM whatever the layout.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/ubench_loads tools/ubench_loads.cu && gpurun_out/ubench_loads
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

// MODE: 0 LDS.64 broadcast   1 LDS.128 broadcast   2 LDS.128 lane-distinct (conflict-free)   3 LDS.64 lane-distinct
//       4 LDG.64 broadcast   5 LDG.128 broadcast   6 LDG.128 lane-distinct contiguous        7 LDG.64 lane-distinct contiguous
//       8 LDG.256 (v4.f64) broadcast   9 LDG.128, two addresses per warp (lanes 0-15 / 16-31 in different lines)
//       10 LDG.128 broadcast + 8 DFMA per load (does the FP64 pipe overlap the load path?)
template <int MODE>
__global__ void __launch_bounds__(256) k(const double *__restrict__ g, double *out, int iters, int stride) {
    extern __shared__ __align__(128) double sm[];
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = g[i];
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double acc0 = 0, acc1 = 0, acc2 = 0, acc3 = 0;
    // per-warp base so that different warps use different lines; (i * stride) walks through a 32 KB window (L1 resident)
    int off = warp * 64;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int base = (off + u * 16) & 4095 & ~15;
            if (MODE == 0) {
                acc0 += sm[base];
            } else if (MODE == 1) {
                const double2 v = *reinterpret_cast<const double2 *>(sm + base);
                acc0 += v.x; acc1 += v.y;
            } else if (MODE == 2) {
                const double2 v = *reinterpret_cast<const double2 *>(sm + ((base + 2 * lane) & 4095));
                acc0 += v.x; acc1 += v.y;
            } else if (MODE == 3) {
                acc0 += sm[(base + lane) & 4095];
            } else if (MODE == 4) {
                acc0 += __ldg(g + base);
            } else if (MODE == 5 || MODE == 10) {
                const double2 v = __ldg(reinterpret_cast<const double2 *>(g + base));
                if (MODE == 10) {
                    acc0 = fma(v.x, acc0, v.y); acc1 = fma(v.x, acc1, v.y); acc2 = fma(v.x, acc2, v.y); acc3 = fma(v.x, acc3, v.y);
                    acc0 = fma(v.y, acc0, v.x); acc1 = fma(v.y, acc1, v.x); acc2 = fma(v.y, acc2, v.x); acc3 = fma(v.y, acc3, v.x);
                } else {
                    acc0 += v.x; acc1 += v.y;
                }
            } else if (MODE == 6) {
                const double2 v = __ldg(reinterpret_cast<const double2 *>(g + ((base + 2 * lane) & 4095)));
                acc0 += v.x; acc1 += v.y;
            } else if (MODE == 7) {
                acc0 += __ldg(g + ((base + lane) & 4095));
            } else if (MODE == 8) {
                double a, b, c, d;
                asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(a), "=d"(b), "=d"(c), "=d"(d) : "l"(g + (base & ~3)));
                acc0 += a; acc1 += b; acc2 += c; acc3 += d;
            } else if (MODE == 9) {
                const double2 v = __ldg(reinterpret_cast<const double2 *>(g + ((base + (lane >> 4) * 16) & 4095)));
                acc0 += v.x; acc1 += v.y;
            }
        }
        off += stride;
    }
    if (acc0 + acc1 + acc2 + acc3 == 12345.678) out[0] = acc0;
}

template <int MODE>
void run(const char *name, const double *g, double *out, int sms, double clk_ghz, int bytes_per_lane) {
    const int iters = 20000, threads = 256, ctas = sms * 2;
    CK(cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768));
    k<MODE><<<ctas, threads, 32768>>>(g, out, 100, 16);
    CK(cudaDeviceSynchronize());
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    CK(cudaEventRecord(a));
    k<MODE><<<ctas, threads, 32768>>>(g, out, iters, 16);
    CK(cudaEventRecord(b));
    CK(cudaEventSynchronize(b));
    float ms;
    CK(cudaEventElapsedTime(&ms, a, b));
    const double warp_loads_per_sm = (double)iters * 8 * (threads / 32) * 2;  // 2 CTAs per SM
    const double clk = ms * 1e-3 * clk_ghz * 1e9;
    printf("%-58s %7.3f clk per warp-load per SM   %7.1f B/clk/SM delivered to registers\n", name, clk / warp_loads_per_sm,
           warp_loads_per_sm * 32 * bytes_per_lane / clk);
}

int main() {
    cudaDeviceProp p;
    CK(cudaGetDeviceProperties(&p, 0));
    int khz = 0;
    CK(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0));
    const double ghz = khz * 1e-6;
    printf("%s, %d SMs, %.3f GHz (attribute; under load the clock may differ)\n", p.name, p.multiProcessorCount, ghz);
    double *g, *out;
    CK(cudaMalloc(&g, 4096 * 8 + 4096));
    CK(cudaMalloc(&out, 64));
    CK(cudaMemset(g, 0, 4096 * 8 + 4096));
    const int sms = p.multiProcessorCount;
    run<0>("LDS.64   broadcast (all lanes one address)", g, out, sms, ghz, 8);
    run<1>("LDS.128  broadcast", g, out, sms, ghz, 16);
    run<2>("LDS.128  lane-distinct, conflict-free", g, out, sms, ghz, 16);
    run<3>("LDS.64   lane-distinct, conflict-free", g, out, sms, ghz, 8);
    run<4>("LDG.64   broadcast, L1 resident", g, out, sms, ghz, 8);
    run<5>("LDG.128  broadcast, L1 resident", g, out, sms, ghz, 16);
    run<6>("LDG.128  lane-distinct contiguous (4 lines)", g, out, sms, ghz, 16);
    run<7>("LDG.64   lane-distinct contiguous (2 lines)", g, out, sms, ghz, 8);
    run<8>("LDG.256  broadcast (ld.global.nc.v4.f64)", g, out, sms, ghz, 32);
    run<9>("LDG.128  two addresses per warp (2 lines)", g, out, sms, ghz, 16);
    run<10>("LDG.128  broadcast + 8 DFMA per load", g, out, sms, ghz, 16);
    return 0;
}
