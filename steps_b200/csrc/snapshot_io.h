// snapshot_io.h -- ASCII snapshots in the reference's format without stalling the GPUs (the HDF5 formats need libhdf5,
// which this build does not have).  Host code only; included by engine.cu.
//
// Format = write_ascii_snapshot (inputoutput.cc:826-909): one line per particle,
//     x*H0_dimless  y*H0_dimless  z*H0_dimless  vx*sqrt(a)*UNIT_V  vy*..  vz*..  M*H0_dimless   each printed "%.16f\t", then "\n"
// (H0_dimless is a REAL and multiplies in REAL precision; the velocity factor multiplies in double; a GLASS_MAKING build prints 0.0
// velocities).  The reference formats 7N numbers with fprintf on rank 0 while every GPU waits (about 10 s at N = 2M, four KDK steps of
// one B200).  Here the state leaves the device by asynchronous copies into pinned staging buffers, and a background thread formats it with
// a pool of workers and writes the file, while the caller goes on stepping.
#pragma once
#include <atomic>
#include <cmath>
#include <cstdio>
#include <string>
#include <thread>
#include <vector>

namespace steps {

constexpr double SNAP_UNIT_V = 20.738652969925447;  // global_variables.h:17

template <typename T>
static void snapshot_format_range(const T *x, const T *v, const T *M, size_t lo, size_t hi, T h0_dimless, double sqrt_a, int zero_v, std::string &out) {
    out.clear();
    out.reserve((hi - lo) * 180);
    char buf[512];
    // v*sqrt(a)*UNIT_V as the reference's build evaluates it: its Makefile compiles with -O3 -ffast-math (Template-LinuxGCC-Makefile:81),
    // which hoists the loop-invariant factor sqrt(a)*UNIT_V; files are byte-identical to that build's (tests/test_snapshot_io.py)
    const double vfac = sqrt_a * SNAP_UNIT_V;
    for (size_t i = lo; i < hi; ++i) {
        int len = 0;
        for (int k = 0; k < 3; ++k) len += snprintf(buf + len, sizeof(buf) - len, "%.16f\t", (double)(T)(x[3 * i + k] * h0_dimless));
        for (int k = 0; k < 3; ++k) {
            const double vv = zero_v ? 0.0 : (double)v[3 * i + k] * vfac;
            len += snprintf(buf + len, sizeof(buf) - len, "%.16f\t", vv);
        }
        len += snprintf(buf + len, sizeof(buf) - len, "%.16f\t\n", (double)(T)(M[i] * h0_dimless));
        out.append(buf, (size_t)len);
    }
}

// formats [0, n) with `nthreads` workers in batches and writes the file in particle order; returns "" or an error text
template <typename T>
static std::string snapshot_write_ascii(const char *path, const T *x, const T *v, const T *M, size_t n, double h0_dimless, double a, int zero_v,
                                        int nthreads) {
    FILE *f = fopen(path, "w");
    if (!f) return std::string("cannot open ") + path;
    const T h0 = (T)h0_dimless;
    const double sqrt_a = std::sqrt(a);
    if (nthreads < 1) nthreads = (int)std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
    const size_t chunk = 16384;
    std::vector<std::string> bufs((size_t)nthreads);
    bool ok = true;
    for (size_t base = 0; base < n && ok; base += chunk * (size_t)nthreads) {
        std::vector<std::thread> th;
        int used = 0;
        for (int t = 0; t < nthreads; ++t) {
            const size_t lo = base + (size_t)t * chunk;
            if (lo >= n) break;
            const size_t hi = std::min(n, lo + chunk);
            ++used;
            if (nthreads == 1) snapshot_format_range<T>(x, v, M, lo, hi, h0, sqrt_a, zero_v, bufs[t]);
            else th.emplace_back([=, &bufs] { snapshot_format_range<T>(x, v, M, lo, hi, h0, sqrt_a, zero_v, bufs[t]); });
        }
        for (auto &w : th) w.join();
        for (int t = 0; t < used && ok; ++t) ok = fwrite(bufs[t].data(), 1, bufs[t].size(), f) == bufs[t].size();
    }
    if (fclose(f) != 0) ok = false;
    return ok ? std::string() : std::string("write error on ") + path;
}

// one asynchronous snapshot in flight per group
struct SnapshotJob {
    std::thread th;
    bool running = false;
    std::string err;                                   // set by the writer thread, read after join
    void *hx = nullptr, *hv = nullptr, *hm = nullptr;  // pinned staging: x[3N], v[3N], M[N]
    size_t n = 0, real_bytes = 0;
};

}  // namespace steps
