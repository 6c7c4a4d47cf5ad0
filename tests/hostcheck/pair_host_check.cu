// Test harness (not product code): runs the __host__ __device__ per-pair functions of steps_b200/csrc/pair_generic.cuh and
// pair_generic_sym.cuh on the HOST, so that the CPU test tier can check, without a GPU, the numerical claims the action-reaction
// kernel of the table-lookup topologies rests on (the product only ever calls these functions from CUDA kernels):
//   (1) unit-mass factoring:  m_j * pair_exact_unit == the increment of pair_exact
//   (2) antisymmetry:         pair_exact_unit(i, j) == -pair_exact_unit(j, i) to rounding of the cell coordinate
//   (3) lean T^3 arithmetic:  pair_t3_fast_unit == pair_exact_unit<., 1> to a few ulp
// usage: pair_host_check t3 <table.bin> <ngrid> <L> <is_periodic> <n_pairs> <seed> <soft>
//        pair_host_check s1r2 <table.bin> <nrho> <nz> <rho_max> <L> <order> <n_pairs> <seed> <soft>
// prints: max relative differences (1) (2) (3), scale = |t| of the pair
//        pair_host_check forces_t3 | forces_s1r2 ...: a whole force evaluation formed pair-symmetrically (see forces_sym)
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <vector>
#include "../../steps_b200/csrc/pair_generic_sym.cuh"
using namespace steps;

static std::vector<double> load(const char *path) {
    FILE *f = fopen(path, "rb");
    if (!f) { perror(path); exit(3); }
    fseek(f, 0, SEEK_END);
    const long n = ftell(f) / (long)sizeof(double);
    fseek(f, 0, SEEK_SET);
    std::vector<double> v((size_t)n);
    if (fread(v.data(), sizeof(double), (size_t)n, f) != (size_t)n) exit(4);
    fclose(f);
    return v;
}

static double rel(const double a[3], const double b[3]) {
    const double d = std::sqrt((a[0] - b[0]) * (a[0] - b[0]) + (a[1] - b[1]) * (a[1] - b[1]) + (a[2] - b[2]) * (a[2] - b[2]));
    const double s = std::sqrt(b[0] * b[0] + b[1] * b[1] + b[2] * b[2]);
    return s > 0 ? d / s : d;
}

// the aligned row copies of a T^3 table (t3_lookup.cuh), built on the host with the function the device kernel runs per (copy, row)
template <int TOPO>
static std::vector<double> zwin_copy(TopoParams &tp) {
    std::vector<double> zw;
    if (TOPO == 1 && tp.is_periodic >= 2 && tp.table) {
        const int N = tp.dim0;
        zw.resize(t3_aligned_elems<double>(N));
        for (int c = 0; c < t3_copies<double>(); ++c)
            for (size_t row = 0; row < (size_t)N * N; ++row) t3_aligned_fill<double>(static_cast<const double *>(tp.table), N, c, row, zw.data());
        tp.table_zwin = zw.data();
    }
    return zw;
}

template <int TOPO>
static int run(TopoParams tp, int n_pairs, unsigned seed, double soft, double box_xy) {
    std::mt19937_64 rng(seed);
    std::uniform_real_distribution<double> U(0.0, 1.0);
    double e1 = 0, e2 = 0, e3 = 0;
    const std::vector<double> zw = zwin_copy<TOPO>(tp);
    const T3Lookup fk = t3_lookup_of<double>(tp);
    for (int p = 0; p < n_pairs; ++p) {
        double xi[3], xj[3];
        for (int k = 0; k < 3; ++k) {
            const double ext = (TOPO == 1 || k == 2) ? tp.L : box_xy;
            xi[k] = U(rng) * ext;
            xj[k] = U(rng) * ext;
        }
        if (p % 7 == 0) for (int k = 0; k < 3; ++k) xj[k] = xi[k] + (U(rng) - 0.5) * 3.0 * soft;  // close pairs: the softened branches
        if (TOPO == 1 && p % 11 == 0) xj[p % 3] = std::fmod(xi[p % 3] + 0.5 * tp.L, tp.L);       // |d| = L/2: the wrap boundary
        const double si = soft * (0.5 + U(rng)), sj = soft * (0.5 + U(rng)), mj = 0.5 + U(rng);
        double t[3], tr[3], ax = 0, ay = 0, az = 0;
        pair_exact_unit<double, TOPO>(tp, xi[0], xi[1], xi[2], si, xj[0], xj[1], xj[2], sj, t[0], t[1], t[2]);
        pair_exact<double, TOPO>(tp, xi[0], xi[1], xi[2], si, xj[0], xj[1], xj[2], mj, sj, ax, ay, az);
        const double inc[3] = {ax, ay, az}, mt[3] = {mj * t[0], mj * t[1], mj * t[2]};
        e1 = std::fmax(e1, rel(mt, inc));
        pair_exact_unit<double, TOPO>(tp, xj[0], xj[1], xj[2], sj, xi[0], xi[1], xi[2], si, tr[0], tr[1], tr[2]);
        const double neg[3] = {-tr[0], -tr[1], -tr[2]};
        e2 = std::fmax(e2, rel(neg, t));
        if (TOPO == 1 && tp.is_periodic >= 2) {
            double tf[3];
            pair_t3_fast_unit<double>(fk, xi[0], xi[1], xi[2], si, xj[0], xj[1], xj[2], sj, tf[0], tf[1], tf[2]);
            e3 = std::fmax(e3, rel(tf, t));
        }
    }
    printf("%.6e %.6e %.6e\n", e1, e2, e3);
    fflush(stdout);
    return 0;
}

// whole force evaluation the way the action-reaction kernel forms it: every unordered pair once (t applied to both particles with
// opposite signs), the self pair one-sidedly; LEAN = the lean T^3 arithmetic.  state.bin = x[3n] M[n] soft[n] (doubles).
template <int TOPO>
static int forces_sym(TopoParams tp, const char *state_path, int n, bool lean, const char *out_path) {
    const std::vector<double> st = load(state_path);
    if ((int)st.size() != 5 * n) return 6;
    const double *x = st.data(), *M = x + 3 * n, *S = M + n;
    const std::vector<double> zw = zwin_copy<TOPO>(tp);
    const T3Lookup fk = t3_lookup_of<double>(tp);
    std::vector<double> F(3 * (size_t)n, 0.0);
    for (int i = 0; i < n; ++i) {
        pair_exact<double, TOPO>(tp, x[3 * i], x[3 * i + 1], x[3 * i + 2], S[i], x[3 * i], x[3 * i + 1], x[3 * i + 2], M[i], S[i], F[3 * i], F[3 * i + 1],
                                 F[3 * i + 2]);
        for (int j = i + 1; j < n; ++j) {
            double t[3];
            if (TOPO == 1 && lean)
                pair_t3_fast_unit<double>(fk, x[3 * i], x[3 * i + 1], x[3 * i + 2], S[i], x[3 * j], x[3 * j + 1],
                                          x[3 * j + 2], S[j], t[0], t[1], t[2]);
            else
                pair_exact_unit<double, TOPO>(tp, x[3 * i], x[3 * i + 1], x[3 * i + 2], S[i], x[3 * j], x[3 * j + 1], x[3 * j + 2], S[j], t[0], t[1], t[2]);
            for (int k = 0; k < 3; ++k) {
                F[3 * i + k] += M[j] * t[k];
                F[3 * j + k] -= M[i] * t[k];
            }
        }
    }
    FILE *f = fopen(out_path, "wb");
    if (!f) return 7;
    fwrite(F.data(), sizeof(double), F.size(), f);
    fclose(f);
    return 0;
}

int main(int argc, char **argv) {
    if (argc < 2) return 2;
    TopoParams tp{};
    // pair_host_check forces_t3 <table.bin> <ngrid> <L> <is_periodic> <state.bin> <n> <lean 0|1> <out.bin>
    if (!strcmp(argv[1], "forces_t3") && argc == 10) {
        const std::vector<double> tab = load(argv[2]);
        tp.topology = 1;
        tp.dim0 = tp.dim1 = atoi(argv[3]);
        tp.L = atof(argv[4]);
        tp.is_periodic = atoi(argv[5]);
        tp.table = tab.data();
        return forces_sym<1>(tp, argv[6], atoi(argv[7]), atoi(argv[8]) != 0, argv[9]);
    }
    // pair_host_check forces_s1r2 <table.bin> <nrho> <nz> <rho_max> <L> <order> <state.bin> <n> <out.bin>
    if (!strcmp(argv[1], "forces_s1r2") && argc == 11) {
        const std::vector<double> tab = load(argv[2]);
        tp.topology = 2;
        tp.dim0 = atoi(argv[3]);
        tp.dim1 = atoi(argv[4]);
        tp.rho_max = atof(argv[5]);
        tp.L = atof(argv[6]);
        tp.order = atoi(argv[7]);
        tp.is_periodic = 2;
        tp.table = tab.data();
        return forces_sym<2>(tp, argv[8], atoi(argv[9]), false, argv[10]);
    }
    if (!strcmp(argv[1], "t3") && argc == 9) {
        const std::vector<double> tab = load(argv[2]);
        tp.topology = 1;
        tp.dim0 = tp.dim1 = atoi(argv[3]);
        tp.L = atof(argv[4]);
        tp.is_periodic = atoi(argv[5]);
        tp.table = tab.data();
        if (tab.size() != (size_t)tp.dim0 * tp.dim0 * tp.dim0 * 3) return 5;
        return run<1>(tp, atoi(argv[6]), (unsigned)atoi(argv[7]), atof(argv[8]), tp.L);
    }
    if (!strcmp(argv[1], "s1r2") && argc == 11) {
        const std::vector<double> tab = load(argv[2]);
        tp.topology = 2;
        tp.dim0 = atoi(argv[3]);
        tp.dim1 = atoi(argv[4]);
        tp.rho_max = atof(argv[5]);
        tp.L = atof(argv[6]);
        tp.order = atoi(argv[7]);
        tp.is_periodic = 2;
        tp.table = tab.data();
        if (tab.size() != (size_t)tp.dim0 * tp.dim1 * 2) return 5;
        return run<2>(tp, atoi(argv[8]), (unsigned)atoi(argv[9]), atof(argv[10]), 0.4 * tp.rho_max);
    }
    return 2;
}
