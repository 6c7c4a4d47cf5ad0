// Test harness (not product code): runs the __host__ __device__ table-producer functions of steps_b200/csrc/ewald_t3.cuh,
// ewald_s1r2.cuh and radial_table.cuh on the HOST, so that the CPU test tier can compare the restated algorithms with the
// reference's own builders (oracle/_ref) without a GPU.  The product only ever launches them as CUDA kernels.
//   table_host_check t3 <ngrid> <L> <rel_cut> <rec_cut> <out.bin>
//   table_host_check s1r2 <nrho> <nz> <rho_max> <Lz> <alpha> <nmax> <mmax> <out.bin>
//   table_host_check radial <R> <Lz> <size> <accuracy> <out.bin>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "../../steps_b200/csrc/ewald_t3.cuh"
#include "../../steps_b200/csrc/ewald_s1r2.cuh"
#include "../../steps_b200/csrc/radial_table.cuh"
using namespace steps;

static int dump(const char *path, const std::vector<double> &v) {
    FILE *f = fopen(path, "wb");
    if (!f) return 1;
    fwrite(v.data(), sizeof(double), v.size(), f);
    fclose(f);
    return 0;
}

int main(int argc, char **argv) {
    if (argc < 2) return 2;
    if (!strcmp(argv[1], "t3") && argc == 7) {
        const int ngrid = atoi(argv[2]);
        const double L = atof(argv[3]), rel = atof(argv[4]), rec = atof(argv[5]);
        T3EwaldParams p{ngrid, L, 2.0 / L, rel, rec, 0, 0};
        std::vector<LatticeShift> re;
        std::vector<RecipMode> rc;
        t3_ewald_prepare(p, re, rc);
        std::vector<double> tab((size_t)ngrid * ngrid * ngrid * 3, 0.0);
#pragma omp parallel for schedule(dynamic)
        for (int i = ngrid / 2; i < ngrid; ++i)
            for (int j = ngrid / 2; j <= i; ++j)
                for (int k = ngrid / 2; k <= j; ++k) t3_ewald_fill(i, j, k, p, re.data(), rc.data(), tab.data());
        return dump(argv[6], tab);
    }
    if (!strcmp(argv[1], "s1r2") && argc == 10) {
        const S1R2EwaldParams p{atoi(argv[2]), atoi(argv[3]), atoi(argv[7]), atoi(argv[8]), atof(argv[4]), atof(argv[5]), atof(argv[6])};
        std::vector<double> tab((size_t)p.nrho * p.nz * 2);
#pragma omp parallel for
        for (int t = 0; t < p.nrho * p.nz; ++t) s1r2_ewald_cell(t / p.nz, t % p.nz, p, tab[2 * (size_t)t], tab[2 * (size_t)t + 1]);
        return dump(argv[9], tab);
    }
    if (!strcmp(argv[1], "radial") && argc == 7) {
        const double R = atof(argv[2]), Lz = atof(argv[3]);
        const int size = atoi(argv[4]), acc = atoi(argv[5]);
        std::vector<double> t(size, 0.0);
#pragma omp parallel for
        for (int i = 1; i < size; ++i) t[i] = radial_table_entry(i, R, Lz, size, acc);
        if (size >= 3) t[0] = radial_table_origin(R, size, t[1], t[2]);
        return dump(argv[6], t);
    }
    return 2;
}
