// ewald_t3.cuh -- GPU builder of the T^3 Ewald force-correction lookup table.
//
// Replaces calculate_t3_ewald_lookup_table() + ewald_force_correction() + ewald_space()
// (StePS/src/ewald_space.cc:288-383, :198-286, :118-196; set up by main.cc:425-494): for every grid point of the
// fundamental wedge i >= j >= k >= Ngrid/2 of the cell-centred Ngrid^3 grid on [-L/2, L/2)^3,
//     D(r) = F_real(r) + F_rec(r) - F_newton(r)          (Hernquist, Bouchet & Suto 1991)
//     F_real = - sum_n (r + nL) [erfc(a R) + 2 a R exp(-a^2 R^2)/sqrt(pi)] / R^3,   R = |r + nL| <= rel_cut L
//     F_rec  = - sum_h k (4 pi / V) exp(-k^2 / 4 a^2) sin(k.r) / k^2,               k = 2 pi h / L, |k| <= rec_cut
//     F_newton = - r / |r|^3
// and the other 47 images of the point are filled by the octahedral symmetry of the cube.
// One thread per wedge point; the two lattice sums run in the reference's index order inside the thread over lists prepared on
// the host (translations n L; wave vectors with their point-independent amplitudes, skipped modes dropped), so the value of
// an entry differs from the reference's only by the last bits of erfc/exp/sin (CUDA math library vs glibc).  The images
// are written in the reference's order (24 proper rotations, each followed by its inversion; last write wins on the
// symmetry planes, where several images coincide).  Note the reference's convention, kept here: rec_cut is compared with
// |k| in physical units while the index list is generated in integer units (ewald_space.cc:253-256, main.cc:468).
#pragma once
#include <cmath>
#include <vector>
#include <cuda_runtime.h>

namespace steps {

struct EwaldIdx {
    int x, y, z, n2;
};

// lattice vectors with |n| < R in the reference's enumeration order (ewald_space.cc:118-196)
inline void build_ewald_space(double R, std::vector<EwaldIdx> &out) {
    out.clear();
    for (int i = 0; i < R; ++i)
        for (int j = 0; j < R; ++j)
            for (int k = 0; k < R; ++k) {
                const int n2 = i * i + j * j + k * k;
                if (!((double)n2 < R * R)) continue;
                out.push_back({i, j, k, n2});
                if (i != 0) out.push_back({-i, j, k, n2});
                if (j != 0) out.push_back({i, -j, k, n2});
                if (k != 0) out.push_back({i, j, -k, n2});
                if (j != 0 && k != 0) out.push_back({i, -j, -k, n2});
                if (i != 0 && k != 0) out.push_back({-i, j, -k, n2});
                if (i != 0 && j != 0) out.push_back({-i, -j, k, n2});
                if (i != 0 && j != 0 && k != 0) out.push_back({-i, -j, -k, n2});
            }
}

// the 24 proper rotations of the cube as (axis permutation, signs), in the reference's order (ewald_space.cc:62-100):
// even permutations x sign triples of product +1, then odd permutations x sign triples of product -1
struct CubeRot {
    signed char perm[3], sign[3];
};
__host__ __device__ inline CubeRot cube_rotation(int q) {
    const signed char even[3][3] = {{0, 1, 2}, {1, 2, 0}, {2, 0, 1}};
    const signed char odd[3][3] = {{0, 2, 1}, {2, 1, 0}, {1, 0, 2}};
    const signed char pos[4][3] = {{1, 1, 1}, {1, -1, -1}, {-1, 1, -1}, {-1, -1, 1}};
    const signed char neg[4][3] = {{-1, -1, -1}, {-1, 1, 1}, {1, -1, 1}, {1, 1, -1}};
    CubeRot r;
    const int p = (q % 12) / 4, s = q % 4;
    for (int c = 0; c < 3; ++c) {
        r.perm[c] = q < 12 ? even[p][c] : odd[p][c];
        r.sign[c] = q < 12 ? pos[s][c] : neg[s][c];
    }
    return r;
}

// What a lattice sum needs per term, hoisted out of the per-point loops and prepared once on the host (t3_ewald_prepare):
//   real space: the translation n L itself;
//   reciprocal space: the wave vector k = 2 pi h / L and the point-independent amplitude vector -k (4 pi / V) exp(-k^2/4a^2) / k^2,
//   with the modes the reference skips (h = 0, |k| beyond the cut) already dropped.  List order = the reference's summation order.
struct LatticeShift {
    double sx, sy, sz;
};
struct RecipMode {
    double kx, ky, kz;
    double ax, ay, az;
};

struct T3EwaldParams {
    int ngrid;
    double L, alpha, rel_cut, rec_cut;
    int n_real, n_rec;  // entries of the prepared lists
};

inline void t3_ewald_prepare(T3EwaldParams &p, std::vector<LatticeShift> &shifts, std::vector<RecipMode> &modes) {
    const double pi = 3.14159265358979323846;
    std::vector<EwaldIdx> real_idx, rec_idx;
    build_ewald_space(p.rel_cut + 1.0, real_idx);  // main.cc:467-468
    build_ewald_space(p.rec_cut + 2.0, rec_idx);
    shifts.clear();
    for (const EwaldIdx &n : real_idx) shifts.push_back({(double)n.x * p.L, (double)n.y * p.L, (double)n.z * p.L});
    modes.clear();
    const double unit_k = 2.0 * pi / p.L;
    const double volume_factor = (4.0 * pi) / (p.L * p.L * p.L);
    for (const EwaldIdx &h : rec_idx) {
        if (h.x == 0 && h.y == 0 && h.z == 0) continue;
        const double kx = unit_k * (double)h.x, ky = unit_k * (double)h.y, kz = unit_k * (double)h.z;
        const double k2 = kx * kx + ky * ky + kz * kz;
        if (k2 == 0.0 || k2 > p.rec_cut * p.rec_cut) continue;
        const double amp = volume_factor * exp(-k2 / (4.0 * p.alpha * p.alpha)) / k2;
        modes.push_back({kx, ky, kz, -kx * amp, -ky * amp, -kz * amp});
    }
    p.n_real = (int)shifts.size();
    p.n_rec = (int)modes.size();
}

// short-range part: the screened Newtonian pull of every lattice image within rel_cut L of the point
__host__ __device__ inline void t3_ewald_real_part(const double r[3], const T3EwaldParams &p, const LatticeShift *__restrict__ shifts, double out[3]) {
    const double two_a_over_sqrtpi = 2.0 * p.alpha / sqrt(3.14159265358979323846);
    const double a2 = p.alpha * p.alpha;
    const double reach2 = (p.rel_cut * p.L) * (p.rel_cut * p.L);
    double fx = 0.0, fy = 0.0, fz = 0.0;
    for (int m = 0; m < p.n_real; ++m) {
        const LatticeShift s = shifts[m];
        const double X = r[0] + s.sx, Y = r[1] + s.sy, Z = r[2] + s.sz;
        const double d2 = X * X + Y * Y + Z * Z;
        if (d2 == 0.0 || d2 > reach2) continue;
        const double d = sqrt(d2);
        const double inv = 1.0 / d;
        const double screened = (inv * inv * inv) * (erfc(p.alpha * d) + two_a_over_sqrtpi * d * exp(-a2 * d2));
        fx += -X * screened;
        fy += -Y * screened;
        fz += -Z * screened;
    }
    out[0] = fx; out[1] = fy; out[2] = fz;
}

// long-range part: one sine per surviving mode times its prepared amplitude vector
__host__ __device__ inline void t3_ewald_recip_part(const double r[3], const T3EwaldParams &p, const RecipMode *__restrict__ modes, double out[3]) {
    double fx = 0.0, fy = 0.0, fz = 0.0;
    for (int m = 0; m < p.n_rec; ++m) {
        const RecipMode q = modes[m];
        const double s = sin(q.kx * r[0] + q.ky * r[1] + q.kz * r[2]);
        fx += q.ax * s;
        fy += q.ay * s;
        fz += q.az * s;
    }
    out[0] = fx; out[1] = fy; out[2] = fz;
}

// D(r) = periodic force - Newtonian force of the nearest image (ewald_space.cc:198-286); zero at the origin by definition
__host__ __device__ inline void t3_ewald_point(const double r[3], const T3EwaldParams &p, const LatticeShift *__restrict__ shifts,
                                               const RecipMode *__restrict__ modes, double D[3]) {
    const double r2 = r[0] * r[0] + r[1] * r[1] + r[2] * r[2];
    if (!(r2 > 0.0)) {
        D[0] = D[1] = D[2] = 0.0;
        return;
    }
    double near_part[3], far_part[3];
    t3_ewald_real_part(r, p, shifts, near_part);
    t3_ewald_recip_part(r, p, modes, far_part);
    const double rn = sqrt(r2);
    const double inv3 = 1.0 / (rn * rn * rn);
    for (int c = 0; c < 3; ++c) D[c] = (near_part[c] + far_part[c]) - (-r[c] * inv3);
}

__host__ __device__ inline double t3_cell_centre(int i, double h, double L) {
#ifdef __CUDA_ARCH__
    return __dadd_rn(__dmul_rn((double)i + 0.5, h), -(L / 2.0));
#else
    volatile double prod = ((double)i + 0.5) * h;  // volatile: no contraction on hosts that have FMA either
    return prod - L / 2.0;
#endif
}

// value at wedge point (i,j,k) and its images under the cube group (ewald_space.cc:316-376)
__host__ __device__ inline void t3_ewald_fill(int i, int j, int k, const T3EwaldParams &p, const LatticeShift *__restrict__ shifts,
                                              const RecipMode *__restrict__ modes, double *__restrict__ table) {
    const int N = p.ngrid;
    const double h = p.L / (double)N;
    // cell-centre coordinate (i + 1/2) h - L/2, product and difference rounded separately as in the reference build.  A fused
    // multiply-add would leave ~1e-15 instead of 0 on the symmetry planes (i = N/2 of the odd grid), and at the centre point
    // that turns D = 0 (ewald_space.cc:280-283) into the difference of two terms of size 1/r^2 = 1e30.
    const double r[3] = {t3_cell_centre(i, h, p.L), t3_cell_centre(j, h, p.L), t3_cell_centre(k, h, p.L)};
    double Dw[3];
    t3_ewald_point(r, p, shifts, modes, Dw);
    auto at = [N](int a, int b, int c) { return ((size_t)((a * N + b) * N + c)) * 3u; };
    {
        const size_t o = at(i, j, k);
        table[o] = Dw[0]; table[o + 1] = Dw[1]; table[o + 2] = Dw[2];
    }
    const int src[3] = {i, j, k};
    for (int q = 0; q < 24; ++q) {
        const CubeRot rot = cube_rotation(q);
        int dst[3];
        double Dr[3];
        for (int d = 0; d < 3; ++d) {
            int v = src[rot.perm[d]];
            if (rot.sign[d] < 0) v = (N - 1) - v;  // reflection about the box centre
            dst[d] = v;
            Dr[d] = (double)rot.sign[d] * Dw[rot.perm[d]];
        }
        size_t o = at(dst[0], dst[1], dst[2]);
        table[o] = Dr[0]; table[o + 1] = Dr[1]; table[o + 2] = Dr[2];
        o = at((N - 1) - dst[0], (N - 1) - dst[1], (N - 1) - dst[2]);  // inversion: D(-r) = -D(r)
        table[o] = -Dr[0]; table[o + 1] = -Dr[1]; table[o + 2] = -Dr[2];
    }
}

// one thread per wedge point; points = packed (i, j, k) triples
__global__ void t3_ewald_table_kernel(const int *__restrict__ points, int n_points, const T3EwaldParams p,
                                      const LatticeShift *__restrict__ shifts, const RecipMode *__restrict__ modes, double *__restrict__ table) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_points) return;
    t3_ewald_fill(points[3 * t], points[3 * t + 1], points[3 * t + 2], p, shifts, modes, table);
}

}  // namespace steps
