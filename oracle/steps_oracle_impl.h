/* steps_oracle_impl.h -- included twice by steps_oracle.c with REAL = double / float and SFX = _f64 / _f32.
 * TEST INFRASTRUCTURE.  Every function cites the reference lines it restates (StePS/src/...). */

#define CAT_(a, b) a##b
#define CAT(a, b) CAT_(a, b)
#define FN(name) CAT(name, SFX)

/* forces.cc:52-87 force_softening: cubic-spline softened 1/r^3 (CPU form: pow) */
REAL FN(oracle_force_softening)(REAL r, REAL beta)
{
    REAL betap2 = beta * (REAL)0.5;
    REAL wij = 0;
    if (r >= beta) {
        wij = (REAL)pow(r, -3);
    } else if (r > betap2 && r < beta) {
        REAL c0 = (REAL)(-32.0 / (3.0 * pow(beta, 6)));
        REAL c1 = (REAL)(38.4 / pow(beta, 5));
        REAL c2 = (REAL)(-48.0 / pow(beta, 4));
        REAL c3 = (REAL)(64.0 / (3.0 * pow(beta, 3)));
        REAL c4 = (REAL)(-1.0 / 15.0);
        wij = (REAL)(c0 * pow(r, 3) + c1 * pow(r, 2) + c2 * r + c3 + c4 / pow(r, 3));
    } else {
        REAL c0 = (REAL)(32.0 / pow(beta, 6));
        REAL c1 = (REAL)(-38.4 / pow(beta, 5));
        REAL c2 = (REAL)(32.0 / (3.0 * pow(beta, 3)));
        wij = (REAL)(c0 * pow(r, 3) + c1 * pow(r, 2) + c2);
    }
    return wij;
}

/* utils.cc:59-82 calculate_softening_length */
void FN(oracle_softening)(const REAL *M, int n, REAL particle_radii, REAL *soft, REAL *M_min_out, REAL *rho_part_out)
{
    REAL M_min = M[0];
    for (int i = 0; i < n; i++) if (M_min > M[i]) M_min = M[i];
    REAL rho_part = (REAL)(M_min / (4.0 * ORACLE_PI * pow(particle_radii, 3.0) / 3.0));
    REAL const_beta = (REAL)(3.0 / rho_part / (4.0 * ORACLE_PI));
    for (int i = 0; i < n; i++) soft[i] = (REAL)cbrt(M[i] * const_beta);
    if (M_min_out) *M_min_out = M_min;
    if (rho_part_out) *rho_part_out = rho_part;
}

/* ewald_space.cc:392-409 imodp / map_to_centered_grid, :439-449 get_cubic_weights */
static void FN(map_grid)(REAL r, REAL L, int Ngrid, int *i0, REAL *fx)
{
    REAL gs = L / (REAL)Ngrid;
    REAL u = (r + L * (REAL)0.5) / gs - (REAL)0.5;
    REAL uf = (REAL)floor(u);
    *i0 = imodp_c((int)uf, Ngrid);
    *fx = (REAL)(u - uf);
}
static void FN(cubic_w)(REAL t, REAL w[4])
{
    REAL t2 = t * t, t3 = t * t2;
    w[0] = (REAL)(-0.5 * t3 + t2 - 0.5 * t);
    w[1] = (REAL)(1.5 * t3 - 2.5 * t2 + 1.0);
    w[2] = (REAL)(-1.5 * t3 + 2.0 * t2 + 0.5 * t);
    w[3] = (REAL)(0.5 * t3 - 0.5 * t2);
}
/* ewald_space.cc:451-523 ewald_interpolate_D with order 4 (tricubic, 4^3 wrapped neighbours from i0-1) */
static void FN(t3_interp)(int Ngrid, REAL L, const REAL *table, REAL dx, REAL dy, REAL dz, REAL D[3])
{
    int ix0, iy0, iz0;
    REAL fx, fy, fz, wx[4], wy[4], wz[4];
    FN(map_grid)(dx, L, Ngrid, &ix0, &fx);
    FN(map_grid)(dy, L, Ngrid, &iy0, &fy);
    FN(map_grid)(dz, L, Ngrid, &iz0, &fz);
    FN(cubic_w)(fx, wx); FN(cubic_w)(fy, wy); FN(cubic_w)(fz, wz);
    REAL s[3] = {0, 0, 0};
    for (int i = 0; i < 4; i++) {
        int ix = imodp_c(ix0 - 1 + i, Ngrid);
        for (int j = 0; j < 4; j++) {
            int iy = imodp_c(iy0 - 1 + j, Ngrid);
            REAL wxy = wx[i] * wy[j];
            for (int k = 0; k < 4; k++) {
                int iz = imodp_c(iz0 - 1 + k, Ngrid);
                REAL w = wxy * wz[k];
                size_t idx = ((size_t)((ix * Ngrid + iy) * Ngrid + iz)) * 3u;
                s[0] += w * table[idx]; s[1] += w * table[idx + 1]; s[2] += w * table[idx + 2];
            }
        }
    }
    D[0] = s[0]; D[1] = s[1]; D[2] = s[2];
}

/* ewald_space.cc:803-838 NGP, :840-895 CIC, :898-1025 TSC on the (rho,z) table; rho clamped, z wrapped */
static void FN(s1r2_table)(const oracle_params *p, REAL rho, REAL z, REAL *Drho, REAL *Dz)
{
    const REAL *T = (const REAL *)p->ewald_table;
    const int Nrho = p->table_dim0, Nz = p->table_dim1;
    const REAL rho_max = (REAL)((REAL)2.25 * (REAL)p->Rsim), Lz = (REAL)p->L;
    if (rho < 0) rho = 0;
    if (rho > rho_max) rho = rho_max;
    const REAL half = (REAL)0.5 * Lz;
    const REAL drho = rho_max / (REAL)(Nrho - 1 > 1 ? Nrho - 1 : 1);
    const REAL dz = Lz / (REAL)Nz;
    const REAL ur = (drho > 0) ? (rho / drho) : 0;
    const REAL uz = (z + half) / dz - (REAL)0.5;
    if (p->interp_order == 0) {
        int ir = (int)floor(ur + (REAL)0.5), iz = (int)floor(uz + (REAL)0.5);
        if (ir < 0) ir = 0;
        if (ir > Nrho - 1) ir = Nrho - 1;
        iz = imodp_c(iz, Nz);
        size_t b = ((size_t)ir * Nz + iz) * 2u;
        *Drho = T[b]; *Dz = T[b + 1];
    } else if (p->interp_order == 2) {
        int ir0 = (int)floor(ur);
        REAL fr = ur - (REAL)ir0;
        if (ir0 < 0) { ir0 = 0; fr = 0; }
        if (ir0 > Nrho - 2) { ir0 = Nrho - 2 > 0 ? Nrho - 2 : 0; fr = 1; }
        int ir1 = ir0 + 1;
        int iz0 = (int)floor(uz);
        REAL fz = uz - (REAL)iz0;
        iz0 = imodp_c(iz0, Nz);
        int iz1 = imodp_c(iz0 + 1, Nz);
        REAL w00 = (1 - fr) * (1 - fz), w10 = fr * (1 - fz), w01 = (1 - fr) * fz, w11 = fr * fz;
#define TG(ir, iz, c) T[((size_t)(ir) * Nz + (iz)) * 2u + (c)]
        *Drho = w00 * TG(ir0, iz0, 0) + w10 * TG(ir1, iz0, 0) + w01 * TG(ir0, iz1, 0) + w11 * TG(ir1, iz1, 0);
        *Dz = w00 * TG(ir0, iz0, 1) + w10 * TG(ir1, iz0, 1) + w01 * TG(ir0, iz1, 1) + w11 * TG(ir1, iz1, 1);
    } else {
        int jr = (int)floor(ur + (REAL)0.5), jz = (int)floor(uz + (REAL)0.5);
        REAL sr = ur - (REAL)jr, sz = uz - (REAL)jz;
        REAL wr[3] = {(REAL)0.5 * ((REAL)0.5 - sr) * ((REAL)0.5 - sr), (REAL)0.75 - sr * sr, (REAL)0.5 * ((REAL)0.5 + sr) * ((REAL)0.5 + sr)};
        REAL wz[3] = {(REAL)0.5 * ((REAL)0.5 - sz) * ((REAL)0.5 - sz), (REAL)0.75 - sz * sz, (REAL)0.5 * ((REAL)0.5 + sz) * ((REAL)0.5 + sz)};
        int ir[3] = {jr - 1, jr, jr + 1};
        if (ir[0] < 0) ir[0] = 0;
        if (ir[1] < 0) ir[1] = 0;
        if (ir[1] > Nrho - 1) ir[1] = Nrho - 1;
        if (ir[2] > Nrho - 1) ir[2] = Nrho - 1;
        int iz[3] = {imodp_c(jz - 1, Nz), imodp_c(jz, Nz), imodp_c(jz + 1, Nz)};
        REAL d0 = 0, d1 = 0;
        for (int q = 0; q < 3; q++) {
            REAL w0 = wr[0] * wz[q], w1 = wr[1] * wz[q], w2 = wr[2] * wz[q];
            d0 += w0 * TG(ir[0], iz[q], 0) + w1 * TG(ir[1], iz[q], 0) + w2 * TG(ir[2], iz[q], 0);
            d1 += w0 * TG(ir[0], iz[q], 1) + w1 * TG(ir[1], iz[q], 1) + w2 * TG(ir[2], iz[q], 1);
        }
#undef TG
        *Drho = d0; *Dz = d1;
    }
}

/* forces.cc:279-316 get_cylindrical_force_correction with ORDER == 1 (utils.cc:24-38 linear_interpolation) */
static REAL FN(cyl_corr)(REAL r, REAL R, const REAL *tab, int size)
{
    REAL step = R / (REAL)size;
    int i = (int)floor(r / R * (size - 1));
    REAL corr = tab[size - 1];
    if (i < size - 1) {
        REAL X1 = step * i, Y1 = tab[i], X2 = step * (i + 1), Y2 = tab[i + 1];
        REAL A = (Y2 - Y1) / (X2 - X1);
        REAL B = Y1 - A * X1;
        corr = A * r + B;
    }
    return corr;
}

/* forces.cc:510-577 forces (R^3), :776-876 forces_periodic (T^3), :1221-1404 forces_periodic_z (S^1xR^2).
 * j summed in order 0..N-1 per i (the reference's atomics add into F[i] in j order). */
void FN(oracle_forces)(const oracle_params *p, const REAL *x, const REAL *M, const REAL *soft, REAL *F, int id_min, int id_max)
{
    const int N = p->n, topo = p->topology, isp = p->is_periodic;
    const REAL L = (REAL)p->L;
    const REAL DE = (REAL)((REAL)p->H0 * p->H0 * p->Omega_lambda);          /* forces.cc:513 */
    const REAL mius = (REAL)p->mass_in_unit_sphere;
    const int ewald_max = isp + 1;                                          /* main.cc:1270 */
    const REAL ewald_cut = ((REAL)ewald_max) - (REAL)0.4;                   /* main.cc:1271 */
    if (p->nthreads > 0) omp_set_num_threads(p->nthreads);
#pragma omp parallel for schedule(dynamic, 16)
    for (int i = id_min; i <= id_max; i++) {
        REAL fx = 0, fy = 0, fz = 0;
        for (int j = 0; j < N; j++) {
            REAL beta = soft[i] + soft[j];
            REAL dx = x[3 * j] - x[3 * i], dy = x[3 * j + 1] - x[3 * i + 1], dz = x[3 * j + 2] - x[3 * i + 2];
            if (topo == 0) {
                REAL r = (REAL)sqrt(dx * dx + dy * dy + dz * dz);
                REAL w = M[j] * FN(oracle_force_softening)(r, beta);
                fx += w * dx; fy += w * dy; fz += w * dz;
            } else if (topo == 1) {
                if (fabs(dx) > 0.5 * L) dx = (REAL)(dx - L * dx / fabs(dx));
                if (fabs(dy) > 0.5 * L) dy = (REAL)(dy - L * dy / fabs(dy));
                if (fabs(dz) > 0.5 * L) dz = (REAL)(dz - L * dz / fabs(dz));
                REAL r = (REAL)sqrt(dx * dx + dy * dy + dz * dz);
                if (isp >= 2) {
                    REAL w = FN(oracle_force_softening)(r, beta), D[3];
                    FN(t3_interp)(p->table_dim0, L, (const REAL *)p->ewald_table, dx, dy, dz, D);
                    fx += M[j] * (w * dx - D[0]); fy += M[j] * (w * dy - D[1]); fz += M[j] * (w * dz - D[2]);
                } else {
                    REAL w = M[j] * FN(oracle_force_softening)(r, beta);
                    fx += w * dx; fy += w * dy; fz += w * dz;
                }
            } else if (topo == 3 && isp >= 2) {
                REAL tx = 0, ty = 0, tz = 0;                                 /* forces.cc:1258-1284 */
                for (int m = -ewald_max; m < ewald_max + 1; m++) {
                    REAL dzi = dz + ((REAL)m) * L;
                    REAL r = (REAL)sqrt(dx * dx + dy * dy + dzi * dzi);
                    if (fabs(dzi) <= ewald_cut * L) {
                        REAL w = M[j] * FN(oracle_force_softening)(r, beta);
                        tx += w * dx; ty += w * dy; tz += w * dzi;
                    }
                }
                fx += tx; fy += ty; fz += tz;
            } else {
                if (fabs(dz) > 0.5 * L) dz = (REAL)(dz - L * dz / fabs(dz));  /* forces.cc:1320-1323, :1370 */
                REAL r = (REAL)sqrt(dx * dx + dy * dy + dz * dz);
                if (topo == 2 && isp >= 2) {
                    REAL w = FN(oracle_force_softening)(r, beta), Drho, Dz;
                    REAL rho = (REAL)sqrt(dx * dx + dy * dy);
                    FN(s1r2_table)(p, rho, dz, &Drho, &Dz);
                    REAL ex = (rho > 0) ? dx / rho : 0, ey = (rho > 0) ? dy / rho : 0;
                    fx += M[j] * (w * dx - Drho * ex); fy += M[j] * (w * dy - Drho * ey); fz += M[j] * (w * dz - Dz);
                } else {
                    REAL w = M[j] * FN(oracle_force_softening)(r, beta);
                    fx += w * dx; fy += w * dy; fz += w * dz;
                }
            }
        }
        /* background terms: forces.cc:557-568 (R^3); :1286-1300, :1338-1349, :1380-1393 (S^1xR^2); none in T^3 */
        const REAL xi = x[3 * i], yi = x[3 * i + 1], zi = x[3 * i + 2];
        if (p->cosmology == 1 && topo == 0) {
            REAL B = p->comoving == 1 ? mius : DE;
            fx += B * xi; fy += B * yi; fz += B * zi;
        } else if (p->cosmology == 1 && (topo == 2 || topo == 3)) {
            if (p->comoving == 1) {
                if (topo == 3 || isp == 1) {
                    REAL rxy = (REAL)sqrt(xi * xi + yi * yi);
                    REAL c = FN(cyl_corr)(rxy, (REAL)p->Rsim, (const REAL *)p->radial_table, p->radial_size);
                    fx += mius * xi * c; fy += mius * yi * c;
                } else {
                    fx += mius * xi; fy += mius * yi;
                }
            } else {
                fx += DE * xi; fy += DE * yi;
            }
        }
        F[3 * (i - id_min)] = fx; F[3 * (i - id_min) + 1] = fy; F[3 * (i - id_min) + 2] = fz;
    }
}

/* step.cc:128-181: first half kick, drift, periodic wrap */
void FN(oracle_kick_drift)(const oracle_params *p, REAL *x, REAL *v, const REAL *F, double a, double hubble, double h)
{
    const REAL a3 = (REAL)pow(a, -3.0), L = (REAL)p->L;
    for (int i = 0; i < p->n; i++)
        for (int k = 0; k < 3; k++) {
            REAL acc = (REAL)(1.0 * F[3 * i + k] * a3 - 2.0 * (REAL)hubble * v[3 * i + k]);
            v[3 * i + k] += acc * (REAL)(h / 2.0);
            REAL xx = x[3 * i + k] + v[3 * i + k] * (REAL)h;
            if (p->topology == 1 || ((p->topology == 2 || p->topology == 3) && k == 2)) {
                if (xx < 0) xx = xx + L;
                else if (xx >= L) xx = xx - L;
            }
            x[3 * i + k] = xx;
        }
}

/* step.cc:254-269 second half kick + errmax (do_kick=1); step.cc:74-86 calculate_init_h (do_kick=0) */
double FN(oracle_kick_errmax)(const oracle_params *p, REAL *v, const REAL *F, const REAL *soft, double a, double hubble, double h, int do_kick)
{
    const REAL a3 = (REAL)pow(a, -3.0);
    REAL errmax = 0;
    for (int i = 0; i < p->n; i++) {
        REAL acc[3];
        for (int k = 0; k < 3; k++) {
            acc[k] = (REAL)(1.0 * F[3 * i + k] * a3 - 2.0 * (REAL)hubble * v[3 * i + k]);
            if (do_kick) v[3 * i + k] += acc[k] * (REAL)(h / 2.0);
        }
        /* cbrt(M[i]*const_beta) == SOFT_LENGTH[i] (utils.cc:75 computes it with the same expression) */
        REAL err = (REAL)sqrt(acc[0] * acc[0] + acc[1] * acc[1] + acc[2] * acc[2]) / soft[i];
        if (err > errmax) errmax = err;
    }
    return (double)errmax;
}

/* ---- GLASS_MAKING build of step(): G = -1.0 (global_variables.h:19-23) and the diagnostics ---- */

/* step.cc:129-148: first half kick + drift (+ wrap :151-180) with the displacement statistics; out2 = {sum of disp, max disp} */
void FN(oracle_glass_kick_drift)(const oracle_params *p, REAL *x, REAL *v, const REAL *F, double a, double hubble, double h, double *out2)
{
    const REAL a3 = (REAL)pow(a, -3.0), L = (REAL)p->L;
    REAL dmean = 0, dmax = 0;
    for (int i = 0; i < p->n; i++) {
        for (int k = 0; k < 3; k++) {
            REAL acc = (REAL)(-1.0 * F[3 * i + k] * a3 - 2.0 * (REAL)hubble * v[3 * i + k]);
            v[3 * i + k] += acc * (REAL)(h / 2.0);
            x[3 * i + k] = x[3 * i + k] + v[3 * i + k] * (REAL)h;
        }
        REAL d0 = v[3 * i] * (REAL)h, d1 = v[3 * i + 1] * (REAL)h, d2 = v[3 * i + 2] * (REAL)h;
        REAL disp = (REAL)sqrt(d0 * d0 + d1 * d1 + d2 * d2);
        dmean += disp;
        if (dmax <= disp) dmax = disp;
    }
    for (int i = 0; i < p->n; i++)
        for (int k = 0; k < 3; k++)
            if (p->topology == 1 || ((p->topology == 2 || p->topology == 3) && k == 2)) {
                REAL xx = x[3 * i + k];
                if (xx < 0) xx = xx + L;
                else if (xx >= L) xx = xx - L;
                x[3 * i + k] = xx;
            }
    out2[0] = (double)dmean;
    out2[1] = (double)dmax;
}

/* step.cc:254-303: second half kick + errmax + force / acceleration / velocity statistics;
 * out6 = {sum |F|, max |F|, sum |A|, max |A|, sum |v|, max |v|} */
double FN(oracle_glass_kick_errmax)(const oracle_params *p, REAL *v, const REAL *F, const REAL *soft, double a, double hubble, double h,
                                    double *out6)
{
    const REAL a3 = (REAL)pow(a, -3.0);
    REAL errmax = 0, Fm = 0, Fx = 0, Am = 0, Ax = 0, Vm = 0, Vx = 0;
    for (int i = 0; i < p->n; i++) {
        REAL acc[3];
        for (int k = 0; k < 3; k++) {
            acc[k] = (REAL)(-1.0 * F[3 * i + k] * a3 - 2.0 * (REAL)hubble * v[3 * i + k]);
            v[3 * i + k] += acc[k] * (REAL)(h / 2.0);
        }
        REAL A_abs = (REAL)sqrt(acc[0] * acc[0] + acc[1] * acc[1] + acc[2] * acc[2]);
        REAL err = A_abs / soft[i];
        if (err > errmax) errmax = err;
        REAL F_abs = (REAL)sqrt(F[3 * i] * F[3 * i] + F[3 * i + 1] * F[3 * i + 1] + F[3 * i + 2] * F[3 * i + 2]);
        Fm += F_abs;
        if (F_abs > Fx) Fx = F_abs;
        Am += A_abs;
        if (A_abs > Ax) Ax = A_abs;
        REAL V_abs = (REAL)sqrt(v[3 * i] * v[3 * i] + v[3 * i + 1] * v[3 * i + 1] + v[3 * i + 2] * v[3 * i + 2]);
        Vm += V_abs;
        if (V_abs > Vx) Vx = V_abs;
    }
    out6[0] = (double)Fm; out6[1] = (double)Fx; out6[2] = (double)Am; out6[3] = (double)Ax; out6[4] = (double)Vm; out6[5] = (double)Vx;
    return (double)errmax;
}

#undef FN
#undef CAT
#undef CAT_
