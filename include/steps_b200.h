/* steps_b200.h -- C ABI of libstepsb200.so, the B200-native direct-summation gravity engine
 * that replaces the O(N^2) force path (and the KDK step wrapped around it) of eltevo/StePS.
 *
 * Everything here is `extern "C"`, plain pointers and sizes.  Citations are to the reference tree
 * (StePS/src/...) and name the interface each entry point replaces.
 *
 * Conventions shared by all calls
 *   - REAL is `double` for the *_f64 entry points and `float` for *_f32 (reference:
 *     global_variables.h:26-32, -DUSE_SINGLE_PRECISION).
 *   - positions/velocities/forces are AoS, x[3*i+k] (reference: global `x`, `v`, `F`).
 *   - every function returns 0 on success, non-zero on failure; steps_b200_last_error() then
 *     returns a message.  The reference's convention (print to stderr, set ForceError=true, return;
 *     forces_cuda.cu:970-974, main.cc:1851-1856) is restored by the C++ shim (shim_steps.cc).
 *   - there is NO CPU fallback: without a CUDA device every compute call fails with an error.
 */
#ifndef STEPS_B200_H
#define STEPS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define STEPS_B200_ABI_VERSION 1

/* topology = which reference build is being replaced (Template-LinuxGCC-Makefile:23-47) */
enum {
    STEPS_TOPO_R3 = 0,          /* no flag           : forces()            forces.cc:510 / forces_cuda.cu:522 */
    STEPS_TOPO_T3 = 1,          /* -DPERIODIC        : forces_periodic()   forces.cc:776 / forces_cuda.cu:567 */
    STEPS_TOPO_S1R2_LOOKUP = 2, /* -DPERIODIC_Z      : forces_periodic_z() forces.cc:1305 / forces_cuda.cu:764 */
    STEPS_TOPO_S1R2_NOLOOKUP = 3/* -DPERIODIC_Z -DPERIODIC_Z_NOLOOKUP      forces.cc:1250 / forces_cuda.cu:652 */
};

/* The globals the reference force path reads at link time (`nm -u forces.o`),
 * packed into one POD.  Table pointers are HOST pointers; the engine uploads them once. */
typedef struct steps_b200_params {
    int32_t abi_version;        /* = STEPS_B200_ABI_VERSION */
    int32_t topology;           /* STEPS_TOPO_* */
    int32_t n;                  /* N */
    int32_t cosmology;          /* COSMOLOGY */
    int32_t comoving;           /* COMOVING_INTEGRATION */
    int32_t is_periodic;        /* IS_PERIODIC */
    int32_t s1r2_interp_order;  /* EWALD_INTERPOLATION_ORDER for the S^1xR^2 lookup build: 0 NGP, 2 CIC, 4 TSC
                                   (forces_cuda.cu:435-447); T^3 is always tricubic (main.cc:498-502) */
    int32_t table_dim0;         /* T^3: N_EWALD_FORCE_GRID;  S^1xR^2 lookup: Nrho_EWALD_FORCE_GRID */
    int32_t table_dim1;         /* S^1xR^2 lookup: Nz_EWALD_FORCE_GRID */
    int32_t radial_table_size;  /* RADIAL_FORCE_TABLE_SIZE */
    double L;                   /* L (box size / z period) */
    double Rsim;                /* Rsim */
    double mass_in_unit_sphere; /* mass_in_unit_sphere (main.cc:1269,1288,1313) */
    double H0;                  /* H0 (internal units) */
    double Omega_lambda;        /* Omega_lambda; DE = (REAL)H0*H0*Omega_lambda (forces.cc:513) */
    const void *ewald_table;    /* T3_EWALD_FORCE_TABLE [Ng^3][3] or S1R2_EWALD_FORCE_TABLE [Nrho][Nz][2], REAL */
    const void *radial_table;   /* RADIAL_FORCE_TABLE [radial_table_size], REAL */
} steps_b200_params;

const char *steps_b200_last_error(void);
int steps_b200_abi_version(void);
/* number of visible CUDA devices (0 when there is no GPU; never an error) */
int steps_b200_device_count(void);

/* ------------------------------------------------------------------------------------------
 * (1) Stateless parity path.  Replaces
 *        void forces(REAL*x, REAL*F, int ID_min, int ID_max)            forces_cuda.cu:74-78
 *        void forces_periodic(REAL*x, REAL*F, int ID_min, int ID_max)   forces_cuda.cu:169-173
 *        void forces_periodic_z(REAL*x, REAL*F, int ID_min, int ID_max) forces_cuda.cu:457-461
 *     x: host AoS [3N]; M, soft: host [N] (globals M, SOFT_LENGTH); F: host [3*(id_max-id_min+1)],
 *     fully overwritten, index relative to id_min.  Device buffers are cached between calls
 *     (the reference re-mallocs and re-uploads per call, forces_cuda.cu:976-1050).
 * ------------------------------------------------------------------------------------------ */
int steps_b200_forces_f64(const steps_b200_params *p, const double *x, const double *M, const double *soft,
                          double *F, int id_min, int id_max);
int steps_b200_forces_f32(const steps_b200_params *p, const float *x, const float *M, const float *soft,
                          float *F, int id_min, int id_max);

/* The same call with the i-range split over n_gpu devices driven from the calling thread, starting at device
 * $STEPS_B200_DEVICE (default 0): replaces forces_cuda(x, F, n_GPU, ID_min, ID_max) (forces_cuda.cu:878-1117),
 * whose n_GPU OpenMP threads each own one device.  Fewer visible devices than asked: warn and clamp
 * (forces_cuda.cu:897-907). */
int steps_b200_forces_multi_f64(const steps_b200_params *p, const double *x, const double *M, const double *soft,
                                double *F, int id_min, int id_max, int n_gpu);
int steps_b200_forces_multi_f32(const steps_b200_params *p, const float *x, const float *M, const float *soft,
                                float *F, int id_min, int id_max, int n_gpu);
/* frees the device buffers the stateless calls cache between invocations */
void steps_b200_release_cached(void);

/* void calculate_softening_length(REAL*SOFT_LENGTH, REAL*M, int N)  utils.cc:59-82.
 * Host O(N) helper kept in the library so a caller does not need the reference's utils.o.
 * Outputs M_min and rho_part as the reference's globals of the same name. */
int steps_b200_softening_f64(const double *M, int n, double particle_radii, double *soft_out, double *M_min_out,
                             double *rho_part_out);
int steps_b200_softening_f32(const float *M, int n, float particle_radii, float *soft_out, float *M_min_out,
                             float *rho_part_out);

/* ------------------------------------------------------------------------------------------
 * (1b) Table producer: the T^3 Ewald force-correction lookup table on the GPU.
 *     Replaces calculate_t3_ewald_lookup_table() + ewald_force_correction() + ewald_space()
 *     (ewald_space.cc:288-383, :198-286, :118-196) as called at main.cc:467-494.  table_host: [ngrid^3][3] doubles,
 *     layout of ewald_space.cc:46-49 -- exactly what steps_b200_params.ewald_table expects.
 *     t3_ewald_defaults restates main.cc:425-446 (IS_PERIODIC 2/3/4 -> 63/127/255 grid, cuts 2.6/3.6/4.6 L and
 *     8/10/12, alpha = 2/L).  ewald_space_count(R) = number of lattice vectors ewald_space(R, ...) enumerates (host only).
 * ------------------------------------------------------------------------------------------ */
int steps_b200_t3_ewald_defaults(int is_periodic, double L, int *ngrid, double *alpha, double *rel_cut, double *rec_cut);
int steps_b200_ewald_space_count(double R);
int steps_b200_t3_ewald_table_f64(int ngrid, double L, double alpha, double rel_cut, double rec_cut, double *table_host, int device);
/* The S^1 x R^2 (rho, z) Ewald correction table of the lookup build: replaces calculate_S1R2ewald_correction_table() +
 * S1R2ewald_force_pair() (ewald_space.cc:754-798, :618-752; Ewald variant) as called at main.cc:562-705.  table_host:
 * [nrho][nz][2] doubles = (D_rho, D_z), layout of ewald_space.cc:611-614.  s1r2_ewald_defaults restates main.cc:575-605
 * (IS_PERIODIC 2/3/>=4 -> Nz 128/256/512, nmax 4/5/IS_PERIODIC+2, mmax 10/12/IS_PERIODIC+9, alpha 0.787875/0.71805/0.6642 / L,
 * rho_max = 2.25 Rsim, Nrho = floor(Nz * 2.25 Rsim / L)). */
int steps_b200_s1r2_ewald_defaults(int is_periodic, double L, double Rsim, int *nrho, int *nz, double *rho_max, double *alpha, int *nmax,
                                   int *mmax);
int steps_b200_s1r2_ewald_table_f64(int nrho, int nz, double rho_max, double Lz, double alpha, int nmax, int mmax, double *table_host,
                                    int device);
/* RADIAL_FORCE_TABLE of the S^1 x R^2 builds: replaces get_cylindrical_force_table(FORCE_TABLE, R, Lz, TABLE_SIZE,
 * RADIAL_FORCE_ACCURACY) (utils.cc:162-228), called with R = Rsim and Lz = L/2 (IS_PERIODIC == 1) or Lz = L * ewald_cut,
 * ewald_cut = IS_PERIODIC + 1 - 0.4 (NOLOOKUP build, IS_PERIODIC >= 2) at main.cc:1263-1310. */
int steps_b200_radial_force_table_f64(double R, double Lz, int table_size, int accuracy, double *table_host, int device);

/* ------------------------------------------------------------------------------------------
 * (2) Device-resident engine (north_star item 3): x, v, F, M, s stay in HBM across KDK steps.
 *     One engine per process per GPU.  Replaces step() (step.cc:100-312), calculate_init_h()
 *     (step.cc:35-98) and the per-step host<->device traffic of forces_cuda.cu:976-1105.
 * ------------------------------------------------------------------------------------------ */
typedef struct steps_b200_engine steps_b200_engine;

/* real_bytes: 8 or 4.  device: CUDA ordinal.  [i_lo, i_hi): the i-particles this engine owns
 * (pass 0, n for a single GPU); see steps_b200_partition(). */
int steps_b200_engine_create(steps_b200_engine **out, const steps_b200_params *p, int real_bytes, int device);
void steps_b200_engine_destroy(steps_b200_engine *e);

/* contiguous i-partition, remainder spread one-each over the first ranks (replaces main.cc:1581-1607
 * and forces_cuda.cu:942-951, which give the whole remainder to rank/GPU 0). */
void steps_b200_partition(int n, int nranks, int rank, int *i_lo, int *i_hi);

/* Action-reaction evaluation of the R^3 path (FP64: pair_r3_sym.cuh, FP32: pair_r3_sym_f32.cuh): every unordered pair is evaluated once and
 * applied to both particles (F_i += m_j w d, F_j -= m_i w d), the same force law as forces() (forces.cc:510-577).
 * It serves force calls for exactly the engine's own rows; multi-GPU engines then partition rows on i-block
 * boundaries and exchange the j-side sums with one all-reduce per evaluation.  On by default; environment variable
 * STEPS_B200_SYM=0 turns it off.  set_symmetric() must precede comm_init() and is collective in spirit: every rank of a job
 * must make the same choice.  engine_range() returns the rows [i_lo, i_hi) the engine owns under the partition
 * in force.  sym_rules() is the host-only rule builder (no GPU needed): for rank `rank` of `nranks`, i-block size
 * ib_size (a multiple of the 128-record j-tile), it writes 16 ints per local i-block
 * {diag_lo, diag_hi, n_sym, sym_lo[5], sym_hi[5], pad[3]} (tile indices) and returns the number of blocks, or -1
 * when the geometry does not admit the symmetric path. */
int steps_b200_engine_set_symmetric(steps_b200_engine *e, int on);
int steps_b200_engine_is_symmetric(steps_b200_engine *e);
int steps_b200_engine_range(steps_b200_engine *e, int *i_lo, int *i_hi);
int steps_b200_sym_rules(int n, int nranks, int rank, int ib_size, int *i_lo, int *i_hi, int *rules_out, int max_blocks);
/* Host-only (no GPU needed): rules -> launch plan -> schedule of rank `rank` of `nranks` for an n-particle job of the given topology
 * (STEPS_TOPO_*) and precision; row_budget_bytes = memory for the j-side rows (0: 16 GB).  plan_out[8] = {ib_size, n_ib, sb, n_sb,
 * n_chunks, tiles_per_chunk, n_tiles, n_passes}; order_out = (superblock within its pass, chunk) int pairs of all passes one after
 * the other, pass_off_out[n_passes + 1] their offsets; cmask_out = one bit per (local i-block, chunk) that gets a partial sum written,
 * *words_out 64-bit words per block.  Returns the number of CTAs of the evaluation, -1 when the action-reaction path does not apply,
 * -2 when an output array is too small. */
int steps_b200_sym_schedule_host(int n, int nranks, int rank, int topology, int real_bytes, long long row_budget_bytes, int *plan_out,
                                 int *order_out, int max_ctas, int *pass_off_out, int max_passes, unsigned long long *cmask_out,
                                 long long max_mask_words, int *words_out);
/* Host-only, pure: the number of j-chunks the action-reaction launch aims for, given the j-tiles of the problem, the i-blocks of the
 * call, the j-side rows one pass holds and the resident-CTA slots of the GPU (see engine.cu; 56 unless a pass would have < 24 waves). */
int steps_b200_sym_chunk_target(int n_tiles, int n_ib, long long rows_per_pass, int slots);
/* Test hooks (tests/test_gpu_sym.py): one GPU plays every rank of a P-GPU action-reaction job in turn.
 * debug_set_rank gives the engine the rows and rules of `rank` of `nranks` without a communicator (its evaluation
 * then skips the all-reduce); debug_fsym reads the engine's j-side sums ([3][n_pad] REALs of the engine's precision) and/or replaces them by
 * the caller's total and redoes the final reduction. */
int steps_b200_engine_debug_set_rank(steps_b200_engine *e, int rank, int nranks, int symmetric);
int steps_b200_engine_debug_fsym(steps_b200_engine *e, void *fsym_out, const void *fsym_in, int *n_pad_out);

/* NCCL bootstrap for one-process-per-GPU runs: rank 0 calls unique_id() and ships the 128 bytes to
 * the others by any means (MPI_Bcast in StePS, torch.distributed in bench.py); then every rank
 * calls comm_init().  Without comm_init the engine is single-GPU and owns [0, n). */
int steps_b200_nccl_unique_id(void *id128);
int steps_b200_engine_comm_init(steps_b200_engine *e, const void *id128, int rank, int nranks);

/* upload the full state (host AoS x[3N], v[3N]; M[N], soft[N]); REAL = engine precision */
int steps_b200_engine_upload(steps_b200_engine *e, const void *x, const void *v, const void *M, const void *soft);
/* only positions (stateless-style use of a resident engine) */
int steps_b200_engine_upload_x(steps_b200_engine *e, const void *x);

/* resident forces [3N] from host (a caller that evaluated the initial forces elsewhere, main.cc:1581-1607) */
int steps_b200_engine_upload_forces(steps_b200_engine *e, const void *F);

/* forces for i in [id_min, id_max] from the resident positions into the resident F (asynchronous) */
int steps_b200_engine_forces(steps_b200_engine *e, int id_min, int id_max);
/* copy the resident F slice [id_min, id_max] to host (synchronises) */
int steps_b200_engine_download_forces(steps_b200_engine *e, void *F, int id_min, int id_max);
/* full state to host: any of x, v, F may be NULL.  x, v: [3N]; F: [3N] (rows outside the owned
 * range are only meaningful after a gather -- multi-GPU callers read their own range). */
int steps_b200_engine_download(steps_b200_engine *e, void *x, void *v, void *F);

/* calculate_init_h() step.cc:35-98: wraps positions into the box, returns
 * errmax = max_i |G F a^-3 - 2 H v| / s_i  over ALL particles (max-all-reduced over ranks). */
int steps_b200_engine_init_errmax(steps_b200_engine *e, double a, double hubble, double *errmax_out);

/* one KDK step, step.cc:100-312:
 *   kick(h/2; a_old, H_old) -> drift(h) -> periodic wrap -> [position all-gather over NCCL]
 *   -> forces -> kick(h/2; a_new, H_new) + errmax reduction (device, then max over ranks).
 * The caller advances the scale factor itself (friedmann_solver_step, step.cc:237-240) -- use
 * steps_b200_friedmann_step()/steps_b200_hubble() or the reference's own. */
int steps_b200_engine_kdk_step(steps_b200_engine *e, double h, double a_old, double hubble_old, double a_new,
                               double hubble_new, double *errmax_out);

/* GLASS_MAKING builds of the reference: G = -1 (global_variables.h:19-23) and the diagnostics step() accumulates in
 * that mode (step.cc:107-121, :143-148, :270-303).  set_glass_making(e, 1) switches kdk_step / init_errmax of the engine to that
 * build's arithmetic; after a step glass_stats() returns, in the argument order of Log_write_glass (inputoutput.cc:974),
 * {F_mean, Fmax, A_mean, A_max, dmean, dmax, V_mean, V_max} over ALL N particles (all-reduced in a multi-GPU job), internal units.
 * The velocities are zeroed by the caller before the upload, as main.cc:1240-1254 does. */
int steps_b200_engine_set_glass_making(steps_b200_engine *e, int on);
int steps_b200_engine_glass_stats(steps_b200_engine *e, double *out8);

/* device timers (CUDA events on the engine's stream): milliseconds of the last force evaluation
 * (pack + pair kernels + reduce) and of the last complete kdk_step. */
int steps_b200_engine_timings(steps_b200_engine *e, double *force_ms, double *step_ms);
/* milliseconds of the last pair-kernel launch alone (CUDA events around it; the roofline numerator) */
int steps_b200_engine_pair_kernel_ms(steps_b200_engine *e, double *ms_out);
/* caller-placed CUDA events on the engine's stream (slot 0..7) and the time between two of them;
 * elapsed_ms synchronises on slot_b.  bench.py brackets its timed region with these. */
int steps_b200_engine_mark(steps_b200_engine *e, int slot);
int steps_b200_engine_elapsed_ms(steps_b200_engine *e, int slot_a, int slot_b, double *ms_out);
/* number of kernel launches issued by the engine so far (bench.py's gpu_launches) */
long long steps_b200_engine_launch_count(steps_b200_engine *e);
int steps_b200_engine_sync(steps_b200_engine *e);
/* launch-shape report for DESIGN/bench: out[0]=i per CTA, out[1]=j chunks, out[2]=CTAs, out[3]=j tile */
int steps_b200_engine_launch_shape(steps_b200_engine *e, int id_min, int id_max, int *out4);

/* ------------------------------------------------------------------------------------------
 * (2b) In-process multi-GPU group: n_gpu resident engines in ONE process (devices first_device ...),
 *     i-partitioned, NCCL between them, one library-owned host thread per device inside every call --
 *     the caller stays single-threaded.  This is what replaces the reference's `StePS_CUDA <param> <nGPU>`
 *     model (one OpenMP thread per GPU, forces_cuda.cu:933-941; step.cc:122-125) for the resident path.
 *     One-process-per-GPU callers (MPI ranks, torchrun) use steps_b200_engine_comm_init instead.
 * ------------------------------------------------------------------------------------------ */
typedef struct steps_b200_group steps_b200_group;
int steps_b200_group_create(steps_b200_group **out, const steps_b200_params *p, int real_bytes, int n_gpu, int first_device);
void steps_b200_group_destroy(steps_b200_group *g);
int steps_b200_group_size(steps_b200_group *g);
steps_b200_engine *steps_b200_group_engine(steps_b200_group *g, int d);
/* full host state to every engine; F may be NULL (then call steps_b200_group_forces) */
int steps_b200_group_upload(steps_b200_group *g, const void *x, const void *v, const void *M, const void *soft, const void *F);
/* every engine evaluates the forces of its own i-range (synchronises) */
int steps_b200_group_forces(steps_b200_group *g);
int steps_b200_group_init_errmax(steps_b200_group *g, double a, double hubble, double *errmax_out);
int steps_b200_group_kdk_step(steps_b200_group *g, double h, double a_old, double hubble_old, double a_new, double hubble_new,
                              double *errmax_out);
/* gather the state to host arrays [3N]: x from the replica, v and F from each owner; any may be NULL */
int steps_b200_group_set_glass_making(steps_b200_group *g, int on);
int steps_b200_group_glass_stats(steps_b200_group *g, double *out8);
/* Spatial order of the resident particles (opt-in; tools/t3_wavefront_model.py: the table gathers of the T^3 / S^1xR^2-lookup pair kernels
 * touch 6x fewer cache lines when the 32 particles of a warp are neighbours in space, which a caller's array order does not guarantee).
 * spatial_order(): host-only; perm_out[k] = index of the k-th particle when sorted by cell (bounding box of x, ngrid cells per axis, z
 * fastest, stable).  permute(): host-only gather (dst[k] = src[perm[k]]) or scatter (dst[perm[k]] = src[k]) of n rows.
 * group_set_spatial_order(g, ngrid > 0) before group_upload(): the group keeps its resident copy in that order -- uploads gather, downloads and
 * snapshots scatter back, so the caller keeps seeing its own order; forces change only by the order of summation.  Engines obtained by
 * group_engine() see the resident order; group_permutation() returns it. */
int steps_b200_spatial_order(const void *x, int n, int real_bytes, int ngrid, int *perm_out);
int steps_b200_permute(const void *src, void *dst, const int *perm, int n, int width, int elem_bytes, int scatter);
int steps_b200_group_set_spatial_order(steps_b200_group *g, int ngrid);
int steps_b200_group_permutation(steps_b200_group *g, int *perm_out);
/* T^3 with the Ewald table decides by itself when the caller has not (no set_spatial_order call, not even with 0): group_upload() measures
 * order_incoherence() -- median nearest-image distance between array neighbours (sample of 4096) in units of the mean interparticle
 * spacing, ~1 for lattice-ordered input, ~N^(1/3)/2 for a shuffled one -- and keeps the resident copy sorted by table cell when it exceeds 4.
 * STEPS_B200_SPATIAL_ORDER=0 switches the automatic decision off. */
double steps_b200_order_incoherence(const void *x, int n, int real_bytes, double L);
/* ASCII snapshots in the reference's format (write_ascii_snapshot, inputoutput.cc:826-909; SURVEY.md 8f.2 -- the HDF5 formats need
 * libhdf5): per particle "x y z vx vy vz M", each "%.16f\t", x and M times H0_dimless (in REAL precision), v times sqrt(a)*UNIT_V (in
 * double), zero velocities in a GLASS_MAKING build.  snapshot_ascii_host formats host arrays with a pool of workers (nthreads <= 0: all
 * cores, at most 16).  group_snapshot_ascii_async copies the resident state out asynchronously (in stream order, pinned staging) and returns;
 * a background thread formats and writes while the caller goes on stepping; group_snapshot_wait joins it and reports its status (the next
 * snapshot call and group_destroy wait too). */
int steps_b200_snapshot_ascii_host(const char *path, const void *x, const void *v, const void *M, int n, int real_bytes,
                                   double h0_dimless, double a, int zero_velocities, int nthreads);
int steps_b200_group_snapshot_ascii_async(steps_b200_group *g, const char *path, double h0_dimless, double a, int zero_velocities);
int steps_b200_group_snapshot_wait(steps_b200_group *g);
/* Redshift-cone output (write_redshift_cone, inputoutput.cc:314-405; called at main.cc:1785) with the state resident on the
 * devices.  cone_select(): every resident particle with r_min <= |x| (all != 0: every particle) that no earlier call selected is flagged
 * (the reference's IN_CONE) and its row is brought to the host; *count_out of them, kept in ascending order of the caller's particle index.
 * cone_rows(): rows_out[count][8] REAL = x y z vx vy vz M pad, index_out[count]; either may be NULL.  cone_reset(): clear the flags.
 * redshift_cone_ascii_host(): host-only formatter, appends to `path` exactly the lines the reference writes for these particles -- all == 0:
 * "x y z vx vy vz M D z index" with out_list[z_index]; all != 0 (end of the run): the redshift of the radial bin the distance falls into
 * (limits = r_bin_limits, descending) and D * H0_dimless.  cone_write_ascii(): the same for the rows of the last selection. */
int steps_b200_group_cone_select(steps_b200_group *g, double r_min, int all, int *count_out);
int steps_b200_group_cone_rows(steps_b200_group *g, void *rows_out, int *index_out);
int steps_b200_group_cone_reset(steps_b200_group *g);
int steps_b200_redshift_cone_ascii_host(const char *path, const void *rows, const int *index, int count, int real_bytes, double h0_dimless,
                                        int all, const double *limits, int n_limits, const double *out_list, int z_index);
int steps_b200_group_cone_write_ascii(steps_b200_group *g, const char *path, double h0_dimless, int all, const double *limits, int n_limits,
                                      const double *out_list, int z_index);
int steps_b200_group_download(steps_b200_group *g, void *x, void *v, void *F);

/* ------------------------------------------------------------------------------------------
 * (3) Host scalars of the integrator, restated so a C/C++ driver needs nothing else:
 *     friedmann_solver_step / CALCULATE_Hubble_param (friedmann_solver.cc:100-164, LCDM),
 *     the timestep rule of main.cc:1834-1846.
 * ------------------------------------------------------------------------------------------ */
typedef struct steps_b200_cosmo {
    double H0, Omega_m, Omega_r, Omega_lambda, Omega_k;
} steps_b200_cosmo;
double steps_b200_friedmann_step(const steps_b200_cosmo *c, double a0, double h);
double steps_b200_hubble(const steps_b200_cosmo *c, double a);
double steps_b200_next_timestep(double acc_param, double errmax, double h_min, double h_max);
/* the same with the output-time clamp of main.cc:1843-1846: with outputs scheduled in time (OUTPUT_TIME_VARIABLE == 0) a step that would pass
 * t_next ends 1e-9 h_min after it */
double steps_b200_next_timestep_to_output(double acc_param, double errmax, double h_min, double h_max, double T, double t_next,
                                          int output_time_variable);

/* FP64 / FP32 FMA-pipe microbenchmark on `device`: returns measured TFLOP/s (2 flop per FMA) --
 * the roofline denominator for this path. */
int steps_b200_fma_peak(int device, int real_bytes, double *tflops_out, double *sm_clock_mhz_out);
/* same kernel launched back to back for `seconds`: the sustained figure under the power cap, the
 * denominator for a pair kernel that runs for seconds */
int steps_b200_fma_peak_sustained(int device, int real_bytes, double seconds, double *tflops_out);

#ifdef __cplusplus
}
#endif
#endif /* STEPS_B200_H */
