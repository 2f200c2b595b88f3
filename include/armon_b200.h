/*
 * armon_b200.h -- C ABI of libarmon_b200.so, the B200-native (sm_100a) backend for Armon.jl's
 * axis-split Lagrange+remap time step.
 *
 * This is the drop-in boundary: every entry point is what a Julia `ccall` from the reference's backend
 * extension (the ArmonB200.jl stub shown in INTEGRATION.md, modelled on ext/ArmonCUDA.jl and
 * ext/ArmonKokkos.jl) binds.  Plain pointers and sizes only; no torch / C++ types.  Each declaration
 * cites the reference interface it replaces (paths relative to the reference root).
 *
 * Conventions
 *  - Float type is double (`armon_flt_size() == 8`), index type int64 (`armon_idx_size() == 8`);
 *    cf. the ABI self-description of ext/ArmonKokkos.jl:122-140.
 *  - Every function returns an int status: 0 = ok, non-zero = ARMON_ERR_*; the message of the last
 *    failure of the calling thread is `armon_last_error()`.  The Julia side turns it into
 *    `solver_error(:cpp, msg)` (src/utils.jl:102-117, precedent ext/ArmonKokkos.jl:72-76).
 *  - All device pointers are arrays of (nx+2g)*(ny+2g) doubles laid out like BlockData
 *    (src/blocking/blocks.jl:18-44; "contiguous rows": 0-based offset of real cell (ix,iy), 1-based,
 *    = (iy+g-1)*(nx+2g) + (ix+g-1), src/blocking/blocking.jl:129-131).
 *  - Calls are asynchronous on the context's stream and ordered; only `armon_ctx_sync`, the `*_d2h`
 *    copies, `armon_dtCFL`, `armon_conservation_vars` and `armon_solver_state` block.
 *  - One caller thread per context.  The library never frees or retains caller memory beyond the device
 *    arrays bound with `armon_solver_bind`.
 */
#ifndef ARMON_B200_H
#define ARMON_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ARMON_B200_ABI_VERSION 2

enum {
    ARMON_OK = 0,
    ARMON_ERR_INVALID = 1,   /* bad argument / unsupported configuration  -> SolverException(:config) */
    ARMON_ERR_CUDA = 2,      /* CUDA runtime failure                      -> SolverException(:cpp)    */
    ARMON_ERR_NCCL = 3,      /* NCCL failure                              -> SolverException(:cpp)    */
    ARMON_ERR_TIME = 4,      /* invalid time step, src/solver_state.jl:123-124 -> SolverException(:time) */
    ARMON_ERR_NO_DEVICE = 5, /* no CUDA device: the backend has no CPU fallback */
    ARMON_ERR_RANGE = 6      /* math_mode strict: the work list of the IEEE fix-up overflowed (more than 4 Mi column
                                chunks of one sweep held division/sqrt operands outside the range in which the
                                branch-free correctly rounded routines are proven: divisors in [2^-120, 2^120],
                                dividends 0 or in [2^-900, 2^900]); rerun with math_mode ieee -> SolverException(:cpp) */
};

/* src/utils.jl:15-78 (Axis.X=1.. in Julia; 0-based here) */
enum { ARMON_AXIS_X = 0, ARMON_AXIS_Y = 1 };
enum { ARMON_SIDE_LEFT = 0, ARMON_SIDE_RIGHT = 1, ARMON_SIDE_BOTTOM = 2, ARMON_SIDE_TOP = 3 };
/* test cases: same codes as ext/ArmonKokkos.jl:61-69, +5 for DebugIndexes (src/tests.jl:217-233) */
enum { ARMON_TEST_SOD = 0, ARMON_TEST_SOD_Y = 1, ARMON_TEST_SOD_CIRC = 2, ARMON_TEST_BIZARRIUM = 3,
       ARMON_TEST_SEDOV = 4, ARMON_TEST_DEBUG_INDEXES = 5 };
enum { ARMON_RIEMANN_GODUNOV = 0, ARMON_RIEMANN_GAD = 1 };                      /* src/riemann_schemes.jl:2-3 */
enum { ARMON_LIMITER_NONE = 0, ARMON_LIMITER_MINMOD = 1, ARMON_LIMITER_SUPERBEE = 2 }; /* ext/ArmonKokkos.jl:50-58 */
enum { ARMON_PROJ_EULER = 0, ARMON_PROJ_EULER_2ND = 1 };                        /* src/projection_schemes.jl:2-3 */
enum { ARMON_SPLIT_SEQUENTIAL = 0, ARMON_SPLIT_GODUNOV = 1, ARMON_SPLIT_STRANG = 2,
       ARMON_SPLIT_X_ONLY = 3, ARMON_SPLIT_Y_ONLY = 4 };                        /* src/axis_splitting.jl:2-5 */
enum { ARMON_EOS_PERFECT_GAS = 0, ARMON_EOS_BIZARRIUM = 1 };                    /* src/kernels.jl:151-161 */
/* arithmetic mode of the fused sweep kernels */
enum { ARMON_MATH_STRICT = 0,   /* IEEE operation order of the reference source, no FMA contraction, correctly rounded
                                   branch-free division/sqrt: bit-exact vs the oracle; column chunks whose operands leave
                                   the proven range (divisors [2^-120, 2^120], dividends 0 or [2^-900, 2^900]) are
                                   recomputed with the full IEEE division by a fix-up kernel */
       ARMON_MATH_FAST = 1,     /* FMA contraction + reciprocal-based division (<= 2 ulp), like the reference's own
                                   @fastmath kernels (src/generic_kernel.jl:2-4) */
       ARMON_MATH_IEEE = 2 };   /* as STRICT but with nvcc's full IEEE division/sqrt (slow paths for every operand) */

/* marching kernel of the fused sweep.  AUTO: TMA for math_mode fast, ASYNC for strict, SINGLE for ieee (and for
 * strict when the input pitch is odd: the tensor maps of its TMA staging need 16-byte aligned rows; the fast kernels
 * fall back to 8-byte staging copies by themselves). */
enum { ARMON_KERNEL_AUTO = 0,
       ARMON_KERNEL_SINGLE = 1,     /* register prefetch, no shared-memory staging (any math mode) */
       ARMON_KERNEL_ASYNC = 4,      /* strict: the strict arithmetic on the fast kernel's four-chain schedule (TMA staging)
                                       + IEEE fix-up kernel; the name is the ABI's from round 1 (cp.async staging) */
       ARMON_KERNEL_ASYNC2 = 5,     /* fast: explicit-arithmetic software-pipelined kernel, cp.async (16-byte) staging */
       ARMON_KERNEL_TMA = 6         /* fast: the same kernel staged by the TMA (cp.async.bulk.tensor.2d + mbarrier) */ };

typedef struct armon_ctx armon_ctx;
typedef struct armon_solver armon_solver;
typedef struct armon_group armon_group;

/* Block geometry: StaticBSize/DynamicBSize (src/blocking/blocking.jl:19-58) with one block per GPU. */
typedef struct { int64_t nx, ny, g; } armon_dims;

/* Inclusive rectangle in 1-based real-cell coordinates; what block_domain_range builds from a
 * StepsRanges entry (src/blocking/blocking.jl:71-85, src/domain_ranges.jl:39-42,96-105). */
typedef struct { int64_t ix0, ix1, iy0, iy1; } armon_domain;

/* Two-state test case, evaluated on the host (src/tests.jl:59-121,150-211). */
typedef struct {
    int32_t test;                /* ARMON_TEST_* */
    int32_t eos;                 /* ARMON_EOS_* */
    double  high_rho, low_rho, high_E, low_E, high_u, low_u, high_v, low_v;   /* InitTestParamsTwoState */
    double  sedov_r;             /* Sedov{T}.r */
    double  gamma;               /* specific_heat_ratio */
    double  bc_u[4], bc_v[4];    /* boundary_condition(test, side) -> (u_factor, v_factor), indexed by ARMON_SIDE_* */
} armon_test_case;

/* Everything the fused solver needs: the subset of ArmonParameters (src/parameters.jl:267-389) that the
 * hot path reads. */
typedef struct {
    armon_dims dims;                 /* local sub-domain: params.N, params.nghost */
    int64_t global_nx, global_ny;    /* params.global_grid */
    int64_t origin_ix, origin_iy;    /* params.N_origin (1-based) */
    double  domain_size[2];          /* params.domain_size */
    double  origin[2];               /* params.origin */
    int32_t riemann, limiter, projection, splitting;
    double  cfl, maxtime;
    int64_t maxcycle;
    int32_t cst_dt;
    double  Dt;
    int32_t neighbours[4];           /* rank of the neighbour per side, -1 = global edge (MPI.PROC_NULL), params.neighbours */
    int32_t math_mode;               /* ARMON_MATH_* */
    int32_t march_segment;           /* cells per marching segment along the swept axis (0 = auto) */
    int32_t kernel_variant;          /* ARMON_KERNEL_*: 0 = auto */
    int32_t cuda_graph;              /* 0 = auto (grids of <= 512x512 cells on one rank), 1 = replay captured cycle
                                        pairs with a CUDA graph, 2 = never */
    armon_test_case tc;
} armon_solver_desc;

/* GlobalTimeStep scalars (src/solver_state.jl:26-47) as kept on the device. */
typedef struct {
    int64_t cycle;
    double  time;
    double  current_dt;
    double  next_cycle_dt;
    int32_t error;                   /* ARMON_ERR_TIME when an invalid time step was met */
    int32_t done;                    /* 1 once time >= maxtime or cycle >= maxcycle (src/solver.jl:333) */
    int64_t error_cycle;             /* cycle the reference's message would name (src/solver_state.jl:123-124) */
} armon_time_state;

/* One line of the reference's per-cycle log (src/solver.jl:359-371), produced on the device. */
typedef struct {
    int64_t cycle;                   /* global_dt.cycle after next_cycle! */
    double  time, dt;                /* global_dt.time, global_dt.current_dt */
    double  mass, energy;            /* conservation_vars of this rank's blocks (src/reductions.jl:202-298) */
} armon_cycle_diag;

/* ---------------------------------------------------------------------------------------------------
 * ABI self-description and errors (ext/ArmonKokkos.jl:72-76,119-140)
 * ------------------------------------------------------------------------------------------------- */
int         armon_b200_abi_version(void);
int         armon_flt_size(void);
int         armon_idx_size(void);
const char *armon_last_error(void);

/* ---------------------------------------------------------------------------------------------------
 * Device / context: create_device(::Val{:B200}) + init_backend (src/parameters.jl:738-778),
 * Base.wait(params) (src/parameters.jl:1031-1038), device_memory_info (src/parameters.jl:916-926),
 * print_device_info (src/parameters.jl:787-802)
 * ------------------------------------------------------------------------------------------------- */
int armon_device_count(int *count);
int armon_ctx_create(int device, armon_ctx **ctx);
int armon_ctx_destroy(armon_ctx *ctx);
int armon_ctx_sync(armon_ctx *ctx);
int armon_device_memory_info(armon_ctx *ctx, uint64_t *free_bytes, uint64_t *total_bytes);
int armon_device_name(armon_ctx *ctx, char *buf, int len);
/* number of CUDA kernels launched by this library on this context since creation (bench.py's gpu_launches) */
int armon_ctx_launch_count(armon_ctx *ctx, uint64_t *count);

/* ---------------------------------------------------------------------------------------------------
 * Device arrays backing the Julia `B200Array{T,1}`: device_array_type (src/parameters.jl:950-951),
 * `V(undef, n)` (src/blocking/blocks.jl:36-43), copyto! host<->device (src/blocking/blocks.jl:121-143)
 * ------------------------------------------------------------------------------------------------- */
int armon_alloc(armon_ctx *ctx, uint64_t n_elems, double **dptr);
int armon_free(armon_ctx *ctx, double *dptr);
int armon_copy_h2d(armon_ctx *ctx, double *dst_dev, const double *src_host, uint64_t n_elems);
int armon_copy_d2h(armon_ctx *ctx, double *dst_host, const double *src_dev, uint64_t n_elems);
int armon_copy_d2d(armon_ctx *ctx, double *dst_dev, const double *src_dev, uint64_t n_elems);
/* `ArmonParameters{Float32}`: the caller's arrays are Float32, the device arrays (and all arithmetic) stay Float64; the
 * conversion runs on the device (h2d: exact widening; d2h: round to nearest), through a staging buffer of the context.
 * A Float32 run therefore carries no Float32 rounding noise of its own: it agrees with the reference's Float64 result
 * to Float32 precision (tests/test_gpu_parity.py::test_float32_boundary). */
int armon_copy_h2d_f32(armon_ctx *ctx, double *dst_dev, const float *src_host, uint64_t n_elems);
int armon_copy_d2h_f32(armon_ctx *ctx, float *dst_host, const double *src_dev, uint64_t n_elems);
int armon_fill(armon_ctx *ctx, double *dst_dev, double value, uint64_t n_elems);
/* fills every ghost cell of one array (test/convergence.jl:67-102 poisons ghosts with 1e100) */
int armon_fill_ghosts(armon_ctx *ctx, armon_dims d, double *arr, double value);

/* ---------------------------------------------------------------------------------------------------
 * Kernel seam: one entry point per `@generic_kernel` of the hot path.  These are the per-step overloads
 * (debug / `compare=true` path, src/io.jl:185-227); same arguments as the generated Julia functions.
 * ------------------------------------------------------------------------------------------------- */
/* perfect_gas_EOS!  src/kernels.jl:4-13, wrapper :151-155 */
int armon_perfect_gas_EOS(armon_ctx *ctx, armon_dims d, armon_domain dom, double gamma,
                          const double *rho, const double *E, const double *u, const double *v,
                          double *p, double *c, double *g);
/* bizarrium_EOS!  src/kernels.jl:16-55, wrapper :158-161 */
int armon_bizarrium_EOS(armon_ctx *ctx, armon_dims d, armon_domain dom,
                        const double *rho, const double *u, const double *v, const double *E,
                        double *p, double *c, double *g);
/* boundary_conditions!  src/halo_exchange.jl:2-36 (domain = border_domain(side)) */
int armon_boundary_conditions(armon_ctx *ctx, armon_dims d, int side, double u_factor, double v_factor,
                              double *rho, double *u, double *v, double *p, double *c, double *g, double *E);
/* acoustic!  src/riemann_schemes.jl:33-52 */
int armon_acoustic(armon_ctx *ctx, armon_dims d, armon_domain dom, int axis,
                   double *us, double *ps, const double *rho, const double *ua, const double *p, const double *c);
/* acoustic_GAD!  src/riemann_schemes.jl:55-113 */
int armon_acoustic_GAD(armon_ctx *ctx, armon_dims d, armon_domain dom, int axis, double dt, double dx, int limiter,
                       double *us, double *ps, const double *rho, const double *ua, const double *p, const double *c);
/* cell_update!  src/kernels.jl:58-68, wrapper :217-223 */
int armon_cell_update(armon_ctx *ctx, armon_dims d, armon_domain dom, int axis, double dx, double dt,
                      const double *us, const double *ps, double *rho, double *ua, double *E);
/* advection_first_order!  src/projection_schemes.jl:62-89 */
int armon_advection_first_order(armon_ctx *ctx, armon_dims d, armon_domain dom, int axis, double dt,
                                const double *us, const double *rho, const double *u, const double *v, const double *E,
                                double *adv_rho, double *adv_urho, double *adv_vrho, double *adv_Erho);
/* advection_second_order!  src/projection_schemes.jl:92-136 */
int armon_advection_second_order(armon_ctx *ctx, armon_dims d, armon_domain dom, int axis, double dx, double dt,
                                 const double *us, const double *rho, const double *u, const double *v, const double *E,
                                 double *adv_rho, double *adv_urho, double *adv_vrho, double *adv_Erho);
/* euler_projection!  src/projection_schemes.jl:23-52 */
int armon_euler_projection(armon_ctx *ctx, armon_dims d, armon_domain dom, int axis, double dx, double dt,
                           const double *us, double *rho, double *u, double *v, double *E,
                           const double *adv_rho, const double *adv_urho, const double *adv_vrho, const double *adv_Erho);
/* dtCFL_kernel  src/reductions.jl:2-88: min over real cells; returns the host value like `mapreduce` */
int armon_dtCFL(armon_ctx *ctx, armon_dims d, const double *u, const double *v, const double *c,
                double dx, double dy, double *result);
/* conservation_vars  src/reductions.jl:202-298: (sum rho, sum rho*E) * ds over real cells, fixed summation tree */
int armon_conservation_vars(armon_ctx *ctx, armon_dims d, const double *rho, const double *E, double ds,
                            double *mass, double *energy);
/* init_test kernel  src/kernels.jl:106-145, wrapper :176-214: full domain incl. ghosts.  Any of x, y, mask,
 * p, c, g, us, ps, work_* may be NULL (skipped).  n_local_x is params.N[1] (DebugIndexes stride). */
int armon_init_test(armon_ctx *ctx, armon_dims d, int64_t origin_ix, int64_t origin_iy,
                    int64_t global_nx, int64_t global_ny, const double domain_size[2], const double origin[2],
                    const armon_test_case *tc,
                    double *x, double *y, double *mask, double *rho, double *E, double *u, double *v,
                    double *p, double *c, double *g,
                    double *us, double *ps, double *work_1, double *work_2, double *work_3, double *work_4);

/* ---------------------------------------------------------------------------------------------------
 * Fused solver: the overloads of `solver_cycle(params::ArmonParameters{T,<:B200Device}, grid)`
 * (src/solver.jl:288-320), `next_time_step` (src/reductions.jl:164-199), `next_cycle!`
 * (src/solver_state.jl:145-166) and `time_loop`'s stop condition (src/solver.jl:333).  One marching kernel
 * per axis sweep replaces update_EOS! + block_ghost_exchange/boundary_conditions! + numerical_fluxes! +
 * cell_update! + projection_remap!, and folds the dtCFL reduction of the next cycle into the last sweep.
 * ------------------------------------------------------------------------------------------------- */
int armon_solver_create(armon_ctx *ctx, const armon_solver_desc *desc, armon_solver **solver);
int armon_solver_destroy(armon_solver *solver);
/* Bind the caller-owned BlockData arrays: main = (rho, u, v, E), work = (work_1..work_4) used as the
 * second buffer set of the ping-pong.  pcg = (p, c, g) may be NULL pointers (then never written). */
int armon_solver_bind(armon_solver *solver, double *const main_vars[4], double *const work_vars[4],
                      double *const pcg[3]);
/* init_test for the fused path (rho, u, v, E only) + reset!(global_dt) (src/solver_state.jl:58-67) */
int armon_solver_init(armon_solver *solver);
/* reset!(global_dt) alone, after the caller filled the bound arrays itself (e.g. h2d of an initial state) */
int armon_solver_reset(armon_solver *solver);
/* Enqueue up to `n_cycles` solver cycles (asynchronous).  Cycles past maxtime/maxcycle are device-side no-ops,
 * exactly reproducing `while time < maxtime && cycle < maxcycle`. */
int armon_solver_run(armon_solver *solver, int64_t n_cycles);
/* Run until done (blocking): time_loop (src/solver.jl:323-403). */
int armon_solver_time_loop(armon_solver *solver);
/* Blocking read of the GlobalTimeStep scalars. */
int armon_solver_state(armon_solver *solver, armon_time_state *out);
/* Bring rho,u,v,E back to the canonical row-major layout in `main_vars` and, if bound, write the stale
 * p, c, g the reference would hold (EOS at the start of the last sweep).
 * Must precede any device_to_host! / per-step kernel on the bound arrays.  Asynchronous. */
int armon_solver_finalize(armon_solver *solver);
/* Halo exchange of the current state along `axis` alone (block_ghost_exchange with a RemoteTaskBlock,
 * src/halo_exchange.jl:286-310): used by the DebugIndexes halo test (test/mpi.jl:272-360). */
int armon_solver_halo_exchange(armon_solver *solver, int axis);
/* device time (ms, CUDA events on the solver's stream) spent in the cycles enqueued by the last
 * armon_solver_run / armon_solver_time_loop call; blocks until they finished */
int armon_solver_elapsed_ms(armon_solver *solver, float *ms);
/* Per-kernel timing of the sweep launches (CUDA events recorded on the solver's stream around every sweep kernel
 * while enabled): `armon_solver_profile(s, 1)` starts a new measurement, `armon_solver_sweep_time_ms` blocks and
 * returns the summed device time and the number of launches measured.  Used for bench.py's roofline figure. */
int armon_solver_profile(armon_solver *solver, int enable);
int armon_solver_sweep_time_ms(armon_solver *solver, double *total_ms, uint64_t *count);
/* kernel + launch statistics of the fused path */
int armon_solver_sweep_launches(armon_solver *solver, uint64_t *count);
/* Layout the state is kept in between the sweeps of the fused path (no reference counterpart: the reference keeps one
 * row-major layout and strides through it, src/blocking/blocking.jl:148-172): *tiled = 1 when the solver's group runs
 * the band-tiled layout (fast mode, TMA staging, every block and every rank with extents that are multiples of 8;
 * ARMON_B200_TILED=0 disables), 0 when it runs the row-major / transposed pair, -1 before the first cycle decided it.
 * The arrays the caller sees (armon_solver_bind, after armon_solver_finalize) are always in the canonical layout. */
int armon_solver_tiled(armon_solver *solver, int32_t *tiled);
/* Kernel of the bit-exact arithmetic mode (no reference counterpart): *chains = 1 when the solver's strict sweeps run
 * on the four-chain schedule of the fast kernel (sweep_fast_kernel<..., MATH_STRICT>, TMA staging; the default for
 * math_mode strict -- sweeps whose rows are not 16-byte aligned still take the register-prefetch kernel), 0 when they
 * all run the register-prefetch kernel (kernel_variant single, ARMON_B200_STRICT=single) or the solver is not in strict
 * mode.  Same bits either way; informational (bench.py names the kernel it timed). */
int armon_solver_strict_chains(armon_solver *solver, int32_t *chains);

/* Per-cycle diagnostics without a host round trip: the reference's `silent <= 1` log line (src/solver.jl:359-371:
 * wait + conservation_vars + print after every cycle).  While enabled, every cycle enqueues a fixed-tree reduction of
 * (sum rho, sum rho*E) * dx*dy over the real cells of the state it produced, in whatever layout the last sweep left
 * it (no finalize, no copy), and appends {cycle, time, dt, mass, energy} to a device ring of `capacity` lines.
 * `armon_solver_read_diagnostics` blocks, copies out the lines appended since the last read (oldest first, at most
 * `max_lines`) and returns their number; lines of cycles enqueued past the end of the run are dropped. */
int armon_solver_diagnostics(armon_solver *solver, int32_t capacity);     /* 0 disables */
int armon_solver_read_diagnostics(armon_solver *solver, armon_cycle_diag *lines, int64_t max_lines, int64_t *n_lines);

/* ---------------------------------------------------------------------------------------------------
 * Several blocks per GPU: a BlockGrid of LocalTaskBlocks (src/blocking/block_grid.jl:46-183) advanced in lock step.
 * `blocks` are nbx*nby bound solvers created on the same context, block (bx, by) at index by*nbx + bx, each
 * describing its own sub-domain (dims, origin_ix/iy) of the same global grid; sides that face another block of the
 * group must have neighbours[side] == -1 in the descriptor.  The group replaces, between those blocks,
 * `block_ghost_exchange(params, state, blk1::LocalTaskBlock, blk2::LocalTaskBlock, side)` (src/halo_exchange.jl:
 * 107-121,172-186: the g innermost real strips of one block become the g ghost strips of the other) by one
 * device-to-device copy per variable and side -- in the marching layout the strips are g contiguous array rows -- and
 * shares one device-resident GlobalTimeStep: the CFL maxima of all blocks meet in the same accumulators
 * (`next_time_step`'s loop over `all_blocks`, src/reductions.jl:91-110).  Results are bit-identical to one block
 * covering the whole sub-domain.  A grouped solver may only be driven through its group.
 * ------------------------------------------------------------------------------------------------- */
int armon_group_create(armon_ctx *ctx, int32_t nbx, int32_t nby, armon_solver *const blocks[], armon_group **group);
int armon_group_destroy(armon_group *group);          /* the solvers survive and become standalone again */
int armon_group_init(armon_group *group);             /* armon_solver_init of every block */
int armon_group_reset(armon_group *group);
int armon_group_run(armon_group *group, int64_t n_cycles);
int armon_group_time_loop(armon_group *group);
int armon_group_state(armon_group *group, armon_time_state *out);
int armon_group_finalize(armon_group *group);
int armon_group_elapsed_ms(armon_group *group, float *ms);
int armon_group_diagnostics(armon_group *group, int32_t capacity);
int armon_group_read_diagnostics(armon_group *group, armon_cycle_diag *lines, int64_t max_lines, int64_t *n_lines);

/* On-device self test: compares the branch-free division / sqrt of the strict mode with nvcc's IEEE div.rn.f64 /
 * sqrt.rn.f64 on `n_samples` pseudo-random operand pairs.  mismatch = {division, sqrt, shared-reciprocal division,
 * spurious range flags, missed range flags}; all must be 0. */
int armon_selftest_math(armon_ctx *ctx, uint64_t seed, uint64_t n_samples, uint64_t mismatch[5]);

/* ---------------------------------------------------------------------------------------------------
 * Multi-GPU plumbing, replacing MPI.Cart_create / Startall / Iallreduce (src/parameters.jl:408-467,
 * src/halo_exchange.jl:229-283, src/utils.jl:126-134) with NCCL over NVLink.
 * ------------------------------------------------------------------------------------------------- */
int armon_comm_unique_id(char id[128]);                                   /* ncclGetUniqueId on rank 0 */
int armon_ctx_comm_init(armon_ctx *ctx, const char id[128], int rank, int nranks);   /* ncclCommInitRank */
int armon_ctx_comm_destroy(armon_ctx *ctx);   /* ncclCommDestroy: collective in effect, every rank calls it at the same
                                                  point; armon_ctx_destroy on a live communicator aborts it instead */

#ifdef __cplusplus
}
#endif
#endif /* ARMON_B200_H */
