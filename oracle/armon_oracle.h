/*
 * armon_oracle.h -- CPU restatement of Armon.jl's axis-split Lagrange+remap time step.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: it is imported,
 * linked or executed only by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * `--impl reference` legs, and there only as the checker / CPU baseline.  The product path
 * (armon.jl_b200/) never calls into it and fails loudly when its CUDA library is missing.
 *
 * Parity status: PINNED.  The restatement is checked (tests/test_oracle_golden.py) against all five
 * 64-bit golden vectors of the reference (test/reference_data/ref_{Sod,Sod_y,Sod_circ,Bizarrium,
 * Sedov}_64bits.csv, committed as tests/golden/ npz files by tests/golden/make_golden.py): identical cycle
 * counts, final dt to >= 13 digits, fields within the reference's own atol=1e-13/rtol=4eps for the Sod
 * family.  Variants no reference test pins (Godunov, superbee, euler 1st order, other splittings)
 * are pinned only by this oracle.
 *
 * Every function cites the reference file:line it follows (paths relative to the reference root).
 * Build flavours (oracle/Makefile): strict IEEE (-ffp-contract=off, the parity oracle) and
 * FMA-contracted (-ffp-contract=fast -mfma, brackets the @fastmath noise floor of the golden data).
 */
#ifndef ARMON_ORACLE_H
#define ARMON_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* src/utils.jl:15-78 */
enum { ORC_AXIS_X = 0, ORC_AXIS_Y = 1 };
enum { ORC_SIDE_LEFT = 0, ORC_SIDE_RIGHT = 1, ORC_SIDE_BOTTOM = 2, ORC_SIDE_TOP = 3 };
/* src/tests.jl:2-11 (same integer codes as ext/ArmonKokkos.jl:61-69) */
enum { ORC_TEST_SOD = 0, ORC_TEST_SOD_Y = 1, ORC_TEST_SOD_CIRC = 2, ORC_TEST_BIZARRIUM = 3,
       ORC_TEST_SEDOV = 4, ORC_TEST_DEBUG_INDEXES = 5 };
/* src/riemann_schemes.jl:2-3 */
enum { ORC_RIEMANN_GODUNOV = 0, ORC_RIEMANN_GAD = 1 };
/* src/limiters.jl:2-4 (same codes as ext/ArmonKokkos.jl:50-58) */
enum { ORC_LIMITER_NONE = 0, ORC_LIMITER_MINMOD = 1, ORC_LIMITER_SUPERBEE = 2 };
/* src/projection_schemes.jl:2-3 */
enum { ORC_PROJ_EULER = 0, ORC_PROJ_EULER_2ND = 1 };
/* src/axis_splitting.jl:2-5 */
enum { ORC_SPLIT_SEQUENTIAL = 0, ORC_SPLIT_GODUNOV = 1, ORC_SPLIT_STRANG = 2,
       ORC_SPLIT_X_ONLY = 3, ORC_SPLIT_Y_ONLY = 4 };
/* EOS selection: src/kernels.jl:151-161 (Bizarrium test -> bizarrium_EOS!, else perfect gas) */
enum { ORC_EOS_PERFECT_GAS = 0, ORC_EOS_BIZARRIUM = 1 };

/* Inclusive rectangle of cells, 1-based real-cell coordinates (ghosts are <= 0 or > N).
 * Restates DomainRange built by block_domain_range (src/blocking/blocking.jl:71-85). */
typedef struct { int ix0, ix1, iy0, iy1; } orc_domain;

/* Host-computed test-case description (src/tests.jl:59-121,150-211); shared verbatim with the
 * CUDA library so that both sides start from bit-identical initial states. */
typedef struct {
    int    test;           /* ORC_TEST_* */
    double high_rho, low_rho, high_E, low_E, high_u, low_u, high_v, low_v;
    double sedov_r;        /* Sedov{T}.r (src/tests.jl:15-19), unused otherwise */
    double gamma;          /* specific_heat_ratio, 7/5 (src/tests.jl:46) */
    int    eos;            /* ORC_EOS_* */
    double bc_u[4], bc_v[4];  /* (u_factor, v_factor) per side, boundary_condition (src/tests.jl:150-211) */
} orc_test_case;

typedef struct {
    /* sub-domain geometry (src/parameters.jl:673-697) */
    int    nx, ny, g;              /* local real cells and ghost width */
    int    global_nx, global_ny;   /* global grid */
    int    origin_ix, origin_iy;   /* N_origin: 1-based global index of the first local real cell */
    double domain_size[2], origin[2];
    /* schemes (src/parameters.jl:577-629) */
    int    riemann, limiter, projection, splitting;
    double cfl, maxtime;
    int    maxcycle;
    int    cst_dt;
    double Dt;
    /* neighbours: 1 if a remote sub-domain lies on that side (no BC there), src/halo_exchange.jl:286-294 */
    int    has_neighbour[4];
    orc_test_case tc;
    int    nthreads;               /* OpenMP threads for the CPU-baseline build (ignored when built without OpenMP) */
} orc_params;

/* All 16 arrays of BlockData (src/blocking/blocks.jl:18-44), each (nx+2g)*(ny+2g) doubles. */
typedef struct {
    double *x, *y, *rho, *u, *v, *E, *p, *c, *g, *us, *ps, *work_1, *work_2, *work_3, *work_4, *mask;
} orc_data;

/* GlobalTimeStep scalars that matter to the sync path (src/solver_state.jl:26-47). */
typedef struct {
    int    cycle;
    double time, current_dt, next_cycle_dt;
} orc_dt_state;

typedef void (*orc_halo_fn)(void *user, int axis);
typedef double (*orc_min_fn)(void *user, double local_min);

typedef struct {
    orc_params   p;
    orc_data     d;
    orc_dt_state t;
    /* optional hooks standing for MPI (src/halo_exchange.jl:229-283, src/solver_state.jl:102-119) */
    orc_halo_fn  halo_exchange; void *halo_user;
    orc_min_fn   allreduce_min; void *min_user;
    int          error;            /* 1: invalid time step (SolverException(:time), solver_state.jl:123-124) */
} orc_solver;

/* ---- kernels (one C function per reference kernel) ---- */
void orc_perfect_gas_EOS(int nx, int ny, int g, orc_domain dom, double gamma,
                         const double *rho, const double *E, const double *u, const double *v,
                         double *p, double *c, double *gg);
void orc_bizarrium_EOS(int nx, int ny, int g, orc_domain dom,
                       const double *rho, const double *u, const double *v, const double *E,
                       double *p, double *c, double *gg);
void orc_boundary_conditions(int nx, int ny, int g, int side, double u_factor, double v_factor,
                             double *rho, double *u, double *v, double *p, double *c, double *gg, double *E);
void orc_acoustic(int nx, int ny, int g, orc_domain dom, int axis,
                  double *us, double *ps, const double *rho, const double *ua, const double *p, const double *c);
void orc_acoustic_GAD(int nx, int ny, int g, orc_domain dom, int axis, double dt, double dx, int limiter,
                      double *us, double *ps, const double *rho, const double *ua, const double *p, const double *c);
void orc_cell_update(int nx, int ny, int g, orc_domain dom, int axis, double dx, double dt,
                     const double *us, const double *ps, double *rho, double *ua, double *E);
void orc_advection_first_order(int nx, int ny, int g, orc_domain dom, int axis, double dt,
                               const double *us, const double *rho, const double *u, const double *v, const double *E,
                               double *a_rho, double *a_urho, double *a_vrho, double *a_Erho);
void orc_advection_second_order(int nx, int ny, int g, orc_domain dom, int axis, double dx, double dt,
                                const double *us, const double *rho, const double *u, const double *v, const double *E,
                                double *a_rho, double *a_urho, double *a_vrho, double *a_Erho);
void orc_euler_projection(int nx, int ny, int g, orc_domain dom, int axis, double dx, double dt,
                          const double *us, double *rho, double *u, double *v, double *E,
                          const double *a_rho, const double *a_urho, const double *a_vrho, const double *a_Erho);
double orc_dtCFL(int nx, int ny, int g, const double *u, const double *v, const double *c, double dx, double dy);
void orc_conservation_vars(int nx, int ny, int g, const double *rho, const double *E, double ds,
                           double *mass, double *energy);
void orc_init_test(const orc_params *p, orc_data *d);
void orc_steps_ranges(int nx, int ny, int g, int axis, int projection,
                      orc_domain *eos, orc_domain *fluxes, orc_domain *cell_update,
                      orc_domain *advection, orc_domain *proj);

/* ---- solver (src/solver.jl:288-403, src/solver_state.jl:58-166, src/reductions.jl:164-199) ---- */
orc_solver *orc_solver_create(const orc_params *p);
void        orc_solver_destroy(orc_solver *s);
void        orc_solver_init(orc_solver *s);              /* init_test + reset! */
int         orc_split_axes(int splitting, int cycle, int axes[3], double factors[3]);
int         orc_solver_cycle(orc_solver *s);             /* one solver_cycle + next_cycle!; returns error flag */
int         orc_time_loop(orc_solver *s);                /* until maxtime / maxcycle; returns error flag */
/* single steps, for step-by-step comparisons (the reference's `compare=true` checkpoints, src/io.jl:185-227) */
void        orc_step_EOS(orc_solver *s, int axis);
void        orc_step_BC(orc_solver *s, int axis);
void        orc_step_fluxes(orc_solver *s, int axis, double dt);
void        orc_step_cell_update(orc_solver *s, int axis, double dt);
void        orc_step_remap(orc_solver *s, int axis, double dt);
double      orc_local_time_step(orc_solver *s);
int         orc_next_time_step(orc_solver *s);
void        orc_sweep(orc_solver *s, int axis, double dt);
int         orc_num_threads(void);

#ifdef __cplusplus
}
#endif
#endif
