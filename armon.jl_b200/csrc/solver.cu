// solver.cu -- the fused solver behind `solver_cycle` / `next_time_step` / `next_cycle!` / `time_loop`
// (src/solver.jl:288-403, src/reductions.jl:164-199, src/solver_state.jl:58-166) for ArmonParameters{T,<:B200Device}.
//
// Objects.  An `armon_solver` is one block: a sub-domain with its 8 bound arrays (two rotating sets of rho,u,v,E), its
// kernel selection and its halo plumbing.  An `armon_group` is what advances in time: one or several blocks of one
// context sharing one device-resident GlobalTimeStep (DeviceTimeState).  Every solver owns a private group of one (the
// `armon_solver_*` entry points drive it); `armon_group_create` ties several blocks of one GPU together, the
// BlockGrid-of-LocalTaskBlocks of the reference (src/blocking/block_grid.jl:46-183).
//
// Per cycle the stream receives, for every axis of the splitting: [halo: NCCL send/recv with the neighbour ranks,
// device-to-device row copies between local blocks, k_bc_fill on global edges ->] one marching sweep kernel per block
// (sweep_*_kernel.cuh); then [NCCL all-reduce(max) of the CFL accumulators ->] one single-thread kernel that performs
// next_cycle! and the next cycle's time-step update on the device.  Nothing synchronises with the host: dt, time,
// cycle count and the stop condition live in DeviceTimeState.  Launch-bound grids replay two captured cycles (one
// period of every per-cycle parity) as a CUDA graph.
#include "sweep_dispatch.h"

#include <nvtx3/nvToolsExt.h>

#include <cmath>
#include <cstring>
#include <cstdlib>
#include <map>
#include <string>
#include <tuple>
#include <vector>

namespace {

constexpr int TPB = 256;

// NVTX ranges named after the reference's `@section`s (src/solver.jl:293-316, ext/ArmonNVTX.jl): host-side enqueue
// ranges; a no-op unless a tool is attached.
struct NvtxRange {
    explicit NvtxRange(const char *name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
};

// ---- device-side time step state machine ----------------------------------------------------------------------
__global__ void k_ts_reset(DeviceTimeState *ts, int cst_dt, double Dt)
{
    // reset!(global_dt), src/solver_state.jl:58-67
    ts->cycle = 0;
    ts->time = 0.0;
    ts->current_dt = cst_dt ? Dt : 0.0;
    ts->next_cycle_dt = __longlong_as_double(0x7FF0000000000000LL);
    ts->error = 0;
    ts->done = 0;
    ts->range_error = 0;
    ts->redo_count = 0u;
    ts->error_cycle = 0;
    ts->past_end = 0;
    for (int k = 0; k < 3; k++)
        for (int j = 0; j < 4; j++) ts->acc[k][j] = 0ULL;
}

struct CycleStepArgs {
    int first;            // 1: called before cycle 0 (no next_cycle! to apply)
    int read_slot;        // accumulator slot holding the CFL maxima to consume (-1: none, i.e. after cycle 0)
    int acc_is_xy;        // 1: that slot holds (x, y); 0: (y, x)
    double dx, dy;        // GLOBAL cell sizes, src/reductions.jl:91-94
    double cfl, maxtime;
    long long maxcycle;
    int cst_dt;
    double Dt;
};

// next_cycle! (src/solver_state.jl:145-166) of the cycle that just ran, next_time_step + update_dt!
// (src/reductions.jl:164-199, src/solver_state.jl:102-142) and the loop condition of time_loop (src/solver.jl:333).
// SURVEY.md section 3.3 gives the recurrence this reproduces: D0 = cfl L(0) is used by cycles 0 and 1;
// D_k = min(cfl L(k), 1.05 D_{k-1}) is used by cycle k+1, where L(k) comes from the state at the start of cycle k, i.e.
// from the maxima accumulated by the last sweep of cycle k-1.  Like the reference's MPI_Iallreduce (waited one cycle
// later, src/solver_state.jl:89-119,154-156) the reduction of those maxima has the whole of cycle k to complete: the
// step after cycle k consumes the slot written by cycle k-1.
__global__ void k_cycle_step(DeviceTimeState *ts, CycleStepArgs a)
{
    unsigned long long bx = 0ULL, by = 0ULL, berr = 0ULL;
    if (a.read_slot >= 0) {
        bx = ts->acc[a.read_slot][a.acc_is_xy ? 0 : 1];
        by = ts->acc[a.read_slot][a.acc_is_xy ? 1 : 0];
        berr = ts->acc[a.read_slot][2];
        ts->acc[a.read_slot][0] = ts->acc[a.read_slot][1] = 0ULL;
        // error channel: this rank's sticky flag rides on the next all-reduce(max) of the slot, so that every rank
        // sees it at the same cycle step
        ts->acc[a.read_slot][2] = ts->range_error ? 1ULL : 0ULL;
    }
    ts->acc[2][0] = ts->acc[2][1] = 0ULL;
    if (ts->done) {   // a cycle enqueued past the end of the run: a no-op (src/solver.jl:333)
        ts->past_end += 1;
        return;
    }

    if (!a.first) {
        ts->cycle += 1;
        ts->time = __dadd_rn(ts->time, ts->current_dt);
    }
    if (berr) {
        ts->error = ARMON_ERR_RANGE;
        ts->error_cycle = ts->cycle;
        ts->done = 1;
        return;
    }
    ts->next_cycle_dt = __longlong_as_double(0x7FF0000000000000LL);   // typemax(T) after next_cycle!
    if (a.cst_dt) {   // src/reductions.jl:165-167
        ts->current_dt = a.Dt;
        ts->next_cycle_dt = a.Dt;
    } else if (a.read_slot >= 0) {
        // local_time_step: min over cells of min(dx/max(|u+c|,|u-c|), dy/...) == min(dx/max_cells(|u|+c), dy/...)
        // (division by a positive number is monotone, so the min commutes with the correctly rounded quotient)
        const double ax = __longlong_as_double((long long)bx), ay = __longlong_as_double((long long)by);
        double new_dt = fmin(__ddiv_rn(a.dx, ax), __ddiv_rn(a.dy, ay));
        if (ax != ax || ay != ay) new_dt = ax + ay;   // NaN in the fields: propagate
        const double previous_dt = ts->current_dt;
        if (!isfinite(new_dt) || new_dt <= 0.0) {   // src/solver_state.jl:123-124
            // the reference meets this value in next_time_step at the start of the cycle whose state produced it: the
            // maxima consumed here are one cycle old
            ts->error = ARMON_ERR_TIME;
            ts->error_cycle = a.first ? 0 : ts->cycle - 1;
            ts->done = 1;
            return;
        } else if (previous_dt == 0.0) {
            new_dt = __dmul_rn(a.cfl, new_dt);
        } else {
            new_dt = fmin(__dmul_rn(a.cfl, new_dt), __dmul_rn(1.05, previous_dt));
        }
        ts->current_dt = new_dt;
    }
    if (!(ts->time < a.maxtime && ts->cycle < a.maxcycle)) ts->done = 1;
}

// EOS_init + the first local_time_step (src/solver.jl:291-297): CFL maxima of the initial state.
// Works on either layout: the reduction does not care about cell order.
constexpr int INIT_DT_ROWS = 32;
template <int EOS>
__global__ void k_init_dt(long long n_rows, long long n_cols, long long pitch, int g, const double *rho,
                          const double *u, const double *v, const double *E, double gamma, DeviceTimeState *ts)
{
    const long long col = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long bx = 0ULL, by = 0ULL;
    // INIT_DT_ROWS rows per block: one pair of atomics per warp and 32 rows instead of per row (the maxima are
    // order-independent, so the grouping does not change the result)
    for (long long r = (long long)blockIdx.y * INIT_DT_ROWS; r < n_rows && r < ((long long)blockIdx.y + 1) * INIT_DT_ROWS; r++) {
        if (col >= n_cols) break;
        const long long i = (r + g) * pitch + (col + g);
        sd p, c, gg;
        RangeFlag f;
        if (EOS == ARMON_EOS_BIZARRIUM) eos_bizarrium<sd, DIV_IEEE, false>(sd(rho[i]), sd(u[i]), sd(v[i]), sd(E[i]), p, c, gg, f);
        else eos_perfect_gas<sd, DIV_IEEE>(sd(gamma), sd(rho[i]), sd(u[i]), sd(v[i]), sd(E[i]), p, c, f);
        const unsigned long long cx = (unsigned long long)__double_as_longlong(__dadd_rn(fabs(u[i]), c.v));
        const unsigned long long cy = (unsigned long long)__double_as_longlong(__dadd_rn(fabs(v[i]), c.v));
        bx = cx > bx ? cx : bx;
        by = cy > by ? cy : by;
    }
    for (int off = 16; off > 0; off >>= 1) {
        const unsigned long long ox = __shfl_xor_sync(0xffffffffu, bx, off);
        const unsigned long long oy = __shfl_xor_sync(0xffffffffu, by, off);
        bx = ox > bx ? ox : bx;
        by = oy > by ? oy : by;
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMax(&ts->acc[1][0], bx);   // slot of "cycle -1": consumed by the step before cycle 0
        atomicMax(&ts->acc[1][1], by);
    }
}

// boundary_conditions! (src/halo_exchange.jl:2-36) for the two sides along the march axis of the sweep about to run:
// ghost row -1-k <- real row k (low side), ghost row nm+k <- real row nm-1-k (high side), k = 0..g-1, real columns
// only (no corners, src/blocking/blocking.jl:148-172); rho and E copied, the velocities multiplied by the factors of
// boundary_condition(test, side) (src/tests.jl:150-211).  O(perimeter); the marching kernels then read every ghost
// row as plain data, whether it came from here or from the halo exchange.
struct BcFillArgs {
    double *rho, *ua, *ut, *E;
    long long nm, nw, pitch;
    int g, lo, hi;
    int tiled;              // band-tiled layout (common.cuh): the ghost rows of a side are one band
    double fa_lo, ft_lo, fa_hi, ft_hi;
};

__global__ void k_bc_fill(BcFillArgs B)
{
    const long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int k = blockIdx.y, side = blockIdx.z;
    if (w >= B.nw || !(side == 0 ? B.lo : B.hi)) return;
    const long long src_row = side == 0 ? k : B.nm - 1 - k, dst_row = side == 0 ? -1 - k : B.nm + k;
    const long long is = layout_index(B.tiled != 0, src_row + B.g, w + B.g, B.pitch);
    const long long id = layout_index(B.tiled != 0, dst_row + B.g, w + B.g, B.pitch);
    const double fa = side == 0 ? B.fa_lo : B.fa_hi, ft = side == 0 ? B.ft_lo : B.ft_hi;
    B.rho[id] = B.rho[is];
    B.E[id] = B.E[is];
    B.ua[id] = __dmul_rn(B.ua[is], fa);
    B.ut[id] = __dmul_rn(B.ut[is], ft);
}

// ---- layout helpers ---------------------------------------------------------------------------------------------
struct Ptr4 { const double *in[4]; double *out[4]; };

// Change of layout of 4 arrays at once: out[c][r] = in[r][c] (transpose) or out[r][c] = in[r][c], each side row-major
// or band-tiled (common.cuh); rows x cols are the full extents of the INPUT including ghosts.  Runs where the state
// enters or leaves the marching layouts (first sweep, finalize, an unpredicted change of axis), never inside the loop.
__global__ void k_relayout4(Ptr4 P, long long rows, long long cols, int in_tiled, int out_tiled, int transpose)
{
    __shared__ double tile[4][32][33];
    const long long c0 = (long long)blockIdx.x * 32, r0 = (long long)blockIdx.y * 32;
    const int tx = threadIdx.x, ty = threadIdx.y;   // 32 x 8
#pragma unroll
    for (int k = 0; k < 4; k++)
        for (int j = ty; j < 32; j += 8) {
            const long long r = r0 + j, c = c0 + tx;
            if (r < rows && c < cols) tile[k][j][tx] = P.in[k][layout_index(in_tiled != 0, r, c, cols)];
        }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 4; k++)
        for (int j = ty; j < 32; j += 8) {
            if (transpose) {
                const long long c = c0 + j, r = r0 + tx;
                if (r < rows && c < cols) P.out[k][layout_index(out_tiled != 0, c, r, rows)] = tile[k][tx][j];
            } else {
                const long long r = r0 + j, c = c0 + tx;
                if (r < rows && c < cols) P.out[k][layout_index(out_tiled != 0, r, c, cols)] = tile[k][j][tx];
            }
        }
}

// Stale p, c, g of the reference (EOS of the state at the start of the last sweep), canonical layout out.
// Fixed-size grid with strided loops, so that the conditional launches cost nothing when they do not apply.
//   mode 0: always; mode 1 (after every cycle step): only when the run has just ended (done, nothing enqueued past the
//   end yet: the other buffer set still holds the input of the last real sweep); mode 2 (finalize): unless cycles were
//   enqueued past the end (then mode 1 already saved them before its source was overwritten).
template <int EOS>
__global__ void k_eos_pcg(long long nx, long long ny, int g, int in_transposed, int in_tiled, const double *rho, const double *u,
                          const double *v, const double *E, double gamma, double *p, double *c, double *gg,
                          const DeviceTimeState *ts, int mode)
{
    if (mode == 1 && !(ts->done && ts->past_end == 0)) return;
    if (mode == 2 && ts->past_end != 0) return;
    for (long long iy = blockIdx.y; iy < ny; iy += gridDim.y)
        for (long long ix = (long long)blockIdx.x * blockDim.x + threadIdx.x; ix < nx; ix += (long long)gridDim.x * blockDim.x) {
            const long long io = (iy + g) * (nx + 2 * g) + (ix + g);
            const long long ii = in_transposed ? layout_index(in_tiled != 0, ix + g, iy + g, ny + 2 * g)
                                               : layout_index(in_tiled != 0, iy + g, ix + g, nx + 2 * g);
            sd pp, cc, g_;
            RangeFlag f;
            if (EOS == ARMON_EOS_BIZARRIUM) {
                eos_bizarrium<sd, DIV_IEEE, true>(sd(rho[ii]), sd(u[ii]), sd(v[ii]), sd(E[ii]), pp, cc, g_, f);
            } else {
                eos_perfect_gas<sd, DIV_IEEE>(sd(gamma), sd(rho[ii]), sd(u[ii]), sd(v[ii]), sd(E[ii]), pp, cc, f);
                g_ = (sd(1.) + sd(gamma)) / sd(2.);
            }
            if (p) p[io] = pp.v;
            if (c) c[io] = cc.v;
            if (gg) gg[io] = g_.v;
        }
}

// ---- per-cycle diagnostics (conservation_vars, src/reductions.jl:202-298) on the current buffers ---------------
// Fixed summation tree: every array row of the current layout is summed by one CTA (strided per-thread partial sums,
// then a binary tree), the row partials by one CTA in the same way: deterministic for a given grid and layout.
__global__ void k_diag_rows(long long n_rows, long long n_cols, long long pitch, int g, int tiled, const double *rho,
                            const double *E, double *row_m, double *row_e)
{
    __shared__ double sm[TPB], se[TPB];
    for (long long r = blockIdx.x; r < n_rows; r += gridDim.x) {
        double m = 0.0, e = 0.0;
        for (long long c = threadIdx.x; c < n_cols; c += TPB) {
            const long long i = layout_index(tiled != 0, r + g, c + g, pitch);
            m = __dadd_rn(m, rho[i]);
            e = __dadd_rn(e, __dmul_rn(rho[i], E[i]));
        }
        sm[threadIdx.x] = m; se[threadIdx.x] = e;
        __syncthreads();
        for (int off = TPB / 2; off > 0; off >>= 1) {
            if (threadIdx.x < off) {
                sm[threadIdx.x] = __dadd_rn(sm[threadIdx.x], sm[threadIdx.x + off]);
                se[threadIdx.x] = __dadd_rn(se[threadIdx.x], se[threadIdx.x + off]);
            }
            __syncthreads();
        }
        if (threadIdx.x == 0) { row_m[r] = sm[0]; row_e[r] = se[0]; }
        __syncthreads();
    }
}

__global__ void k_diag_final(long long n_parts, const double *row_m, const double *row_e, double ds,
                             const DeviceTimeState *ts, armon_cycle_diag *ring, unsigned long long *head, unsigned cap)
{
    __shared__ double sm[TPB], se[TPB];
    double m = 0.0, e = 0.0;
    for (long long r = threadIdx.x; r < n_parts; r += TPB) {
        m = __dadd_rn(m, row_m[r]);
        e = __dadd_rn(e, row_e[r]);
    }
    sm[threadIdx.x] = m; se[threadIdx.x] = e;
    __syncthreads();
    for (int off = TPB / 2; off > 0; off >>= 1) {
        if (threadIdx.x < off) {
            sm[threadIdx.x] = __dadd_rn(sm[threadIdx.x], sm[threadIdx.x + off]);
            se[threadIdx.x] = __dadd_rn(se[threadIdx.x], se[threadIdx.x + off]);
        }
        __syncthreads();
    }
    if (threadIdx.x == 0 && ts->past_end == 0) {   // cycles enqueued past the end of the run log nothing
        armon_cycle_diag &d = ring[*head % cap];
        d.cycle = ts->cycle;
        d.time = ts->time;
        d.dt = ts->current_dt;
        d.mass = __dmul_rn(sm[0], ds);
        d.energy = __dmul_rn(se[0], ds);
        *head += 1ULL;
    }
}

bool is_pow2_double(double x)
{
    if (!(x > 0.0) || !std::isfinite(x)) return false;
    int e;
    return std::frexp(x, &e) == 0.5;
}

int split_axes(int splitting, long long cycle, int axes[3], double factors[3])
{
    // src/axis_splitting.jl:24-46
    const bool even = (cycle % 2) == 0;
    switch (splitting) {
    case ARMON_SPLIT_SEQUENTIAL:
        axes[0] = ARMON_AXIS_X; axes[1] = ARMON_AXIS_Y; factors[0] = factors[1] = 1.0; return 2;
    case ARMON_SPLIT_GODUNOV:
        axes[0] = even ? ARMON_AXIS_X : ARMON_AXIS_Y; axes[1] = even ? ARMON_AXIS_Y : ARMON_AXIS_X;
        factors[0] = factors[1] = 1.0; return 2;
    case ARMON_SPLIT_STRANG:
        axes[0] = axes[2] = even ? ARMON_AXIS_X : ARMON_AXIS_Y; axes[1] = even ? ARMON_AXIS_Y : ARMON_AXIS_X;
        factors[0] = factors[2] = 0.5; factors[1] = 1.0; return 3;
    case ARMON_SPLIT_X_ONLY: axes[0] = ARMON_AXIS_X; factors[0] = 1.0; return 1;
    default:                 axes[0] = ARMON_AXIS_Y; factors[0] = 1.0; return 1;
    }
}

}   // namespace

struct armon_solver {
    armon_ctx        *ctx = nullptr;
    armon_solver_desc d{};
    double           *buf[2][4] = {{nullptr, nullptr, nullptr, nullptr}, {nullptr, nullptr, nullptr, nullptr}};
    double           *pcg[3] = {nullptr, nullptr, nullptr};
    bool              bound = false;
    int               cur = 0;                 // buffer set holding the current state (0 = main_vars, 1 = work_vars)
    bool              cur_transposed = false;  // false: canonical rows = y; true: rows = x
    bool              have_prev = false;       // the other set still holds the state at the start of the last sweep
    bool              prev_transposed = false;
    // band-tiled marching layout (sweep_fast_kernel.cuh, 5.): `tiled_ok` = this block can run it (fast mode, TMA staging,
    // extents multiples of 8); whether it is used is decided per group and across ranks (decide_tiling)
    bool              tiled_ok = false, cur_tiled = false, prev_tiled = false;
    sweep_fast_fn_t   fast_tiled_kernel[2] = {nullptr, nullptr}, fast_tiled_cons_kernel[2] = {nullptr, nullptr};
    DeviceTimeState  *own_ts = nullptr;        // the time-step state of the solver's private group
    DeviceTimeState  *ts = nullptr;            // the state in use: own_ts, or the one of the block group it belongs to
    armon_group      *self = nullptr;          // private group of one
    armon_group      *group = nullptr;         // group driving this solver (self unless armon_group_create claimed it)
    armon_solver     *local_nb[4] = {nullptr, nullptr, nullptr, nullptr};   // neighbouring block of the same group per side
    cudaEvent_t       ev_state = nullptr, ev_halo = nullptr;   // compute -> comm (state ready), comm -> compute (ghosts ready)
    cudaStream_t      edge_stream = nullptr;                   // the edge segments of an overlapped sweep
    cudaEvent_t       ev_edge = nullptr;                       // edge segments done
    uint64_t          sweep_launches = 0;
    sweep_fn_t        kernel = nullptr;                        // register-prefetch marching kernel (always available)
    sweep_fixup_fn_t  fixup_kernel = nullptr;  // IEEE fix-up of the strict mode's staged kernel (strict4_kernel)
    // fast-mode marching kernels (sweep_fast_kernel.cuh), [staging variant][transposed output]; `fast_stg` is the
    // staging used for even pitches (STG_TMA or STG_CPA16), odd pitches always take STG_CPA8
    sweep_fast_fn_t   fast_kernel[3][2] = {{nullptr, nullptr}, {nullptr, nullptr}, {nullptr, nullptr}};
    bool              use_fast = false;
    int               fast_stg = STG_TMA;
    sweep_fast_fn_t   fast_cons_kernel[2] = {nullptr, nullptr};   // TMA staging + conservation sums, [transposed output]
    // strict arithmetic on the fast kernel's schedule (sweep_fast_kernel.cuh 6.), [transposed output]; even pitches
    sweep_fast_fn_t   strict4_kernel[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};   // [cell size is a power of two][transposed output]
    bool              use_strict4 = false;
    // per-cycle diagnostics: this block's region of the group's partial-sum scratch, and whether the last sweep of the
    // cycle being enqueued filled it itself (fused) or k_diag_rows has to
    long long         diag_base = 0, diag_cap = 0;
    double           *cons_m = nullptr, *cons_e = nullptr;   // set by group_sweep for the last sweep of a cycle
    bool              cons_done = false;
    std::map<std::tuple<const double *, long long, long long, int>, CUtensorMap> tmaps;   // (array, rows, pitch, box rows) -> tensor map
    bool              overlap = true;          // interior / edge split of a sweep around the halo exchange (ARMON_B200_OVERLAP=0 disables)
    unsigned         *fix_count = nullptr;     // two counters, used alternately by successive sweeps
    unsigned long long *fix_list = nullptr;
    unsigned          fix_cap = 0;             // entries
    uint64_t          sweep_index = 0;
    // optional per-sweep-kernel timing (CUDA events on the launching stream), for the roofline figure
    bool              profile = false;
    std::vector<cudaEvent_t> prof_events;     // pairs (before, after) of each profiled sweep launch
    size_t            prof_used = 0;
};

struct armon_group {
    armon_ctx                 *ctx = nullptr;
    std::vector<armon_solver *> blocks;       // blocks[0] is the master: its descriptor holds the run parameters
    int                        nbx = 1, nby = 1;
    bool                       is_self = true;
    DeviceTimeState           *ts = nullptr;  // = blocks[0]->own_ts
    long long                  host_cycle = 0;          // cycles enqueued since the last reset
    bool                       started = false;         // initial time step enqueued
    cudaEvent_t                ev_start = nullptr, ev_stop = nullptr;
    cudaEvent_t                ev_dt[2] = {nullptr, nullptr};    // all-reduce of accumulator slot 0 / 1 done
    bool                       last_axis_is_x[2] = {true, true}; // axis of the sweep that filled accumulator slot 0 / 1
    bool                       timed = false;
    int                        tiled = -1;              // band-tiled layout between the sweeps: -1 undecided, 0 no, 1 yes
    int                       *agree = nullptr;         // device word of the cross-rank agreement on `tiled`
    // CUDA graph of two consecutive cycles (one period of the buffer rotation, the accumulator slots and the
    // alternating splittings)
    int                        graph_mode = 0;          // 0 auto, 1 on, 2 off (armon_solver_desc.cuda_graph)
    cudaGraphExec_t            graph_exec = nullptr;
    uint64_t                   graph_launches = 0;      // kernel launches inside the captured pair
    uint64_t                   graph_sweeps = 0;
    uint64_t                   graph_replays = 0;
    // per-cycle diagnostics ring
    armon_cycle_diag          *diag_ring = nullptr;
    unsigned long long        *diag_head = nullptr;
    double                    *diag_rows = nullptr;     // row partials of every block, (mass, energy) halves
    long long                  diag_rows_cap = 0;
    unsigned                   diag_cap = 0;
    unsigned long long         diag_read = 0;
};

namespace {

int group_check(armon_group *G)
{
    ARMON_CHECK_ARG(G != nullptr && G->ctx != nullptr && !G->blocks.empty(), "null group");
    for (armon_solver *b : G->blocks) ARMON_CHECK_ARG(b->bound, "armon_solver_bind was not called");
    return armon_ctx_activate(G->ctx);
}

int solver_check(armon_solver *s, bool need_bound = true, bool need_standalone = true)
{
    ARMON_CHECK_ARG(s != nullptr && s->ctx != nullptr, "null solver");
    if (need_bound) ARMON_CHECK_ARG(s->bound, "armon_solver_bind was not called");
    if (need_standalone) ARMON_CHECK_ARG(s->group == s->self, "the solver belongs to a block group: drive it with armon_group_*");
    return armon_ctx_activate(s->ctx);
}

long long n_elems(const armon_solver *s)
{
    return (s->d.dims.nx + 2 * s->d.dims.g) * (s->d.dims.ny + 2 * s->d.dims.g);
}

void drop_graph(armon_group *G)
{
    if (G->graph_exec) {
        cudaGraphExecDestroy(G->graph_exec);
        G->graph_exec = nullptr;
    }
}

// Rewrite the current state into the other buffer set in another layout (orientation x row-major / band-tiled).
int relayout_current(armon_solver *s, bool transposed, bool tiled)
{
    if (transposed == s->cur_transposed && tiled == s->cur_tiled) return ARMON_OK;
    const armon_dims &D = s->d.dims;
    const long long rows = s->cur_transposed ? D.nx + 2 * D.g : D.ny + 2 * D.g;
    const long long cols = s->cur_transposed ? D.ny + 2 * D.g : D.nx + 2 * D.g;
    Ptr4 P;
    for (int k = 0; k < 4; k++) { P.in[k] = s->buf[s->cur][k]; P.out[k] = s->buf[1 - s->cur][k]; }
    const dim3 grid((unsigned)((cols + 31) / 32), (unsigned)((rows + 31) / 32), 1), block(32, 8, 1);
    k_relayout4<<<grid, block, 0, s->ctx->stream>>>(P, rows, cols, s->cur_tiled, tiled, transposed != s->cur_transposed);
    ARMON_LAUNCH_CHECK(s->ctx);
    s->cur = 1 - s->cur;
    s->cur_transposed = transposed;
    s->cur_tiled = tiled;
    s->have_prev = false;
    return ARMON_OK;
}

// A sweep along `axis` marches along the strided dimension: X needs the transposed orientation, Y the canonical one.
int ensure_layout(armon_solver *s, int axis, bool tiled)
{
    return relayout_current(s, axis == ARMON_AXIS_X, tiled);
}

// block_ghost_exchange with RemoteTaskBlocks (src/halo_exchange.jl:286-354): the two sides along `axis`.  In the
// marching layout both sides are `g` contiguous rows of each array, so there is no pack/unpack kernel: the g
// innermost real rows are sent, the g ghost rows received (translation, same orientation as the reference's
// pack_to_array!/unpack_from_array!, src/halo_exchange.jl:187-216).  Only rho, u, v, E travel: p, c, g are
// recomputed by the receiver from the same values, bit for bit.
int halo_exchange(armon_solver *s, int axis, cudaStream_t stream)
{
    const armon_dims &D = s->d.dims;
    const int lo_side = axis == ARMON_AXIS_X ? ARMON_SIDE_LEFT : ARMON_SIDE_BOTTOM;
    const int lo = s->d.neighbours[lo_side], hi = s->d.neighbours[lo_side + 1];
    if (lo < 0 && hi < 0) return ARMON_OK;
    if (!s->ctx->comm) {
        armon_set_error("a neighbour rank is set but the context has no communicator (armon_ctx_comm_init)");
        return ARMON_ERR_INVALID;
    }
    const long long nm = axis == ARMON_AXIS_X ? D.nx : D.ny;
    const long long pitch = (axis == ARMON_AXIS_X ? D.ny : D.nx) + 2 * D.g;
    const size_t count = (size_t)(D.g * pitch);
    ARMON_NCCL(ncclGroupStart());
    for (int k = 0; k < 4; k++) {
        double *a = s->buf[s->cur][k];
        if (lo >= 0) {
            ARMON_NCCL(ncclSend(a + D.g * pitch, count, ncclDouble, lo, s->ctx->comm, stream));
            ARMON_NCCL(ncclRecv(a, count, ncclDouble, lo, s->ctx->comm, stream));
        }
        if (hi >= 0) {
            ARMON_NCCL(ncclSend(a + nm * pitch, count, ncclDouble, hi, s->ctx->comm, stream));
            ARMON_NCCL(ncclRecv(a + (nm + D.g) * pitch, count, ncclDouble, hi, s->ctx->comm, stream));
        }
    }
    ARMON_NCCL(ncclGroupEnd());
    return ARMON_OK;
}

// block_ghost_exchange between two LocalTaskBlocks (src/halo_exchange.jl:107-121,172-186): the g innermost real rows of
// the neighbouring block become the g ghost rows of this one (each block pulls its own ghosts; the pair of pulls is
// the reference's symmetric exchange).  All blocks of a group are in the same layout and rotation state, and two
// blocks facing each other along `axis` have the same extent along the other axis, hence the same pitch: one
// contiguous device-to-device copy per variable and side, on the compute stream (ordered after the sweeps that
// produced the rows and before the sweeps that read them).
int local_halo(armon_solver *s, int axis)
{
    const armon_dims &D = s->d.dims;
    const int lo_side = axis == ARMON_AXIS_X ? ARMON_SIDE_LEFT : ARMON_SIDE_BOTTOM;
    const long long nm = axis == ARMON_AXIS_X ? D.nx : D.ny;
    const long long pitch = (axis == ARMON_AXIS_X ? D.ny : D.nx) + 2 * D.g;
    const size_t bytes = (size_t)(D.g * pitch) * sizeof(double);
    if (const armon_solver *nb = s->local_nb[lo_side]) {
        const long long nb_nm = axis == ARMON_AXIS_X ? nb->d.dims.nx : nb->d.dims.ny;
        for (int k = 0; k < 4; k++)   // real rows nb_nm-g .. nb_nm-1 of the neighbour = its array rows nb_nm .. nb_nm+g-1
            ARMON_CUDA(cudaMemcpyAsync(s->buf[s->cur][k], nb->buf[nb->cur][k] + nb_nm * pitch, bytes,
                                       cudaMemcpyDeviceToDevice, s->ctx->stream));
    }
    if (const armon_solver *nb = s->local_nb[lo_side + 1]) {
        for (int k = 0; k < 4; k++)   // real rows 0 .. g-1 of the neighbour = its array rows g .. 2g-1
            ARMON_CUDA(cudaMemcpyAsync(s->buf[s->cur][k] + (nm + D.g) * pitch, nb->buf[nb->cur][k] + D.g * pitch, bytes,
                                       cudaMemcpyDeviceToDevice, s->ctx->stream));
    }
    return ARMON_OK;
}

// Every NCCL call of the solver goes to the context's communication stream, fenced by events on both sides: the
// compute stream's work so far is visible to it (ev_state), and the compute stream continues after it when
// `wait_after` (else the caller waits on ev_halo itself, after launching the work that overlaps the exchange).
int comm_begin(armon_solver *s)
{
    ARMON_CUDA(cudaEventRecord(s->ev_state, s->ctx->stream));
    ARMON_CUDA(cudaStreamWaitEvent(s->ctx->comm_stream, s->ev_state, 0));
    return ARMON_OK;
}

int comm_end(armon_solver *s, bool wait_after)
{
    ARMON_CUDA(cudaEventRecord(s->ev_halo, s->ctx->comm_stream));
    if (wait_after) ARMON_CUDA(cudaStreamWaitEvent(s->ctx->stream, s->ev_halo, 0));
    return ARMON_OK;
}

int pick_segment(const armon_solver *s, long long nm, long long nw)
{
    if (s->d.march_segment > 0) {   // a multiple of the staging chunks of every kernel (16: aligned 128-byte pieces)
        int seg = (s->d.march_segment + 15) / 16 * 16;
        return seg;
    }
    if (s->use_fast || (s->use_strict4 && ((nw + 2 * s->d.dims.g) % 2) == 0)) {
        // Fast kernels (8 warps = 2 CTAs per SM): every CTA marches seg + 12 rows (the warm-up rows of a segment are
        // redundant work) and the CTAs run in waves of 2 x SMs, so the sweep costs about waves x (seg + 12) row times.
        // Take the multiple of 16 that minimises it -- a power of two can sit just above a whole number of waves (the
        // tiled layout has one more, mostly empty, column of CTAs: 65 x 32 CTAs at 8192^2 are 7.03 waves).
        const long long g = s->tiled_ok ? s->d.dims.g : 0;
        const long long ncol = (nw + g + ASYNC_TPB - 1) / ASYNC_TPB;
        const long long slots = 2LL * s->ctx->sm_count;
        long long best_cost = -1;
        int best = 16;
        for (int seg = 16; seg <= 4096; seg += 16) {
            const long long nseg = (nm + seg - 1) / seg;
            const long long waves = (ncol * nseg + slots - 1) / slots;
            const long long cost = waves * (seg + 12);
            if (best_cost < 0 || cost <= best_cost) { best_cost = cost; best = seg; }   // ties: the longer segment
            if (seg >= nm) break;
        }
        return best;
    }
    // register-prefetch kernels: as long as possible (the warm-up rows of every segment are redundant work) while
    // keeping >= 6 waves of CTAs
    const long long cols_per_cta = SWEEP_TPB;
    const long long ctas_per_sm = 2;
    const long long ncol = (nw + cols_per_cta - 1) / cols_per_cta;
    const long long want = 6LL * ctas_per_sm * s->ctx->sm_count;
    const int cands[] = {2048, 1024, 512, 256, 128, 64, 32, 16};
    for (int seg : cands) {
        if (seg > nm && seg != 16) continue;
        if (ncol * ((nm + seg - 1) / seg) >= want) return seg;
    }
    return nm >= 64 ? 32 : 16;
}

// Tensor map of one input array in the marching layout: a 2-D Float64 tensor [rows][pitch], box = 4 rows x 32 columns
// (one staging group of one warp, sweep_fast_kernel.cuh); band-tiled arrays are described as [bands][4 pitch] with a box
// of 1 band x 128 elements (the 4 adjacent tiles of the same 4 rows x 32 columns).  cuTensorMapEncodeTiled is reached through the runtime's
// driver entry point table, so the library does not link against libcuda.  Maps are cached per (array, extent).
typedef CUresult (*encode_tiled_fn_t)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                      const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                      CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int tensor_map_for(armon_solver *s, const double *arr, long long rows, long long pitch, int box_rows, int box_cols,
                   CUtensorMap *out)
{
    const auto key = std::make_tuple(arr, rows, pitch, box_rows);
    const auto it = s->tmaps.find(key);
    if (it != s->tmaps.end()) { *out = it->second; return ARMON_OK; }
    static encode_tiled_fn_t encode = nullptr;
    if (!encode) {
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        ARMON_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
        if (!fn || qres != cudaDriverEntryPointSuccess) {
            armon_set_error("cuTensorMapEncodeTiled is not available in this driver");
            return ARMON_ERR_CUDA;
        }
        encode = reinterpret_cast<encode_tiled_fn_t>(fn);
    }
    alignas(64) CUtensorMap map;
    const cuuint64_t dims[2] = {(cuuint64_t)pitch, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)pitch * sizeof(double)};   // bytes, multiple of 16: even pitch
    const cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1u, 1u};
    const CUresult rc = encode(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double *>(arr), dims, strides, box,
                               estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                               CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc != CUDA_SUCCESS) {
        armon_set_error("cuTensorMapEncodeTiled failed (%d) for a %lld x %lld array", (int)rc, rows, pitch);
        return ARMON_ERR_CUDA;
    }
    s->tmaps[key] = map;
    *out = map;
    return ARMON_OK;
}

bool is_mirror(const armon_solver *s, int side) { return s->d.neighbours[side] < 0 && s->local_nb[side] == nullptr; }

// One axis sweep of one block: [NCCL halo exchange with the neighbour ranks] + k_bc_fill on global edges + the
// marching kernel.  The block must already be in the layout of `axis`, and the ghost rows that come from local
// neighbour blocks must already be in place (local_halo).
int launch_sweep(armon_solver *s, int axis, double dt_factor, int acc_slot, int next_axis)
{
    const armon_dims &D = s->d.dims;
    const armon_test_case &tc = s->d.tc;
    const bool x = axis == ARMON_AXIS_X;
    SweepArgs A;
    const int in_set = s->cur, out_set = 1 - s->cur;
    // roles: (rho, ua, ut, E); storage order of a set: (rho, u, v, E)
    const int role[4] = {0, x ? 1 : 2, x ? 2 : 1, 3};
    for (int k = 0; k < 4; k++) { A.in[k] = s->buf[in_set][role[k]]; A.out[k] = s->buf[out_set][role[k]]; }
    A.nm = x ? D.nx : D.ny;
    A.nw = x ? D.ny : D.nx;
    A.g = (int)D.g;
    A.pitch_in = A.nw + 2 * D.g;
    A.transpose_out = (next_axis != axis) ? 1 : 0;
    A.pitch_out = A.transpose_out ? A.nm + 2 * D.g : A.nw + 2 * D.g;
    A.seg = pick_segment(s, A.nm, A.nw);
    const int lo_side = x ? ARMON_SIDE_LEFT : ARMON_SIDE_BOTTOM, hi_side = lo_side + 1;
    A.mirror_lo = is_mirror(s, lo_side);
    A.mirror_hi = is_mirror(s, hi_side);
    A.bc_a_lo = x ? tc.bc_u[lo_side] : tc.bc_v[lo_side];
    A.bc_t_lo = x ? tc.bc_v[lo_side] : tc.bc_u[lo_side];
    A.bc_a_hi = x ? tc.bc_u[hi_side] : tc.bc_v[hi_side];
    A.bc_t_hi = x ? tc.bc_v[hi_side] : tc.bc_u[hi_side];
    A.dx = s->d.domain_size[axis] / (double)(x ? s->d.global_nx : s->d.global_ny);   // update_solver_state!
    A.dx_pow2 = is_pow2_double(A.dx) ? 1 : 0;
    A.inv_dx = 1.0 / A.dx;
    A.dt_factor = dt_factor;
    A.gamma = tc.gamma;
    A.gm1 = tc.gamma - 1.0;
    A.ggm1 = tc.gamma * (tc.gamma - 1.0);
    A.ts = s->ts;
    A.acc_slot = acc_slot;

    // block_ghost_exchange with the neighbour ranks (src/halo_exchange.jl:286-354) runs on the communication stream.
    // Only the march segments next to the two ends read ghost rows: the interior segments are launched right away and
    // overlap the exchange, the edge segments follow once the ghost rows have arrived.  Segment n-2 also reads ghost
    // rows when the last segment is shorter than the ghost width (its cells need rows up to m1+3): it then belongs to
    // the edge launches.
    const long long nseg = (A.nm + A.seg - 1) / A.seg;
    A.nseg = (int)nseg;
    const bool has_nb = s->d.neighbours[lo_side] >= 0 || s->d.neighbours[hi_side] >= 0;
    // (the fast kernels cut their segments 4 cells earlier -- 64-byte aligned transposed stores -- so their last
    // segment is never shorter than 5 cells)
    // strict arithmetic on the fast kernel's schedule: TMA staging needs 16-byte aligned rows (even pitch)
    const bool strict4_launch = s->use_strict4 && (A.pitch_in % 2) == 0;
    const bool short_tail = !(s->use_fast || strict4_launch) && A.nm - (nseg - 1) * A.seg < A.g;
    const long long n_interior = nseg - 2 - (short_tail ? 1 : 0);
    const bool overlap = has_nb && n_interior >= 1 && s->overlap;
    if (has_nb) {
        NvtxRange r("BC");
        if (int rc = comm_begin(s)) return rc;
        if (int rc = halo_exchange(s, axis, s->ctx->comm_stream)) return rc;
        if (int rc = comm_end(s, !overlap)) return rc;
    }

    const bool fast_launch = s->use_fast || strict4_launch;
    const bool tiled = s->cur_tiled;   // layout of the input and of the output (ensure_layout ran before)
    if (tiled && !(fast_launch && s->tiled_ok)) {
        armon_set_error("internal: band-tiled state without the tiled kernels");
        return ARMON_ERR_INVALID;
    }
    const int stg = (tiled || strict4_launch) ? STG_TMA : (A.pitch_in % 2) == 0 ? s->fast_stg : STG_CPA8;
    SweepTmaMaps maps;
    if (fast_launch && stg == STG_TMA) {
        for (int k = 0; k < 4; k++) {
            const int rc = tiled ? tensor_map_for(s, A.in[k], (A.nm + 2 * D.g) / FK_GROUP, FK_GROUP * A.pitch_in, 1, FK_GROUP * 32, &maps.m[k])
                                 : tensor_map_for(s, A.in[k], A.nm + 2 * D.g, A.pitch_in, FK_GROUP, 32, &maps.m[k]);
            if (rc) return rc;
        }
    } else {
        memset(&maps, 0, sizeof(maps));
    }
    const int tr = A.transpose_out ? 1 : 0;
    // per-cycle diagnostics: the last sweep of the cycle accumulates the conservation sums itself when it can
    const bool cons = fast_launch && !strict4_launch && stg == STG_TMA && s->cons_m != nullptr &&
                      (tiled ? s->fast_tiled_cons_kernel[tr] : s->fast_cons_kernel[tr]) != nullptr;
    A.cons_m = cons ? s->cons_m : nullptr;
    A.cons_e = cons ? s->cons_e : nullptr;
    s->cons_done = cons;
    const long long cols_per_cta = fast_launch ? ASYNC_TPB : SWEEP_TPB;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    if (s->profile) {
        if (s->prof_used + 2 > s->prof_events.size()) {
            for (int k = 0; k < 2; k++) {
                cudaEvent_t e;
                ARMON_CUDA(cudaEventCreate(&e));
                s->prof_events.push_back(e);
            }
        }
        ev0 = s->prof_events[s->prof_used];
        ev1 = s->prof_events[s->prof_used + 1];
        s->prof_used += 2;
        ARMON_CUDA(cudaEventRecord(ev0, s->ctx->stream));
    }
    const bool with_fixup = strict4_launch && s->fixup_kernel;
    FixupArgs F;
    F.count = s->fix_count ? s->fix_count + (s->sweep_index & 1) : nullptr;
    F.count_next = s->fix_count ? s->fix_count + ((s->sweep_index + 1) & 1) : nullptr;
    F.list = s->fix_list;
    A.fix_count = F.count;
    A.fix_list = F.list;
    A.fix_cap = s->fix_cap;
    A.fix_rows = with_fixup ? FIX_CHUNKS * SWEEP_CHUNK : 0;
    // one launch over `ny` march segments y_base, y_base + y_jump, ...
    auto launch = [&](long long y_base, long long y_jump, long long ny, cudaStream_t st) -> int {
        A.y_base = (int)y_base;
        A.y_jump = (int)y_jump;
        // (tiled: the threads are shifted by the g ghost columns, sweep_fast_kernel.cuh 5.)
        const dim3 grid((unsigned)((A.nw + (tiled ? D.g : 0) + cols_per_cta - 1) / cols_per_cta), (unsigned)ny, 1);
        if (fast_launch)
            (strict4_launch ? s->strict4_kernel[A.dx_pow2 ? 1 : 0][tr]
             : tiled ? (cons ? s->fast_tiled_cons_kernel[tr] : s->fast_tiled_kernel[tr])
                     : (cons ? s->fast_cons_kernel[tr] : s->fast_kernel[stg][tr]))
                <<<grid, ASYNC_TPB, ASYNC_TPB / 32 * sizeof(FastWarpShared), st>>>(A, maps);
        else
            s->kernel<<<grid, SWEEP_TPB, 0, st>>>(A);
        ARMON_LAUNCH_CHECK(s->ctx);
        return ARMON_OK;
    };
    auto bc_fill = [&](cudaStream_t st) -> int {
        if (!(A.mirror_lo || A.mirror_hi)) return ARMON_OK;
        BcFillArgs B;
        B.rho = const_cast<double *>(A.in[0]); B.ua = const_cast<double *>(A.in[1]);
        B.ut = const_cast<double *>(A.in[2]); B.E = const_cast<double *>(A.in[3]);
        B.nm = A.nm; B.nw = A.nw; B.pitch = A.pitch_in; B.g = A.g; B.lo = A.mirror_lo; B.hi = A.mirror_hi;
        B.tiled = tiled ? 1 : 0;
        B.fa_lo = A.bc_a_lo; B.ft_lo = A.bc_t_lo; B.fa_hi = A.bc_a_hi; B.ft_hi = A.bc_t_hi;
        const dim3 bgrid((unsigned)((A.nw + TPB - 1) / TPB), (unsigned)A.g, 2);
        k_bc_fill<<<bgrid, TPB, 0, st>>>(B);
        ARMON_LAUNCH_CHECK(s->ctx);
        return ARMON_OK;
    };
    {
        NvtxRange r("EOS+fluxes+update+remap");
        if (overlap) {
            // interior segments on the compute stream, overlapping the exchange; the edge segments on their own stream
            // once the ghost rows are there (ev_halo also carries ev_state: the input state is complete), so that they
            // fill the last, partial wave of the interior launch instead of running after it
            if (int rc = launch(1, 1, n_interior, s->ctx->stream)) return rc;
            ARMON_CUDA(cudaStreamWaitEvent(s->edge_stream, s->ev_halo, 0));
            if (int rc = bc_fill(s->edge_stream)) return rc;
            if (int rc = launch(0, nseg - 1, 2, s->edge_stream)) return rc;
            if (short_tail)
                if (int rc = launch(nseg - 2, 1, 1, s->edge_stream)) return rc;
            ARMON_CUDA(cudaEventRecord(s->ev_edge, s->edge_stream));
            ARMON_CUDA(cudaStreamWaitEvent(s->ctx->stream, s->ev_edge, 0));
        } else {
            if (int rc = bc_fill(s->ctx->stream)) return rc;
            if (int rc = launch(0, 1, nseg, s->ctx->stream)) return rc;
        }
        if (with_fixup) {
            A.y_base = 0; A.y_jump = 1;
            s->fixup_kernel<<<2 * s->ctx->sm_count, 32, 0, s->ctx->stream>>>(A, F);
            ARMON_LAUNCH_CHECK(s->ctx);
            s->sweep_index++;
        }
    }
    if (s->profile) ARMON_CUDA(cudaEventRecord(ev1, s->ctx->stream));
    s->sweep_launches++;

    s->have_prev = true;
    s->prev_transposed = s->cur_transposed;
    s->prev_tiled = s->cur_tiled;
    s->cur = out_set;
    if (A.transpose_out) s->cur_transposed = !s->cur_transposed;
    return ARMON_OK;
}

// One axis sweep of every block of the group, in lock step.
int group_sweep(armon_group *G, int axis, double dt_factor, int acc_slot, int next_axis, bool last_of_cycle)
{
    NvtxRange r(axis == ARMON_AXIS_X ? "X" : "Y");
    for (armon_solver *b : G->blocks) {
        if (int rc = ensure_layout(b, axis, G->tiled == 1)) return rc;
        const bool diag = last_of_cycle && G->diag_ring != nullptr;
        b->cons_m = diag ? G->diag_rows + b->diag_base : nullptr;
        b->cons_e = diag ? G->diag_rows + G->diag_rows_cap + b->diag_base : nullptr;
        b->cons_done = false;
    }
    if (G->blocks.size() > 1) {
        NvtxRange r2("BC");
        for (armon_solver *b : G->blocks)
            if (int rc = local_halo(b, axis)) return rc;
    }
    for (armon_solver *b : G->blocks)
        if (int rc = launch_sweep(b, axis, dt_factor, acc_slot, next_axis)) return rc;
    return ARMON_OK;
}

bool multi_rank(const armon_group *G) { return G->ctx->comm && G->ctx->nranks > 1; }

// MPI_Iallreduce(MIN) of the local dt (src/utils.jl:126-134, src/solver_state.jl:107-111) becomes an all-reduce(max) of
// the CFL maxima of one accumulator slot (the min of the quotients is the quotient of the max) and of the slot's error
// flag, issued on the communication stream.  `wait`: the compute stream waits for it right away (initial time step);
// otherwise it is waited one cycle later (wait_allreduce), so that it overlaps the sweeps of the next cycle.
int allreduce_acc(armon_group *G, int slot, bool wait)
{
    if (multi_rank(G)) {
        armon_solver *s = G->blocks[0];
        if (int rc = comm_begin(s)) return rc;
        ARMON_NCCL(ncclAllReduce(&G->ts->acc[slot][0], &G->ts->acc[slot][0], 3, ncclUint64, ncclMax, G->ctx->comm,
                                 G->ctx->comm_stream));
        ARMON_CUDA(cudaEventRecord(G->ev_dt[slot], G->ctx->comm_stream));
        if (wait) ARMON_CUDA(cudaStreamWaitEvent(G->ctx->stream, G->ev_dt[slot], 0));
    }
    return ARMON_OK;
}

int wait_allreduce(armon_group *G, int slot)
{
    if (multi_rank(G)) ARMON_CUDA(cudaStreamWaitEvent(G->ctx->stream, G->ev_dt[slot], 0));
    return ARMON_OK;
}

int launch_cycle_step(armon_group *G, bool first, int read_slot, bool acc_is_xy)
{
    const armon_solver_desc &d = G->blocks[0]->d;
    CycleStepArgs a;
    a.first = first ? 1 : 0;
    a.read_slot = read_slot;
    a.acc_is_xy = acc_is_xy ? 1 : 0;
    a.dx = d.domain_size[0] / (double)d.global_nx;
    a.dy = d.domain_size[1] / (double)d.global_ny;
    a.cfl = d.cfl;
    a.maxtime = d.maxtime;
    a.maxcycle = d.maxcycle;
    a.cst_dt = d.cst_dt;
    a.Dt = d.Dt;
    k_cycle_step<<<1, 1, 0, G->ctx->stream>>>(G->ts, a);
    ARMON_LAUNCH_CHECK(G->ctx);
    return ARMON_OK;
}

int launch_init_dt(armon_solver *s)
{
    const armon_dims &D = s->d.dims;
    const long long n_rows = s->cur_transposed ? D.nx : D.ny, n_cols = s->cur_transposed ? D.ny : D.nx;
    const long long pitch = n_cols + 2 * D.g;
    double *const *b = s->buf[s->cur];
    const dim3 grid((unsigned)((n_cols + TPB - 1) / TPB), (unsigned)((n_rows + INIT_DT_ROWS - 1) / INIT_DT_ROWS), 1);
    if (s->d.tc.eos == ARMON_EOS_BIZARRIUM)
        k_init_dt<ARMON_EOS_BIZARRIUM><<<grid, TPB, 0, s->ctx->stream>>>(n_rows, n_cols, pitch, (int)D.g, b[0], b[1],
                                                                         b[2], b[3], s->d.tc.gamma, s->ts);
    else
        k_init_dt<ARMON_EOS_PERFECT_GAS><<<grid, TPB, 0, s->ctx->stream>>>(n_rows, n_cols, pitch, (int)D.g, b[0], b[1],
                                                                           b[2], b[3], s->d.tc.gamma, s->ts);
    ARMON_LAUNCH_CHECK(s->ctx);
    return ARMON_OK;
}

// Stale p, c, g from the other buffer set of a block (the input of its last sweep); see k_eos_pcg for `mode`.
int launch_eos_pcg(armon_solver *s, int mode)
{
    if (!(s->pcg[0] || s->pcg[1] || s->pcg[2]) || !s->have_prev) return ARMON_OK;
    const armon_dims &D = s->d.dims;
    double *const *b = s->buf[1 - s->cur];
    const long long gx = (D.nx + TPB - 1) / TPB;
    const dim3 grid((unsigned)(gx < 64 ? gx : 64), (unsigned)(D.ny < 64 ? D.ny : 64), 1);
    if (s->d.tc.eos == ARMON_EOS_BIZARRIUM)
        k_eos_pcg<ARMON_EOS_BIZARRIUM><<<grid, TPB, 0, s->ctx->stream>>>(
            D.nx, D.ny, (int)D.g, s->prev_transposed, s->prev_tiled, b[0], b[1], b[2], b[3], s->d.tc.gamma, s->pcg[0], s->pcg[1],
            s->pcg[2], s->ts, mode);
    else
        k_eos_pcg<ARMON_EOS_PERFECT_GAS><<<grid, TPB, 0, s->ctx->stream>>>(
            D.nx, D.ny, (int)D.g, s->prev_transposed, s->prev_tiled, b[0], b[1], b[2], b[3], s->d.tc.gamma, s->pcg[0], s->pcg[1],
            s->pcg[2], s->ts, mode);
    ARMON_LAUNCH_CHECK(s->ctx);
    return ARMON_OK;
}

int launch_diagnostics(armon_group *G)
{
    if (!G->diag_ring) return ARMON_OK;
    NvtxRange r("conservation_vars");
    const armon_solver_desc &d0 = G->blocks[0]->d;
    double *row_m = G->diag_rows, *row_e = G->diag_rows + G->diag_rows_cap;
    for (armon_solver *b : G->blocks) {
        if (b->cons_done) continue;   // the last sweep left one partial per warp in the block's region
        const armon_dims &D = b->d.dims;
        const long long n_rows = b->cur_transposed ? D.nx : D.ny, n_cols = b->cur_transposed ? D.ny : D.nx;
        const unsigned nblk = (unsigned)(n_rows < 4096 ? n_rows : 4096);
        k_diag_rows<<<nblk, TPB, 0, G->ctx->stream>>>(n_rows, n_cols, n_cols + 2 * D.g, (int)D.g, b->cur_tiled, b->buf[b->cur][0],
                                                      b->buf[b->cur][3], row_m + b->diag_base, row_e + b->diag_base);
        ARMON_LAUNCH_CHECK(G->ctx);
    }
    const double ds = (d0.domain_size[0] / (double)d0.global_nx) * (d0.domain_size[1] / (double)d0.global_ny);
    k_diag_final<<<1, TPB, 0, G->ctx->stream>>>(G->diag_rows_cap, row_m, row_e, ds, G->ts, G->diag_ring, G->diag_head,
                                                G->diag_cap);
    ARMON_LAUNCH_CHECK(G->ctx);
    return ARMON_OK;
}

// Band-tiled layout between the sweeps (sweep_fast_kernel.cuh, 5.): used when every block of the group can run it and,
// since the ghost rows travel between ranks as raw 4-row bands, when every rank of the communicator says the same
// (one all-reduce at the first cycle of the group's first run; every rank enqueues the same sequence of calls).
int decide_tiling(armon_group *G)
{
    if (G->tiled >= 0) return ARMON_OK;
    int ok = 1;
    for (const armon_solver *b : G->blocks) ok = ok && b->tiled_ok;
    if (multi_rank(G)) {
        if (!G->agree) ARMON_CUDA(cudaMalloc(&G->agree, sizeof(int)));
        cudaStream_t cs = G->ctx->comm_stream;
        ARMON_CUDA(cudaMemcpyAsync(G->agree, &ok, sizeof(int), cudaMemcpyHostToDevice, cs));
        ARMON_NCCL(ncclAllReduce(G->agree, G->agree, 1, ncclInt, ncclMin, G->ctx->comm, cs));
        ARMON_CUDA(cudaMemcpyAsync(&ok, G->agree, sizeof(int), cudaMemcpyDeviceToHost, cs));
        ARMON_CUDA(cudaStreamSynchronize(cs));
    }
    G->tiled = ok ? 1 : 0;
    if (getenv("ARMON_B200_VERBOSE")) fprintf(stderr, "[armon_b200] band-tiled layout: %s\n", G->tiled ? "on" : "off");
    return ARMON_OK;
}

int enqueue_cycle(armon_group *G)
{
    NvtxRange r("solver_cycle");
    const armon_solver_desc &d = G->blocks[0]->d;
    if (!G->started) {
        if (int rc = decide_tiling(G)) return rc;
        // cycle 0: EOS_init + first time step (src/solver.jl:291-297); its maxima go to slot 1 ("cycle -1")
        NvtxRange r2("time_step");
        for (armon_solver *b : G->blocks)
            if (int rc = launch_init_dt(b)) return rc;
        if (int rc = allreduce_acc(G, 1, true)) return rc;
        if (int rc = launch_cycle_step(G, true, 1, true)) return rc;
        G->started = true;
    }
    if (G->diag_ring)   // partial sums of this cycle's log line: unused slots must read 0
        ARMON_CUDA(cudaMemsetAsync(G->diag_rows, 0, (size_t)(2 * G->diag_rows_cap) * sizeof(double), G->ctx->stream));
    int axes[3], next_axes[3];
    double factors[3], next_factors[3];
    const long long k = G->host_cycle;
    const int n = split_axes(d.splitting, k, axes, factors);
    split_axes(d.splitting, k + 1, next_axes, next_factors);
    for (int i = 0; i < n; i++) {
        const bool last = i == n - 1;
        const int next_axis = last ? next_axes[0] : axes[i + 1];
        if (int rc = group_sweep(G, axes[i], factors[i], last ? (int)(k & 1) : 2, next_axis, last)) return rc;
    }
    {
        NvtxRange r2("time_step");
        // the last sweep ran along axes[n-1]: slot k & 1 = (march axis, transverse axis).  Its all-reduce overlaps cycle k+1.
        if (int rc = allreduce_acc(G, (int)(k & 1), false)) return rc;
        G->last_axis_is_x[k & 1] = axes[n - 1] == ARMON_AXIS_X;
        // next_cycle! of cycle k; the time step D_k it installs comes from the maxima of cycle k-1 (none after cycle 0)
        if (k >= 1) {
            const int rs = (int)((k - 1) & 1);
            if (int rc = wait_allreduce(G, rs)) return rc;
            if (int rc = launch_cycle_step(G, false, rs, G->last_axis_is_x[rs])) return rc;
        } else {
            if (int rc = launch_cycle_step(G, false, -1, true)) return rc;
        }
    }
    // the stale p, c, g of the reference must be saved before a cycle enqueued past the end overwrites their source
    for (armon_solver *b : G->blocks)
        if (int rc = launch_eos_pcg(b, 1)) return rc;
    if (int rc = launch_diagnostics(G)) return rc;
    G->host_cycle++;
    return ARMON_OK;
}

bool graph_wanted(const armon_group *G)
{
    int mode = G->graph_mode;
    if (const char *env = getenv("ARMON_B200_GRAPH")) mode = atoi(env) ? 1 : 2;
    if (mode == 2 || multi_rank(G)) return false;
    for (const armon_solver *b : G->blocks)
        if (b->profile) return false;
    if (mode == 1) return true;
    for (const armon_solver *b : G->blocks)
        if (b->d.dims.nx * b->d.dims.ny > 512LL * 512LL) return false;
    return true;
}

// Enqueue n cycles.  From cycle 2 on, pairs of cycles starting at an even index are identical launch sequences (buffer
// rotation, layout, accumulator slots and the alternating splittings all have period 2): the first such pair is
// captured into a CUDA graph while it is enqueued, the following ones replay it -- one launch instead of 10-20 for the
// launch-bound grids (the 100x100 cases of test/reference_data).
int enqueue_cycles(armon_group *G, long long n)
{
    const bool use_graph = graph_wanted(G);
    cudaStream_t st = G->ctx->stream;
    while (n > 0) {
        if (use_graph && n >= 2 && G->host_cycle >= 2 && (G->host_cycle & 1) == 0) {
            if (!G->graph_exec) {
                const uint64_t l0 = G->ctx->launches, s0 = G->blocks[0]->sweep_launches;
                ARMON_CUDA(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
                int rc = enqueue_cycle(G);
                if (!rc) rc = enqueue_cycle(G);
                cudaGraph_t graph = nullptr;
                const cudaError_t ce = cudaStreamEndCapture(st, &graph);
                if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
                ARMON_CUDA(ce);
                const cudaError_t ie = cudaGraphInstantiate(&G->graph_exec, graph, 0);
                cudaGraphDestroy(graph);
                ARMON_CUDA(ie);
                G->graph_launches = G->ctx->launches - l0;
                G->graph_sweeps = G->blocks[0]->sweep_launches - s0;
                ARMON_CUDA(cudaGraphLaunch(G->graph_exec, st));   // the capture recorded the pair, this runs it
            } else {
                ARMON_CUDA(cudaGraphLaunch(G->graph_exec, st));
                G->ctx->launches += G->graph_launches;
                for (armon_solver *b : G->blocks) b->sweep_launches += G->graph_sweeps;
                G->host_cycle += 2;
            }
            G->graph_replays++;
            n -= 2;
        } else {
            if (int rc = enqueue_cycle(G)) return rc;
            n -= 1;
        }
    }
    return ARMON_OK;
}

int read_state(armon_group *G, armon_time_state *out)
{
    DeviceTimeState *h = reinterpret_cast<DeviceTimeState *>(G->ctx->pinned);
    ARMON_CUDA(cudaMemcpyAsync(h, G->ts, sizeof(DeviceTimeState), cudaMemcpyDeviceToHost, G->ctx->stream));
    ARMON_CUDA(cudaStreamSynchronize(G->ctx->stream));
    out->cycle = h->cycle;
    out->time = h->time;
    out->current_dt = h->current_dt;
    out->next_cycle_dt = h->next_cycle_dt;
    out->error = h->error;
    out->done = h->done;
    out->error_cycle = h->error_cycle;
    if (getenv("ARMON_B200_VERBOSE"))   // strict mode: column chunks handed to the IEEE fix-up so far
        fprintf(stderr, "[armon_b200] cycle %lld: %u column chunks recomputed by the IEEE fix-up\n", h->cycle, h->redo_count);
    return ARMON_OK;
}

int group_reset(armon_group *G)
{
    // the last cycle's all-reduce (communication stream) and edge launches must not outlive the state they touch
    ARMON_CUDA(cudaStreamSynchronize(G->ctx->comm_stream));
    for (armon_solver *b : G->blocks) ARMON_CUDA(cudaStreamSynchronize(b->edge_stream));
    const armon_solver_desc &d = G->blocks[0]->d;
    k_ts_reset<<<1, 1, 0, G->ctx->stream>>>(G->ts, d.cst_dt, d.Dt);
    ARMON_LAUNCH_CHECK(G->ctx);
    if (G->diag_head) ARMON_CUDA(cudaMemsetAsync(G->diag_head, 0, sizeof(unsigned long long), G->ctx->stream));
    G->diag_read = 0;
    G->host_cycle = 0;
    G->started = false;
    G->timed = false;
    drop_graph(G);
    for (armon_solver *b : G->blocks) {
        b->cur = 0;
        b->cur_transposed = false;
        b->cur_tiled = false;
        b->have_prev = false;
    }
    return ARMON_OK;
}

int group_run(armon_group *G, int64_t n_cycles)
{
    ARMON_CHECK_ARG(n_cycles >= 0, "negative cycle count");
    ARMON_CUDA(cudaEventRecord(G->ev_start, G->ctx->stream));
    if (int rc = enqueue_cycles(G, n_cycles)) return rc;
    ARMON_CUDA(cudaEventRecord(G->ev_stop, G->ctx->stream));
    G->timed = true;
    return ARMON_OK;
}

int group_time_loop(armon_group *G)
{
    const armon_solver_desc &d = G->blocks[0]->d;
    ARMON_CUDA(cudaEventRecord(G->ev_start, G->ctx->stream));
    armon_time_state st;
    if (d.maxcycle <= 0 || !(0.0 < d.maxtime)) {   // `while time < maxtime && cycle < maxcycle` never entered
        ARMON_CUDA(cudaEventRecord(G->ev_stop, G->ctx->stream));
        G->timed = true;
        return ARMON_OK;
    }
    if (int rc = enqueue_cycle(G)) return rc;
    for (;;) {
        if (int rc = read_state(G, &st)) return rc;
        if (st.error == ARMON_ERR_RANGE) {
            armon_set_error("cycle %lld: the work list of the strict mode's IEEE fix-up overflowed (division operands "
                            "outside the proven range all over the domain); rerun with math_mode ieee",
                            (long long)st.error_cycle);
            return ARMON_ERR_RANGE;
        }
        if (st.error) {
            armon_set_error("Invalid time step for cycle %lld", (long long)st.error_cycle);
            return ARMON_ERR_TIME;
        }
        if (st.done) break;
        // Lower bound of the cycles still to run: the time step grows by at most 5% per cycle
        // (src/solver_state.jl:127-130), so n cycles advance the time by at most dt*(1.05^n - 1)/0.05.
        long long batch = 1;
        const double remaining = d.maxtime - st.time;
        if (st.current_dt > 0.0 && remaining > 0.0) {
            const double n = d.cst_dt ? remaining / st.current_dt
                                      : std::log1p(0.05 * remaining / st.current_dt) / std::log(1.05);
            batch = (long long)std::floor(n) - 1;
        }
        const long long left = d.maxcycle - st.cycle;
        if (batch > left) batch = left;
        if (batch > 4096) batch = 4096;
        if (G->diag_cap && batch > (long long)G->diag_cap) batch = G->diag_cap;
        if (batch < 1) batch = 1;
        if (int rc = enqueue_cycles(G, batch)) return rc;
    }
    ARMON_CUDA(cudaEventRecord(G->ev_stop, G->ctx->stream));
    G->timed = true;
    return ARMON_OK;
}

int solver_finalize(armon_solver *s)
{
    // 1. stale p, c, g from the state at the start of the last sweep (still intact in the other buffer set unless cycles
    //    were enqueued past the end of the run, in which case they were saved when the run ended)
    if (int rc = launch_eos_pcg(s, 2)) return rc;
    s->have_prev = false;
    // 2. canonical layout, in main_vars
    if (int rc = relayout_current(s, false, false)) return rc;
    if (s->cur != 0) {
        for (int k = 0; k < 4; k++)
            ARMON_CUDA(cudaMemcpyAsync(s->buf[0][k], s->buf[1][k], (size_t)n_elems(s) * sizeof(double),
                                       cudaMemcpyDeviceToDevice, s->ctx->stream));
        s->cur = 0;
    }
    return ARMON_OK;
}

int group_finalize(armon_group *G)
{
    drop_graph(G);   // the rotation state changes: a captured pair no longer applies
    for (armon_solver *b : G->blocks)
        if (int rc = solver_finalize(b)) return rc;
    return ARMON_OK;
}

int group_elapsed_ms(armon_group *G, float *ms)
{
    ARMON_CHECK_ARG(ms != nullptr, "null result");
    ARMON_CHECK_ARG(G->timed, "no run / time_loop call to time");
    ARMON_CUDA(cudaEventSynchronize(G->ev_stop));
    ARMON_CUDA(cudaEventElapsedTime(ms, G->ev_start, G->ev_stop));
    return ARMON_OK;
}

int group_diagnostics(armon_group *G, int32_t capacity)
{
    ARMON_CHECK_ARG(capacity >= 0 && capacity <= (1 << 20), "ring capacity");
    ARMON_CUDA(cudaStreamSynchronize(G->ctx->stream));
    drop_graph(G);
    if (G->diag_ring) { cudaFree(G->diag_ring); G->diag_ring = nullptr; }
    if (G->diag_head) { cudaFree(G->diag_head); G->diag_head = nullptr; }
    if (G->diag_rows) { cudaFree(G->diag_rows); G->diag_rows = nullptr; }
    G->diag_cap = 0;
    G->diag_read = 0;
    if (capacity == 0) return ARMON_OK;
    // per block: one partial per array row (k_diag_rows) or per warp of the last sweep (fused), whichever is larger
    long long rows = 0;
    for (armon_solver *b : G->blocks) {
        const armon_dims &D = b->d.dims;
        long long cap = D.nx > D.ny ? D.nx : D.ny;
        for (int axis = 0; axis < 2; axis++) {
            const long long nm = axis == ARMON_AXIS_X ? D.nx : D.ny, nw = axis == ARMON_AXIS_X ? D.ny : D.nx;
            const long long seg = pick_segment(b, nm, nw);
            const long long parts = ((nm + seg - 1) / seg) * ((nw + D.g + ASYNC_TPB - 1) / ASYNC_TPB) * (ASYNC_TPB / 32);
            cap = parts > cap ? parts : cap;
        }
        b->diag_base = rows;
        b->diag_cap = cap;
        rows += cap;
    }
    ARMON_CUDA(cudaMalloc(&G->diag_ring, (size_t)capacity * sizeof(armon_cycle_diag)));
    ARMON_CUDA(cudaMalloc(&G->diag_head, sizeof(unsigned long long)));
    ARMON_CUDA(cudaMalloc(&G->diag_rows, (size_t)(2 * rows) * sizeof(double)));
    ARMON_CUDA(cudaMemsetAsync(G->diag_head, 0, sizeof(unsigned long long), G->ctx->stream));
    G->diag_rows_cap = rows;
    G->diag_cap = (unsigned)capacity;
    return ARMON_OK;
}

int group_read_diagnostics(armon_group *G, armon_cycle_diag *lines, int64_t max_lines, int64_t *n_lines)
{
    ARMON_CHECK_ARG(lines && n_lines && max_lines >= 0, "null result");
    *n_lines = 0;
    ARMON_CHECK_ARG(G->diag_ring != nullptr, "diagnostics are not enabled");
    unsigned long long head = 0;
    ARMON_CUDA(cudaMemcpyAsync(&head, G->diag_head, sizeof(head), cudaMemcpyDeviceToHost, G->ctx->stream));
    ARMON_CUDA(cudaStreamSynchronize(G->ctx->stream));
    unsigned long long first = G->diag_read;
    if (head - first > G->diag_cap) first = head - G->diag_cap;   // the oldest lines were overwritten
    int64_t n = 0;
    for (unsigned long long k = first; k < head && n < max_lines; k++, n++)
        ARMON_CUDA(cudaMemcpyAsync(&lines[n], &G->diag_ring[k % G->diag_cap], sizeof(armon_cycle_diag),
                                   cudaMemcpyDeviceToHost, G->ctx->stream));
    ARMON_CUDA(cudaStreamSynchronize(G->ctx->stream));
    G->diag_read = first + (unsigned long long)n;
    *n_lines = n;
    return ARMON_OK;
}

int group_alloc_common(armon_group *G)
{
    ARMON_CUDA(cudaEventCreate(&G->ev_start));
    ARMON_CUDA(cudaEventCreate(&G->ev_stop));
    ARMON_CUDA(cudaEventCreateWithFlags(&G->ev_dt[0], cudaEventDisableTiming));
    ARMON_CUDA(cudaEventCreateWithFlags(&G->ev_dt[1], cudaEventDisableTiming));
    return ARMON_OK;
}

void group_free_common(armon_group *G)
{
    drop_graph(G);
    if (G->ev_start) cudaEventDestroy(G->ev_start);
    if (G->ev_stop) cudaEventDestroy(G->ev_stop);
    for (int k = 0; k < 2; k++) if (G->ev_dt[k]) cudaEventDestroy(G->ev_dt[k]);
    if (G->diag_ring) cudaFree(G->diag_ring);
    if (G->diag_head) cudaFree(G->diag_head);
    if (G->diag_rows) cudaFree(G->diag_rows);
    if (G->agree) cudaFree(G->agree);
}

// Select the marching kernels of a solver from its descriptor.
int select_kernels(armon_solver *s)
{
    const armon_solver_desc *desc = &s->d;
    const armon_dims &D = desc->dims;
    const int rl = desc->riemann == ARMON_RIEMANN_GODUNOV ? 0 : 1 + desc->limiter;
    const bool biz = desc->tc.eos == ARMON_EOS_BIZARRIUM;
    if (desc->math_mode == ARMON_MATH_STRICT)
        s->kernel = biz ? sweep_table_strict_biz(rl, desc->projection) : sweep_table_strict_pg(rl, desc->projection);
    else if (desc->math_mode == ARMON_MATH_IEEE)
        s->kernel = biz ? sweep_table_ieee_biz(rl, desc->projection) : sweep_table_ieee_pg(rl, desc->projection);
    else
        s->kernel = biz ? sweep_table_fast_biz(rl, desc->projection) : sweep_table_fast_pg(rl, desc->projection);
    if (!s->kernel) {
        armon_set_error("no sweep kernel for this scheme combination");
        return ARMON_ERR_INVALID;
    }
    // kernel_variant / ARMON_B200_KERNEL: auto | single | async | async2 | tma (include/armon_b200.h,
    // ARMON_KERNEL_*).  auto: fast mode -> the explicit-arithmetic kernel with TMA staging (cp.async for odd pitches);
    // strict mode -> `async`: the strict arithmetic on the fast kernel's schedule + IEEE fix-up (see below); ieee -> `single`.
    int variant = desc->kernel_variant;
    if (const char *env = getenv("ARMON_B200_KERNEL")) {
        const std::string e(env);
        variant = e == "single" ? ARMON_KERNEL_SINGLE : e == "async" ? ARMON_KERNEL_ASYNC
                : e == "async2" ? ARMON_KERNEL_ASYNC2 : e == "tma" ? ARMON_KERNEL_TMA : ARMON_KERNEL_AUTO;
    }
    if (variant == ARMON_KERNEL_AUTO)
        variant = desc->math_mode == ARMON_MATH_FAST ? ARMON_KERNEL_TMA
                : desc->math_mode == ARMON_MATH_STRICT ? ARMON_KERNEL_ASYNC : ARMON_KERNEL_SINGLE;
    const char *cv = getenv("ARMON_B200_CARVEOUT");   // percent of shared memory, -1 = driver default
    if (desc->math_mode == ARMON_MATH_FAST && (variant == ARMON_KERNEL_TMA || variant == ARMON_KERNEL_ASYNC2)) {
        s->fast_stg = variant == ARMON_KERNEL_TMA ? STG_TMA : STG_CPA16;
        bool ok = true;
        for (int stg = 0; stg < 3; stg++)
            for (int tr = 0; tr < 2; tr++) {
                sweep_fast_fn_t fn = nullptr;
                if (stg == STG_TMA) fn = biz ? sweep_fast_table_tma_biz(rl, desc->projection, tr) : sweep_fast_table_tma_pg(rl, desc->projection, tr);
                else if (stg == STG_CPA16) fn = biz ? sweep_fast_table_cpa16_biz(rl, desc->projection, tr) : sweep_fast_table_cpa16_pg(rl, desc->projection, tr);
                else fn = biz ? sweep_fast_table_cpa8_biz(rl, desc->projection, tr) : sweep_fast_table_cpa8_pg(rl, desc->projection, tr);
                s->fast_kernel[stg][tr] = fn;
                ok = ok && fn != nullptr;
                if (!fn) continue;
                ARMON_CUDA(cudaFuncSetAttribute((const void *)fn, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                (int)(ASYNC_TPB / 32 * sizeof(FastWarpShared))));
                // everything is staged through shared memory, L1 is of no use: all of the unified array to shared memory
                ARMON_CUDA(cudaFuncSetAttribute((const void *)fn, cudaFuncAttributePreferredSharedMemoryCarveout,
                                                cv ? atoi(cv) : (int)cudaSharedmemCarveoutMaxShared));
                if (getenv("ARMON_B200_VERBOSE")) {
                    int nb = 0;
                    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, (const void *)fn, ASYNC_TPB,
                                                                  ASYNC_TPB / 32 * sizeof(FastWarpShared));
                    fprintf(stderr, "[armon_b200] fast kernel stg=%d tr=%d: %d resident CTAs/SM, %zu B shared/CTA\n", stg,
                            tr, nb, ASYNC_TPB / 32 * sizeof(FastWarpShared));
                }
            }
        s->use_fast = ok;
        // band-tiled layout: extents multiples of 8 (whole tiles and bands in both orientations), TMA staging;
        // ARMON_B200_TILED=0 keeps the row-major layouts (comparison runs, tests)
        const char *tl = getenv("ARMON_B200_TILED");
        bool tiled = ok && s->fast_stg == STG_TMA && D.nx % 8 == 0 && D.ny % 8 == 0 && !(tl && atoi(tl) == 0);
        for (int tr = 0; tr < 2 && ok; tr++) {
            sweep_fast_fn_t fns[3];
            fns[0] = biz ? sweep_fast_table_tma_cons_biz(rl, desc->projection, tr) : sweep_fast_table_tma_cons_pg(rl, desc->projection, tr);
            fns[1] = biz ? sweep_fast_table_tiled_biz(rl, desc->projection, tr) : sweep_fast_table_tiled_pg(rl, desc->projection, tr);
            fns[2] = biz ? sweep_fast_table_tiled_cons_biz(rl, desc->projection, tr) : sweep_fast_table_tiled_cons_pg(rl, desc->projection, tr);
            s->fast_cons_kernel[tr] = fns[0];
            s->fast_tiled_kernel[tr] = fns[1];
            s->fast_tiled_cons_kernel[tr] = fns[2];
            tiled = tiled && fns[1] != nullptr;
            for (sweep_fast_fn_t fn : fns) {
                if (!fn) continue;
                ARMON_CUDA(cudaFuncSetAttribute((const void *)fn, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                (int)(ASYNC_TPB / 32 * sizeof(FastWarpShared))));
                ARMON_CUDA(cudaFuncSetAttribute((const void *)fn, cudaFuncAttributePreferredSharedMemoryCarveout,
                                                cv ? atoi(cv) : (int)cudaSharedmemCarveoutMaxShared));
            }
        }
        s->tiled_ok = tiled;
    }
    // math_mode strict, variant async (the default): the strict arithmetic on the four-chain schedule of the fast kernel
    // (sweep_fast_kernel<..., MATH_STRICT>, TMA staging: even pitches; odd pitches take s->kernel) + the IEEE fix-up of
    // the chunks whose division operands left the proven range.  ARMON_B200_STRICT=single keeps the register-prefetch
    // kernel everywhere (comparison runs, tests).
    if (variant == ARMON_KERNEL_ASYNC && desc->math_mode == ARMON_MATH_STRICT) {
        const char *sk = getenv("ARMON_B200_STRICT");
        const bool want = !(sk && std::string(sk) == "single");
        bool ok = want;
        for (int k = 0; k < 4 && want; k++) {
            const int dxp = k >> 1, tr = k & 1;
            sweep_fast_fn_t fn = dxp ? (biz ? sweep_fast_table_strict_dxp_biz(rl, desc->projection, tr) : sweep_fast_table_strict_dxp_pg(rl, desc->projection, tr))
                                     : (biz ? sweep_fast_table_strict_biz(rl, desc->projection, tr) : sweep_fast_table_strict_pg(rl, desc->projection, tr));
            s->strict4_kernel[dxp][tr] = fn;
            ok = ok && fn != nullptr;
            if (!fn) continue;
            ARMON_CUDA(cudaFuncSetAttribute((const void *)fn, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            (int)(ASYNC_TPB / 32 * sizeof(FastWarpShared))));
            ARMON_CUDA(cudaFuncSetAttribute((const void *)fn, cudaFuncAttributePreferredSharedMemoryCarveout,
                                            cv ? atoi(cv) : (int)cudaSharedmemCarveoutMaxShared));
        }
        s->use_strict4 = ok;
    }
    if (s->use_strict4) {
        s->fixup_kernel = biz ? sweep_fixup_table_biz(rl, desc->projection) : sweep_fixup_table_pg(rl, desc->projection);
        // work list of the IEEE fix-up: one entry per (column, chunk of 8 rows) at most; capped at 32 MB, an overflow (a
        // domain full of out-of-range values) raises ARMON_ERR_RANGE
        const long long lo = D.nx > D.ny ? D.ny : D.nx, hi = D.nx > D.ny ? D.nx : D.ny;
        const long long worst = hi * ((lo + SWEEP_CHUNK - 1) / SWEEP_CHUNK + 1);
        long long cap = worst < (4LL << 20) ? worst : (4LL << 20);
        if (cap < 1024) cap = 1024;
        s->fix_cap = (unsigned)cap;
        ARMON_CUDA(cudaMalloc(&s->fix_count, 2 * sizeof(unsigned)));
        ARMON_CUDA(cudaMemsetAsync(s->fix_count, 0, 2 * sizeof(unsigned), s->ctx->stream));
        ARMON_CUDA(cudaMalloc(&s->fix_list, (size_t)cap * sizeof(unsigned long long)));
    }
    return ARMON_OK;
}

}   // namespace

extern "C" {

int armon_solver_create(armon_ctx *ctx, const armon_solver_desc *desc, armon_solver **out)
{
    ARMON_CHECK_ARG(ctx && desc && out, "null argument");
    *out = nullptr;
    if (int rc = armon_ctx_activate(ctx)) return rc;
    const armon_dims &D = desc->dims;
    ARMON_CHECK_ARG(D.nx > 0 && D.ny > 0, "empty sub-domain");
    ARMON_CHECK_ARG(D.g == 4, "the fused sweep needs nghost == 4 (dependency cone of GAD + euler_2nd, SURVEY.md 8a)");
    ARMON_CHECK_ARG(D.nx >= D.g && D.ny >= D.g, "sub-domain smaller than the ghost width (src/parameters.jl:684-690)");
    ARMON_CHECK_ARG(desc->riemann == ARMON_RIEMANN_GODUNOV || desc->riemann == ARMON_RIEMANN_GAD, "riemann scheme");
    ARMON_CHECK_ARG(desc->limiter >= 0 && desc->limiter <= 2, "limiter");
    ARMON_CHECK_ARG(desc->projection == ARMON_PROJ_EULER || desc->projection == ARMON_PROJ_EULER_2ND, "projection");
    ARMON_CHECK_ARG(desc->splitting >= 0 && desc->splitting <= 4, "axis splitting");
    ARMON_CHECK_ARG(desc->tc.eos == ARMON_EOS_PERFECT_GAS || desc->tc.eos == ARMON_EOS_BIZARRIUM, "EOS");
    ARMON_CHECK_ARG(desc->math_mode == ARMON_MATH_STRICT || desc->math_mode == ARMON_MATH_FAST ||
                    desc->math_mode == ARMON_MATH_IEEE, "math mode");
    ARMON_CHECK_ARG(desc->kernel_variant == ARMON_KERNEL_AUTO || desc->kernel_variant == ARMON_KERNEL_SINGLE ||
                    desc->kernel_variant == ARMON_KERNEL_ASYNC || desc->kernel_variant == ARMON_KERNEL_ASYNC2 ||
                    desc->kernel_variant == ARMON_KERNEL_TMA,
                    "kernel variant");
    ARMON_CHECK_ARG(desc->cuda_graph >= 0 && desc->cuda_graph <= 2, "cuda_graph");
    ARMON_CHECK_ARG(!desc->cst_dt || desc->Dt != 0.0, "Dt == 0 with constant step enabled");

    armon_solver *s = new armon_solver();
    s->ctx = ctx;
    s->d = *desc;
    if (const char *ov = getenv("ARMON_B200_OVERLAP")) s->overlap = atoi(ov) != 0;
    armon_group *G = new armon_group();
    G->ctx = ctx;
    G->blocks.push_back(s);
    G->graph_mode = desc->cuda_graph;
    s->self = s->group = G;
    auto fail = [&](int rc) { armon_solver_destroy(s); return rc; };
    if (int rc = select_kernels(s)) return fail(rc);
    if (cudaMalloc(&s->own_ts, sizeof(DeviceTimeState)) != cudaSuccess) {
        armon_set_error("cudaMalloc of the time-step state failed");
        return fail(ARMON_ERR_CUDA);
    }
    s->ts = G->ts = s->own_ts;
    if (int rc = group_alloc_common(G)) return fail(rc);
    if (cudaEventCreateWithFlags(&s->ev_state, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&s->ev_halo, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&s->ev_edge, cudaEventDisableTiming) != cudaSuccess ||
        cudaStreamCreateWithFlags(&s->edge_stream, cudaStreamNonBlocking) != cudaSuccess) {
        armon_set_error("event / stream creation failed: %s", cudaGetErrorString(cudaGetLastError()));
        return fail(ARMON_ERR_CUDA);
    }
    k_ts_reset<<<1, 1, 0, ctx->stream>>>(s->ts, desc->cst_dt, desc->Dt);
    ctx->launches++;
    if (cudaGetLastError() != cudaSuccess) {
        armon_set_error("k_ts_reset launch failed");
        return fail(ARMON_ERR_CUDA);
    }
    *out = s;
    return ARMON_OK;
}

int armon_solver_destroy(armon_solver *s)
{
    if (!s) return ARMON_OK;
    cudaSetDevice(s->ctx->device);
    // nothing may still reference the streams, events and buffers released below
    cudaStreamSynchronize(s->ctx->comm_stream);
    if (s->edge_stream) cudaStreamSynchronize(s->edge_stream);
    cudaStreamSynchronize(s->ctx->stream);
    if (s->group && s->group != s->self) {   // leave the block group: it cannot run without this block
        armon_group *G = s->group;
        for (armon_solver *b : G->blocks) {
            b->group = b->self;
            b->ts = b->own_ts;
            for (int k = 0; k < 4; k++) b->local_nb[k] = nullptr;
        }
        G->blocks.clear();
    }
    if (s->self) {
        group_free_common(s->self);
        delete s->self;
    }
    if (s->own_ts) cudaFree(s->own_ts);
    if (s->fix_count) cudaFree(s->fix_count);
    if (s->fix_list) cudaFree(s->fix_list);
    if (s->ev_state) cudaEventDestroy(s->ev_state);
    if (s->ev_halo) cudaEventDestroy(s->ev_halo);
    if (s->ev_edge) cudaEventDestroy(s->ev_edge);
    if (s->edge_stream) cudaStreamDestroy(s->edge_stream);
    for (cudaEvent_t e : s->prof_events) cudaEventDestroy(e);
    delete s;
    return ARMON_OK;
}

int armon_solver_bind(armon_solver *s, double *const main_vars[4], double *const work_vars[4], double *const pcg[3])
{
    if (int rc = solver_check(s, false, false)) return rc;
    ARMON_CHECK_ARG(main_vars && work_vars, "null array lists");
    for (int k = 0; k < 4; k++) {
        ARMON_CHECK_ARG(main_vars[k] && work_vars[k], "null device array");
        s->buf[0][k] = main_vars[k];
        s->buf[1][k] = work_vars[k];
    }
    for (int k = 0; k < 3; k++) s->pcg[k] = pcg ? pcg[k] : nullptr;
    s->bound = true;
    s->cur = 0;
    s->cur_transposed = false;
    s->cur_tiled = false;
    s->have_prev = false;
    s->tmaps.clear();
    drop_graph(s->group);
    return ARMON_OK;
}

int armon_solver_reset(armon_solver *s)
{
    if (int rc = solver_check(s)) return rc;
    return group_reset(s->group);
}

static int solver_init_fields(armon_solver *s)
{
    const armon_solver_desc &d = s->d;
    return armon_init_test(s->ctx, d.dims, d.origin_ix, d.origin_iy, d.global_nx, d.global_ny, d.domain_size, d.origin,
                           &d.tc, nullptr, nullptr, nullptr, s->buf[0][0], s->buf[0][3], s->buf[0][1], s->buf[0][2],
                           s->pcg[0], s->pcg[1], s->pcg[2], nullptr, nullptr, nullptr, nullptr, nullptr, nullptr);
}

int armon_solver_init(armon_solver *s)
{
    if (int rc = solver_check(s)) return rc;
    if (int rc = solver_init_fields(s)) return rc;
    return group_reset(s->group);
}

int armon_solver_run(armon_solver *s, int64_t n_cycles)
{
    if (int rc = solver_check(s)) return rc;
    return group_run(s->group, n_cycles);
}

int armon_solver_state(armon_solver *s, armon_time_state *out)
{
    if (int rc = solver_check(s, false)) return rc;
    ARMON_CHECK_ARG(out != nullptr, "null state");
    return read_state(s->group, out);
}

int armon_solver_time_loop(armon_solver *s)
{
    if (int rc = solver_check(s)) return rc;
    return group_time_loop(s->group);
}

int armon_solver_finalize(armon_solver *s)
{
    if (int rc = solver_check(s)) return rc;
    return group_finalize(s->group);
}

int armon_solver_halo_exchange(armon_solver *s, int axis)
{
    if (int rc = solver_check(s)) return rc;
    ARMON_CHECK_ARG(axis == ARMON_AXIS_X || axis == ARMON_AXIS_Y, "axis");
    if (int rc = ensure_layout(s, axis, false)) return rc;
    if (int rc = comm_begin(s)) return rc;
    if (int rc = halo_exchange(s, axis, s->ctx->comm_stream)) return rc;
    return comm_end(s, true);
}

int armon_solver_elapsed_ms(armon_solver *s, float *ms)
{
    if (int rc = solver_check(s, false)) return rc;
    return group_elapsed_ms(s->group, ms);
}

int armon_solver_profile(armon_solver *s, int enable)
{
    if (int rc = solver_check(s, false, false)) return rc;
    s->profile = enable != 0;
    s->prof_used = 0;
    if (s->profile) drop_graph(s->group);   // profiled sweeps are enqueued one by one
    return ARMON_OK;
}

int armon_solver_sweep_time_ms(armon_solver *s, double *total_ms, uint64_t *count)
{
    if (int rc = solver_check(s, false, false)) return rc;
    ARMON_CHECK_ARG(total_ms && count, "null result");
    ARMON_CUDA(cudaStreamSynchronize(s->ctx->stream));
    double total = 0.0;
    for (size_t k = 0; k + 1 < s->prof_used; k += 2) {
        float ms = 0.f;
        ARMON_CUDA(cudaEventElapsedTime(&ms, s->prof_events[k], s->prof_events[k + 1]));
        total += ms;
    }
    *total_ms = total;
    *count = s->prof_used / 2;
    return ARMON_OK;
}

int armon_solver_sweep_launches(armon_solver *s, uint64_t *count)
{
    if (int rc = solver_check(s, false, false)) return rc;
    ARMON_CHECK_ARG(count != nullptr, "null result");
    *count = s->sweep_launches;
    return ARMON_OK;
}

int armon_solver_tiled(armon_solver *s, int32_t *tiled)
{
    if (int rc = solver_check(s, false, false)) return rc;
    ARMON_CHECK_ARG(tiled != nullptr, "null result");
    *tiled = s->group->tiled;
    return ARMON_OK;
}

int armon_solver_strict_chains(armon_solver *s, int32_t *chains)
{
    if (int rc = solver_check(s, false, false)) return rc;
    ARMON_CHECK_ARG(chains != nullptr, "null result");
    *chains = s->use_strict4 ? 1 : 0;
    return ARMON_OK;
}

int armon_solver_diagnostics(armon_solver *s, int32_t capacity)
{
    if (int rc = solver_check(s)) return rc;
    return group_diagnostics(s->group, capacity);
}

int armon_solver_read_diagnostics(armon_solver *s, armon_cycle_diag *lines, int64_t max_lines, int64_t *n_lines)
{
    if (int rc = solver_check(s)) return rc;
    return group_read_diagnostics(s->group, lines, max_lines, n_lines);
}

// ---- block groups ----------------------------------------------------------------------------------------------
int armon_group_create(armon_ctx *ctx, int32_t nbx, int32_t nby, armon_solver *const blocks[], armon_group **out)
{
    ARMON_CHECK_ARG(ctx && blocks && out, "null argument");
    *out = nullptr;
    ARMON_CHECK_ARG(nbx >= 1 && nby >= 1 && (long long)nbx * nby <= 4096, "block grid");
    if (int rc = armon_ctx_activate(ctx)) return rc;
    const int n = nbx * nby;
    const armon_solver_desc &d0 = blocks[0] ? blocks[0]->d : armon_solver_desc{};
    for (int i = 0; i < n; i++) {
        armon_solver *b = blocks[i];
        ARMON_CHECK_ARG(b != nullptr && b->ctx == ctx, "blocks must be solvers of this context");
        ARMON_CHECK_ARG(b->bound, "armon_solver_bind was not called on a block");
        ARMON_CHECK_ARG(b->group == b->self, "a block already belongs to a group");
        for (int j = 0; j < i; j++) ARMON_CHECK_ARG(blocks[j] != b, "a block appears twice");
        const armon_solver_desc &d = b->d;
        ARMON_CHECK_ARG(d.global_nx == d0.global_nx && d.global_ny == d0.global_ny && d.riemann == d0.riemann &&
                        d.limiter == d0.limiter && d.projection == d0.projection && d.splitting == d0.splitting &&
                        d.cfl == d0.cfl && d.maxtime == d0.maxtime && d.maxcycle == d0.maxcycle &&
                        d.cst_dt == d0.cst_dt && d.Dt == d0.Dt && d.math_mode == d0.math_mode &&
                        d.tc.test == d0.tc.test && d.dims.g == d0.dims.g,
                        "the blocks of a group must share the run parameters");
        const int bx = i % nbx, by = i / nbx;
        // a Cartesian grid of blocks: same height along a row of blocks, same width along a column, contiguous origins
        ARMON_CHECK_ARG(d.dims.ny == blocks[by * nbx]->d.dims.ny && d.dims.nx == blocks[bx]->d.dims.nx,
                        "block sizes do not form a Cartesian grid");
        if (bx > 0) ARMON_CHECK_ARG(d.origin_ix == blocks[i - 1]->d.origin_ix + blocks[i - 1]->d.dims.nx, "block origins (x)");
        if (by > 0) ARMON_CHECK_ARG(d.origin_iy == blocks[i - nbx]->d.origin_iy + blocks[i - nbx]->d.dims.ny, "block origins (y)");
        // sides facing another block of the group are wired here; a group spans one rank for now
        ARMON_CHECK_ARG(n == 1 || (d.neighbours[0] < 0 && d.neighbours[1] < 0 && d.neighbours[2] < 0 && d.neighbours[3] < 0),
                        "blocks of a multi-block group cannot have neighbour ranks");
    }
    armon_group *G = new armon_group();
    G->ctx = ctx;
    G->is_self = false;
    G->nbx = nbx;
    G->nby = nby;
    G->graph_mode = d0.cuda_graph;
    if (int rc = group_alloc_common(G)) { group_free_common(G); delete G; return rc; }
    for (int i = 0; i < n; i++) G->blocks.push_back(blocks[i]);
    G->ts = blocks[0]->own_ts;
    for (int i = 0; i < n; i++) {
        armon_solver *b = blocks[i];
        const int bx = i % nbx, by = i / nbx;
        b->group = G;
        b->ts = G->ts;
        b->local_nb[ARMON_SIDE_LEFT] = bx > 0 ? blocks[i - 1] : nullptr;
        b->local_nb[ARMON_SIDE_RIGHT] = bx < nbx - 1 ? blocks[i + 1] : nullptr;
        b->local_nb[ARMON_SIDE_BOTTOM] = by > 0 ? blocks[i - nbx] : nullptr;
        b->local_nb[ARMON_SIDE_TOP] = by < nby - 1 ? blocks[i + nbx] : nullptr;
    }
    if (int rc = group_reset(G)) { armon_group_destroy(G); return rc; }
    *out = G;
    return ARMON_OK;
}

int armon_group_destroy(armon_group *G)
{
    if (!G) return ARMON_OK;
    if (G->is_self) {
        armon_set_error("armon_group_destroy on the private group of a solver");
        return ARMON_ERR_INVALID;
    }
    cudaSetDevice(G->ctx->device);
    cudaStreamSynchronize(G->ctx->comm_stream);
    cudaStreamSynchronize(G->ctx->stream);
    for (armon_solver *b : G->blocks) {
        b->group = b->self;
        b->ts = b->own_ts;
        for (int k = 0; k < 4; k++) b->local_nb[k] = nullptr;
    }
    group_free_common(G);
    delete G;
    return ARMON_OK;
}

int armon_group_init(armon_group *G)
{
    if (int rc = group_check(G)) return rc;
    for (armon_solver *b : G->blocks)
        if (int rc = solver_init_fields(b)) return rc;
    return group_reset(G);
}

int armon_group_reset(armon_group *G)
{
    if (int rc = group_check(G)) return rc;
    return group_reset(G);
}

int armon_group_run(armon_group *G, int64_t n_cycles)
{
    if (int rc = group_check(G)) return rc;
    return group_run(G, n_cycles);
}

int armon_group_time_loop(armon_group *G)
{
    if (int rc = group_check(G)) return rc;
    return group_time_loop(G);
}

int armon_group_state(armon_group *G, armon_time_state *out)
{
    if (int rc = group_check(G)) return rc;
    ARMON_CHECK_ARG(out != nullptr, "null state");
    return read_state(G, out);
}

int armon_group_finalize(armon_group *G)
{
    if (int rc = group_check(G)) return rc;
    return group_finalize(G);
}

int armon_group_elapsed_ms(armon_group *G, float *ms)
{
    if (int rc = group_check(G)) return rc;
    return group_elapsed_ms(G, ms);
}

int armon_group_diagnostics(armon_group *G, int32_t capacity)
{
    if (int rc = group_check(G)) return rc;
    return group_diagnostics(G, capacity);
}

int armon_group_read_diagnostics(armon_group *G, armon_cycle_diag *lines, int64_t max_lines, int64_t *n_lines)
{
    if (int rc = group_check(G)) return rc;
    return group_read_diagnostics(G, lines, max_lines, n_lines);
}

}   // extern "C"
