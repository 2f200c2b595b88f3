// solver.cu -- the fused solver object behind `solver_cycle` / `next_time_step` / `next_cycle!` / `time_loop`
// (src/solver.jl:288-403, src/reductions.jl:164-199, src/solver_state.jl:58-166) for ArmonParameters{T,<:B200Device}.
//
// Per cycle the stream receives: [halo exchange ->] one marching sweep kernel per axis of the splitting
// (sweep_kernel.cuh), [NCCL all-reduce(max) of the two CFL accumulators ->] one single-thread kernel that performs
// next_cycle! and the next cycle's time-step update on the device.  Nothing synchronises with the host: dt, time,
// cycle count and the stop condition live in DeviceTimeState.
#include "sweep_dispatch.h"

#include <cmath>
#include <cstdlib>
#include <string>
#include <vector>

namespace {

constexpr int TPB = 256;

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
    for (int k = 0; k < 3; k++) ts->acc[k][0] = ts->acc[k][1] = 0ULL;
}

struct CycleStepArgs {
    int first;            // 1: called before cycle 0 (no next_cycle! to apply)
    long long k;          // index of the cycle that just ran (when !first)
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
    unsigned long long bx = 0ULL, by = 0ULL;
    if (a.read_slot >= 0) {
        bx = ts->acc[a.read_slot][a.acc_is_xy ? 0 : 1];
        by = ts->acc[a.read_slot][a.acc_is_xy ? 1 : 0];
        ts->acc[a.read_slot][0] = ts->acc[a.read_slot][1] = 0ULL;
    }
    ts->acc[2][0] = ts->acc[2][1] = 0ULL;
    if (ts->done) return;

    if (!a.first) {
        ts->cycle += 1;
        ts->time = __dadd_rn(ts->time, ts->current_dt);
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
            ts->error = ARMON_ERR_TIME;
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
    double fa_lo, ft_lo, fa_hi, ft_hi;
};

__global__ void k_bc_fill(BcFillArgs B)
{
    const long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int k = blockIdx.y, side = blockIdx.z;
    if (w >= B.nw || !(side == 0 ? B.lo : B.hi)) return;
    const long long src_row = side == 0 ? k : B.nm - 1 - k, dst_row = side == 0 ? -1 - k : B.nm + k;
    const long long is = (src_row + B.g) * B.pitch + w + B.g, id = (dst_row + B.g) * B.pitch + w + B.g;
    const double fa = side == 0 ? B.fa_lo : B.fa_hi, ft = side == 0 ? B.ft_lo : B.ft_hi;
    B.rho[id] = B.rho[is];
    B.E[id] = B.E[is];
    B.ua[id] = __dmul_rn(B.ua[is], fa);
    B.ut[id] = __dmul_rn(B.ut[is], ft);
}

// ---- layout helpers ---------------------------------------------------------------------------------------------
struct Ptr4 { const double *in[4]; double *out[4]; };

// out[c][r] = in[r][c] for 4 arrays at once; rows x cols are the full array extents including ghosts
__global__ void k_transpose4(Ptr4 P, long long rows, long long cols)
{
    __shared__ double tile[4][32][33];
    const long long c0 = (long long)blockIdx.x * 32, r0 = (long long)blockIdx.y * 32;
    const int tx = threadIdx.x, ty = threadIdx.y;   // 32 x 8
#pragma unroll
    for (int k = 0; k < 4; k++)
        for (int j = ty; j < 32; j += 8) {
            const long long r = r0 + j, c = c0 + tx;
            if (r < rows && c < cols) tile[k][j][tx] = P.in[k][r * cols + c];
        }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 4; k++)
        for (int j = ty; j < 32; j += 8) {
            const long long c = c0 + j, r = r0 + tx;
            if (r < rows && c < cols) P.out[k][c * rows + r] = tile[k][tx][j];
        }
}

// Stale p, c, g of the reference (EOS of the state at the start of the last sweep), canonical layout out.
template <int EOS>
__global__ void k_eos_pcg(long long nx, long long ny, int g, int in_transposed, const double *rho, const double *u,
                          const double *v, const double *E, double gamma, double *p, double *c, double *gg)
{
    const long long ix = (long long)blockIdx.x * blockDim.x + threadIdx.x;   // 0-based real cell
    const long long iy = blockIdx.y;
    if (ix >= nx) return;
    const long long io = (iy + g) * (nx + 2 * g) + (ix + g);
    const long long ii = in_transposed ? (ix + g) * (ny + 2 * g) + (iy + g) : io;
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
    DeviceTimeState  *ts = nullptr;
    long long         host_cycle = 0;          // cycles enqueued since the last reset
    bool              started = false;         // initial time step enqueued
    cudaEvent_t       ev_start = nullptr, ev_stop = nullptr;
    cudaEvent_t       ev_state = nullptr, ev_halo = nullptr;   // compute -> comm (state ready), comm -> compute (ghosts ready)
    cudaEvent_t       ev_dt[2] = {nullptr, nullptr};           // all-reduce of accumulator slot 0 / 1 done
    bool              last_axis_is_x[2] = {true, true};        // axis of the sweep that filled accumulator slot 0 / 1
    cudaStream_t      edge_stream = nullptr;                   // the two edge segments of an overlapped sweep
    cudaEvent_t       ev_edge = nullptr;                       // edge segments done
    bool              timed = false;
    uint64_t          sweep_launches = 0;
    sweep_fn_t        kernel = nullptr;
    // warp-specialised path (sweep_ws_kernel.cuh): main kernel + IEEE fix-up kernel (strict mode only)
    sweep_ws_fn_t     ws_kernel = nullptr, fixup_kernel = nullptr;
    bool              use_ws = false;
    bool              overlap = true;      // interior / edge split of a sweep around the halo exchange (ARMON_B200_OVERLAP=0 disables)
    // TMA-staged marching kernel (sweep_tma_kernel.cuh); falls back to `kernel` per launch when the bulk-copy
    // alignment rules do not hold (odd input pitch)
    sweep_fn_t        tma_kernel = nullptr;
    bool              use_tma = false;
    // cp.async-staged marching kernel (sweep_async_kernel.cuh)
    sweep_fn_t        async_kernel[2] = {nullptr, nullptr};   // [transposed output]
    bool              use_async = false;
    size_t            async_smem = 0;      // dynamic shared memory per CTA of the selected async kernel
    unsigned         *fix_count = nullptr;            // two counters, used alternately by successive sweeps
    unsigned long long *fix_list = nullptr;
    unsigned          fix_cap = 0;                    // entries
    uint64_t          sweep_index = 0;
    // optional per-sweep-kernel timing (CUDA events on the launching stream), for the roofline figure
    bool              profile = false;
    std::vector<cudaEvent_t> prof_events;     // pairs (before, after) of each profiled sweep launch
    size_t            prof_used = 0;
};

namespace {

int solver_check(armon_solver *s, bool need_bound = true)
{
    ARMON_CHECK_ARG(s != nullptr && s->ctx != nullptr, "null solver");
    if (need_bound) ARMON_CHECK_ARG(s->bound, "armon_solver_bind was not called");
    return armon_ctx_activate(s->ctx);
}

long long n_elems(const armon_solver *s)
{
    return (s->d.dims.nx + 2 * s->d.dims.g) * (s->d.dims.ny + 2 * s->d.dims.g);
}

// Transpose the current state into the other buffer set.
int transpose_current(armon_solver *s)
{
    const armon_dims &D = s->d.dims;
    const long long rows = s->cur_transposed ? D.nx + 2 * D.g : D.ny + 2 * D.g;
    const long long cols = s->cur_transposed ? D.ny + 2 * D.g : D.nx + 2 * D.g;
    Ptr4 P;
    for (int k = 0; k < 4; k++) { P.in[k] = s->buf[s->cur][k]; P.out[k] = s->buf[1 - s->cur][k]; }
    const dim3 grid((unsigned)((cols + 31) / 32), (unsigned)((rows + 31) / 32), 1), block(32, 8, 1);
    k_transpose4<<<grid, block, 0, s->ctx->stream>>>(P, rows, cols);
    ARMON_LAUNCH_CHECK(s->ctx);
    s->cur = 1 - s->cur;
    s->cur_transposed = !s->cur_transposed;
    s->have_prev = false;
    return ARMON_OK;
}

// A sweep along `axis` marches along the strided dimension: X needs the transposed layout, Y the canonical one.
int ensure_layout(armon_solver *s, int axis)
{
    const bool need_transposed = (axis == ARMON_AXIS_X);
    if (need_transposed != s->cur_transposed) return transpose_current(s);
    return ARMON_OK;
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
    if (s->d.march_segment > 0) {
        int seg = (s->d.march_segment + SWEEP_CHUNK - 1) / SWEEP_CHUNK * SWEEP_CHUNK;
        return seg;
    }
    // as long as possible (8 warm-up rows per segment are redundant work) while keeping >= 6 waves of CTAs
    // TMA_TPB == SWEEP_TPB; the async kernels run 8 warps per SM whatever their CTA size
    const bool async_pitch_ok = s->use_async && ((nw + 2 * s->d.dims.g) % 2) == 0;
    const long long cols_per_cta = s->use_ws ? 32 : (async_pitch_ok ? ASYNC_TPB : SWEEP_TPB);
    const long long ctas_per_sm = s->use_ws ? 7 : (async_pitch_ok ? 256 / ASYNC_TPB : 2);
    const long long ncol = (nw + cols_per_cta - 1) / cols_per_cta;
    const long long want = 6LL * ctas_per_sm * s->ctx->sm_count;
    const int cands[] = {2048, 1024, 512, 256, 128, 64, 32, 16};
    for (int seg : cands) {
        if (seg > nm && seg != 16) continue;
        if (ncol * ((nm + seg - 1) / seg) >= want) return seg;
    }
    return nm >= 64 ? 32 : 16;
}

int launch_sweep(armon_solver *s, int axis, double dt_factor, bool last_of_cycle, int next_axis)
{
    if (int rc = ensure_layout(s, axis)) return rc;

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
    A.mirror_lo = s->d.neighbours[lo_side] < 0;
    A.mirror_hi = s->d.neighbours[hi_side] < 0;
    A.bc_a_lo = x ? tc.bc_u[lo_side] : tc.bc_v[lo_side];
    A.bc_t_lo = x ? tc.bc_v[lo_side] : tc.bc_u[lo_side];
    A.bc_a_hi = x ? tc.bc_u[hi_side] : tc.bc_v[hi_side];
    A.bc_t_hi = x ? tc.bc_v[hi_side] : tc.bc_u[hi_side];
    A.dx = s->d.domain_size[axis] / (double)(x ? s->d.global_nx : s->d.global_ny);   // update_solver_state!
    A.dx_pow2 = is_pow2_double(A.dx) ? 1 : 0;
    A.inv_dx = 1.0 / A.dx;
    A.dt_factor = dt_factor;
    A.gamma = tc.gamma;
    A.ts = s->ts;
    A.acc_slot = last_of_cycle ? (int)(s->host_cycle & 1) : 2;

    // block_ghost_exchange with the neighbour ranks (src/halo_exchange.jl:286-354) runs on the communication stream.
    // Only the first and the last march segment read ghost rows: the interior segments are launched right away and
    // overlap the exchange, the two edge segments follow once the ghost rows have arrived.
    const long long nseg = (A.nm + A.seg - 1) / A.seg;
    const bool has_nb = !A.mirror_lo || !A.mirror_hi;
    const bool overlap = has_nb && nseg >= 3 && s->overlap;
    if (has_nb) {
        if (int rc = comm_begin(s)) return rc;
        if (int rc = halo_exchange(s, axis, s->ctx->comm_stream)) return rc;
        if (int rc = comm_end(s, !overlap)) return rc;
    }

    const bool async_launch = s->use_async && (A.pitch_in % 2) == 0;
    const long long cols_per_cta = s->use_ws ? 32 : (async_launch ? ASYNC_TPB : SWEEP_TPB);
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
    FixupArgs F;
    F.count = s->fix_count ? s->fix_count + (s->sweep_index & 1) : nullptr;
    F.count_next = s->fix_count ? s->fix_count + ((s->sweep_index + 1) & 1) : nullptr;
    F.list = s->fix_list;
    A.fix_count = F.count;
    A.fix_list = F.list;
    A.fix_cap = s->fix_cap;
    A.fix_rows = (async_launch && s->fixup_kernel) ? FIX_CHUNKS * SWEEP_CHUNK : 0;
    // one launch over `ny` march segments y_base, y_base + y_jump, ...
    auto launch = [&](int y_base, int y_jump, long long ny, cudaStream_t st) -> int {
        A.y_base = y_base;
        A.y_jump = y_jump;
        const dim3 grid((unsigned)((A.nw + cols_per_cta - 1) / cols_per_cta), (unsigned)ny, 1);
        if (s->use_ws)
            s->ws_kernel<<<grid, WS_TPB, 0, st>>>(A, F);
        else if (async_launch)
            s->async_kernel[A.transpose_out ? 1 : 0]<<<grid, ASYNC_TPB, s->async_smem, st>>>(A);
        else if (s->use_tma && (A.pitch_in % 2) == 0)
            s->tma_kernel<<<grid, TMA_TPB, TMA_TPB / 32 * sizeof(TmaWarpShared), st>>>(A);
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
        B.fa_lo = A.bc_a_lo; B.ft_lo = A.bc_t_lo; B.fa_hi = A.bc_a_hi; B.ft_hi = A.bc_t_hi;
        const dim3 bgrid((unsigned)((A.nw + TPB - 1) / TPB), (unsigned)A.g, 2);
        k_bc_fill<<<bgrid, TPB, 0, st>>>(B);
        ARMON_LAUNCH_CHECK(s->ctx);
        return ARMON_OK;
    };
    if (overlap) {
        // interior segments on the compute stream, overlapping the exchange; the two edge segments on their own stream
        // once the ghost rows are there (ev_halo also carries ev_state: the input state is complete), so that they
        // fill the last, partial wave of the interior launch instead of running after it
        if (int rc = launch(1, 1, nseg - 2, s->ctx->stream)) return rc;
        ARMON_CUDA(cudaStreamWaitEvent(s->edge_stream, s->ev_halo, 0));
        if (int rc = bc_fill(s->edge_stream)) return rc;
        if (int rc = launch(0, (int)(nseg - 1), 2, s->edge_stream)) return rc;
        ARMON_CUDA(cudaEventRecord(s->ev_edge, s->edge_stream));
        ARMON_CUDA(cudaStreamWaitEvent(s->ctx->stream, s->ev_edge, 0));
    } else {
        if (int rc = bc_fill(s->ctx->stream)) return rc;
        if (int rc = launch(0, 1, nseg, s->ctx->stream)) return rc;
    }
    if (s->use_ws || (async_launch && s->fixup_kernel)) {
        if (s->fixup_kernel) {
            A.y_base = 0; A.y_jump = 1;
            s->fixup_kernel<<<2 * s->ctx->sm_count, 32, 0, s->ctx->stream>>>(A, F);
            ARMON_LAUNCH_CHECK(s->ctx);
        }
        s->sweep_index++;
    }
    if (s->profile) ARMON_CUDA(cudaEventRecord(ev1, s->ctx->stream));
    s->sweep_launches++;

    s->have_prev = true;
    s->prev_transposed = s->cur_transposed;
    s->cur = out_set;
    if (A.transpose_out) s->cur_transposed = !s->cur_transposed;
    return ARMON_OK;
}

// MPI_Iallreduce(MIN) of the local dt (src/utils.jl:126-134, src/solver_state.jl:107-111) becomes an all-reduce(max) of
// the two CFL maxima of one accumulator slot (the min of the quotients is the quotient of the max), issued on the
// communication stream.  `wait`: the compute stream waits for it right away (initial time step); otherwise it is waited
// one cycle later (wait_allreduce), so that it overlaps the sweeps of the next cycle.
int allreduce_acc(armon_solver *s, int slot, bool wait)
{
    if (s->ctx->comm && s->ctx->nranks > 1) {
        if (int rc = comm_begin(s)) return rc;
        ARMON_NCCL(ncclAllReduce(&s->ts->acc[slot][0], &s->ts->acc[slot][0], 2, ncclUint64, ncclMax, s->ctx->comm,
                                 s->ctx->comm_stream));
        ARMON_CUDA(cudaEventRecord(s->ev_dt[slot], s->ctx->comm_stream));
        if (wait) ARMON_CUDA(cudaStreamWaitEvent(s->ctx->stream, s->ev_dt[slot], 0));
    }
    return ARMON_OK;
}

int wait_allreduce(armon_solver *s, int slot)
{
    if (s->ctx->comm && s->ctx->nranks > 1) ARMON_CUDA(cudaStreamWaitEvent(s->ctx->stream, s->ev_dt[slot], 0));
    return ARMON_OK;
}

int launch_cycle_step(armon_solver *s, bool first, long long k, int read_slot, bool acc_is_xy)
{
    CycleStepArgs a;
    a.first = first ? 1 : 0;
    a.k = k;
    a.read_slot = read_slot;
    a.acc_is_xy = acc_is_xy ? 1 : 0;
    a.dx = s->d.domain_size[0] / (double)s->d.global_nx;
    a.dy = s->d.domain_size[1] / (double)s->d.global_ny;
    a.cfl = s->d.cfl;
    a.maxtime = s->d.maxtime;
    a.maxcycle = s->d.maxcycle;
    a.cst_dt = s->d.cst_dt;
    a.Dt = s->d.Dt;
    k_cycle_step<<<1, 1, 0, s->ctx->stream>>>(s->ts, a);
    ARMON_LAUNCH_CHECK(s->ctx);
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

int enqueue_cycle(armon_solver *s)
{
    if (!s->started) {
        // cycle 0: EOS_init + first time step (src/solver.jl:291-297); its maxima go to slot 1 ("cycle -1")
        if (int rc = launch_init_dt(s)) return rc;
        if (int rc = allreduce_acc(s, 1, true)) return rc;
        if (int rc = launch_cycle_step(s, true, -1, 1, true)) return rc;
        s->started = true;
    }
    int axes[3], next_axes[3];
    double factors[3], next_factors[3];
    const long long k = s->host_cycle;
    const int n = split_axes(s->d.splitting, k, axes, factors);
    split_axes(s->d.splitting, k + 1, next_axes, next_factors);
    for (int i = 0; i < n; i++) {
        const bool last = i == n - 1;
        const int next_axis = last ? next_axes[0] : axes[i + 1];
        if (int rc = launch_sweep(s, axes[i], factors[i], last, next_axis)) return rc;
    }
    // the last sweep ran along axes[n-1]: slot k & 1 = (march axis, transverse axis).  Its all-reduce overlaps cycle k+1.
    if (int rc = allreduce_acc(s, (int)(k & 1), false)) return rc;
    s->last_axis_is_x[k & 1] = axes[n - 1] == ARMON_AXIS_X;
    // next_cycle! of cycle k; the time step D_k it installs comes from the maxima of cycle k-1 (none after cycle 0)
    if (k >= 1) {
        const int rs = (int)((k - 1) & 1);
        if (int rc = wait_allreduce(s, rs)) return rc;
        if (int rc = launch_cycle_step(s, false, k, rs, s->last_axis_is_x[rs])) return rc;
    } else {
        if (int rc = launch_cycle_step(s, false, k, -1, true)) return rc;
    }
    s->host_cycle++;
    return ARMON_OK;
}

int read_state(armon_solver *s, armon_time_state *out)
{
    DeviceTimeState *h = reinterpret_cast<DeviceTimeState *>(s->ctx->pinned);
    ARMON_CUDA(cudaMemcpyAsync(h, s->ts, sizeof(DeviceTimeState), cudaMemcpyDeviceToHost, s->ctx->stream));
    ARMON_CUDA(cudaStreamSynchronize(s->ctx->stream));
    out->cycle = h->cycle;
    out->time = h->time;
    out->current_dt = h->current_dt;
    out->next_cycle_dt = h->next_cycle_dt;
    out->error = h->error;
    out->done = h->done;
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
    ARMON_CHECK_ARG(!desc->cst_dt || desc->Dt != 0.0, "Dt == 0 with constant step enabled");

    armon_solver *s = new armon_solver();
    s->ctx = ctx;
    s->d = *desc;
    if (const char *ov = getenv("ARMON_B200_OVERLAP")) s->overlap = atoi(ov) != 0;
    const int rl = desc->riemann == ARMON_RIEMANN_GODUNOV ? 0 : 1 + desc->limiter;
    const bool biz = desc->tc.eos == ARMON_EOS_BIZARRIUM;
    if (desc->math_mode == ARMON_MATH_STRICT)
        s->kernel = biz ? sweep_table_strict_biz(rl, desc->projection) : sweep_table_strict_pg(rl, desc->projection);
    else if (desc->math_mode == ARMON_MATH_IEEE)
        s->kernel = biz ? sweep_table_ieee_biz(rl, desc->projection) : sweep_table_ieee_pg(rl, desc->projection);
    else
        s->kernel = biz ? sweep_table_fast_biz(rl, desc->projection) : sweep_table_fast_pg(rl, desc->projection);
    if (!s->kernel) {
        delete s;
        armon_set_error("no sweep kernel for this scheme combination");
        return ARMON_ERR_INVALID;
    }
    // Kernel variant: the warp-specialised kernel pays off once the grid fills the GPU; tiny grids (the 100x100
    // golden cases) are launch-bound and keep the single-role kernel.  ARMON_B200_KERNEL=single|ws overrides.
    {
        const char *env = getenv("ARMON_B200_KERNEL");
        // measured at 8192^2 (profiles/): strict 3.25 ms (ws) vs 3.48 ms (single); fast 2.23 ms (ws) vs 1.88 ms (single)
        // kernel_variant / ARMON_B200_KERNEL: 0 auto (= async2), 1 single (register prefetch), 2 ws, 3 tma, 4 async,
        // 5 async2 (cp.async staging + software-pipelined step).  Measured at 8192^2, fast mode (profiles/README.md):
        // async2 0.99 ms, async 1.03 ms, tma 1.55 ms, single 1.68 ms, ws 2.2 ms per sweep.
        const bool want_tma = env ? (std::string(env) == "tma") : (desc->kernel_variant == 3);
        if (want_tma && desc->math_mode != ARMON_MATH_IEEE) {
            if (desc->math_mode == ARMON_MATH_STRICT)
                s->tma_kernel = biz ? sweep_tma_table_strict_biz(rl, desc->projection) : sweep_tma_table_strict_pg(rl, desc->projection);
            else
                s->tma_kernel = biz ? sweep_tma_table_fast_biz(rl, desc->projection) : sweep_tma_table_fast_pg(rl, desc->projection);
            if (s->tma_kernel) {
                ARMON_CUDA(cudaFuncSetAttribute((const void *)s->tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                (int)(TMA_TPB / 32 * sizeof(TmaWarpShared))));
                s->use_tma = true;
            }
        }
        // auto: the software-pipelined kernel for the fast mode (0.96 vs 1.03 ms per sweep at 8192^2); the strict mode is
        // register-bound (255 registers either way) and measures the same or better without the skew (1.92 ms)
        const bool is_auto = env ? std::string(env) == "auto" : desc->kernel_variant == 0;
        const bool want_async2 = !s->use_tma && ((env ? std::string(env) == "async2" : desc->kernel_variant == 5) ||
                                                 (is_auto && desc->math_mode == ARMON_MATH_FAST));
        const bool want_async = !s->use_tma && (want_async2 || (env ? std::string(env) == "async" : desc->kernel_variant == 4) ||
                                                (is_auto && desc->math_mode != ARMON_MATH_FAST));
        if (want_async && desc->math_mode != ARMON_MATH_IEEE) {
            bool ok = true;
            s->async_smem = ASYNC_TPB / 32 * (want_async2 ? sizeof(Async2WarpShared) : sizeof(AsyncWarpShared));
            for (int tr = 0; tr < 2; tr++) {
                if (want_async2) {
                    if (desc->math_mode == ARMON_MATH_STRICT)
                        s->async_kernel[tr] = biz ? sweep_async2_table_strict_biz(rl, desc->projection, tr)
                                                  : sweep_async2_table_strict_pg(rl, desc->projection, tr);
                    else
                        s->async_kernel[tr] = biz ? sweep_async2_table_fast_biz(rl, desc->projection, tr)
                                                  : sweep_async2_table_fast_pg(rl, desc->projection, tr);
                } else if (desc->math_mode == ARMON_MATH_STRICT)
                    s->async_kernel[tr] = biz ? sweep_async_table_strict_biz(rl, desc->projection, tr)
                                              : sweep_async_table_strict_pg(rl, desc->projection, tr);
                else
                    s->async_kernel[tr] = biz ? sweep_async_table_fast_biz(rl, desc->projection, tr)
                                              : sweep_async_table_fast_pg(rl, desc->projection, tr);
                ok = ok && s->async_kernel[tr] != nullptr;
                if (s->async_kernel[tr]) {
                    ARMON_CUDA(cudaFuncSetAttribute((const void *)s->async_kernel[tr],
                                                    cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s->async_smem));
                    // the kernel stages everything through shared memory and has no use for L1: give the whole
                    // unified array to shared memory so that ASYNC_MIN_BLOCKS CTAs are resident per SM
                    const char *cv = getenv("ARMON_B200_CARVEOUT");   // percent of shared memory, -1 = driver default
                    ARMON_CUDA(cudaFuncSetAttribute((const void *)s->async_kernel[tr],
                                                    cudaFuncAttributePreferredSharedMemoryCarveout,
                                                    cv ? atoi(cv) : (int)cudaSharedmemCarveoutMaxShared));
                    if (getenv("ARMON_B200_VERBOSE")) {
                        int nb = 0;
                        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, (const void *)s->async_kernel[tr], ASYNC_TPB,
                                                                      s->async_smem);
                        fprintf(stderr, "[armon_b200] async%s kernel tr=%d: %d resident CTAs/SM, %zu B shared/CTA\n",
                                want_async2 ? "2" : "", tr, nb, s->async_smem);
                    }
                }
            }
            s->use_async = ok;
        }
        const bool want_ws = !s->use_tma && !s->use_async && (env ? (std::string(env) == "ws") : (desc->kernel_variant == 2));
        if (want_ws && desc->math_mode != ARMON_MATH_IEEE && !(env && std::string(env) == "single")) {
            if (desc->math_mode == ARMON_MATH_STRICT) {
                s->ws_kernel = biz ? sweep_ws_table_strict_biz(rl, desc->projection) : sweep_ws_table_strict_pg(rl, desc->projection);
                s->fixup_kernel = biz ? sweep_fixup_table_biz(rl, desc->projection) : sweep_fixup_table_pg(rl, desc->projection);
            } else {
                s->ws_kernel = biz ? sweep_ws_table_fast_biz(rl, desc->projection) : sweep_ws_table_fast_pg(rl, desc->projection);
            }
            s->use_ws = s->ws_kernel != nullptr;
        }
        if (s->use_async && desc->math_mode == ARMON_MATH_STRICT)
            s->fixup_kernel = biz ? sweep_fixup_table_biz(rl, desc->projection) : sweep_fixup_table_pg(rl, desc->projection);
        if (s->use_ws || (s->use_async && s->fixup_kernel)) {
            // work list of the IEEE fix-up: one entry per (column, segment) at most
            long long cap = 0;
            for (int axis = 0; axis < 2; axis++) {
                const long long nm = axis == ARMON_AXIS_X ? D.nx : D.ny, nw = axis == ARMON_AXIS_X ? D.ny : D.nx;
                const long long seg = pick_segment(s, nm, nw);
                const long long n = nw * ((nm + seg - 1) / seg);
                cap = n > cap ? n : cap;
            }
            if (s->use_async) {
                // cp.async kernels: one entry per (column, chunk of 8 rows) at most; capped at 32 MB, an overflow (a
                // domain full of out-of-range values) raises ARMON_ERR_RANGE
                const long long worst = (D.nx > D.ny ? D.nx : D.ny) * (((D.nx > D.ny ? D.ny : D.nx) + SWEEP_CHUNK - 1) / SWEEP_CHUNK + 1);
                cap = worst < (4LL << 20) ? worst : (4LL << 20);
                if (cap < 1024) cap = 1024;
            }
            s->fix_cap = (unsigned)cap;
            ARMON_CUDA(cudaMalloc(&s->fix_count, 2 * sizeof(unsigned)));
            ARMON_CUDA(cudaMemsetAsync(s->fix_count, 0, 2 * sizeof(unsigned), ctx->stream));
            ARMON_CUDA(cudaMalloc(&s->fix_list, (size_t)cap * sizeof(unsigned long long)));
        }
    }
    ARMON_CUDA(cudaMalloc(&s->ts, sizeof(DeviceTimeState)));
    ARMON_CUDA(cudaEventCreate(&s->ev_start));
    ARMON_CUDA(cudaEventCreate(&s->ev_stop));
    ARMON_CUDA(cudaEventCreateWithFlags(&s->ev_state, cudaEventDisableTiming));
    ARMON_CUDA(cudaEventCreateWithFlags(&s->ev_halo, cudaEventDisableTiming));
    ARMON_CUDA(cudaEventCreateWithFlags(&s->ev_edge, cudaEventDisableTiming));
    ARMON_CUDA(cudaEventCreateWithFlags(&s->ev_dt[0], cudaEventDisableTiming));
    ARMON_CUDA(cudaEventCreateWithFlags(&s->ev_dt[1], cudaEventDisableTiming));
    ARMON_CUDA(cudaStreamCreateWithFlags(&s->edge_stream, cudaStreamNonBlocking));
    k_ts_reset<<<1, 1, 0, ctx->stream>>>(s->ts, desc->cst_dt, desc->Dt);
    ARMON_LAUNCH_CHECK(ctx);
    *out = s;
    return ARMON_OK;
}

int armon_solver_destroy(armon_solver *s)
{
    if (!s) return ARMON_OK;
    cudaSetDevice(s->ctx->device);
    cudaStreamSynchronize(s->ctx->stream);
    if (s->ts) cudaFree(s->ts);
    if (s->fix_count) cudaFree(s->fix_count);
    if (s->fix_list) cudaFree(s->fix_list);
    if (s->ev_start) cudaEventDestroy(s->ev_start);
    if (s->ev_stop) cudaEventDestroy(s->ev_stop);
    if (s->ev_state) cudaEventDestroy(s->ev_state);
    if (s->ev_halo) cudaEventDestroy(s->ev_halo);
    if (s->ev_edge) cudaEventDestroy(s->ev_edge);
    for (int k = 0; k < 2; k++) if (s->ev_dt[k]) cudaEventDestroy(s->ev_dt[k]);
    if (s->edge_stream) cudaStreamDestroy(s->edge_stream);
    for (cudaEvent_t e : s->prof_events) cudaEventDestroy(e);
    delete s;
    return ARMON_OK;
}

int armon_solver_bind(armon_solver *s, double *const main_vars[4], double *const work_vars[4], double *const pcg[3])
{
    if (int rc = solver_check(s, false)) return rc;
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
    s->have_prev = false;
    return ARMON_OK;
}

int armon_solver_reset(armon_solver *s)
{
    if (int rc = solver_check(s)) return rc;
    k_ts_reset<<<1, 1, 0, s->ctx->stream>>>(s->ts, s->d.cst_dt, s->d.Dt);
    ARMON_LAUNCH_CHECK(s->ctx);
    s->host_cycle = 0;
    s->started = false;
    s->cur = 0;
    s->cur_transposed = false;
    s->have_prev = false;
    s->timed = false;
    return ARMON_OK;
}

int armon_solver_init(armon_solver *s)
{
    if (int rc = solver_check(s)) return rc;
    const armon_solver_desc &d = s->d;
    if (int rc = armon_init_test(s->ctx, d.dims, d.origin_ix, d.origin_iy, d.global_nx, d.global_ny, d.domain_size,
                                 d.origin, &d.tc, nullptr, nullptr, nullptr, s->buf[0][0], s->buf[0][3], s->buf[0][1],
                                 s->buf[0][2], s->pcg[0], s->pcg[1], s->pcg[2], nullptr, nullptr, nullptr, nullptr,
                                 nullptr, nullptr))
        return rc;
    return armon_solver_reset(s);
}

int armon_solver_run(armon_solver *s, int64_t n_cycles)
{
    if (int rc = solver_check(s)) return rc;
    ARMON_CHECK_ARG(n_cycles >= 0, "negative cycle count");
    ARMON_CUDA(cudaEventRecord(s->ev_start, s->ctx->stream));
    for (int64_t c = 0; c < n_cycles; c++)
        if (int rc = enqueue_cycle(s)) return rc;
    ARMON_CUDA(cudaEventRecord(s->ev_stop, s->ctx->stream));
    s->timed = true;
    return ARMON_OK;
}

int armon_solver_state(armon_solver *s, armon_time_state *out)
{
    if (int rc = solver_check(s, false)) return rc;
    ARMON_CHECK_ARG(out != nullptr, "null state");
    return read_state(s, out);
}

int armon_solver_time_loop(armon_solver *s)
{
    if (int rc = solver_check(s)) return rc;
    ARMON_CUDA(cudaEventRecord(s->ev_start, s->ctx->stream));
    armon_time_state st;
    if (s->d.maxcycle <= 0 || !(0.0 < s->d.maxtime)) {   // `while time < maxtime && cycle < maxcycle` never entered
        ARMON_CUDA(cudaEventRecord(s->ev_stop, s->ctx->stream));
        s->timed = true;
        return ARMON_OK;
    }
    if (int rc = enqueue_cycle(s)) return rc;
    for (;;) {
        if (int rc = read_state(s, &st)) return rc;
        if (st.error == ARMON_ERR_RANGE) {
            armon_set_error("cycle %lld: a division/sqrt operand left [2^-500, 2^500]; rerun with math_mode ieee",
                            (long long)st.cycle);
            return ARMON_ERR_RANGE;
        }
        if (st.error) {
            armon_set_error("Invalid time step for cycle %lld", (long long)st.cycle);
            return ARMON_ERR_TIME;
        }
        if (st.done) break;
        // Lower bound of the cycles still to run: the time step grows by at most 5% per cycle
        // (src/solver_state.jl:127-130), so n cycles advance the time by at most dt*(1.05^n - 1)/0.05.
        long long batch = 1;
        const double remaining = s->d.maxtime - st.time;
        if (st.current_dt > 0.0 && remaining > 0.0) {
            const double n = s->d.cst_dt ? remaining / st.current_dt
                                         : std::log1p(0.05 * remaining / st.current_dt) / std::log(1.05);
            batch = (long long)std::floor(n) - 1;
        }
        const long long left = s->d.maxcycle - st.cycle;
        if (batch > left) batch = left;
        if (batch > 4096) batch = 4096;
        if (batch < 1) batch = 1;
        for (long long c = 0; c < batch; c++)
            if (int rc = enqueue_cycle(s)) return rc;
    }
    ARMON_CUDA(cudaEventRecord(s->ev_stop, s->ctx->stream));
    s->timed = true;
    return ARMON_OK;
}

int armon_solver_finalize(armon_solver *s)
{
    if (int rc = solver_check(s)) return rc;
    const armon_dims &D = s->d.dims;
    // 1. stale p, c, g from the state at the start of the last sweep (still intact in the other buffer set)
    if (s->have_prev && (s->pcg[0] || s->pcg[1] || s->pcg[2])) {
        double *const *b = s->buf[1 - s->cur];
        const dim3 grid((unsigned)((D.nx + TPB - 1) / TPB), (unsigned)D.ny, 1);
        if (s->d.tc.eos == ARMON_EOS_BIZARRIUM)
            k_eos_pcg<ARMON_EOS_BIZARRIUM><<<grid, TPB, 0, s->ctx->stream>>>(
                D.nx, D.ny, (int)D.g, s->prev_transposed, b[0], b[1], b[2], b[3], s->d.tc.gamma, s->pcg[0], s->pcg[1], s->pcg[2]);
        else
            k_eos_pcg<ARMON_EOS_PERFECT_GAS><<<grid, TPB, 0, s->ctx->stream>>>(
                D.nx, D.ny, (int)D.g, s->prev_transposed, b[0], b[1], b[2], b[3], s->d.tc.gamma, s->pcg[0], s->pcg[1], s->pcg[2]);
        ARMON_LAUNCH_CHECK(s->ctx);
    }
    s->have_prev = false;
    // 2. canonical layout, in main_vars
    if (s->cur_transposed) {
        if (int rc = transpose_current(s)) return rc;
    }
    if (s->cur != 0) {
        for (int k = 0; k < 4; k++)
            ARMON_CUDA(cudaMemcpyAsync(s->buf[0][k], s->buf[1][k], (size_t)n_elems(s) * sizeof(double),
                                       cudaMemcpyDeviceToDevice, s->ctx->stream));
        s->cur = 0;
    }
    return ARMON_OK;
}

int armon_solver_halo_exchange(armon_solver *s, int axis)
{
    if (int rc = solver_check(s)) return rc;
    ARMON_CHECK_ARG(axis == ARMON_AXIS_X || axis == ARMON_AXIS_Y, "axis");
    if (int rc = ensure_layout(s, axis)) return rc;
    if (int rc = comm_begin(s)) return rc;
    if (int rc = halo_exchange(s, axis, s->ctx->comm_stream)) return rc;
    return comm_end(s, true);
}

int armon_solver_elapsed_ms(armon_solver *s, float *ms)
{
    if (int rc = solver_check(s, false)) return rc;
    ARMON_CHECK_ARG(ms != nullptr, "null result");
    ARMON_CHECK_ARG(s->timed, "no armon_solver_run / armon_solver_time_loop call to time");
    ARMON_CUDA(cudaEventSynchronize(s->ev_stop));
    ARMON_CUDA(cudaEventElapsedTime(ms, s->ev_start, s->ev_stop));
    return ARMON_OK;
}

int armon_solver_profile(armon_solver *s, int enable)
{
    if (int rc = solver_check(s, false)) return rc;
    s->profile = enable != 0;
    s->prof_used = 0;
    return ARMON_OK;
}

int armon_solver_sweep_time_ms(armon_solver *s, double *total_ms, uint64_t *count)
{
    if (int rc = solver_check(s, false)) return rc;
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
    if (int rc = solver_check(s, false)) return rc;
    ARMON_CHECK_ARG(count != nullptr, "null result");
    *count = s->sweep_launches;
    return ARMON_OK;
}

}   // extern "C"
