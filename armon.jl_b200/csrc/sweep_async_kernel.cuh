// sweep_async_kernel.cuh -- the fused axis-sweep marching kernel with inputs staged through shared memory by
// per-thread asynchronous copies (cp.async, LDGSTS).
//
// Same mathematics, data layout and HBM traffic as sweep_kernel.cuh (read it first; march_compute is shared).  The
// register-prefetch kernel keeps the 4 rows in flight in 32 registers and the round-1 profile shows 15 % of its cycles
// waiting on those loads (HBM latency under load exceeds a 4-row lead); the bulk-copy (TMA) variant of
// sweep_tma_kernel.cuh removes the wait but pays ~45 issue slots per step for the elected-lane producer code.  Here
// the warp copies the 4 x 256 bytes (rho, ua, ut, E of its 32 columns) of the row ASYNC_NS - 1 steps ahead into a
// per-warp ring [slot][variable][lane] with two 16-byte `cp.async.cg` per thread (lanes 0-15: variables 0 and 2, lanes
// 16-31: variables 1 and 3) and one commit per step; the consumer side is one `cp.async.wait_group`, a warp barrier and
// four conflict-free 8-byte shared loads.  `.cg` bypasses L1: with the 8-byte `.ca` form every row in flight pins L1
// lines (57 KB per SM at 7 rows x 8 warps), and the measured time followed the L1 size left over by the shared-memory
// carve-out (1.05 ms at the default carve-out, 1.23 ms with all of the array given to shared memory).  The lead is a
// compile-time constant that costs shared memory instead of registers.  16-byte copies need an even pitch and
// 16-byte aligned arrays; the host falls back to sweep_kernel otherwise.
#pragma once

#include "sweep_kernel.cuh"

#ifndef ASYNC_NS
#define ASYNC_NS 8          // ring slots per warp; rows in flight = ASYNC_NS - 1
#endif
#ifndef ASYNC_TPB_VALUE
#define ASYNC_TPB_VALUE 128
#endif
constexpr int ASYNC_TPB = ASYNC_TPB_VALUE;   // threads (= columns) per CTA; warps are independent of one another

struct AsyncWarpShared {
    double ring[ASYNC_NS][4][32];                          // [slot][variable][lane]
    double stage[4 * 32 * SWEEP_STAGE_PITCH];              // transposed-store staging (flush_stage)
};

__device__ __forceinline__ void async_copy16(unsigned dst, const double *src)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// Per-thread copy plan: 16 bytes (2 columns) of two variables per array row.
struct AsyncLane {
    const double *src[2];   // A.in[2j + (lane >> 4)] + w0 + g + 2 * (lane & 15)
    unsigned dst[2];        // shared address of ring[0][2j + (lane >> 4)][2 * (lane & 15)]
    bool active;            // the two columns exist (ragged last warp of a row)
};

// copies of the array row whose first cell has element offset `off` into ring slot `s`
__device__ __forceinline__ void async_issue_row(const AsyncLane &L, long long off, int s, bool row_ok = true)
{
    if (L.active && row_ok) {
        async_copy16(L.dst[0] + 1024u * (unsigned)s, L.src[0] + off);
        async_copy16(L.dst[1] + 1024u * (unsigned)s, L.src[1] + off);
    }
}

// Strict-mode bookkeeping of the cp.async kernels, evaluated once per chunk of SWEEP_CHUNK emitted cells.  An operand
// outside the proven range of the branch-free division (common.cuh) met while chunk k is being emitted can only reach
// cells emitted in chunks k .. k+2 (dependency cone of 9 cells, emission lag <= 5 steps): the thread appends
// (first row of chunk k, column) to the work list and keeps the CFL maxima of those three chunks out of its totals;
// sweep_fixup_kernel recomputes FIX_CHUNKS * SWEEP_CHUNK rows of that column with nvcc's full IEEE division afterwards
// (bit-identical for the cells that were in range), densely packed, one thread per entry.
constexpr int FIX_CHUNKS = 3;
struct ChunkFix {
    unsigned long long tot_a, tot_t;   // CFL maxima over the clean chunks
    int taint;                         // chunks still reached by an out-of-range operand met earlier
    bool always;                       // dt or dx themselves are out of range: every chunk goes to the fix-up
};

template <int DIV>
__device__ __forceinline__ void chunk_end(const SweepArgs &A, SweepThread &T, ChunkFix &C, long long mb, long long w)
{
    if (DIV != DIV_FLAGGED) return;
    if (T.flag.bad() || C.always) {
        C.taint = FIX_CHUNKS;
        if (T.valid) {
            const unsigned e = atomicAdd(A.fix_count, 1u);
            if (e < A.fix_cap) A.fix_list[e] = ((unsigned long long)mb << 32) | (unsigned long long)(unsigned)w;
            else A.ts->range_error = 1;   // reaches every rank through the error channel of the dt all-reduce
        }
    }
    if (C.taint > 0) {
        C.taint--;
    } else {
        C.tot_a = T.amax > C.tot_a ? T.amax : C.tot_a;
        C.tot_t = T.tmax > C.tot_t ? T.tmax : C.tot_t;
    }
    T.amax = 0ULL; T.tmax = 0ULL;
    T.flag = RangeFlag();
}

#ifndef ASYNC_MIN_BLOCKS
#define ASYNC_MIN_BLOCKS (256 / ASYNC_TPB_VALUE)   // 8 warps per SM
#endif

// TR: 1 = the output is written transposed (through the staging tile), 0 = in the layout it was read (A.transpose_out
// must agree).
template <class R, int DIV, int RL, int PROJ, int EOS, int TR>
__global__ void __launch_bounds__(ASYNC_TPB, ASYNC_MIN_BLOCKS) sweep_async_kernel(const SweepArgs A)
{
    extern __shared__ __align__(128) unsigned char async_smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    AsyncWarpShared &S = reinterpret_cast<AsyncWarpShared *>(async_smem_raw)[warp];

    const long long w = (long long)blockIdx.x * ASYNC_TPB + threadIdx.x;
    const long long w0 = (long long)blockIdx.x * ASYNC_TPB + (threadIdx.x & ~31);
    const long long m0 = sweep_segment_index(A) * A.seg;
    const long long m1 = (m0 + A.seg < A.nm) ? m0 + A.seg : A.nm;

    SweepThread T;
    T.valid = w < A.nw;
    T.col = (T.valid ? w : A.nw - 1) + A.g;
#pragma unroll
    for (int k = 0; k < 4; k++) T.base[k] = A.in[k] + T.col;
    T.amax = 0ULL; T.tmax = 0ULL;

    const DeviceTimeState *ts = A.ts;
    if (ts->done) {   // see sweep_kernel: copy the state through so that the host's buffer rotation stays valid
        if (T.valid) {
            for (long long m = m0; m < m1; m++) {
                const long long i = (m + A.g) * A.pitch_in + T.col;
                const long long o = A.transpose_out ? T.col * A.pitch_out + (m + A.g) : (m + A.g) * A.pitch_out + T.col;
#pragma unroll
                for (int k = 0; k < 4; k++) A.out[k][o] = A.in[k][i];
            }
        }
        return;
    }
    if (w0 >= A.nw) return;   // warp entirely outside the domain (warps are independent: no CTA barrier below)

    const R dt = R(ts->current_dt) * R(A.dt_factor);   // update_solver_state!, src/solver_state.jl:339-345
    const long long nchunks = (m1 - m0 + SWEEP_CHUNK - 1) / SWEEP_CHUNK;
    const long long a_begin = m0 - 4;

    // prologue: rows a_begin .. a_begin + ASYNC_NS - 2, one commit group per row (a segment has >= 16 steps);
    // afterwards step t fetches row a_begin + t + ASYNC_NS - 1 into the slot consumed at step t - 1
    static_assert(ASYNC_NS >= 2 && ASYNC_NS <= 16 && (ASYNC_NS & (ASYNC_NS - 1)) == 0, "ring size");
    AsyncLane L;
    {
        const int h = lane >> 4, piece = lane & 15;
        const long long cols = A.nw - w0 < 32 ? A.nw - w0 : 32;
        L.active = 2 * piece < cols;
#pragma unroll
        for (int j = 0; j < 2; j++) {
            L.src[j] = (h ? A.in[2 * j + 1] : A.in[2 * j]) + w0 + A.g + 2 * piece;
            L.dst[j] = (unsigned)__cvta_generic_to_shared(&S.ring[0][2 * j + h][2 * piece]);
        }
        // columns past the end of the row are never copied: give those lanes a benign finite state (rho = E = 1, u = v = 0)
        for (int k = lane; k < ASYNC_NS * 4 * 32; k += 32) (&S.ring[0][0][0])[k] = ((k >> 5) & 3) == 0 || ((k >> 5) & 3) == 3 ? 1.0 : 0.0;
        __syncwarp();
    }
#pragma unroll 1
    for (int s = 0; s < ASYNC_NS - 1; s++) {
        async_issue_row(L, march_row_offset(A, a_begin + s), s);
        async_commit();
    }
    long long off_run = march_row_offset(A, a_begin + ASYNC_NS - 1);   // offset of row a + ASYNC_NS - 1
    const long long off_max = (A.nm + 2 * A.g - 1) * A.pitch_in;

    const typename Div<R, DIV>::Rcp inv_dx = Div<R, DIV>::prepare(R(A.dx), T.flag);
    ChunkFix C;
    C.tot_a = 0ULL; C.tot_t = 0ULL; C.taint = 0;
    if (DIV == DIV_FLAGGED) range_check_dividend(dt.v, T.flag);
    C.always = DIV == DIV_FLAGGED && T.flag.bad();
    Pipe<R> P;
#pragma unroll
    for (int j = 0; j < 4; j++) {
        P.cu[j] = R(0.); P.cp[j] = R(1.); P.crc[j] = R(1.); P.cdm[j] = R(1.); P.cut[j] = R(0.); P.cE[j] = R(1.); P.cc[j] = R(1.);
        P.Gu[j] = R(0.); P.Gp[j] = R(1.); P.Fu[j] = R(0.); P.Fp[j] = R(1.); P.FpFu[j] = R(0.); P.disp[j] = R(0.);
        P.dxl[j] = R(1.); P.Lr[j] = R(1.); P.Lu[j] = R(0.); P.Lt[j] = R(0.); P.LE[j] = R(1.);
        P.Lru[j] = R(0.); P.Lrt[j] = R(0.); P.LrE[j] = R(1.);
    }
    P.Ar = R(0.); P.Aru = R(0.); P.Art = R(0.); P.ArE = R(0.);
    P.Sr = R(0.); P.Sru = R(0.); P.Srt = R(0.); P.SrE = R(0.); P.S2b = R(2.); P.S2r = R(0.5);

    double *stage = S.stage;
    long long a = a_begin;
    unsigned step = 0;   // a - a_begin

    // Group accounting: ASYNC_NS - 1 groups before the loop, exactly one per step afterwards, so the group of the row
    // consumed at step t is complete once at most ASYNC_NS - 2 groups are pending.  Rows past the end of the segment
    // are fetched (clamped to the last array row) and never consumed: no branch in the step.
#define ASYNC_STEP(J, EMIT)                                                                                 \
    {                                                                                                       \
        async_wait<ASYNC_NS - 2>();                                                                         \
        __syncwarp();   /* every lane's copies of this row have landed; the slot refilled below was read a step ago */ \
        const double *slot = &S.ring[step & (ASYNC_NS - 1)][0][lane];                                       \
        const R rho(slot[0]), ua(slot[32]), ut(slot[64]), E(slot[96]);                                      \
        async_issue_row(L, off_run, (int)((step + ASYNC_NS - 1) & (ASYNC_NS - 1)));                         \
        async_commit();                                                                                     \
        off_run = off_run < off_max ? off_run + A.pitch_in : off_run;   /* clamped at the last array row */ \
        march_compute<R, DIV, RL, PROJ, EOS, true, J, TR, EMIT>(A, T, P, rho, ua, ut, E, a, dt, inv_dx, EMIT != 0, \
                                                                kc + J, m1, stage);                         \
        a++; step++;                                                                                        \
    }

    // warm-up: 8 steps fill the dependency cone of the first output, nothing is emitted
    {
        const int kc = 0;
#pragma unroll 1
        for (int it = 0; it < 2; it++) {
            ASYNC_STEP(0, 0)
            ASYNC_STEP(1, 0)
            ASYNC_STEP(2, 0)
            ASYNC_STEP(3, 0)
        }
    }
    // steady state: every iteration emits 4 cells, every second one flushes the transposed staging tile
#pragma unroll 1
    for (long long it = 0; it < 2 * nchunks; it++) {
        const int kc = (int)(it & 1) * 4;
        ASYNC_STEP(0, 1)
        ASYNC_STEP(1, 1)
        ASYNC_STEP(2, 1)
        ASYNC_STEP(3, 1)
        if (TR == 1 && (it & 1)) flush_stage(A, stage, w0, a - 12, m1);
        if (it & 1) chunk_end<DIV>(A, T, C, a - 12, w);
    }
#undef ASYNC_STEP
    async_wait<0>();

    unsigned long long am = DIV == DIV_FLAGGED ? C.tot_a : T.amax, tm = DIV == DIV_FLAGGED ? C.tot_t : T.tmax;
    if (!T.valid) { am = 0ULL; tm = 0ULL; }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        const unsigned long long oa = __shfl_xor_sync(0xffffffffu, am, off);
        const unsigned long long ot = __shfl_xor_sync(0xffffffffu, tm, off);
        am = oa > am ? oa : am;
        tm = ot > tm ? ot : tm;
    }
    if (lane == 0) {
        atomicMax(&A.ts->acc[A.acc_slot][0], am);
        atomicMax(&A.ts->acc[A.acc_slot][1], tm);
    }
}
