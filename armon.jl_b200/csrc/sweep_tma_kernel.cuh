// sweep_tma_kernel.cuh -- the fused axis-sweep marching kernel with TMA-staged inputs.
//
// Same mathematics, data layout and HBM traffic as sweep_kernel.cuh (read it first; march_compute is shared).  What
// changes is how the cells of the march reach the threads: instead of per-thread 8-byte loads prefetched 4 rows
// ahead in registers (32 registers, and a prefetch distance the compiler shortens under register pressure -- the
// round-1 profile shows 15 % of the cycles waiting on those loads), every warp owns a shared-memory ring of
// TMA_NS array rows.  One lane of the warp issues four bulk asynchronous copies (cp.async.bulk, the 1-D TMA path:
// 32 columns x 8 bytes = 256 contiguous bytes of one variable) per row, TMA_NS - 1 rows ahead of the march, and
// the copies signal a per-slot mbarrier with their byte count.  The consumer side is one mbarrier wait and four
// conflict-free 8-byte shared loads per step.
//
// Requirements of the bulk copy (16-byte aligned source, size multiple of 16): even input pitch and 16-byte aligned
// arrays; the host falls back to sweep_kernel otherwise.  Warps are independent (no CTA-wide barrier in the march).
#pragma once

#include "sweep_kernel.cuh"

#ifndef TMA_NS
#define TMA_NS 8          // rows in flight per warp
#endif
constexpr int TMA_TPB = 128;

struct TmaWarpShared {
    double ring[TMA_NS][4][32];                            // [slot][variable][lane]
    double stage[4 * 32 * SWEEP_STAGE_PITCH];              // transposed-store staging (flush_stage)
    unsigned long long full[TMA_NS];
    unsigned long long pad[16 - (TMA_NS % 16)];            // keep sizeof a multiple of 128
};

__device__ __forceinline__ unsigned tma_smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void tma_mbar_init(unsigned long long *bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(tma_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void tma_mbar_expect_tx(unsigned bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_mbar_wait(unsigned bar, unsigned parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "TMA_WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra TMA_DONE_%=;\n"
        "bra TMA_WAIT_%=;\n"
        "TMA_DONE_%=:\n"
        "}\n" ::"r"(bar), "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(unsigned dst, const void *src, unsigned bytes, unsigned bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}

// Per-warp producer state (warp-uniform).
struct TmaProducer {
    unsigned ring;         // shared address of ring[0][0][0]
    unsigned bar;          // shared address of full[0]
    unsigned bytes;        // bytes per variable and row (<= 256, multiple of 16)
    long long col0;        // w0 + g: first column of the warp inside an array row
};

// One lane issues the four bulk copies (rho, ua, ut, E) of the array row whose first cell has element offset `off`
// (march_row_offset: mirrored / clamped at the edges) into slot `s`; the copies complete on the slot's mbarrier.
__device__ __forceinline__ void tma_issue_row(const SweepArgs &A, const TmaProducer &Q, long long off, int s, int lane)
{
    if (lane == 0) {
        const unsigned bar = Q.bar + 8u * (unsigned)s, dst = Q.ring + 1024u * (unsigned)s;
        tma_mbar_expect_tx(bar, 4u * Q.bytes);
        const long long o = off + Q.col0;
#pragma unroll
        for (int k = 0; k < 4; k++) tma_bulk_g2s(dst + 256u * k, A.in[k] + o, Q.bytes, bar);
    }
}

#ifndef TMA_MIN_BLOCKS
#define TMA_MIN_BLOCKS 2
#endif

template <class R, int DIV, int RL, int PROJ, int EOS>
__global__ void __launch_bounds__(TMA_TPB, TMA_MIN_BLOCKS) sweep_tma_kernel(const SweepArgs A)
{
    extern __shared__ __align__(128) unsigned char tma_smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    TmaWarpShared &S = reinterpret_cast<TmaWarpShared *>(tma_smem_raw)[warp];

    const long long w = (long long)blockIdx.x * TMA_TPB + threadIdx.x;
    const long long w0 = (long long)blockIdx.x * TMA_TPB + (threadIdx.x & ~31);
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

    if (lane == 0) {
        for (int s = 0; s < TMA_NS; s++) tma_mbar_init(&S.full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    // lanes whose column is outside the domain never receive data: give them a benign finite state
    for (int k = lane; k < TMA_NS * 4 * 32; k += 32) (&S.ring[0][0][0])[k] = 1.0;
    __syncwarp();

    const R dt = R(ts->current_dt) * R(A.dt_factor);   // update_solver_state!, src/solver_state.jl:339-345
    const long long nchunks = (m1 - m0 + SWEEP_CHUNK - 1) / SWEEP_CHUNK;
    const long long a_begin = m0 - 4;
    const long long a_last = m0 + nchunks * SWEEP_CHUNK + 3;   // last cell index consumed

    TmaProducer Q;
    Q.ring = tma_smem_u32(&S.ring[0][0][0]);
    Q.bar = tma_smem_u32(&S.full[0]);
    Q.col0 = w0 + A.g;
    {
        const long long cols = A.nw - w0 < 32 ? A.nw - w0 : 32;
        Q.bytes = (unsigned)(cols * 8);
    }
    // prologue: rows a_begin .. a_begin + TMA_NS - 2 (a segment has at least 16 >= TMA_NS steps); afterwards step t
    // fetches row a_begin + t + TMA_NS - 1 into the slot consumed at step t - 1
    static_assert(TMA_NS <= 16 && (TMA_NS & (TMA_NS - 1)) == 0, "ring size");
#pragma unroll 1
    for (int s = 0; s < TMA_NS - 1; s++) tma_issue_row(A, Q, march_row_offset(A, a_begin + s), s, lane);
    long long off_run = march_row_offset(A, a_begin + TMA_NS - 1);   // offset of row a + TMA_NS - 1
    const long long off_max = (A.nm + 2 * A.g - 1) * A.pitch_in;

    const typename Div<R, DIV>::Rcp inv_dx = Div<R, DIV>::prepare(R(A.dx), T.flag);
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
    const long long n_iter = 2 + 2 * nchunks;
    unsigned step = 0;   // a - a_begin

#define TMA_STEP(J)                                                                                         \
    {                                                                                                       \
        const int s = (int)(step & (TMA_NS - 1));                                                           \
        tma_mbar_wait(Q.bar + 8u * s, (step / TMA_NS) & 1u);                                                \
        const double *slot = &S.ring[s][0][lane];                                                           \
        const R rho(slot[0]), ua(slot[32]), ut(slot[64]), E(slot[96]);                                      \
        __syncwarp();                                                                                       \
        {   /* refill the slot consumed at the previous step; never issue a copy that would not be consumed */ \
            const long long an = a + (TMA_NS - 1);                                                          \
            if (an <= a_last)                                                                               \
                tma_issue_row(A, Q, off_run, (int)((step + TMA_NS - 1) & (TMA_NS - 1)), lane);              \
            if (off_run < off_max) off_run += A.pitch_in;   /* clamped at the last array row */             \
        }                                                                                                   \
        march_compute<R, DIV, RL, PROJ, EOS, true, J>(A, T, P, rho, ua, ut, E, a, dt, inv_dx, emit, kc + J, m1, stage); \
        a++; step++;                                                                                        \
    }

#pragma unroll 1
    for (long long it = 0; it < n_iter; it++) {
        const bool emit = it >= 2;
        const int kc = (int)(it & 1) * 4;
        TMA_STEP(0)
        TMA_STEP(1)
        TMA_STEP(2)
        TMA_STEP(3)
        if (A.transpose_out && emit && (it & 1)) flush_stage(A, stage, w0, a - 12, m1);
    }
#undef TMA_STEP

    if (DIV == DIV_FLAGGED) {
        // see sweep_kernel: threads whose operands left the proven range of the branch-free division recompute
        // their segment with nvcc's full IEEE division (register-prefetch path, direct stores)
        range_check_dividend(dt.v, T.flag);
        if (T.flag.bad() && T.valid) {
            T.amax = 0ULL; T.tmax = 0ULL;
            march_segment<R, DIV_IEEE, RL, PROJ, EOS, false>(A, T, dt, m0, m1, w0, stage);
            if ((threadIdx.x & 31) == __ffs(__activemask()) - 1) atomicAdd(&A.ts->redo_count, 1u);
        }
        __syncwarp();
    }

    unsigned long long am = T.amax, tm = T.tmax;
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
