// sweep_staged_common.cuh -- shared pieces of the shared-memory staged marching kernels (sweep_fast_kernel.cuh): the CTA
// size, the per-thread cp.async copy helpers of the staging variants for rows that are not 16-byte aligned
// (STG_CPA16 / STG_CPA8), and the chunk-granular range bookkeeping of the strict (bit-exact) arithmetic.
//
// (Rounds 1 and 2 had a kernel of their own here, the unskewed cp.async-staged strict kernel: one dependency chain of
// ~250 FP64 operations per step, 0.345 of the HBM peak at 16384^2.  The strict arithmetic now runs on the four-chain
// schedule of the fast kernel -- sweep_fast_kernel<..., MATH_STRICT> -- and that kernel is gone.)
#pragma once

#include "sweep_kernel.cuh"

#ifndef ASYNC_TPB_VALUE
#define ASYNC_TPB_VALUE 128
#endif
constexpr int ASYNC_TPB = ASYNC_TPB_VALUE;   // threads (= columns) per CTA; warps are independent of one another

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

// Strict-mode bookkeeping, evaluated once per chunk of SWEEP_CHUNK emitted cells.  An operand outside the proven range of
// the branch-free division (common.cuh) met while chunk k is being emitted can only reach cells emitted in chunks
// k .. k+2 (a step's chains run up to 11 steps ahead of the cell it emits): the thread appends
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
