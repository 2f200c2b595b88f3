// sweep_fixup_kernel.cuh -- IEEE recomputation of the work-listed column chunks of a strict-mode sweep.
//
// The strict mode of the cp.async-staged kernels divides with a branch-free correctly rounded sequence that is only
// proven for operands inside a range (common.cuh).  A thread that meets an operand outside it appends
// (first row << 32 | column) to a work list (chunk_end, sweep_staged_common.cuh); this kernel recomputes `fix_rows`
// rows of each listed column with nvcc's full IEEE division (march_segment<DIV_IEEE>, direct stores), one thread per
// entry, so that the strict mode is bit-identical to IEEE for every operand.
#pragma once

#include "sweep_kernel.cuh"

struct FixupArgs {
    unsigned *count;             // work-list length of THIS sweep
    unsigned *count_next;        // counter of the next sweep, cleared here
    unsigned long long *list;    // (first row << 32) | column
};

template <class R, int RL, int PROJ, int EOS>
__global__ void __launch_bounds__(32) sweep_fixup_kernel(const SweepArgs A, const FixupArgs F)
{
    if (blockIdx.x == 0 && threadIdx.x == 0) *F.count_next = 0u;
    const unsigned raw_count = *F.count;
    const unsigned count = raw_count > A.fix_cap ? A.fix_cap : raw_count;
    const DeviceTimeState *ts = A.ts;
    if (count == 0u || ts->done) return;
    const R dt = R(ts->current_dt) * R(A.dt_factor);
    for (unsigned e = blockIdx.x * blockDim.x + threadIdx.x; e < count; e += gridDim.x * blockDim.x) {
        const unsigned long long entry = F.list[e];
        const long long w = (long long)(entry & 0xffffffffULL), m0 = (long long)(entry >> 32);
        const long long m1 = (m0 + A.fix_rows < A.nm) ? m0 + A.fix_rows : A.nm;
        SweepThread T;
        T.valid = true;
        T.col = w + A.g;
#pragma unroll
        for (int k = 0; k < 4; k++) T.base[k] = A.in[k] + T.col;
        T.amax = 0ULL; T.tmax = 0ULL;
        march_segment<R, DIV_IEEE, RL, PROJ, EOS, false>(A, T, dt, m0, m1, 0, nullptr);
        atomicMax(&A.ts->acc[A.acc_slot][0], T.amax);
        atomicMax(&A.ts->acc[A.acc_slot][1], T.tmax);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(&A.ts->redo_count, count);
}
