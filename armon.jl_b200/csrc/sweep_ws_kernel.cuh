// sweep_ws_kernel.cuh -- warp-specialised version of the fused axis-sweep marching kernel.
//
// Same mathematics, same data layout and same HBM traffic as sweep_kernel.cuh (read it first), but the per-column
// pipeline is cut in two halves that run in two different warps of the same CTA:
//
//   producer warp : loads cell a (prefetched 4 rows ahead), EOS(a), Godunov interface a, GAD flux a-1,
//                   Lagrangian cell a-2, and publishes that cell (dxl, rho, ua, ut, E, rho*{ua,ut,E}, c, dt*flux)
//                   in a shared-memory ring;
//   consumer warp : advection flux a-3 and projection of cell a-4 from the ring, transposed staging / stores,
//                   dtCFL accumulators.
//
// Lane l of the consumer only ever reads what lane l of the producer wrote, so the ring needs no bank-conflict
// care; the two warps hand slots over with one "full" and one "empty" mbarrier per slot (arrive = release, wait =
// acquire).  Splitting halves the live state of each thread (<= 128 registers instead of 255), which doubles the
// resident warps per SM and doubles the independent FP64 work the schedulers can interleave: the single-role
// kernel is latency-bound at 2 warps per scheduler ("wait" stalls), this one runs 4.
//
// Threads whose operands leave the range in which the branch-free division is exact (common.cuh) do not touch the
// CFL accumulators and append (column, segment) to a work list; sweep_fixup_kernel recomputes those columns with
// nvcc's full IEEE division afterwards, so the strict mode stays bit-identical to IEEE for every operand.
#pragma once

#include "sweep_kernel.cuh"

constexpr int WS_NS = 8;     // ring slots (interfaces/cells in flight between the two warps); the consumer keeps 5 live
constexpr int WS_NV = 10;    // doubles per slot and lane
constexpr int WS_TPB = 64;   // one producer warp + one consumer warp, 32 columns per CTA
// Slot k as published by the producer (flux at interface k, cell k) ...
enum { WP_FU = 0, WP_FP = 1, WP_DM = 2, WP_UA = 3, WP_UT = 4, WP_E = 5, WP_C = 6 };
// ... and after the consumer turned it, in place, into the Lagrangian cell k
enum { WV_DISP = 0, WV_DXL = 1, WV_LR = 2, WV_LU = 3, WV_LT = 4, WV_LE = 5, WV_C = 6, WV_LRU = 7, WV_LRT = 8, WV_LRE = 9 };

struct WsShared {
    double ring[WS_NS][WS_NV][32];
    __align__(16) double stage[4 * 32 * SWEEP_STAGE_PITCH];
    unsigned long long full[WS_NS], empty[WS_NS], fin;
    unsigned pflag[32];
};

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WS_WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WS_DONE_%=;\n"
        "bra WS_WAIT_%=;\n"
        "WS_DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity)
        : "memory");
}

// Producer-side rolling window: cells a, a-1, a-2 and their Godunov interfaces.
template <class R> struct PipeP {
    R cu[4], cp[4], crc[4], cdm[4], cut[4], cE[4], cc[4];
    R Gu[4], Gp[4];
};

// One producer step: consumes cell a; EOS(a), Godunov interface a, flux at interface k = a-1, which is published
// together with cell k (publish index j = a - (m0-1) when >= 0).
template <class R, int DIV, int RL, int EOS, int J>
__device__ __forceinline__ void ws_producer_step(const SweepArgs &A, SweepThread &T, PipeP<R> &P, double (&in)[4][4],
                                                 const long long a, const long long a_last, const long long m0,
                                                 const R dt, WsShared &S)
{
    typedef Div<R, DIV> D;
    constexpr int S0 = J & 3, S1 = (J + 3) & 3, S2 = (J + 2) & 3;
    const R dx(A.dx);
    RangeFlag &f = T.flag;
    const int lane = threadIdx.x & 31;

    R rho(in[J][0]), ua(in[J][1]), ut(in[J][2]), E(in[J][3]);
    {
        const long long an = a + 4 > a_last ? a_last : a + 4;
        issue_loads(A, T, an, in[J]);
    }
    R p, c;
    eos_eval<R, DIV, EOS>(A, rho, ua, ut, E, p, c, f);
    const R rc = rho * c;
    P.cu[S0] = ua; P.cp[S0] = p; P.crc[S0] = rc; P.cdm[S0] = rho * dx; P.cut[S0] = ut; P.cE[S0] = E; P.cc[S0] = c;

    acoustic_godunov<R, DIV>(P.crc[S1], rc, P.cu[S1], ua, P.cp[S1], p, P.Gu[S0], P.Gp[S0], f);

    R Fu, Fp;
    if (RL == 0) {   // acoustic!  src/riemann_schemes.jl:33-43
        Fu = P.Gu[S1];
        Fp = P.Gp[S1];
    } else {         // acoustic_GAD!  src/riemann_schemes.jl:55-104
        constexpr int LIM = RL - 1;
        const R u_i = P.cu[S1], u_im = P.cu[S2], p_i = P.cp[S1], p_im = P.cp[S2];
        const R us_i = P.Gu[S1], ps_i = P.Gp[S1];
        R r_um(1.), r_pm(1.), r_up(1.), r_pp(1.);
        if (LIM != ARMON_LIMITER_NONE) {
            r_um = limiter<R, LIM>(D::div(P.Gu[S0] - u_i, (us_i - u_im) + R(1e-6), f));
            r_pm = limiter<R, LIM>(D::div(P.Gp[S0] - p_i, (ps_i - p_im) + R(1e-6), f));
            r_up = limiter<R, LIM>(D::div(u_im - P.Gu[S2], (u_i - us_i) + R(1e-6), f));
            r_pp = limiter<R, LIM>(D::div(p_im - P.Gp[S2], (p_i - ps_i) + R(1e-6), f));
        }
        const R Dm = (P.cdm[S2] + P.cdm[S1]) * R(0.5);
        const R theta = R(0.5) * (R(1.) - ((P.crc[S2] + P.crc[S1]) * R(0.5)) * D::div(dt, Dm, f));
        Fu = us_i + theta * (r_up * (u_i - us_i) - r_um * (us_i - u_im));
        Fp = ps_i + theta * (r_pp * (p_i - ps_i) - r_pm * (ps_i - p_im));
    }

    const long long j = a - (m0 - 1);
    if (j >= 0) {
        const int s = (int)(j & (WS_NS - 1));
        if (j >= WS_NS) mbar_wait(&S.empty[s], (unsigned)(((j >> 3) - 1) & 1));
        double *slot = &S.ring[s][0][lane];
        slot[WP_FU * 32] = Fu.v;
        slot[WP_FP * 32] = Fp.v;
        slot[WP_DM * 32] = P.cdm[S1].v;
        slot[WP_UA * 32] = P.cu[S1].v;
        slot[WP_UT * 32] = P.cut[S1].v;
        slot[WP_E * 32] = P.cE[S1].v;
        slot[WP_C * 32] = P.cc[S1].v;
        mbar_arrive(&S.full[s]);
    }
}

// One consumer step, kn = newest interface needed (publish index jn = kn - (m0-2)):
//   Lagrangian update of cell kn-1 (src/kernels.jl:58-68), written back in place into its ring slot,
//   advection flux at interface kn-2 (src/projection_schemes.jl:62-124) once kn-2 >= m0,
//   projection of cell kn-3 (src/projection_schemes.jl:23-41) once kn-3 >= m0.
// Fk*: flux of interface kn-1 kept from the previous step (Fu, Fp, Fp*Fu).
template <class R, int DIV, int PROJ>
__device__ __forceinline__ void ws_consumer_step(const SweepArgs &A, SweepThread &T, const long long kn,
                                                 const long long m0, const long long m1, const R dt,
                                                 R &Fku, R &Fkp, R &Fkpu, R &Ar, R &Aru, R &Art, R &ArE,
                                                 const typename Div<R, DIV>::Rcp &inv_dx, WsShared &S)
{
    typedef Div<R, DIV> D;
    const R dx(A.dx);
    RangeFlag &f = T.flag;
    const int lane = threadIdx.x & 31;
    const long long jn = kn - (m0 - 2);           // publish index of interface/cell kn (>= 1)
    mbar_wait(&S.full[jn & (WS_NS - 1)], (unsigned)((jn >> 3) & 1));
#define WS_SLOT(off) (&S.ring[(jn + (off)) & (WS_NS - 1)][0][lane])

    // ---- Lagrangian cell k = kn-1 ----
    {
        const double *sn = WS_SLOT(0);
        double *sk = WS_SLOT(-1);
        const R Fnu(sn[WP_FU * 32]), Fnp(sn[WP_FP * 32]);
        const R Fnpu = Fnp * Fnu;
        const R dm(sk[WP_DM * 32]), ua(sk[WP_UA * 32]), Lt(sk[WP_UT * 32]), E(sk[WP_E * 32]);
        const R dxl = dx + dt * (Fnu - Fku);
        const R dtdm = D::div(dt, dm, f);
        const R Lr = D::div(dm, dxl, f);
        const R Lu = ua + dtdm * (Fkp - Fnp);
        const R LE = E + dtdm * (Fkpu - Fnpu);
        sk[WV_DISP * 32] = (dt * Fku).v;
        sk[WV_DXL * 32] = dxl.v;
        sk[WV_LR * 32] = Lr.v;
        sk[WV_LU * 32] = Lu.v;
        sk[WV_LE * 32] = LE.v;
        sk[WV_LRU * 32] = (Lr * Lu).v;
        sk[WV_LRT * 32] = (Lr * Lt).v;
        sk[WV_LRE * 32] = (Lr * LE).v;
        Fku = Fnu; Fkp = Fnp; Fkpu = Fnpu;
    }
    const long long is = kn - 2;                  // advection interface; cells is-2 .. is+1 are Lagrangian now
    if (is >= m0) {
        const R d(WS_SLOT(-2)[WV_DISP * 32]);
        const bool pos = d.v > 0.0;
        const int io = pos ? -3 : -2;             // slot offset of the upwind cell i = is-1 or is
        R Anr, Anru, Anrt, AnrE;
        if (PROJ == ARMON_PROJ_EULER_2ND) {
            const R dxe = pos ? -(dx - R(WS_SLOT(-3)[WV_DISP * 32])) : dx + R(WS_SLOT(-1)[WV_DISP * 32]);
            const double *cm = WS_SLOT(io - 1), *c0 = WS_SLOT(io), *cp = WS_SLOT(io + 1);
            const R dxl_m(cm[WV_DXL * 32]), dxl_0(c0[WV_DXL * 32]), dxl_p(cp[WV_DXL * 32]);
            const R two_dxl = R(2.) * dxl_0;
            const R r_m = D::div(two_dxl, dxl_0 + dxl_m, f);
            const R r_p = D::div(two_dxl, dxl_0 + dxl_p, f);
            const R lf = D::div(dxe, two_dxl, f);
#define WS_ADVECT(var, res)                                                                           \
            {                                                                                         \
                const R q0(c0[(var) * 32]);                                                           \
                res = d * (q0 - slope_minmod_fused<R>(R(cm[(var) * 32]), q0, R(cp[(var) * 32]), r_m, r_p) * lf); \
            }
            WS_ADVECT(WV_LR, Anr)
            WS_ADVECT(WV_LRU, Anru)
            WS_ADVECT(WV_LRT, Anrt)
            WS_ADVECT(WV_LRE, AnrE)
#undef WS_ADVECT
        } else {
            const double *c0 = WS_SLOT(io);
            Anr = d * R(c0[WV_LR * 32]);
            Anru = d * R(c0[WV_LRU * 32]);
            Anrt = d * R(c0[WV_LRT * 32]);
            AnrE = d * R(c0[WV_LRE * 32]);
        }

        const long long mm = kn - 3;              // projected cell
        if (mm >= m0) {
            const double *cm_ = WS_SLOT(-3);
            const R dXr = R(cm_[WV_DXL * 32]) * R(cm_[WV_LR * 32]);
            R t_r = dXr - (Anr - Ar);
            R t_ru = dXr * R(cm_[WV_LU * 32]) - (Anru - Aru);
            R t_rt = dXr * R(cm_[WV_LT * 32]) - (Anrt - Art);
            R t_rE = dXr * R(cm_[WV_LE * 32]) - (AnrE - ArE);
            if (A.dx_pow2) {
                const R idx(A.inv_dx);
                t_r = t_r * idx; t_ru = t_ru * idx; t_rt = t_rt * idx; t_rE = t_rE * idx;
            } else {
                t_r = D::quot(t_r, inv_dx, f); t_ru = D::quot(t_ru, inv_dx, f);
                t_rt = D::quot(t_rt, inv_dx, f); t_rE = D::quot(t_rE, inv_dx, f);
            }
            const typename D::Rcp inv_r = D::prepare(t_r, f);
            const R o_ua = D::quot(t_ru, inv_r, f), o_ut = D::quot(t_rt, inv_r, f), o_E = D::quot(t_rE, inv_r, f);
            const bool store = T.valid && mm < m1;
            if (store) {
                const R c_out(cm_[WV_C * 32]);
                const unsigned long long ba = (unsigned long long)__double_as_longlong((rabs(o_ua) + c_out).v);
                const unsigned long long bt = (unsigned long long)__double_as_longlong((rabs(o_ut) + c_out).v);
                T.amax = ba > T.amax ? ba : T.amax;
                T.tmax = bt > T.tmax ? bt : T.tmax;
            }
            if (A.transpose_out) {
                double *s = S.stage + lane * SWEEP_STAGE_PITCH + (int)((mm - m0) & (SWEEP_CHUNK - 1));
                s[0 * 32 * SWEEP_STAGE_PITCH] = t_r.v;
                s[1 * 32 * SWEEP_STAGE_PITCH] = o_ua.v;
                s[2 * 32 * SWEEP_STAGE_PITCH] = o_ut.v;
                s[3 * 32 * SWEEP_STAGE_PITCH] = o_E.v;
            } else if (store) {
                const long long o = (mm + A.g) * A.pitch_out + T.col;
                A.out[0][o] = t_r.v;
                A.out[1][o] = o_ua.v;
                A.out[2][o] = o_ut.v;
                A.out[3][o] = o_E.v;
            }
        }
        Ar = Anr; Aru = Anru; Art = Anrt; ArE = AnrE;
    }
#undef WS_SLOT
    if (jn >= 4) mbar_arrive(&S.empty[(jn - 4) & (WS_NS - 1)]);   // interface/cell kn-4 is dead from now on
}

struct FixupArgs {
    unsigned *count;             // work-list length of THIS sweep
    unsigned *count_next;        // counter of the next sweep, cleared by the fix-up kernel
    unsigned long long *list;    // (segment << 32) | column
};

#ifndef WS_MIN_BLOCKS
#define WS_MIN_BLOCKS 8
#endif

template <class R, int DIV, int RL, int PROJ, int EOS>
__global__ void __launch_bounds__(WS_TPB, WS_MIN_BLOCKS) sweep_ws_kernel(const SweepArgs A, const FixupArgs F)
{
    __shared__ __align__(16) WsShared S;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long w0 = (long long)blockIdx.x * 32;
    const long long w = w0 + lane;
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
            for (long long m = m0 + warp; m < m1; m += 2) {
                const long long i = (m + A.g) * A.pitch_in + T.col;
                const long long o = A.transpose_out ? T.col * A.pitch_out + (m + A.g) : (m + A.g) * A.pitch_out + T.col;
#pragma unroll
                for (int k = 0; k < 4; k++) A.out[k][o] = A.in[k][i];
            }
        }
        return;
    }
    if (threadIdx.x == 0) {
        for (int s = 0; s < WS_NS; s++) { mbar_init(&S.full[s], 32); mbar_init(&S.empty[s], 32); }
        mbar_init(&S.fin, 32);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const R dt = R(ts->current_dt) * R(A.dt_factor);
    const long long nchunks = (m1 - m0 + SWEEP_CHUNK - 1) / SWEEP_CHUNK;
    const long long m_end = m0 + nchunks * SWEEP_CHUNK;

    if (warp == 0) {
        // ------------------------------------------- producer -------------------------------------------
        PipeP<R> P;
        double in[4][4];
        const long long a_begin = m0 - 4, a_last = m_end + 3;
#pragma unroll
        for (int j = 0; j < 4; j++) issue_loads(A, T, a_begin + j, in[j]);
#pragma unroll
        for (int j = 0; j < 4; j++) {
            P.cu[j] = R(0.); P.cp[j] = R(1.); P.crc[j] = R(1.); P.cdm[j] = R(1.); P.cut[j] = R(0.); P.cE[j] = R(1.);
            P.cc[j] = R(1.); P.Gu[j] = R(0.); P.Gp[j] = R(1.);
        }
        const long long n_iter = 2 + 2 * nchunks;
        long long a = a_begin;
#pragma unroll 1
        for (long long it = 0; it < n_iter; it++) {
            ws_producer_step<R, DIV, RL, EOS, 0>(A, T, P, in, a + 0, a_last, m0, dt, S);
            ws_producer_step<R, DIV, RL, EOS, 1>(A, T, P, in, a + 1, a_last, m0, dt, S);
            ws_producer_step<R, DIV, RL, EOS, 2>(A, T, P, in, a + 2, a_last, m0, dt, S);
            ws_producer_step<R, DIV, RL, EOS, 3>(A, T, P, in, a + 3, a_last, m0, dt, S);
            a += 4;
        }
        if (DIV == DIV_FLAGGED) {
            range_check_dividend(dt.v, T.flag);
            S.pflag[lane] = T.flag.bad() ? 1u : 0u;
            mbar_arrive(&S.fin);
        }
    } else {
        // ------------------------------------------- consumer -------------------------------------------
        const typename Div<R, DIV>::Rcp inv_dx = Div<R, DIV>::prepare(R(A.dx), T.flag);
        R Ar(0.), Aru(0.), Art(0.), ArE(0.);
        // flux of interface m0-2 (publish index 0), the first one that is valid
        mbar_wait(&S.full[0], 0u);
        R Fku(S.ring[0][WP_FU][lane]), Fkp(S.ring[0][WP_FP][lane]);
        R Fkpu = Fkp * Fku;
#pragma unroll 1
        for (long long kn = m0 - 1; kn < m_end + 3; kn++) {
            ws_consumer_step<R, DIV, PROJ>(A, T, kn, m0, m1, dt, Fku, Fkp, Fkpu, Ar, Aru, Art, ArE, inv_dx, S);
            const long long mm = kn - 3;
            if (A.transpose_out && mm >= m0 && ((mm - m0) & (SWEEP_CHUNK - 1)) == SWEEP_CHUNK - 1)
                flush_stage(A, S.stage, w0, mm - (SWEEP_CHUNK - 1), m1);
        }
        bool bad = false;
        if (DIV == DIV_FLAGGED) {
            mbar_wait(&S.fin, 0u);
            bad = T.flag.bad() || S.pflag[lane] != 0u;
            if (bad && T.valid) {
                const unsigned e = atomicAdd(F.count, 1u);
                F.list[e] = ((unsigned long long)sweep_segment_index(A) << 32) | (unsigned long long)(unsigned)w;
            }
        }
        unsigned long long am = bad ? 0ULL : T.amax, tm = bad ? 0ULL : T.tmax;
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
}

// Recomputes the work-listed (column, segment) pairs with nvcc's full IEEE division (one thread per entry).
template <class R, int RL, int PROJ, int EOS>
__global__ void __launch_bounds__(32) sweep_fixup_kernel(const SweepArgs A, const FixupArgs F)
{
    if (blockIdx.x == 0 && threadIdx.x == 0) *F.count_next = 0u;
    const unsigned raw_count = *F.count;
    const unsigned count = (A.fix_rows > 0 && raw_count > A.fix_cap) ? A.fix_cap : raw_count;
    const DeviceTimeState *ts = A.ts;
    if (count == 0u || ts->done) return;
    const R dt = R(ts->current_dt) * R(A.dt_factor);
    for (unsigned e = blockIdx.x * blockDim.x + threadIdx.x; e < count; e += gridDim.x * blockDim.x) {
        const unsigned long long entry = F.list[e];
        const long long w = (long long)(entry & 0xffffffffULL), hi = (long long)(entry >> 32);
        // ws kernel: the entry names a whole march segment; cp.async kernels: fix_rows rows starting at row `hi`
        const long long m0 = A.fix_rows > 0 ? hi : hi * A.seg;
        const long long len = A.fix_rows > 0 ? A.fix_rows : A.seg;
        const long long m1 = (m0 + len < A.nm) ? m0 + len : A.nm;
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
