// sweep_async2_kernel.cuh -- cp.async-staged marching kernel with a software-pipelined (skewed) step.
//
// Same mathematics, data layout, HBM traffic and input staging as sweep_async_kernel.cuh (read it and sweep_kernel.cuh
// first).  What changes is the schedule inside a thread.  In the plain march every stage of a step consumes what the
// previous stage produced in the same step (EOS(a) -> Godunov(a) -> GAD(a-1) -> Lagrange(a-2) -> slopes(a-3) ->
// advection(a-3) -> projection(a-4)): one dependent chain of ~100 FP64 operations per step, and the round-1 profile
// shows the warps waiting on it at the head of every step (shared-memory load -> reciprocal -> square root).  Here the
// head of the chain runs one step ahead of the rest:
//
//   stage A (cell a)        : EOS(a), Godunov interface a                       -- uses nothing produced in this step
//   stages B-E (one behind) : GAD flux a-2, Lagrangian cell a-3, slopes of cell a-4, advection flux a-4,
//                             projection of cell a-5                            -- use Godunov states up to a-1 only
//
// so that the scheduler always has an independent 45-instruction chain to interleave with the other ~260.
// A2_SKEW = 1 additionally runs stage B one step ahead of C-E (three chains; measured slower: 252 registers).
// Results of the stages that run ahead are committed to the register rings at the end of the step.
// Registers: the skew costs one more live slot of (ua, p, rho*c, rho*dx, Godunov state); to pay for it the values that
// merely ride along the pipeline are left in shared memory instead: ut and E of a cell are re-read from the input ring
// when the Lagrangian update (3 steps later) and the projection (5 steps later) need them, and the sound speed goes
// through a small per-warp ring.  The input ring therefore keeps 6 consumed rows and has 16 slots.
#pragma once

#include "sweep_async_kernel.cuh"

#ifndef A2_SKEW
#define A2_SKEW 0                // 0: only EOS+Godunov run ahead (measured best: 0.99 ms); 1: the GAD flux also runs ahead of the Lagrangian update (252 registers, 1.01 ms)
#endif
constexpr int A2_Q = A2_SKEW;
constexpr int A2_NS = 16;            // input ring slots per warp
constexpr int A2_KEEP = 6 + A2_Q;    // consumed rows that must stay readable (rows a .. a-5-q)
constexpr int A2_CS = 8;             // sound-speed ring slots (written at a, read at a-5-q)

struct Async2WarpShared {
    double ring[A2_NS][4][32];                             // [slot][variable][lane]
    double cring[A2_CS][32];                               // c of cells a .. a-7
    double stage[4 * 32 * SWEEP_STAGE_PITCH];              // transposed-store staging (flush_stage)
};

// Rolling window of the skewed march, 4-slot rings indexed by (cell index mod 4) with compile-time slots.
template <class R> struct Pipe2 {
    R cu[4], cp[4], crc[4], cdm[4];                         // cells a .. a-3: ua, p, rho*c, rho*dx
    R Gu[4], Gp[4];                                         // Godunov state of interfaces a .. a-3
    R Fu[4], Fp[4], FpFu[4];                                // flux used (GAD or Godunov) of interfaces a-2, a-3, and p*u
    R disp[4];                                              // dt * Fu of interfaces a-2 .. a-5
    R dxl[4], Lr[4], Lu[4], LE[4], Lru[4], Lrt[4], LrE[4];  // Lagrangian cells a-3 .. a-5
    R rdxl[4];                                              // fast mode: 1 / dxl of the same cells (reused by the remap)
    R rrho[4];                                              // fast mode: 1 / rho of cells a .. a-3 (EOS reciprocal reused for dt/dm)
    R Ar, Aru, Art, ArE;                                    // advection flux of the previous interface
    R Sr, Sru, Srt, SrE, S2b, S2r;                          // slopes, 2*dxl and its reciprocal of cell a-5
    R Rsum;                                                 // fast mode: 1 / (dxl[a-5] + dxl[a-4]), shared by two cells
    R Dr, Dru, Drt, DrE;                                    // fast mode: q[a-4] - q[a-5] of the four remapped quantities
};

// One skewed march step.  `step` = a - a_begin; J = step & 3 is static.  `ring` / `cring` point at this lane's column
// of slot 0.  TR / EMIT as in march_compute (always compile-time here).
template <class R, int DIV, int RL, int PROJ, int EOS, int J, int TR, int EMIT>
__device__ __forceinline__ void march_compute2(const SweepArgs &A, SweepThread &T, Pipe2<R> &P, const double *ring,
                                               double *cring, const unsigned step, const long long a, const R dt,
                                               const typename Div<R, DIV>::Rcp &inv_dx, const int k_chunk,
                                               const long long m1, double *stage)
{
    typedef Div<R, DIV> D;
    // slot of cell / interface a-k in the 4-entry rings
#define ZS(k) ((J + 8 - (k)) & 3)
    constexpr int q = A2_Q;
    constexpr int Z0 = ZS(0), Z1 = ZS(1), Z2 = ZS(2), Z3 = ZS(3);
    constexpr int ZL = ZS(3 + q), ZLn = ZS(2 + q);   // Lagrangian cell a-3-q: its own slot, slot of its right interface
    constexpr int ZM = ZS(5 + q), ZC = ZS(4 + q), ZP = ZS(3 + q);   // slope / advection: cells a-5-q, a-4-q, a-3-q
    const R dx(A.dx);
    RangeFlag &f = T.flag;
    const double *row0 = ring + ((step) & (A2_NS - 1)) * 128;
    const double *rowL = ring + ((step - 3 - q) & (A2_NS - 1)) * 128;
    const double *rowE = ring + ((step - 5 - q) & (A2_NS - 1)) * 128;

    // ---- stage A, cell a: EOS (src/kernels.jl:4-55), Godunov state at interface a (src/riemann_schemes.jl:21-30) ----
    // results are committed to the rings at the end of the step: with q = 1 the slots still hold cell a-4
    R A_ua, A_p, A_rc, A_dm, A_Gu, A_Gp, A_rrho(0.);
    {
        const R rho(row0[0]), ua(row0[32]), ut(row0[64]), E(row0[96]);
        R p, c;
        eos_eval<R, DIV, EOS>(A, rho, ua, ut, E, p, c, f);
        if (DIV == DIV_FAST) A_rrho = R(rcp_fast(rho.v));   // the very reciprocal the EOS just formed (common subexpression)
        const R rc = rho * c;
        cring[(step & (A2_CS - 1)) * 32] = c.v;
        acoustic_godunov<R, DIV>(P.crc[Z1], rc, P.cu[Z1], ua, P.cp[Z1], p, A_Gu, A_Gp, f);
        A_ua = ua; A_p = p; A_rc = rc; A_dm = rho * dx;
    }

    // ---- stage B, flux at interface i = a-2 (cells a-3, a-2); Godunov states a-3, a-2, a-1 come from earlier steps ----
    R B_Fu, B_Fp;
    if (RL == 0) {   // acoustic!  src/riemann_schemes.jl:33-43
        B_Fu = P.Gu[Z2];
        B_Fp = P.Gp[Z2];
    } else {         // acoustic_GAD!  src/riemann_schemes.jl:55-104
        constexpr int LIM = RL - 1;
        const R u_i = P.cu[Z2], u_im = P.cu[Z3], p_i = P.cp[Z2], p_im = P.cp[Z3];
        const R us_i = P.Gu[Z2], ps_i = P.Gp[Z2];
        R r_um(1.), r_pm(1.), r_up(1.), r_pp(1.);
        if (LIM != ARMON_LIMITER_NONE) {   // limiter(r, NoLimiter) == 1 whatever r is (src/limiters.jl:6)
            r_um = limiter<R, LIM>(D::div(P.Gu[Z1] - u_i, (us_i - u_im) + R(1e-6), f));
            r_pm = limiter<R, LIM>(D::div(P.Gp[Z1] - p_i, (ps_i - p_im) + R(1e-6), f));
            r_up = limiter<R, LIM>(D::div(u_im - P.Gu[Z3], (u_i - us_i) + R(1e-6), f));
            r_pp = limiter<R, LIM>(D::div(p_im - P.Gp[Z3], (p_i - ps_i) + R(1e-6), f));
        }
        R theta;
        if (DIV == DIV_FAST) {   // 0.5 * (1 - (a/2) * (dt / (d/2))) == 0.5 - 0.5 * dt * a / d : 7 operations instead of 11
            const double a_ = P.crc[Z3].v + P.crc[Z2].v, d_ = P.cdm[Z3].v + P.cdm[Z2].v;
            theta = R(fma(a_ * (-0.5 * dt.v), rcp_fast(d_), 0.5));
        } else {
            const R Dm = (P.cdm[Z3] + P.cdm[Z2]) * R(0.5);                               // (dm_l + dm_r) / 2
            theta = R(0.5) * (R(1.) - ((P.crc[Z3] + P.crc[Z2]) * R(0.5)) * D::div(dt, Dm, f));
        }
        B_Fu = us_i + theta * (r_up * (u_i - us_i) - r_um * (us_i - u_im));
        B_Fp = ps_i + theta * (r_pp * (p_i - ps_i) - r_pm * (ps_i - p_im));
    }
    const R B_FpFu = B_Fp * B_Fu;
    const R B_disp = dt * B_Fu;
    if (q == 0) {   // the Lagrangian update below uses this flux right away
        P.Fu[Z2] = B_Fu; P.Fp[Z2] = B_Fp; P.FpFu[Z2] = B_FpFu; P.disp[Z2] = B_disp;
    }

    // ---- stage C, Lagrangian update of cell k = a-3-q: src/kernels.jl:58-68 (ut, E of the cell re-read from the ring) ----
    {
        const R dxl = dx + dt * (P.Fu[ZLn] - P.Fu[ZL]);
        const R dm = P.cdm[ZL];
        const R dtdm = DIV == DIV_FAST ? R((dt.v * inv_dx.r) * P.rrho[ZL].v)   // dt / (rho dx) with the EOS's 1 / rho
                                       : D::div(dt, dm, f);
        R Lr;
        if (DIV == DIV_FAST) {
            const double r_ = rcp_fast(dxl.v);
            P.rdxl[ZL] = R(r_);
            Lr = R(dm.v * r_);
        } else {
            Lr = D::div(dm, dxl, f);
        }
        const R Lu = P.cu[ZL] + dtdm * (P.Fp[ZL] - P.Fp[ZLn]);
        const R LE = R(rowL[96]) + dtdm * (P.FpFu[ZL] - P.FpFu[ZLn]);
        const R Lt(rowL[64]);
        P.dxl[ZL] = dxl; P.Lr[ZL] = Lr; P.Lu[ZL] = Lu; P.LE[ZL] = LE;
        P.Lru[ZL] = Lr * Lu; P.Lrt[ZL] = Lr * Lt; P.LrE[ZL] = Lr * LE;
    }

    // ---- stage D, advection flux at interface is = a-4-q: src/projection_schemes.jl:62-124 (see march_compute) ----
    // cells a-5-q -> ZM, a-4-q -> ZC, a-3-q -> ZP ; disp of the same interfaces in the same slots
    R Anr, Anru, Anrt, AnrE;
    {
        const R d = P.disp[ZC];
        const bool pos = d.v > 0.0;
        if (PROJ == ARMON_PROJ_EULER_2ND) {
            const R dxl_m = P.dxl[ZM], dxl_0 = P.dxl[ZC], dxl_p = P.dxl[ZP];
            const R two_dxl = R(2.) * dxl_0;
            typename D::Rcp k2;
            R sr, sru, srt, srE;
            if (DIV == DIV_FAST) {
                // 1 / (dxl_i + dxl_{i+1}) serves r+ of cell i and r- of cell i+1, q_{i+1} - q_i serves both slopes, and
                // 1 / (2 dxl) comes from the Lagrangian stage: 1 reciprocal and 4 differences per step instead of 3 and 8
                const double rsum_p = rcp_fast(dxl_0.v + dxl_p.v);
                const R r_m(two_dxl.v * P.Rsum.v), r_p(two_dxl.v * rsum_p);
                const R dr = P.Lr[ZP] - P.Lr[ZC], dru = P.Lru[ZP] - P.Lru[ZC], drt = P.Lrt[ZP] - P.Lrt[ZC], drE = P.LrE[ZP] - P.LrE[ZC];
                sr = minmod_of(r_p * dr, r_m * P.Dr);
                sru = minmod_of(r_p * dru, r_m * P.Dru);
                srt = minmod_of(r_p * drt, r_m * P.Drt);
                srE = minmod_of(r_p * drE, r_m * P.DrE);
                P.Rsum = R(rsum_p); P.Dr = dr; P.Dru = dru; P.Drt = drt; P.DrE = drE;
                k2.b = two_dxl.v; k2.r = 0.5 * P.rdxl[ZC].v;
            } else {
                const R r_m = D::div(two_dxl, dxl_0 + dxl_m, f);
                const R r_p = D::div(two_dxl, dxl_0 + dxl_p, f);
                k2 = D::prepare(two_dxl, f);
                sr = slope_minmod_fused<R>(P.Lr[ZM], P.Lr[ZC], P.Lr[ZP], r_m, r_p);
                sru = slope_minmod_fused<R>(P.Lru[ZM], P.Lru[ZC], P.Lru[ZP], r_m, r_p);
                srt = slope_minmod_fused<R>(P.Lrt[ZM], P.Lrt[ZC], P.Lrt[ZP], r_m, r_p);
                srE = slope_minmod_fused<R>(P.LrE[ZM], P.LrE[ZC], P.LrE[ZP], r_m, r_p);
            }

            const R dxe = rsel(pos, -(dx - P.disp[ZM]), dx + P.disp[ZP]);
            typename D::Rcp ksel;
            ksel.b = pos ? P.S2b.v : k2.b;
            ksel.r = pos ? P.S2r.v : k2.r;
            const R lf = D::quot(dxe, ksel, f);
            Anr = d * (rsel(pos, P.Lr[ZM], P.Lr[ZC]) - rsel(pos, P.Sr, sr) * lf);
            Anru = d * (rsel(pos, P.Lru[ZM], P.Lru[ZC]) - rsel(pos, P.Sru, sru) * lf);
            Anrt = d * (rsel(pos, P.Lrt[ZM], P.Lrt[ZC]) - rsel(pos, P.Srt, srt) * lf);
            AnrE = d * (rsel(pos, P.LrE[ZM], P.LrE[ZC]) - rsel(pos, P.SrE, srE) * lf);
            P.Sr = sr; P.Sru = sru; P.Srt = srt; P.SrE = srE;
            P.S2b = R(k2.b); P.S2r = R(k2.r);
        } else {
            Anr = d * rsel(pos, P.Lr[ZM], P.Lr[ZC]);
            Anru = d * rsel(pos, P.Lru[ZM], P.Lru[ZC]);
            Anrt = d * rsel(pos, P.Lrt[ZM], P.Lrt[ZC]);
            AnrE = d * rsel(pos, P.LrE[ZM], P.LrE[ZC]);
        }
    }

    // ---- stage E, projection of cell k = a-5-q: src/projection_schemes.jl:23-41 ----
    if (EMIT == 1) {
        const R dXr = P.dxl[ZM] * P.Lr[ZM];
        const R Lt(rowE[64]);
        const R c_out(cring[((step - 5 - q) & (A2_CS - 1)) * 32]);
        R t_r = dXr - (Anr - P.Ar);
        R t_ru = dXr * P.Lu[ZM] - (Anru - P.Aru);
        R t_rt = dXr * Lt - (Anrt - P.Art);
        R t_rE = dXr * P.LE[ZM] - (AnrE - P.ArE);
        if (DIV == DIV_FAST) {
            // u = (t_ru / dx) / (t_r / dx) = t_ru / t_r: only the density needs the 1/dx factor
        } else if (A.dx_pow2) {   // x / dx == x * (1/dx) bit for bit when dx is a power of two
            const R idx(A.inv_dx);
            t_r = t_r * idx; t_ru = t_ru * idx; t_rt = t_rt * idx; t_rE = t_rE * idx;
        } else {
            t_r = D::quot(t_r, inv_dx, f); t_ru = D::quot(t_ru, inv_dx, f);
            t_rt = D::quot(t_rt, inv_dx, f); t_rE = D::quot(t_rE, inv_dx, f);
        }
        const typename D::Rcp inv_r = D::prepare(t_r, f);
        const R o_ua = D::quot(t_ru, inv_r, f), o_ut = D::quot(t_rt, inv_r, f), o_E = D::quot(t_rE, inv_r, f);
        if (DIV == DIV_FAST) t_r = t_r * R(inv_dx.r);
        const long long m = a - 5 - q;
        const bool store = T.valid && m < m1;
        {   // dtCFL accumulators (src/reductions.jl:14-20), branch-free: cells that are not stored do not contribute
            const unsigned long long ba = (unsigned long long)__double_as_longlong((rabs(o_ua) + c_out).v);
            const unsigned long long bt = (unsigned long long)__double_as_longlong((rabs(o_ut) + c_out).v);
            T.amax = (store && ba > T.amax) ? ba : T.amax;   // `store` rides on the predicate input of the compare
            T.tmax = (store && bt > T.tmax) ? bt : T.tmax;
        }
        if (TR == 1) {
            double *s = stage + (threadIdx.x & 31) * SWEEP_STAGE_PITCH + k_chunk;
            s[0 * 32 * SWEEP_STAGE_PITCH] = t_r.v;
            s[1 * 32 * SWEEP_STAGE_PITCH] = o_ua.v;
            s[2 * 32 * SWEEP_STAGE_PITCH] = o_ut.v;
            s[3 * 32 * SWEEP_STAGE_PITCH] = o_E.v;
        } else if (store) {
            const long long o = (m + A.g) * A.pitch_out + T.col;
            A.out[0][o] = t_r.v;
            A.out[1][o] = o_ua.v;
            A.out[2][o] = o_ut.v;
            A.out[3][o] = o_E.v;
        }
    }
    P.Ar = Anr; P.Aru = Anru; P.Art = Anrt; P.ArE = AnrE;

    // ---- commit the results of the stages that ran ahead (their ring slots were still being read above) ----
    P.cu[Z0] = A_ua; P.cp[Z0] = A_p; P.crc[Z0] = A_rc; P.cdm[Z0] = A_dm; P.Gu[Z0] = A_Gu; P.Gp[Z0] = A_Gp;
    if (DIV == DIV_FAST) P.rrho[Z0] = A_rrho;
    if (q == 1) {
        P.Fu[Z2] = B_Fu; P.Fp[Z2] = B_Fp; P.FpFu[Z2] = B_FpFu; P.disp[Z2] = B_disp;
    }
#undef ZS
}

#ifndef ASYNC2_MIN_BLOCKS
#define ASYNC2_MIN_BLOCKS (256 / ASYNC_TPB_VALUE)   // 8 warps per SM
#endif

template <class R, int DIV, int RL, int PROJ, int EOS, int TR>
__global__ void __launch_bounds__(ASYNC_TPB, ASYNC2_MIN_BLOCKS) sweep_async2_kernel(const SweepArgs A)
{
    extern __shared__ __align__(128) unsigned char async2_smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    Async2WarpShared &S = reinterpret_cast<Async2WarpShared *>(async2_smem_raw)[warp];

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
        // benign finite state everywhere (columns past the end of the row are never copied; rows "before" the first one
        // are read by the lagging stages during warm-up): rho = E = 1, u = v = 0, c = 1
        for (int k = lane; k < A2_NS * 4 * 32; k += 32) (&S.ring[0][0][0])[k] = ((k >> 5) & 3) == 0 || ((k >> 5) & 3) == 3 ? 1.0 : 0.0;
        for (int k = lane; k < A2_CS * 32; k += 32) (&S.cring[0][0])[k] = 1.0;
        __syncwarp();
    }
    // Ring protocol: step t consumes row a_begin + t from slot t mod 16 and keeps rows t .. t-KEEP+1 readable; it
    // refills the slot of row t-KEEP with row t + LEAD.  Prologue: rows 0 .. LEAD-1, one commit group per row; one group
    // per step afterwards, so the group of row t is complete once at most LEAD - 1 groups are pending.
    constexpr int LEAD = A2_NS - A2_KEEP;   // 10 - q
#pragma unroll 1
    for (int s = 0; s < LEAD; s++) {
        async_issue_row(L, march_row_offset(A, a_begin + s), s);
        async_commit();
    }
    // rows past the last array row (tail of the last segment) are not fetched: their slots keep older, finite rows,
    // which only feed cells that are never emitted
    long long off_run = (a_begin + LEAD + A.g) * A.pitch_in;            // offset of row a + LEAD
    int rows_left = (int)((A.nm + A.g - 1) - (a_begin + LEAD));         // >= 0 while row a + LEAD exists

    const typename Div<R, DIV>::Rcp inv_dx = Div<R, DIV>::prepare(R(A.dx), T.flag);
    ChunkFix C;
    C.tot_a = 0ULL; C.tot_t = 0ULL; C.taint = 0;
    if (DIV == DIV_FLAGGED) range_check_dividend(dt.v, T.flag);
    C.always = DIV == DIV_FLAGGED && T.flag.bad();
    Pipe2<R> P;
#pragma unroll
    for (int j = 0; j < 4; j++) {
        P.cu[j] = R(0.); P.cp[j] = R(1.); P.crc[j] = R(1.); P.cdm[j] = R(1.);
        P.Gu[j] = R(0.); P.Gp[j] = R(1.); P.Fu[j] = R(0.); P.Fp[j] = R(1.); P.FpFu[j] = R(0.); P.disp[j] = R(0.);
        P.dxl[j] = R(1.); P.Lr[j] = R(1.); P.Lu[j] = R(0.); P.LE[j] = R(1.);
        P.Lru[j] = R(0.); P.Lrt[j] = R(0.); P.LrE[j] = R(1.); P.rdxl[j] = R(1.); P.rrho[j] = R(1.);
    }
    P.Ar = R(0.); P.Aru = R(0.); P.Art = R(0.); P.ArE = R(0.);
    P.Sr = R(0.); P.Sru = R(0.); P.Srt = R(0.); P.SrE = R(0.); P.S2b = R(2.); P.S2r = R(0.5);
    P.Rsum = R(0.5); P.Dr = R(0.); P.Dru = R(0.); P.Drt = R(0.); P.DrE = R(0.);

    double *stage = S.stage;
    const double *ring = &S.ring[0][0][lane];
    double *cring = &S.cring[0][lane];
    long long a = a_begin;
    unsigned step = 0;   // a - a_begin

#define A2_STEP(J, EMIT, KC)                                                                                \
    {                                                                                                       \
        async_wait<LEAD - 1>();                                                                             \
        __syncwarp();   /* every lane's copies of row a have landed; the slot refilled below was last read a step ago */ \
        async_issue_row(L, off_run, (int)((step + LEAD) & (A2_NS - 1)), rows_left >= 0);                    \
        async_commit();                                                                                     \
        off_run += A.pitch_in; rows_left--;                                                                 \
        march_compute2<R, DIV, RL, PROJ, EOS, J, TR, EMIT>(A, T, P, ring, cring, step, a, dt, inv_dx, KC, m1, stage); \
        a++; step++;                                                                                        \
    }

    // warm-up: 9 + q steps fill the dependency cone of the first output (emitted at a = m0 + 5 + q), nothing is emitted
#pragma unroll 1
    for (int it = 0; it < 2; it++) {
        A2_STEP(0, 0, 0)
        A2_STEP(1, 0, 0)
        A2_STEP(2, 0, 0)
        A2_STEP(3, 0, 0)
    }
    A2_STEP(0, 0, 0)
    if (A2_Q == 1) A2_STEP(1, 0, 0)
    // steady state: every iteration emits 4 cells, every second one flushes the transposed staging tile
#pragma unroll 1
    for (long long it = 0; it < 2 * nchunks; it++) {
        const int kc = (int)(it & 1) * 4;
        A2_STEP((1 + A2_Q) & 3, 1, kc + 0)
        A2_STEP((2 + A2_Q) & 3, 1, kc + 1)
        A2_STEP((3 + A2_Q) & 3, 1, kc + 2)
        A2_STEP((4 + A2_Q) & 3, 1, kc + 3)
        if (TR == 1 && (it & 1)) flush_stage(A, stage, w0, a - 13 - A2_Q, m1);
        if (it & 1) chunk_end<DIV>(A, T, C, a - 13 - A2_Q, w);
    }
#undef A2_STEP
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
