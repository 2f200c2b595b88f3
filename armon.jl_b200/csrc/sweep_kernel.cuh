// sweep_kernel.cuh -- the fused axis-sweep "marching" kernel (the product's hot path).
//
// One launch replaces, for one axis sweep of solver_cycle (src/solver.jl:300-317):
//   update_EOS!            (src/kernels.jl:4-55)
//   boundary_conditions!   (src/halo_exchange.jl:2-36)      -- O(perimeter) ghost-row fill kernel launched just before
//   numerical_fluxes!      (src/riemann_schemes.jl:21-123)  -- acoustic / acoustic_GAD + limiter
//   cell_update!           (src/kernels.jl:58-68)
//   advection_fluxes!      (src/projection_schemes.jl:62-124)
//   euler_projection!      (src/projection_schemes.jl:23-41)
// and accumulates the maxima needed by the next cycle's dtCFL reduction (src/reductions.jl:2-20).
//
// Data layout.  Arrays are [march axis][contiguous axis] including g ghosts on every side: the swept axis is always
// the strided one.  The canonical BlockData layout (rows = y) is therefore the right input for a Y sweep; an X
// sweep reads the transposed layout.  A sweep writes its result either in the layout it read or transposed
// (`transpose_out`), so that alternating X/Y sweeps never need a separate transposition pass.
//
// Parallelisation.  thread <-> one column of the contiguous axis (coalesced 8-byte loads, 256 B per warp and
// array row); blockIdx.y <-> a segment of `seg` cells of the march axis.  Each thread marches along its column
// keeping the whole 9-cell dependency cone in a rolling register window: at the step that
// loads cell a it evaluates EOS(a), the Godunov interface a, the GAD flux a-1, the Lagrangian cell a-2, the
// advection flux a-3 and projects cell a-4.  Every input is read once (plus 8 warm-up rows per segment), every
// interface/cell quantity is computed once, nothing intermediate touches memory: 64 B of HBM traffic per cell.
// Inputs are prefetched 4 rows ahead (16 independent 8-byte loads in flight per thread).
//
// The expression order of the reference source is kept; with R = sd every operation is an
// explicitly rounded IEEE operation and the result is bit-identical to the strict CPU oracle.  Only rewrites that
// are exact in IEEE arithmetic are applied: x/2 -> x*0.5, x/dx -> x*(1/dx) when dx is a power of two,
// s*x with s = sign(..) in {-1,0,1} -> sign-bit flips, max(|u+c|,|u-c|) -> |u|+c for c >= 0, one refined reciprocal
// shared by the quotients that have the same divisor, and a branch-free correctly rounded division (common.cuh).
#pragma once

#include "common.cuh"

constexpr int SWEEP_TPB = 128;      // threads per CTA = columns per CTA
constexpr int SWEEP_CHUNK = 8;      // outputs between two flushes of the transposed staging buffer
constexpr int SWEEP_STAGE_PITCH = SWEEP_CHUNK + 2;   // rows stay 16-byte aligned for the 128-bit reads of flush_stage (2-way write conflicts)

struct SweepArgs {
    const double *in[4];    // rho, ua, ut, E (ua: velocity along the march axis, ut: transverse velocity)
    double       *out[4];
    long long nm, nw;       // real cells along the march axis / along the contiguous axis
    long long pitch_in;     // nw + 2g
    long long pitch_out;    // transposed output: nm + 2g, else nw + 2g
    int g;
    int seg;                // outputs per march segment, multiple of SWEEP_CHUNK
    int nseg;               // number of march segments of the sweep, ceil(nm / seg)
    int y_base, y_jump;     // march segment handled by a CTA = y_base + blockIdx.y * y_jump (interior / edge launches)
    int transpose_out;
    int mirror_lo, mirror_hi;   // 1: global edge (ghost rows written by k_bc_fill); 0: ghost rows hold the neighbour's cells
    double bc_a_lo, bc_t_lo, bc_a_hi, bc_t_hi;   // velocity factors of boundary_condition(test, side), used by k_bc_fill
    double dx, inv_dx;      // cell size along the march axis
    int dx_pow2;            // inv_dx is exact: x/dx == x*inv_dx bit for bit
    double dt_factor;       // axis-splitting factor (src/axis_splitting.jl:24-46)
    double gamma;
    double gm1, ggm1;       // gamma - 1, gamma (gamma - 1): host-side constants of the fast-mode perfect-gas EOS
    DeviceTimeState *ts;
    int acc_slot;
    // strict mode of the cp.async kernels: (segment << 32 | column) of the threads whose operands left the proven range
    // of the branch-free division; sweep_fixup_kernel recomputes them with nvcc's full IEEE division afterwards
    unsigned *fix_count;
    unsigned long long *fix_list;
    unsigned fix_cap;       // capacity of fix_list (entries); an overflow raises ARMON_ERR_RANGE
    int fix_rows;           // 0: an entry is (segment << 32 | column); > 0: (first row << 32 | column), fix_rows rows
    // per-cycle diagnostics fused into the last sweep of a cycle (sweep_fast_kernel<..., CONS = 1>): one partial
    // (sum rho, sum rho*E) per warp, slot (segment * gridDim.x + blockIdx.x) * warps per CTA + warp
    double *cons_m, *cons_e;
};

// index of the march segment of this CTA: a sweep is one launch over all segments, or -- when the ghost rows of a
// side are still in flight from a neighbour rank -- an interior launch (segments 1 .. n-2) that overlaps the halo
// exchange and an edge launch (segments 0 and n-1) after it
__device__ __forceinline__ long long sweep_segment_index(const SweepArgs &A)
{
    return (long long)A.y_base + (long long)blockIdx.y * A.y_jump;
}

// exact s*x for s = sign(src) in {-1, +1}: flip the sign bit of x when src is negative
__device__ __forceinline__ double flip_sign_by(double x, double src)
{
    const unsigned long long sb = (unsigned long long)__double_as_longlong(src) & 0x8000000000000000ULL;
    return __longlong_as_double((long long)((unsigned long long)__double_as_longlong(x) ^ sb));
}

// src/projection_schemes.jl:15-20: s * max(0, min(s*d_p, s*d_m)) with s = sign(d_p), i.e. the operand of smaller
// magnitude when d_p and d_m have the same sign and 0 otherwise (d_p == 0 or d_m == 0 give 0 either way).  One FP64
// compare on the magnitudes, a sign test on the high words and two 64-bit selects; same values as the oracle's
// expression for every finite input (the sign of a zero result is not significant anywhere on this path).
template <class R> __device__ __forceinline__ R slope_minmod_fused(R qm, R q0, R qp, R r_m, R r_p)
{
    const R d_p = r_p * (qp - q0);
    const R d_m = r_m * (q0 - qm);
    const bool m_smaller = fabs(d_m.v) < fabs(d_p.v);
    const double m = m_smaller ? d_m.v : d_p.v;
    const int sx = __double2hiint(d_p.v) ^ __double2hiint(d_m.v);
    return R(sx < 0 ? 0.0 : m);
}

// the selection part of slope_minmod_fused on already formed limited differences
template <class R> __device__ __forceinline__ R minmod_of(R d_p, R d_m)
{
    const bool m_smaller = fabs(d_m.v) < fabs(d_p.v);
    const double m = m_smaller ? d_m.v : d_p.v;
    const int sx = __double2hiint(d_p.v) ^ __double2hiint(d_m.v);
    return R(sx < 0 ? 0.0 : m);
}

template <class R, int DIV, int EOS>
__device__ __forceinline__ void eos_eval(const SweepArgs &A, R rho, R ua, R ut, R E, R &p, R &c, RangeFlag &f)
{
    if (EOS == ARMON_EOS_BIZARRIUM) {
        R g;
        eos_bizarrium<R, DIV, false>(rho, ua, ut, E, p, c, g, f);
    } else {
        eos_perfect_gas<R, DIV>(R(A.gamma), rho, ua, ut, E, p, c, f);
    }
}

// Rolling window of the march, 4-slot rings indexed by (cell index mod 4) with compile-time slots.
template <class R> struct Pipe {
    R cu[4], cp[4], crc[4], cdm[4], cut[4], cE[4], cc[4];   // cells: ua, p, rho*c, rho*dx, ut, E, c
    R Gu[4], Gp[4];                                         // first-order (Godunov) interface state
    R Fu[4], Fp[4], FpFu[4];                                // flux actually used (GAD or Godunov), and p*u
    R disp[4];                                              // dt * Fu
    R dxl[4];                                               // Lagrangian cell width dx + dt*(Fu[k+1]-Fu[k])
    R Lr[4], Lu[4], Lt[4], LE[4], Lru[4], Lrt[4], LrE[4];   // Lagrangian cell: rho, ua, ut, E and rho*{ua,ut,E}
    R Ar, Aru, Art, ArE;                                    // advection flux of the previous interface
    R Sr, Sru, Srt, SrE;                                    // limited slopes of rho, rho*{ua,ut,E} of cell a-4 (2nd order remap)
    R S2b, S2r;                                             // 2*dxl of cell a-4 and its reciprocal (division policy)
};

struct SweepThread {
    long long col;         // element offset of this thread's column inside a row (w + g)
    bool      valid;       // column holds a real cell
    const double *base[4]; // A.in[k] + col
    unsigned long long amax, tmax;   // dt accumulators (integer images of non-negative doubles)
    double cm, ce;         // conservation sums of the cells this thread stored: sum rho, sum rho*E (fast kernel, CONS)
    RangeFlag flag;        // range bookkeeping of the branch-free divisions (DIV_FLAGGED)
};

// element offset of the first cell of array row `a` (march index, ghost rows included: -g <= a < nm + g), clamped so
// that prefetches past the last needed row and the tail of a ragged last chunk stay inside the array.  Ghost rows are
// real data when the sweep starts: the neighbour's cells (halo exchange) or the mirrored boundary cells written by
// k_bc_fill (boundary_conditions!, src/halo_exchange.jl:2-29).  Warp-uniform.
__device__ __forceinline__ long long march_row_offset(const SweepArgs &A, long long a)
{
    const long long rmax = A.nm + A.g - 1, rmin = -(long long)A.g;
    const long long r = a > rmax ? rmax : (a < rmin ? rmin : a);
    return (r + A.g) * A.pitch_in;
}

__device__ __forceinline__ void issue_loads(const SweepArgs &A, const SweepThread &T, long long a, double v[4])
{
    const long long off = march_row_offset(A, a);
#pragma unroll
    for (int k = 0; k < 4; k++) v[k] = __ldg(T.base[k] + off);
}

// The arithmetic of one march step: consumes cell a (rho, ua, ut, E as read from memory), emits cell a-4 when `emit`.
// J = (a - a_begin) & 3 is static.  TR / EMIT: -1 = decided at run time (A.transpose_out / `emit`), 0 / 1 = known at
// compile time, which makes the whole step one basic block for the instruction scheduler (no uniform branches).
template <class R, int DIV, int RL, int PROJ, int EOS, bool STAGED, int J, int TR = -1, int EMIT = -1>
__device__ __forceinline__ void march_compute(const SweepArgs &A, SweepThread &T, Pipe<R> &P, R rho, R ua, R ut, R E,
                                              const long long a, const R dt,
                                              const typename Div<R, DIV>::Rcp &inv_dx, const bool emit,
                                              const int k_chunk, const long long m1, double *stage)
{
    typedef Div<R, DIV> D;
    constexpr int S0 = J & 3, S1 = (J + 3) & 3, S2 = (J + 2) & 3, S3 = (J + 1) & 3;
    const R dx(A.dx);
    RangeFlag &f = T.flag;

    // ---- cell a: EOS (src/kernels.jl:4-55) ----
    const R c_out = P.cc[S0];   // c of cell a-4 (EOS at the start of this sweep), read before the slot is reused
    R p, c;
    eos_eval<R, DIV, EOS>(A, rho, ua, ut, E, p, c, f);
    const R rc = rho * c;
    P.cu[S0] = ua; P.cp[S0] = p; P.crc[S0] = rc; P.cdm[S0] = rho * dx; P.cut[S0] = ut; P.cE[S0] = E; P.cc[S0] = c;

    // ---- Godunov state at interface a (cells a-1, a): src/riemann_schemes.jl:21-30 ----
    acoustic_godunov<R, DIV>(P.crc[S1], rc, P.cu[S1], ua, P.cp[S1], p, P.Gu[S0], P.Gp[S0], f);

    // ---- flux at interface i = a-1 (cells a-2, a-1) ----
    if (RL == 0) {   // acoustic!  src/riemann_schemes.jl:33-43
        P.Fu[S1] = P.Gu[S1];
        P.Fp[S1] = P.Gp[S1];
    } else {         // acoustic_GAD!  src/riemann_schemes.jl:55-104
        constexpr int LIM = RL - 1;
        const R u_i = P.cu[S1], u_im = P.cu[S2], p_i = P.cp[S1], p_im = P.cp[S2];
        const R us_i = P.Gu[S1], ps_i = P.Gp[S1];
        R r_um(1.), r_pm(1.), r_up(1.), r_pp(1.);
        if (LIM != ARMON_LIMITER_NONE) {   // limiter(r, NoLimiter) == 1 whatever r is (src/limiters.jl:6)
            r_um = limiter<R, LIM>(D::div(P.Gu[S0] - u_i, (us_i - u_im) + R(1e-6), f));
            r_pm = limiter<R, LIM>(D::div(P.Gp[S0] - p_i, (ps_i - p_im) + R(1e-6), f));
            r_up = limiter<R, LIM>(D::div(u_im - P.Gu[S2], (u_i - us_i) + R(1e-6), f));
            r_pp = limiter<R, LIM>(D::div(p_im - P.Gp[S2], (p_i - ps_i) + R(1e-6), f));
        }
        const R Dm = (P.cdm[S2] + P.cdm[S1]) * R(0.5);                                   // (dm_l + dm_r) / 2
        const R theta = R(0.5) * (R(1.) - ((P.crc[S2] + P.crc[S1]) * R(0.5)) * D::div_pos(dt, Dm, f));
        P.Fu[S1] = us_i + theta * (r_up * (u_i - us_i) - r_um * (us_i - u_im));
        P.Fp[S1] = ps_i + theta * (r_pp * (p_i - ps_i) - r_pm * (ps_i - p_im));
    }
    P.FpFu[S1] = P.Fp[S1] * P.Fu[S1];
    P.disp[S1] = dt * P.Fu[S1];

    // ---- Lagrangian update of cell k = a-2: src/kernels.jl:58-68 ----
    {
        const R dxl = dx + dt * (P.Fu[S1] - P.Fu[S2]);
        const R dm = P.cdm[S2];
        const R dtdm = D::div_pos(dt, dm, f);
        const R Lr = D::div_pos(dm, dxl, f);
        const R Lu = P.cu[S2] + dtdm * (P.Fp[S2] - P.Fp[S1]);
        const R LE = P.cE[S2] + dtdm * (P.FpFu[S2] - P.FpFu[S1]);
        const R Lt = P.cut[S2];
        P.dxl[S2] = dxl; P.Lr[S2] = Lr; P.Lu[S2] = Lu; P.LE[S2] = LE; P.Lt[S2] = Lt;
        P.Lru[S2] = Lr * Lu; P.Lrt[S2] = Lr * Lt; P.LrE[S2] = Lr * LE;
    }

    // ---- advection flux at interface is = a-3: src/projection_schemes.jl:62-124 ----
    // ring slots: cells a-4 -> S0, a-3 -> S3, a-2 -> S2 ; disp(a-4) -> S0, disp(a-3) -> S3, disp(a-2) -> S2
    // The upwind cell i of the interface is a-4 (disp > 0) or a-3.  Everything the reference computes "relative to the
    // shifted i" (projection_schemes.jl:99-121) only depends on the cell i: the width ratios r-, r+, the limited slopes
    // and 2*dxl_i.  They are evaluated once per CELL (for a-3 here; those of a-4 were kept from the previous step),
    // and the interface selects the upwind cell's (q, slope, 2*dxl): 10 selects instead of 16, same expressions.
    R Anr, Anru, Anrt, AnrE;
    {
        const R d = P.disp[S3];
        const bool pos = d.v > 0.0;
        if (PROJ == ARMON_PROJ_EULER_2ND) {
            const R dxl_m = P.dxl[S0], dxl_0 = P.dxl[S3], dxl_p = P.dxl[S2];
            const R two_dxl = R(2.) * dxl_0;
            const R r_m = D::div_pos(two_dxl, dxl_0 + dxl_m, f);
            const R r_p = D::div_pos(two_dxl, dxl_0 + dxl_p, f);
            const typename D::Rcp k2 = D::prepare_pos(two_dxl, f);
            const R sr = slope_minmod_fused<R>(P.Lr[S0], P.Lr[S3], P.Lr[S2], r_m, r_p);
            const R sru = slope_minmod_fused<R>(P.Lru[S0], P.Lru[S3], P.Lru[S2], r_m, r_p);
            const R srt = slope_minmod_fused<R>(P.Lrt[S0], P.Lrt[S3], P.Lrt[S2], r_m, r_p);
            const R srE = slope_minmod_fused<R>(P.LrE[S0], P.LrE[S3], P.LrE[S2], r_m, r_p);

            const R dxe = rsel(pos, -(dx - P.disp[S0]), dx + P.disp[S2]);
            typename D::Rcp ksel;
            ksel.b = pos ? P.S2b.v : k2.b;
            ksel.r = pos ? P.S2r.v : k2.r;
            const R lf = D::quot(dxe, ksel, f);
            Anr = d * (rsel(pos, P.Lr[S0], P.Lr[S3]) - rsel(pos, P.Sr, sr) * lf);
            Anru = d * (rsel(pos, P.Lru[S0], P.Lru[S3]) - rsel(pos, P.Sru, sru) * lf);
            Anrt = d * (rsel(pos, P.Lrt[S0], P.Lrt[S3]) - rsel(pos, P.Srt, srt) * lf);
            AnrE = d * (rsel(pos, P.LrE[S0], P.LrE[S3]) - rsel(pos, P.SrE, srE) * lf);
            P.Sr = sr; P.Sru = sru; P.Srt = srt; P.SrE = srE;
            P.S2b = R(k2.b); P.S2r = R(k2.r);
        } else {
            Anr = d * rsel(pos, P.Lr[S0], P.Lr[S3]);
            Anru = d * rsel(pos, P.Lru[S0], P.Lru[S3]);
            Anrt = d * rsel(pos, P.Lrt[S0], P.Lrt[S3]);
            AnrE = d * rsel(pos, P.LrE[S0], P.LrE[S3]);
        }
    }

    // ---- projection of cell k = a-4: src/projection_schemes.jl:23-41 ----
    if (EMIT == 1 || (EMIT == -1 && emit)) {
        const bool transpose_out = TR == -1 ? (A.transpose_out != 0) : (TR == 1);
        const R dXr = P.dxl[S0] * P.Lr[S0];
        R t_r = dXr - (Anr - P.Ar);
        R t_ru = dXr * P.Lu[S0] - (Anru - P.Aru);
        R t_rt = dXr * P.Lt[S0] - (Anrt - P.Art);
        R t_rE = dXr * P.LE[S0] - (AnrE - P.ArE);
        if (DIV == DIV_FAST) {   // the refined reciprocal of a power of two is exact: no need to tell the cases apart
            const R idx(inv_dx.r);
            t_r = t_r * idx; t_ru = t_ru * idx; t_rt = t_rt * idx; t_rE = t_rE * idx;
        } else if (A.dx_pow2) {   // x / dx == x * (1/dx) bit for bit when dx is a power of two
            const R idx(A.inv_dx);
            t_r = t_r * idx; t_ru = t_ru * idx; t_rt = t_rt * idx; t_rE = t_rE * idx;
        } else {
            t_r = D::quot(t_r, inv_dx, f); t_ru = D::quot(t_ru, inv_dx, f);
            t_rt = D::quot(t_rt, inv_dx, f); t_rE = D::quot(t_rE, inv_dx, f);
        }
        const typename D::Rcp inv_r = D::prepare_pos(t_r, f);
        const R o_ua = D::quot(t_ru, inv_r, f), o_ut = D::quot(t_rt, inv_r, f), o_E = D::quot(t_rE, inv_r, f);
        const long long m = a - 4;
        const bool store = T.valid && m < m1;
        // dtCFL accumulators (src/reductions.jl:14-20): max(|u+c|,|u-c|) == |u|+c, new velocities, c of this sweep's EOS
        {   // branch-free: cells that are not stored do not contribute
            const unsigned long long ba = (unsigned long long)__double_as_longlong((rabs(o_ua) + c_out).v);
            const unsigned long long bt = (unsigned long long)__double_as_longlong((rabs(o_ut) + c_out).v);
            T.amax = (store && ba > T.amax) ? ba : T.amax;   // `store` rides on the predicate input of the compare
            T.tmax = (store && bt > T.tmax) ? bt : T.tmax;
        }
        if (STAGED && transpose_out) {
            // stage[var][lane][k]: flushed as rows of SWEEP_CHUNK contiguous doubles by flush_stage()
            double *s = stage + (threadIdx.x & 31) * SWEEP_STAGE_PITCH + k_chunk;
            s[0 * 32 * SWEEP_STAGE_PITCH] = t_r.v;
            s[1 * 32 * SWEEP_STAGE_PITCH] = o_ua.v;
            s[2 * 32 * SWEEP_STAGE_PITCH] = o_ut.v;
            s[3 * 32 * SWEEP_STAGE_PITCH] = o_E.v;
        } else if (store) {
            const long long o = transpose_out ? T.col * A.pitch_out + (m + A.g) : (m + A.g) * A.pitch_out + T.col;
            A.out[0][o] = t_r.v;
            A.out[1][o] = o_ua.v;
            A.out[2][o] = o_ut.v;
            A.out[3][o] = o_E.v;
        }
    }
    P.Ar = Anr; P.Aru = Anru; P.Art = Anrt; P.ArE = AnrE;
}

// One march step of the register-prefetch variant: cell a is already in `in[J]`; the slot is refilled with the cell
// consumed 4 steps from now.
template <class R, int DIV, int RL, int PROJ, int EOS, bool STAGED, int J>
__device__ __forceinline__ void march_step(const SweepArgs &A, SweepThread &T, Pipe<R> &P, double (&in)[4][4],
                                           const long long a, const long long a_last, const R dt,
                                           const typename Div<R, DIV>::Rcp &inv_dx, const bool emit,
                                           const int k_chunk, const long long m1, double *stage)
{
    const R rho(in[J][0]), ua(in[J][1]), ut(in[J][2]), E(in[J][3]);
    {
        const long long an = a + 4 > a_last ? a_last : a + 4;
        issue_loads(A, T, an, in[J]);
    }
    march_compute<R, DIV, RL, PROJ, EOS, STAGED, J>(A, T, P, rho, ua, ut, E, a, dt, inv_dx, emit, k_chunk, m1, stage);
}

// Transposed store of one chunk: the warp's staging buffer holds, per variable, 32 columns x SWEEP_CHUNK march
// cells; each column becomes a row of the transposed array, written as SWEEP_CHUNK contiguous doubles.
// Fast path (full tile, even output pitch): 128-bit shared loads and global stores, 8 rows per warp instruction,
// 16 store instructions per tile.  Ragged tiles (last columns / last chunk of a segment, odd pitch) take the
// element-wise path.
__device__ __forceinline__ void flush_stage(const SweepArgs &A, const double *stage, long long w0, long long mb, long long m1)
{
    const int lane = threadIdx.x & 31;
    __syncwarp();
    if (w0 + 32 <= A.nw && mb + SWEEP_CHUNK <= m1 && !(A.pitch_out & 1)) {
        const int r0 = lane >> 2, j = lane & 3;
        const double2 *src = reinterpret_cast<const double2 *>(stage + r0 * SWEEP_STAGE_PITCH + 2 * j);
        const long long off = (w0 + r0 + A.g) * A.pitch_out + (mb + A.g) + 2 * j;
        const long long step = 8 * A.pitch_out;
#pragma unroll 1
        for (int v = 0; v < 4; v++) {
            double *dst = A.out[v] + off;
#pragma unroll
            for (int it = 0; it < 4; it++) {
                const double2 val = src[(v * 32 + it * 8) * SWEEP_STAGE_PITCH / 2];
                *reinterpret_cast<double2 *>(dst + it * step) = val;
            }
        }
    } else {
        const int rsub = lane / SWEEP_CHUNK, col = lane % SWEEP_CHUNK;
#pragma unroll
        for (int v = 0; v < 4; v++) {
#pragma unroll
            for (int it = 0; it < 32 / (32 / SWEEP_CHUNK); it++) {
                const int r = it * (32 / SWEEP_CHUNK) + rsub;
                const double val = stage[(v * 32 + r) * SWEEP_STAGE_PITCH + col];
                const long long w = w0 + r, m = mb + col;
                if (w < A.nw && m < m1) A.out[v][(w + A.g) * A.pitch_out + (m + A.g)] = val;
            }
        }
    }
    __syncwarp();
}

// The march of one thread over one segment [m0, m1) of its column.  STAGED: transposed outputs go through the
// warp's shared staging buffer (warp-collective; every lane of the warp must take part); otherwise every store is
// a direct one (used by the per-thread IEEE recomputation).
template <class R, int DIV, int RL, int PROJ, int EOS, bool STAGED>
__device__ __forceinline__ void march_segment(const SweepArgs &A, SweepThread &T, const R dt, const long long m0,
                                              const long long m1, const long long w0, double *stage)
{
    const typename Div<R, DIV>::Rcp inv_dx = Div<R, DIV>::prepare(R(A.dx), T.flag);
    Pipe<R> P;
    double in[4][4];
    const long long a_begin = m0 - 4;
    const long long nchunks = (m1 - m0 + SWEEP_CHUNK - 1) / SWEEP_CHUNK;
    const long long a_last = m0 + nchunks * SWEEP_CHUNK + 3;   // last cell index consumed (clamped inside march_row_offset)

#pragma unroll
    for (int j = 0; j < 4; j++) issue_loads(A, T, a_begin + j, in[j]);

    // the pipeline registers start with finite dummies: warm-up results are never emitted
#pragma unroll
    for (int j = 0; j < 4; j++) {
        P.cu[j] = R(0.); P.cp[j] = R(1.); P.crc[j] = R(1.); P.cdm[j] = R(1.); P.cut[j] = R(0.); P.cE[j] = R(1.); P.cc[j] = R(1.);
        P.Gu[j] = R(0.); P.Gp[j] = R(1.); P.Fu[j] = R(0.); P.Fp[j] = R(1.); P.FpFu[j] = R(0.); P.disp[j] = R(0.);
        P.dxl[j] = R(1.); P.Lr[j] = R(1.); P.Lu[j] = R(0.); P.Lt[j] = R(0.); P.LE[j] = R(1.);
        P.Lru[j] = R(0.); P.Lrt[j] = R(0.); P.LrE[j] = R(1.);
    }
    P.Ar = R(0.); P.Aru = R(0.); P.Art = R(0.); P.ArE = R(0.);
    P.Sr = R(0.); P.Sru = R(0.); P.Srt = R(0.); P.SrE = R(0.); P.S2b = R(2.); P.S2r = R(0.5);

    // One loop body of 4 steps (the ring period).  The first 2 iterations only fill the dependency cone of the
    // first output (8 warm-up cells); afterwards every iteration emits 4 cells and every second one flushes the
    // transposed staging buffer.
    long long a = a_begin;
    const long long n_iter = 2 + 2 * nchunks;
#pragma unroll 1
    for (long long it = 0; it < n_iter; it++) {
        const bool emit = it >= 2;
        const int kc = (int)(it & 1) * 4;
        march_step<R, DIV, RL, PROJ, EOS, STAGED, 0>(A, T, P, in, a + 0, a_last, dt, inv_dx, emit, kc + 0, m1, stage);
        march_step<R, DIV, RL, PROJ, EOS, STAGED, 1>(A, T, P, in, a + 1, a_last, dt, inv_dx, emit, kc + 1, m1, stage);
        march_step<R, DIV, RL, PROJ, EOS, STAGED, 2>(A, T, P, in, a + 2, a_last, dt, inv_dx, emit, kc + 2, m1, stage);
        march_step<R, DIV, RL, PROJ, EOS, STAGED, 3>(A, T, P, in, a + 3, a_last, dt, inv_dx, emit, kc + 3, m1, stage);
        a += 4;
        if (STAGED && A.transpose_out && emit && (it & 1)) flush_stage(A, stage, w0, a - 12, m1);
    }
}

#ifndef SWEEP_MIN_BLOCKS
#define SWEEP_MIN_BLOCKS 2
#endif

template <class R, int DIV, int RL, int PROJ, int EOS>
__global__ void __launch_bounds__(SWEEP_TPB, SWEEP_MIN_BLOCKS) sweep_kernel(const SweepArgs A)
{
    __shared__ __align__(16) double stage_all[(SWEEP_TPB / 32) * 4 * 32 * SWEEP_STAGE_PITCH];
    double *stage = stage_all + (threadIdx.x / 32) * (4 * 32 * SWEEP_STAGE_PITCH);

    const long long w = (long long)blockIdx.x * SWEEP_TPB + threadIdx.x;
    const long long w0 = (long long)blockIdx.x * SWEEP_TPB + (threadIdx.x & ~31);
    const long long m0 = sweep_segment_index(A) * A.seg;
    const long long m1 = (m0 + A.seg < A.nm) ? m0 + A.seg : A.nm;

    SweepThread T;
    T.valid = w < A.nw;
    T.col = (T.valid ? w : A.nw - 1) + A.g;
#pragma unroll
    for (int k = 0; k < 4; k++) T.base[k] = A.in[k] + T.col;
    T.amax = 0ULL; T.tmax = 0ULL;

    const DeviceTimeState *ts = A.ts;
    if (ts->done) {
        // Past maxtime / maxcycle: the cycle is a no-op for the physics (src/solver.jl:333); keep the buffer
        // rotation of the host bookkeeping consistent by copying the state through.
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
    const R dt = R(ts->current_dt) * R(A.dt_factor);   // update_solver_state!, src/solver_state.jl:339-345

    march_segment<R, DIV, RL, PROJ, EOS, true>(A, T, dt, m0, m1, w0, stage);

    if (DIV == DIV_FLAGGED) {
        // A thread whose operands left the range in which the branch-free division is proven exact (in practice:
        // a tiny non-zero dividend in the decaying tail of the numerical domain of influence) recomputes its
        // whole segment with nvcc's full IEEE division and overwrites its column: the strict mode is IEEE for
        // every operand, the common path just never pays for the slow path.
        range_check_dividend(dt.v, T.flag);
        if (T.flag.bad()) {
            T.amax = 0ULL; T.tmax = 0ULL;
            march_segment<R, DIV_IEEE, RL, PROJ, EOS, false>(A, T, dt, m0, m1, w0, stage);
            if ((threadIdx.x & 31) == __ffs(__activemask()) - 1) atomicAdd(&A.ts->redo_count, 1u);
        }
        __syncwarp();
    }

    // dtCFL partial maxima: warp shuffle, then one atomicMax per warp (max is order-independent: exact)
    unsigned long long am = T.amax, tm = T.tmax;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        const unsigned long long oa = __shfl_xor_sync(0xffffffffu, am, off);
        const unsigned long long ot = __shfl_xor_sync(0xffffffffu, tm, off);
        am = oa > am ? oa : am;
        tm = ot > tm ? ot : tm;
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMax(&A.ts->acc[A.acc_slot][0], am);
        atomicMax(&A.ts->acc[A.acc_slot][1], tm);
    }
}
