// sweep_fast_kernel.cuh -- the fast-mode marching kernel (math_mode fast: the product's default and the benched path).
//
// Same sweep, data layout and HBM traffic as sweep_kernel.cuh (read it first; sweep_staged_common.cuh has the helpers: thread <->
// column, march along the strided axis, rolling register window, inputs staged through a per-warp shared-memory
// ring).  What is specific to this file:
//
// 1. EXPLICIT ARITHMETIC.  Every floating-point operation of the step is written as an explicitly rounded intrinsic
//    (__dadd_rn / __dmul_rn / __fma_rn): where a multiply-add is fused is decided here, not by the compiler's
//    contraction pass.  The compiler contracts differently in different instantiations of the same source (warm-up vs
//    emitting steps, staging variants), which made the previous fast kernels' last bit depend on where a march segment
//    or a sub-domain starts.  With explicit arithmetic a cell's result is a function of its dependency cone only: fast
//    mode is bit-identical across march-segment lengths, block / rank decompositions and staging variants, like strict.
//
// 2. FEWER OPERATIONS (fast mode may regroup sums; results stay within 1e-12 of the oracle, tested at bench size):
//    * Lagrangian update in conserved form: rho' = dm/dxl, (rho u)' = (dm u + dt dFp)/dxl, (rho E)' = (dm E + dt dFpFu)/dxl
//      with one reciprocal of dxl -- no 1/rho, no dt/dm;
//    * projection from the same quantities: dxl (rho q)' - (A_{i+1} - A_i), one fma each;
//    * second-order remap with s' = minmod(T+_j, T+_{j-1}), T+_j = (q_{j+1} - q_j) / (dxl_j + dxl_{j+1}): the factors
//      2 dxl_j of the reference's r-, r+ and of its length fraction cancel (they are positive, so they commute with
//      minmod); each T+ is computed once and serves two cells;
//    * minmod Riemann limiter as an integer clamp of the ratio's bit pattern.
//    128 FP64 operations per cell and sweep for GAD + minmod + euler_2nd, perfect gas (139 before).
//
// 3. INPUT STAGING BY TMA.  The rows a warp consumes are fetched by the tensor-map form of the bulk copy engine
//    (cp.async.bulk.tensor.2d, SASS UTMALDG): one elected lane per warp issues, per group of 4 array rows, one copy per
//    variable of a [4 rows x 32 columns] box into the warp's shared-memory ring and arms an mbarrier with the 4 KB it
//    expects; the warp waits on the barrier's phase once per 4 steps.  No per-thread copy instructions, commit groups or
//    address arithmetic in the step (the cp.async staging costs 2 LDGSTS + commit + wait + warp barrier + bookkeeping
//    per step).  Out-of-range rows / columns of a ragged edge are zero-filled by the copy engine and only ever feed cells
//    that are not stored.  Tensor maps need 16-byte aligned rows, i.e. an even pitch; for odd pitches the same kernel is
//    instantiated with per-thread 8-byte cp.async copies (STG_CPA8), and STG_CPA16 keeps the 16-byte cp.async staging of
//    the previous round for comparison.  All staging variants share the ring layout and the step, hence the bits.
//
// 4. ALIGNED TRANSPOSED STORES.  A sweep that changes axis writes its output transposed, 8 cells (64 bytes) per output
//    row and flush.  Real cell m sits at column m + 4 of the output rows (4 ghost columns first), so chunks cut at
//    multiples of 8 cells straddle the 64-byte units of the output (measured: a transposing copy with straddling
//    64-byte pieces runs 18 % slower than with aligned ones, scratch microbenchmark of round 2).  The march segments
//    of this kernel therefore start 4 cells early: segment k covers cells [k seg - 4, (k+1) seg - 4) (the first one
//    starts with 4 masked virtual cells, the last one runs to the end of the domain), and every full chunk is one
//    aligned 64-byte unit per output row whenever the output pitch is a multiple of 8.
//
// 5. BAND-TILED LAYOUT BETWEEN SWEEPS (LAY_TILED; grids whose extents are multiples of 8).  With row-major arrays the
//    transposed output of 4. leaves a warp as isolated 64-byte pieces 128 KB apart; measured on the Godunov sweep
//    (memory-bound), that store pattern caps a transposing sweep at 0.77 of the copy bandwidth whatever the arithmetic
//    (pieces of 128 B: 0.79, 256 B: 0.865, 512 B: 0.87, 2 KB: 0.89 -- profiles/README.md).  In the tiled layout
//    (common.cuh, tiled_index: bands of 4 rows, tiles of [4 rows][8 columns]) 4 consecutive march cells of a thread are
//    one aligned 32-byte unit of its output row and the 4 rows of a band x 8 cells one contiguous 256-byte tile: the
//    transposed output leaves the REGISTERS as one predicated 256-bit store per variable every 4 steps (no staging tile,
//    no flush phase -- a staged flush of the same tiles stalled the in-order warp ~280 clocks every 8 steps and was no
//    faster than row-major on the FP64-heavy sweeps), and the next sweep fetches a [4 rows x 32 columns] staging group as
//    one contiguous 1 KB run (4 adjacent tiles) instead of 4 row pieces.  0.77 -> 0.92 of the copy bandwidth.  Threads are shifted by 4 columns (lane 0 of a warp <-> array
//    column 32 k, i.e. cell 32 k - 4) so that a warp's columns are whole tiles on the input side and whole bands on the
//    output side; the 4 ghost columns this brings into the first warp are masked like the ragged tail.  The arithmetic
//    is untouched: results are bit-identical to the row-major path.  The ghost rows of a side are exactly one band, so
//    the boundary-condition fill only changes its indexing and the halo copies (NCCL, local blocks) do not change at all.
//
// 6. THE BIT-EXACT MODE ON THE SAME SKELETON (MATH_STRICT; details above strict_step).  math_mode strict -- the reference's
//    operation order, no contraction, correctly rounded branch-free divisions, bit-identical to the CPU oracle -- runs
//    this kernel's staging and four-chain schedule with its own step function; row-major layouts, because the IEEE fix-up
//    of out-of-range division operands (sweep_fixup_kernel.cuh) works on them.  Three staging groups instead of four: the
//    fourth group's shared memory holds the Lagrangian rings, which is what leaves ptxas the registers to interleave the
//    chains (0.345 -> 0.417 of the copy bandwidth at 16384^2 against the one-chain kernel of round 1).
#pragma once

#include <cuda.h>

#include <type_traits>

#include "sweep_staged_common.cuh"

enum { STG_TMA = 0, STG_CPA16 = 1, STG_CPA8 = 2 };
enum { LAY_ROWS = 0, LAY_TILED = 1 };          // layout of the input AND of the output of a sweep
enum { MATH_FAST = 0, MATH_STRICT = 1 };       // arithmetic of the step: fast_step (section 2.) or strict_step (section 6.)

// the four input arrays of a sweep (rho, ua, ut, E) as 2-D tensors: LAY_ROWS [array rows][pitch], box = 4 rows x 32
// columns; LAY_TILED [bands][4 pitch], box = 1 band x 128 elements (4 adjacent tiles = the same 4 rows x 32 columns)
struct SweepTmaMaps { CUtensorMap m[4]; };

constexpr int FK_GROUP = 4;                    // rows per staging group (= one TMA box per variable)
constexpr int FK_NG = 4;                       // groups in the ring: the one being consumed and three in flight
constexpr int FK_ROWS = FK_GROUP * FK_NG;      // 16 ring rows
constexpr int FK_VS = FK_GROUP * 32;           // doubles between two variables of the same row
constexpr int FK_GS = 4 * FK_VS;               // doubles per group
constexpr int FK_CS = 8;                       // sound-speed ring slots (written at a, read at a-7)
#ifndef FK_K
#define FK_K 8                                 // cells per output row and flush of the transposed stores (8 or 16)
#endif
constexpr int FK_PITCH = FK_K + 2;             // staging tile row pitch: rows stay 16-byte aligned, 2-way write conflicts
static_assert(FK_K == 8 || FK_K == 16, "transposed staging chunk");

// Ring slot and mbarrier phase of staging group G.  Fast arithmetic: the 4 groups of the ring.  Strict arithmetic: 3 groups
// (the previous one, kept for the re-reads of chain C, the current one and one in flight -- at its pace 4 to 8 rows in
// flight cover the memory latency several times); the fourth group's 4 KB hold the Lagrangian rings (strict_step).
template <int MATH> __device__ __forceinline__ int fk_slot(int G) { return MATH == 1 ? G % 3 : G & (4 - 1); }
template <int MATH> __device__ __forceinline__ unsigned fk_phase(int G) { return (unsigned)((MATH == 1 ? G / 3 : G >> 2) & 1); }

struct FastWarpShared {
    double ring[FK_NG][4][FK_GROUP][32];                   // [group][variable][row in group][lane]
    double cring[FK_CS][32];
    double stage[4 * 32 * FK_PITCH];                       // transposed-store staging (fast_flush)
    unsigned long long full[FK_NG];                        // mbarriers: the group's 4 boxes have landed (STG_TMA)
    unsigned long long pad_[(128 - (4 * 32 * FK_PITCH * 8 + FK_NG * 8) % 128) / 8 % 16];   // next warp's ring 128-byte aligned
};
static_assert(sizeof(FastWarpShared) % 128 == 0, "per-warp shared block must keep 128-byte alignment");

// ---- explicitly rounded arithmetic -----------------------------------------------------------------------------
__device__ __forceinline__ double xadd(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double xsub(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double xmul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double xfma(double a, double b, double c) { return __fma_rn(a, b, c); }

// reciprocal to ~1 ulp: 20-bit seed + one cubic step
__device__ __forceinline__ double xrcp(double b)
{
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(b));
    double e = xfma(-b, r, 1.0);
    e = xfma(e, e, e);
    return xfma(r, e, r);
}

// sqrt to ~1 ulp: coupled iteration on g ~ sqrt(a), h ~ 1/(2 sqrt(a)) + one residual correction; sqrt(0) = 0
__device__ __forceinline__ double xsqrt(double a)
{
    double y0;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(a));
    const double g0 = xmul(a, y0), h0 = xmul(0.5, y0);
    const double r = xfma(-g0, h0, 0.5);
    const double g1 = xfma(g0, r, g0), h1 = xfma(h0, r, h0);
    const double res = xfma(xfma(-g1, g1, a), h1, g1);
    return a == 0.0 ? a : res;
}

// max(0, min(1, r)) as a clamp of the bit pattern (src/limiters.jl:7): negative, -0 -> +0; >= 1, NaN -> 1
__device__ __forceinline__ double xclamp01(double r)
{
    const int hi = __double2hiint(r), lo = __double2loint(r);
    const bool out = (unsigned)hi >= 0x3ff00000u;           // sign bit set, or r >= 1
    return __hiloint2double(min(max(hi, 0), 0x3ff00000), out ? 0 : lo);
}

template <int LIMITER> __device__ __forceinline__ double xlimiter(double r)
{
    if (LIMITER == ARMON_LIMITER_MINMOD) return xclamp01(r);
    if (LIMITER == ARMON_LIMITER_SUPERBEE) {
        // max(0, min(2r, 1), min(r, 2)) (src/limiters.jl:8) selected by the bit pattern like the minmod clamp, off the
        // FP64 pipe:  r < 0, -0 -> +0 | r < 1/2 -> 2r | r < 1 -> 1 | r < 2 -> r | r >= 2, NaN -> 2
        const int hi = __double2hiint(r), lo = __double2loint(r);
        const double r2 = xadd(r, r);
        const bool small = hi < 0x3fe00000;                            // signed: r < 1/2, negative r included
        const bool mid = (unsigned)(hi - 0x3ff00000) < 0x00100000u;    // 1 <= r < 2
        const bool neg = hi < 0;
        int oh = min(max(hi, 0x3ff00000), 0x40000000);                 // r >= 1/2: clamped to [1, 2]
        int ol = mid ? lo : 0;
        oh = small ? __double2hiint(r2) : oh;
        ol = small ? __double2loint(r2) : ol;
        return __hiloint2double(neg ? 0 : oh, neg ? 0 : ol);
    }
    return 1.0;
}

// the operand of smaller magnitude when both have the same sign, else 0 (src/projection_schemes.jl:15-20)
__device__ __forceinline__ double xminmod(double d_p, double d_m)
{
    const bool m_smaller = fabs(d_m) < fabs(d_p);
    const double m = m_smaller ? d_m : d_p;
    const int sx = __double2hiint(d_p) ^ __double2hiint(d_m);
    return sx < 0 ? 0.0 : m;
}

// p, c and rho c of one cell (src/kernels.jl:4-55), fast-mode algebra
template <int EOS>
__device__ __forceinline__ void xeos(const SweepArgs &A, double rho, double ua, double ut, double E, double &p, double &c)
{
    const double e = xfma(-0.5, xfma(ut, ut, xmul(ua, ua)), E);
    if (EOS == ARMON_EOS_PERFECT_GAS) {
        p = xmul(xmul(A.gm1, rho), e);                      // (gamma - 1) and gamma (gamma - 1) are formed on the host
        c = xsqrt(xmul(A.ggm1, e));                         // gamma p / rho == gamma (gamma - 1) e: no division
    } else {
        // src/kernels.jl:16-55.  Horner forms; one reciprocal of 1 - s x serves f0, f1, f2; p - pk0 is formed directly
        // (G0 rho0 (e - epsk0)) instead of by cancellation.
        constexpr double rho0 = 10000., K0 = 1e+11, Cv0 = 1000., T0 = 300., G0 = 1.5, s = 1.5;
        constexpr double q = -42080895. / 14941154., r = 727668333. / 149411540.;
        constexpr double s3m2 = 1.5 / 3 - 2, CvT = Cv0 * T0, G0r0 = G0 * rho0;
        const double inv_rho = xrcp(rho);
        const double x = xfma(rho, 1.0 / rho0, -1.0);
        const double G = xfma(-G0r0, inv_rho, G0);          // G0 (1 - rho0 / rho)
        const double iden = xrcp(xfma(-s, x, 1.0));
        const double f0 = xmul(xfma(xfma(xfma(r, x, q), x, s3m2), x, 1.0), iden);
        const double f1 = xmul(xfma(s, f0, xfma(xfma(3 * r, x, 2 * q), x, s3m2)), iden);
        const double f2 = xmul(xfma(2 * s, f1, xfma(6 * r, x, 2 * q)), iden);
        const double x2 = xmul(x, x), opx = xadd(1.0, x), opx2 = xmul(opx, opx), opx3 = xmul(opx2, opx);
        const double epsk0 = xfma(xmul(0.5 * K0 / rho0, x2), f0, xfma(-CvT, G, -CvT));
        const double pk0 = xfma(xmul(xmul(0.5 * K0, x), opx2), xfma(x, f1, xadd(f0, f0)), -CvT * G0 * rho0);
        const double sum = xfma(xmul(x2, opx), f2, xfma(xmul(x, xfma(6.0, x, 4.0)), f1, xmul(xfma(6.0, x, 2.0), f0)));
        const double pk0prime = xmul(xmul(-0.5 * K0 * rho0, opx3), sum);
        const double w = xmul(G0r0, xsub(e, epsk0));        // p - pk0
        p = xadd(pk0, w);
        c = xmul(xsqrt(xfma(G0r0, w, -pk0prime)), inv_rho);
    }
}

// Rolling window of the march.  Quantities that live for several steps sit in 4-slot rings indexed by (cell or interface
// index & 3); quantities handed from one step to the next in 2-slot rings indexed by the step's parity.  All slots are
// compile-time (the loop is unrolled by 4), so every value stays in one register for its whole life: no moves.
//
// Schedule of step a (cell a is read from the ring): FOUR chains that only use results of EARLIER steps, so that the
// scheduler always has independent work to cover the FP64 pipe's latency (an in-order warp with one long dependent
// chain per step spends its time in fixed-latency waits):
//   A  EOS(a) + Godunov state of interface a                     <- cells a-1, a
//   B  flux (GAD or Godunov) of interface a-2                     <- Godunov states a-3, a-2, a-1, cells a-3, a-2
//   C  Lagrangian cell a-4                                        <- fluxes a-4, a-3 (steps a-2, a-1)
//   D  advection flux of interface a-6, E  projection of cell a-7 <- Lagrangian cells a-7, a-6, a-5
// Each chain commits its results at the end of the step.
struct PipeF {
    double cu[4], cp[4], crc[4], cdm[4], cut[4], cE[4];     // cells a-1 .. a-4: ua, p, rho c, rho dx, ut, E
    double Gu[4], Gp[4];                                    // Godunov states of interfaces a-1 .. a-3
    double Fu[4], Fp[4], FpFu[4];                           // flux used (GAD or Godunov) of interfaces a-3, a-4, and p u
    double dl[4], dxl[4], Lr[4], Lru[4], Lrt[4], LrE[4];    // Lagrangian cells a-5 .. a-7: dt * flux velocity of the left
                                                            // interface, width, rho, rho {ua, ut, E}
    double T[2][4];                                         // T+ of the cell pairs (a-6, a-5) / (a-7, a-6)
    double S[2][4];                                         // s' of cells a-6 / a-7
    double Adv[2][4];                                       // advection fluxes of interfaces a-6 / a-7
    double Q[4][4];                                         // tiled transposed output: the 4 cells of a 32-byte unit, per variable
};

// Per-iteration (4 steps) addressing, hoisted out of the steps: everything a step touches is at a compile-time offset
// from one of these.
struct FastIter {
    const double *gb0;           // this lane's column of the ring group holding rows a-J .. a-J+3
    double *cw;                  // sound-speed ring: slots of the 4 cells of this iteration
    const double *cr;            // ... and of the 4 cells of the previous one
    double *s0, *s3;             // transposed staging tile: slot of the cell emitted at J = 0 (J = 1, 2 follow) / at J = 3
    long long o_it;              // direct stores: element offset of the cell emitted at J = 0
    int rem;                     // cells of the segment still to emit, counting the one of J = 0 (<= 0: none)
    long long q_off;             // tiled transposed stores: element offset of the 4-cell unit completed at J = 2
    bool qok;                    // ... and whether it is stored (real column, unit inside the segment)
    unsigned row_mask;           // strict arithmetic: bit J set <=> row J of the current group exists in the array (the others
                                 // were zero-filled by the copy engine and are given a benign state before use)
    const double *gbm;           // strict_step: this lane's column of the ring group holding rows a-4-J .. (the previous group)
    volatile double *lr;         // strict_step: this lane's column of the Lagrangian rings [rho, ua, ut, E][slot] (fourth ring group)
};

__device__ __forceinline__ void fk_store4_if(bool ok, double *p, double a, double b, double c, double d);

// One march step at cell a; J = (a - a_begin) & 3 static.  EMIT / TR compile-time as in march_compute2; `ok`: the
// thread's column holds a real cell (false also for the steps that run the emitting code on cells before the segment,
// see the kernel).
template <int RL, int PROJ, int EOS, int J, int TR, int EMIT, int CONS, int LAY>
__device__ __forceinline__ void fast_step(const SweepArgs &A, SweepThread &T, PipeF &P, const FastIter &I,
                                          const double dt, const bool ok)
{
#define ZS(k) ((J + 8 - (k)) & 3)
    constexpr int Z0 = ZS(0), Z1 = ZS(1), Z2 = ZS(2), Z3 = ZS(3);   // also the slots of a-4, a-5, a-6, a-7
    constexpr int CUR = J & 1, PRV = CUR ^ 1;
    const double dx = A.dx;
    // row a of this lane's column: ring group = [variable][row][32 lanes], or [variable][tile][row][8 lanes] (tiled)
    const double *row0 = I.gb0 + J * (LAY == LAY_TILED ? 8 : 32);

    // ---- chain A, cell a: EOS, Godunov state of interface a (cells a-1, a), src/riemann_schemes.jl:21-30 ----
    double A_ua, A_p, A_rc, A_dm, A_Gu, A_Gp, A_ut, A_E;
    {
        const double rho = row0[0], ua = row0[FK_VS], ut = row0[2 * FK_VS], E = row0[3 * FK_VS];
        A_ut = ut; A_E = E;
        double p, c;
        xeos<EOS>(A, rho, ua, ut, E, p, c);
        I.cw[J * 32] = c;
        const double rc = xmul(rho, c);
        const double rc_l = P.crc[Z1], u_l = P.cu[Z1], p_l = P.cp[Z1];
        const double iden = xrcp(xadd(rc_l, rc));
        A_Gu = xmul(xfma(rc_l, u_l, xfma(rc, ua, xsub(p_l, p))), iden);
        A_Gp = xmul(xfma(xmul(rc_l, rc), xsub(u_l, ua), xfma(rc, p_l, xmul(rc_l, p))), iden);
        A_ua = ua; A_p = p; A_rc = rc; A_dm = xmul(rho, dx);
    }

    // ---- chain B, flux at interface i = a-2 (cells a-3, a-2); Godunov states a-3, a-2, a-1 ----
    double B_Fu, B_Fp;
    if (RL == 0) {   // acoustic!  src/riemann_schemes.jl:33-43
        B_Fu = P.Gu[Z2];
        B_Fp = P.Gp[Z2];
    } else {         // acoustic_GAD!  src/riemann_schemes.jl:55-104
        constexpr int LIM = RL - 1;
        const double u_i = P.cu[Z2], u_im = P.cu[Z3], p_i = P.cp[Z2], p_im = P.cp[Z3];
        const double us_i = P.Gu[Z2], ps_i = P.Gp[Z2];
        const double au = xsub(u_i, us_i), bu = xsub(us_i, u_im), ap = xsub(p_i, ps_i), bp = xsub(ps_i, p_im);
        double du, dp;
        if (LIM != ARMON_LIMITER_NONE) {   // limiter(r, NoLimiter) == 1 whatever r is (src/limiters.jl:6)
            const double r_um = xlimiter<LIM>(xmul(xsub(P.Gu[Z1], u_i), xrcp(xadd(bu, 1e-6))));
            const double r_pm = xlimiter<LIM>(xmul(xsub(P.Gp[Z1], p_i), xrcp(xadd(bp, 1e-6))));
            const double r_up = xlimiter<LIM>(xmul(xsub(u_im, P.Gu[Z3]), xrcp(xadd(au, 1e-6))));
            const double r_pp = xlimiter<LIM>(xmul(xsub(p_im, P.Gp[Z3]), xrcp(xadd(ap, 1e-6))));
            du = xfma(r_up, au, -xmul(r_um, bu));
            dp = xfma(r_pp, ap, -xmul(r_pm, bp));
        } else {
            du = xsub(au, bu);
            dp = xsub(ap, bp);
        }
        // theta = 0.5 (1 - (rc_l + rc_r)/2 * dt / ((dm_l + dm_r)/2)) = 0.5 - 0.5 dt (rc_l + rc_r) / (dm_l + dm_r)
        const double theta = xfma(xmul(xadd(P.crc[Z3], P.crc[Z2]), xmul(-0.5, dt)),
                                  xrcp(xadd(P.cdm[Z3], P.cdm[Z2])), 0.5);
        B_Fu = xfma(theta, du, us_i);
        B_Fp = xfma(theta, dp, ps_i);
    }
    const double B_FpFu = xmul(B_Fp, B_Fu);

    // ---- chain C, Lagrangian cell k = a-4 in conserved form (src/kernels.jl:58-68): its interfaces a-4 (left) and a-3
    //      (right) were computed two steps / one step ago; ua, ut, E, dm of the cell are still in slot Z0 of the cell
    //      rings (chain A commits cell a there at the end of the step) ----
    double C_dl, C_dxl, C_Lr, C_Lru, C_Lrt, C_LrE;
    {
        C_dl = xmul(dt, P.Fu[Z0]);
        C_dxl = xfma(dt, xsub(P.Fu[Z3], P.Fu[Z0]), dx);
        const double rdxl = xrcp(C_dxl);
        C_Lr = xmul(P.cdm[Z0], rdxl);
        const double kdt = xmul(dt, rdxl);
        C_Lru = xfma(kdt, xsub(P.Fp[Z0], P.Fp[Z3]), xmul(C_Lr, P.cu[Z0]));
        C_Lrt = xmul(C_Lr, P.cut[Z0]);
        C_LrE = xfma(kdt, xsub(P.FpFu[Z0], P.FpFu[Z3]), xmul(C_Lr, P.cE[Z0]));
    }

    // ---- chain D, advection flux at interface is = a-6 (src/projection_schemes.jl:62-124): upwind cell a-7 (slot Z3,
    //      disp > 0) or a-6 (slot Z2); the Lagrangian cell a-5 (slot Z1) completes the stencil of cell a-6 ----
    {
        const double d = P.dl[Z2];
        const bool pos = d > 0.0;
        const double *qm[4] = {&P.Lr[Z3], &P.Lru[Z3], &P.Lrt[Z3], &P.LrE[Z3]};   // compile-time addresses: registers
        const double *q0[4] = {&P.Lr[Z2], &P.Lru[Z2], &P.Lrt[Z2], &P.LrE[Z2]};
        const double *qp[4] = {&P.Lr[Z1], &P.Lru[Z1], &P.Lrt[Z1], &P.LrE[Z1]};
        if (PROJ == ARMON_PROJ_EULER_2ND) {
            // T+ of the pair (a-6, a-5); the pair (a-7, a-6) was formed one step ago
            const double rs = xrcp(xadd(P.dxl[Z2], P.dxl[Z1]));
            // distance from the centre of the upwind cell to the middle of the swept length:
            // disp > 0: -(dx - disp[is-1]) ; else dx + disp[is+1]   (the reference's dxe; its 1/(2 dxl) is inside s')
            const double dxe = pos ? xsub(P.dl[Z3], dx) : xadd(dx, P.dl[Z1]);
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const double t = xmul(rs, xsub(*qp[k], *q0[k]));
                const double sl = xminmod(t, P.T[PRV][k]);
                P.T[CUR][k] = t;
                P.S[CUR][k] = sl;
                P.Adv[CUR][k] = xmul(d, xfma(-(pos ? P.S[PRV][k] : sl), dxe, pos ? *qm[k] : *q0[k]));
            }
        } else {
#pragma unroll
            for (int k = 0; k < 4; k++) P.Adv[CUR][k] = xmul(d, pos ? *qm[k] : *q0[k]);
        }
    }

    // ---- chain E (continues D), projection of cell k = a-7 (src/projection_schemes.jl:23-41):
    //      dxl (rho q)' - (A_{k+1} - A_k) ----
    if (EMIT == 1) {
        const double dl = P.dxl[Z3];
        const double t_r = xfma(dl, P.Lr[Z3], xsub(P.Adv[PRV][0], P.Adv[CUR][0]));
        const double t_ru = xfma(dl, P.Lru[Z3], xsub(P.Adv[PRV][1], P.Adv[CUR][1]));
        const double t_rt = xfma(dl, P.Lrt[Z3], xsub(P.Adv[PRV][2], P.Adv[CUR][2]));
        const double t_rE = xfma(dl, P.LrE[Z3], xsub(P.Adv[PRV][3], P.Adv[CUR][3]));
        const double inv_r = xrcp(t_r);      // u = (t_ru / dx) / (t_r / dx): only the density needs 1/dx
        const double o_r = xmul(t_r, A.inv_dx), o_ua = xmul(t_ru, inv_r), o_ut = xmul(t_rt, inv_r), o_E = xmul(t_rE, inv_r);
        const double c_out = J == 3 ? I.cr[0] : I.cw[(J + 1) * 32];   // c of cell a-7 (EOS of this sweep)
        const bool store = ok && J < I.rem;
        {   // dtCFL accumulators (src/reductions.jl:14-20), branch-free: cells that are not stored do not contribute
            const unsigned long long ba = (unsigned long long)__double_as_longlong(xadd(fabs(o_ua), c_out));
            const unsigned long long bt = (unsigned long long)__double_as_longlong(xadd(fabs(o_ut), c_out));
            T.amax = (store && ba > T.amax) ? ba : T.amax;
            T.tmax = (store && bt > T.tmax) ? bt : T.tmax;
        }
        if (CONS && store) {   // conservation_vars (src/reductions.jl:202-259) of the state this sweep produces
            T.cm = xadd(T.cm, o_r);
            T.ce = xfma(o_r, o_E, T.ce);
        }
        if (TR == 1 && LAY == LAY_TILED) {
            // Transposed output in the tiled layout: the lane's output row holds the march cells of a tile 8 by 8, so 4
            // consecutive cells of a thread are one aligned 32-byte unit.  They are collected in registers (the unit is
            // {J = 3 of the previous iteration, J = 0, 1, 2}: segments start on a multiple of 4) and leave as one 256-bit
            // store per variable: no shared-memory staging, no flush phase that stalls the warp every 8 steps.
            constexpr int SL = (J + 1) & 3;
            P.Q[0][SL] = o_r; P.Q[1][SL] = o_ua; P.Q[2][SL] = o_ut; P.Q[3][SL] = o_E;
            if (J == 2) {   // predicated stores, not a branch around them: the step stays one basic block
#pragma unroll
                for (int v = 0; v < 4; v++) fk_store4_if(I.qok, A.out[v] + I.q_off, P.Q[v][0], P.Q[v][1], P.Q[v][2], P.Q[v][3]);
            }
        } else if (TR == 1) {
            double *s = J == 3 ? I.s3 : I.s0 + J;
            s[0 * 32 * FK_PITCH] = o_r;
            s[1 * 32 * FK_PITCH] = o_ua;
            s[2 * 32 * FK_PITCH] = o_ut;
            s[3 * 32 * FK_PITCH] = o_E;
        } else if (store) {
            // tiled: the cell of J = 0 is row 1 of its band (segments start on a band), J = 3 row 0 of the next band
            const long long o = LAY == LAY_TILED ? I.o_it + (J < 3 ? J * 8 : 4 * A.pitch_out - 8) : I.o_it + J * A.pitch_out;
            A.out[0][o] = o_r;
            A.out[1][o] = o_ua;
            A.out[2][o] = o_ut;
            A.out[3][o] = o_E;
        }
    }

    // ---- commit the chains: A -> cell a and interface a (slot Z0), B -> interface a-2 (slot Z2), C -> cell a-4 (slot Z0) ----
    P.cu[Z0] = A_ua; P.cp[Z0] = A_p; P.crc[Z0] = A_rc; P.cdm[Z0] = A_dm; P.cut[Z0] = A_ut; P.cE[Z0] = A_E;
    P.Gu[Z0] = A_Gu; P.Gp[Z0] = A_Gp;
    P.Fu[Z2] = B_Fu; P.Fp[Z2] = B_Fp; P.FpFu[Z2] = B_FpFu;
    P.dl[Z0] = C_dl; P.dxl[Z0] = C_dxl; P.Lr[Z0] = C_Lr; P.Lru[Z0] = C_Lru; P.Lrt[Z0] = C_Lrt; P.LrE[Z0] = C_LrE;
#undef ZS
}

// ---- 6. STRICT ARITHMETIC ON THE SAME SCHEDULE (MATH_STRICT) ---------------------------------------------------------
// The bit-exact mode (math_mode strict) on the skeleton of this file: TMA staging, four independent chains per step.
// The arithmetic is the reference's, operation for operation -- the expressions of march_compute<sd, DIV_FLAGGED>
// (sweep_kernel.cuh), only evaluated at other steps: a chain reads what EARLIER steps committed, never a result of its
// own step, so the in-order warp always has four dependency chains to interleave (the unskewed strict kernel runs EOS ->
// Godunov -> GAD -> cell update -> advection -> projection as ONE chain of ~250 dependent FP64 operations per step and
// spends two thirds of its issue slots waiting).  A cell's result is a function of its dependency cone only, so the
// bits are those of the unskewed kernel and of the CPU oracle.  Row-major layouts (LAY_ROWS) only: the IEEE fix-up of
// the out-of-range division operands (sweep_fixup_kernel.cuh) works on them.
//
// Rows outside the array (the 4 virtual cells before the first segment, the rows the last segment consumes past the
// last ghost row) are zero-filled by the copy engine; they only feed cells that are never stored, but a zero density
// would raise the thread's range flag and send real chunks to the fix-up: chain A replaces them by a benign state.
// pos ? a : b as a bit blend with a precomputed all-ones / all-zeros mask: one LOP3 per 32-bit half, whereas the
// compiler lowers the ten upwind selections of a step to pairs of (predicated) register moves at this register pressure
__device__ __forceinline__ sd sblend(unsigned long long m, sd a, sd b)
{
    const unsigned long long ua = (unsigned long long)__double_as_longlong(a.v), ub = (unsigned long long)__double_as_longlong(b.v);
    return sd(__longlong_as_double((long long)((ua & m) | (ub & ~m))));
}

struct PipeS {
    sd cp[4], crc[4];                                       // cells a-1 .. a-3: p, rho c (what the EOS computed; rho, ua, ut
                                                            // and E of the cells a-1 .. a-4 are re-read from the staging
                                                            // ring, which keeps the previous group for that)
    sd Gu[4], Gp[4];                                        // Godunov states of interfaces a-1 .. a-3
    sd Fu[4], Fp[4], FpFu[4];                               // flux used (GAD or Godunov) of interfaces a-3, a-4, and p u
    sd disp[4], dxl[4];                                     // interfaces / cells a-5 .. a-7: dt Fu, Lagrangian width
    // (the Lagrangian cells a-5 .. a-7 -- rho, ua, ut, E -- live in shared-memory rings, FastIter::lr: 24 registers less;
    // the products rho {ua, ut, E} are formed where they are used: same operands, same bits)
    sd Ar, Aru, Art, ArE;                                   // advection flux of interface a-7
    sd Sr, Sru, Srt, SrE;                                   // limited slopes of cell a-7
    sd S2b, S2r;                                            // 2 dxl of cell a-7 and its refined reciprocal
};

template <int RL, int PROJ, int EOS, int J, int TR, int EMIT, int DXP>
__device__ __forceinline__ void strict_step(const SweepArgs &A, SweepThread &T, PipeS &P, const FastIter &I, const sd dt,
                                            const typename Div<sd, DIV_FLAGGED>::Rcp &inv_dx, const bool ok)
{
    typedef sd R;
    typedef Div<sd, DIV_FLAGGED> D;
#define ZS(k) ((J + 8 - (k)) & 3)
    constexpr int Z0 = ZS(0), Z1 = ZS(1), Z2 = ZS(2), Z3 = ZS(3);   // also the slots of a-4, a-5, a-6, a-7
    const R dx(A.dx);
    RangeFlag &f = T.flag;
    // rows a, a-1 .. a-4 of this lane's column: row J of the current group, or of the previous one (compile-time choice)
    const double *row0 = I.gb0 + J * 32;
    const double *row1 = J >= 1 ? I.gb0 + (J - 1) * 32 : I.gbm + (J + 3) * 32;
    const double *row2 = J >= 2 ? I.gb0 + (J - 2) * 32 : I.gbm + (J + 2) * 32;
    const double *row3 = J >= 3 ? I.gb0 + (J - 3) * 32 : I.gbm + (J + 1) * 32;
    const double *row4 = I.gbm + J * 32;

    // ---- chain A, cell a: EOS (src/kernels.jl:4-55), Godunov state of interface a (src/riemann_schemes.jl:21-30) ----
    R A_p, A_rc, A_Gu, A_Gp;
    {
        const R rho(row0[0]), ua(row0[FK_VS]), ut(row0[2 * FK_VS]), E(row0[3 * FK_VS]);
        R p, c;
        eos_eval<R, DIV_FLAGGED, EOS>(A, rho, ua, ut, E, p, c, f);
        I.cw[J * 32] = c.v;
        const R rc = rho * c;
        acoustic_godunov<R, DIV_FLAGGED>(P.crc[Z1], rc, R(row1[FK_VS]), ua, P.cp[Z1], p, A_Gu, A_Gp, f);
        A_p = p; A_rc = rc;
    }

    // ---- chain B, flux at interface i = a-2 (cells a-3, a-2); Godunov states a-3, a-2, a-1 ----
    R B_Fu, B_Fp;
    if (RL == 0) {   // acoustic!  src/riemann_schemes.jl:33-43
        B_Fu = P.Gu[Z2];
        B_Fp = P.Gp[Z2];
    } else {         // acoustic_GAD!  src/riemann_schemes.jl:55-104
        constexpr int LIM = RL - 1;
        const R u_i(row2[FK_VS]), u_im(row3[FK_VS]), p_i = P.cp[Z2], p_im = P.cp[Z3];
        const R us_i = P.Gu[Z2], ps_i = P.Gp[Z2];
        R r_um(1.), r_pm(1.), r_up(1.), r_pp(1.);
        if (LIM != ARMON_LIMITER_NONE) {   // limiter(r, NoLimiter) == 1 whatever r is (src/limiters.jl:6)
            r_um = limiter<R, LIM>(D::div(P.Gu[Z1] - u_i, (us_i - u_im) + R(1e-6), f));
            r_pm = limiter<R, LIM>(D::div(P.Gp[Z1] - p_i, (ps_i - p_im) + R(1e-6), f));
            r_up = limiter<R, LIM>(D::div(u_im - P.Gu[Z3], (u_i - us_i) + R(1e-6), f));
            r_pp = limiter<R, LIM>(D::div(p_im - P.Gp[Z3], (p_i - ps_i) + R(1e-6), f));
        }
        const R Dm = (R(row3[0]) * dx + R(row2[0]) * dx) * R(0.5);                       // (dm_l + dm_r) / 2, dm = rho dx
        const R theta = R(0.5) * (R(1.) - ((P.crc[Z3] + P.crc[Z2]) * R(0.5)) * D::div_pos(dt, Dm, f));
        B_Fu = us_i + theta * (r_up * (u_i - us_i) - r_um * (us_i - u_im));
        B_Fp = ps_i + theta * (r_pp * (p_i - ps_i) - r_pm * (ps_i - p_im));
    }
    const R B_FpFu = B_Fp * B_Fu;

    // ---- chain C, Lagrangian update of cell k = a-4 (src/kernels.jl:58-68): interfaces a-4 (left, slot Z0) and a-3
    //      (right, slot Z3) were committed two steps / one step ago; the cell itself is still in slot Z0 of the cell rings ----
    R C_disp, C_dxl, C_Lr, C_Lu, C_Lt, C_LE;
    {
        C_disp = dt * P.Fu[Z0];
        C_dxl = dx + dt * (P.Fu[Z3] - P.Fu[Z0]);
        const R dm = R(row4[0]) * dx;
        const R dtdm = D::div_pos(dt, dm, f);
        C_Lr = D::div_pos(dm, C_dxl, f);
        C_Lu = R(row4[FK_VS]) + dtdm * (P.Fp[Z0] - P.Fp[Z3]);
        C_LE = R(row4[3 * FK_VS]) + dtdm * (P.FpFu[Z0] - P.FpFu[Z3]);
        C_Lt = R(row4[2 * FK_VS]);
    }

    // ---- chain D, advection flux at interface is = a-6 (src/projection_schemes.jl:62-124): cells a-7 (slot Z3), a-6
    //      (Z2), a-5 (Z1); the per-cell quantities (width ratios, limited slopes, 2 dxl) are formed for cell a-6, those
    //      of cell a-7 were kept from the previous step (see march_compute) ----
    R Anr, Anru, Anrt, AnrE;
    R sr(0.), sru(0.), srt(0.), srE(0.);
    typename D::Rcp k2;
    k2.b = 2.0; k2.r = 0.5;
    // Lagrangian cells from the shared rings [variable][slot]; cell a-7 also serves the projection below
#define LRING(var, Z) I.lr[((var) * 4 + (Z)) * 32]
    const R Lr3(LRING(0, Z3)), Lu3(LRING(1, Z3)), Lt3(LRING(2, Z3)), LE3(LRING(3, Z3));
    {
        const R d = P.disp[Z2];
        const unsigned long long pm = d.v > 0.0 ? ~0ULL : 0ULL;   // disp > 0: the upwind cell is a-7
        const R Lr2(LRING(0, Z2)), Lr1(LRING(0, Z1));
        // rho {ua, ut, E} (cell_update!'s products, formed here: same operands, same bits)
        const R Lru3 = Lr3 * Lu3, Lrt3 = Lr3 * Lt3, LrE3 = Lr3 * LE3;
        const R Lru2 = Lr2 * R(LRING(1, Z2)), Lrt2 = Lr2 * R(LRING(2, Z2)), LrE2 = Lr2 * R(LRING(3, Z2));
        if (PROJ == ARMON_PROJ_EULER_2ND) {
            const R dxl_m = P.dxl[Z3], dxl_0 = P.dxl[Z2], dxl_p = P.dxl[Z1];
            const R two_dxl = R(2.) * dxl_0;
            const R r_m = D::div_pos(two_dxl, dxl_0 + dxl_m, f);
            const R r_p = D::div_pos(two_dxl, dxl_0 + dxl_p, f);
            k2 = D::prepare_pos(two_dxl, f);
            sr = slope_minmod_fused<R>(Lr3, Lr2, Lr1, r_m, r_p);
            sru = slope_minmod_fused<R>(Lru3, Lru2, Lr1 * R(LRING(1, Z1)), r_m, r_p);
            srt = slope_minmod_fused<R>(Lrt3, Lrt2, Lr1 * R(LRING(2, Z1)), r_m, r_p);
            srE = slope_minmod_fused<R>(LrE3, LrE2, Lr1 * R(LRING(3, Z1)), r_m, r_p);

            const R dxe = sblend(pm, -(dx - P.disp[Z3]), dx + P.disp[Z1]);
            typename D::Rcp ksel;
            ksel.b = sblend(pm, P.S2b, R(k2.b)).v;
            ksel.r = sblend(pm, P.S2r, R(k2.r)).v;
            const R lf = D::quot(dxe, ksel, f);
            Anr = d * (sblend(pm, Lr3, Lr2) - sblend(pm, P.Sr, sr) * lf);
            Anru = d * (sblend(pm, Lru3, Lru2) - sblend(pm, P.Sru, sru) * lf);
            Anrt = d * (sblend(pm, Lrt3, Lrt2) - sblend(pm, P.Srt, srt) * lf);
            AnrE = d * (sblend(pm, LrE3, LrE2) - sblend(pm, P.SrE, srE) * lf);
        } else {
            Anr = d * sblend(pm, Lr3, Lr2);
            Anru = d * sblend(pm, Lru3, Lru2);
            Anrt = d * sblend(pm, Lrt3, Lrt2);
            AnrE = d * sblend(pm, LrE3, LrE2);
        }
    }

    // ---- chain E (continues D), projection of cell k = a-7 (src/projection_schemes.jl:23-41) ----
    if (EMIT == 1) {
        const R dXr = P.dxl[Z3] * Lr3;
        R t_r = dXr - (Anr - P.Ar);
        R t_ru = dXr * Lu3 - (Anru - P.Aru);
        R t_rt = dXr * Lt3 - (Anrt - P.Art);
        R t_rE = dXr * LE3 - (AnrE - P.ArE);
        if (DXP) {   // x / dx == x * (1/dx) bit for bit when dx is a power of two (compile-time: no branch in the step)
            const R idx(A.inv_dx);
            t_r = t_r * idx; t_ru = t_ru * idx; t_rt = t_rt * idx; t_rE = t_rE * idx;
        } else {
            t_r = D::quot(t_r, inv_dx, f); t_ru = D::quot(t_ru, inv_dx, f);
            t_rt = D::quot(t_rt, inv_dx, f); t_rE = D::quot(t_rE, inv_dx, f);
        }
        const typename D::Rcp inv_r = D::prepare_pos(t_r, f);
        const R o_ua = D::quot(t_ru, inv_r, f), o_ut = D::quot(t_rt, inv_r, f), o_E = D::quot(t_rE, inv_r, f);
        const R c_out(J == 3 ? I.cr[0] : I.cw[(J + 1) * 32]);   // c of cell a-7 (EOS of this sweep)
        const bool store = ok && J < I.rem;
        {   // dtCFL accumulators (src/reductions.jl:14-20), branch-free: cells that are not stored do not contribute
            const unsigned long long ba = (unsigned long long)__double_as_longlong((rabs(o_ua) + c_out).v);
            const unsigned long long bt = (unsigned long long)__double_as_longlong((rabs(o_ut) + c_out).v);
            T.amax = (store && ba > T.amax) ? ba : T.amax;
            T.tmax = (store && bt > T.tmax) ? bt : T.tmax;
        }
        if (TR == 1) {
            double *s = J == 3 ? I.s3 : I.s0 + J;
            s[0 * 32 * FK_PITCH] = t_r.v;
            s[1 * 32 * FK_PITCH] = o_ua.v;
            s[2 * 32 * FK_PITCH] = o_ut.v;
            s[3 * 32 * FK_PITCH] = o_E.v;
        } else if (store) {
            const long long o = I.o_it + J * A.pitch_out;
            A.out[0][o] = t_r.v;
            A.out[1][o] = o_ua.v;
            A.out[2][o] = o_ut.v;
            A.out[3][o] = o_E.v;
        }
    }

    // ---- commit ----
    P.Ar = Anr; P.Aru = Anru; P.Art = Anrt; P.ArE = AnrE;
    if (PROJ == ARMON_PROJ_EULER_2ND) {
        P.Sr = sr; P.Sru = sru; P.Srt = srt; P.SrE = srE;
        P.S2b = R(k2.b); P.S2r = R(k2.r);
    }
    P.cp[Z0] = A_p; P.crc[Z0] = A_rc;
    P.Gu[Z0] = A_Gu; P.Gp[Z0] = A_Gp;
    P.Fu[Z2] = B_Fu; P.Fp[Z2] = B_Fp; P.FpFu[Z2] = B_FpFu;
    P.disp[Z0] = C_disp; P.dxl[Z0] = C_dxl;
    LRING(0, Z0) = C_Lr.v; LRING(1, Z0) = C_Lu.v; LRING(2, Z0) = C_Lt.v; LRING(3, Z0) = C_LE.v;
#undef LRING
#undef ZS
}

__device__ __forceinline__ void fk_pipe_init(PipeF &P)
{
#pragma unroll
    for (int j = 0; j < 4; j++) {
        P.cu[j] = 0.; P.cp[j] = 1.; P.crc[j] = 1.; P.cdm[j] = 1.; P.cut[j] = 0.; P.cE[j] = 1.; P.Gu[j] = 0.; P.Gp[j] = 1.;
        P.Fu[j] = 0.; P.Fp[j] = 1.; P.FpFu[j] = 0.;
        P.dl[j] = 0.; P.dxl[j] = 1.; P.Lr[j] = 1.; P.Lru[j] = 0.; P.Lrt[j] = 0.; P.LrE[j] = 1.;
    }
#pragma unroll
    for (int j = 0; j < 2; j++)
#pragma unroll
        for (int k = 0; k < 4; k++) { P.T[j][k] = 0.; P.S[j][k] = 0.; P.Adv[j][k] = 0.; }
#pragma unroll
    for (int j = 0; j < 4; j++)
#pragma unroll
        for (int k = 0; k < 4; k++) P.Q[j][k] = 0.;
}

// the finite dummies of march_segment (sweep_kernel.cuh): warm-up results are never emitted
__device__ __forceinline__ void fk_pipe_init(PipeS &P)
{
    typedef sd R;
#pragma unroll
    for (int j = 0; j < 4; j++) {
        P.cp[j] = R(1.); P.crc[j] = R(1.);
        P.Gu[j] = R(0.); P.Gp[j] = R(1.); P.Fu[j] = R(0.); P.Fp[j] = R(1.); P.FpFu[j] = R(0.); P.disp[j] = R(0.);
        P.dxl[j] = R(1.);   // (Lagrangian rings: rho = E = 1, ua = ut = 0 from the initial fill of the staging ring)
    }
    P.Ar = R(0.); P.Aru = R(0.); P.Art = R(0.); P.ArE = R(0.);
    P.Sr = R(0.); P.Sru = R(0.); P.Srt = R(0.); P.SrE = R(0.); P.S2b = R(2.); P.S2r = R(0.5);
}

// L2 eviction policies (the encodings of createpolicy.fractional.L2::evict_first / evict_last with fraction 1.0).
// The transposed output leaves a warp as 64-byte pieces, one per output row every 8 steps; stored evict-last they stay
// in L2 until their neighbours along the row have arrived and go to DRAM as long runs instead of isolated lines
// (measured at 8192^2 / 16384^2: 0.885 -> 0.848 ms, 3.48 -> 3.40 ms per sweep).  Fetching the inputs evict-first, to
// leave the cache to those pieces, measured slower (0.98 ms) and is off.  In the tiled layout, where a store already
// completes half a tile, neither hint changes the time (2.94 ms either way): it keeps the same store helper.
#ifndef FK_LOAD_HINT
#define FK_LOAD_HINT 0
#endif
#ifndef FK_STORE_HINT
#define FK_STORE_HINT 1
#endif
constexpr unsigned long long FK_EVICT_FIRST = 0x12F0000000000000ULL, FK_EVICT_LAST = 0x14F0000000000000ULL;

// one aligned 32-byte sector per thread (256-bit store, sm_100), predicated on `ok`
__device__ __forceinline__ void fk_store4_if(bool ok, double *p, double a, double b, double c, double d)
{
#if FK_STORE_HINT
    asm volatile("{\n.reg .pred pq;\nsetp.ne.s32 pq, %6, 0;\n"
                 "@pq st.global.L2::cache_hint.v4.f64 [%0], {%1, %2, %3, %4}, %5;\n}"
                 ::"l"(p), "d"(a), "d"(b), "d"(c), "d"(d), "l"(FK_EVICT_LAST), "r"((int)ok) : "memory");
#else
    asm volatile("{\n.reg .pred pq;\nsetp.ne.s32 pq, %5, 0;\n@pq st.global.v4.f64 [%0], {%1, %2, %3, %4};\n}"
                 ::"l"(p), "d"(a), "d"(b), "d"(c), "d"(d), "r"((int)ok) : "memory");
#endif
}

__device__ __forceinline__ void fk_store2(double *p, double2 v)
{
#if FK_STORE_HINT
    asm volatile("st.global.L2::cache_hint.v2.f64 [%0], {%1, %2}, %3;" ::"l"(p), "d"(v.x), "d"(v.y), "l"(FK_EVICT_LAST) : "memory");
#else
    *reinterpret_cast<double2 *>(p) = v;
#endif
}

// Transposed store of one chunk: per variable 32 columns x FK_K march cells -> FK_K contiguous doubles of 32 output
// rows.  Fast path (full tile, even output pitch): 128-bit shared loads and global stores, 64 / FK_K output rows per
// warp instruction.  Ragged tiles (last columns, first / last chunk of a segment, odd pitch) go element by element.
__device__ __forceinline__ void fast_flush(const SweepArgs &A, const double *stage, long long w0, long long mb,
                                           long long m_lo, long long m1)
{
    const int lane = threadIdx.x & 31;
    __syncwarp();
    if (w0 + 32 <= A.nw && mb >= m_lo && mb + FK_K <= m1 && !(A.pitch_out & 1)) {
        constexpr int PIECES = FK_K / 2, ROWS = 32 / PIECES;   // 16-byte pieces per row, rows per warp instruction
        const int r0 = lane / PIECES, j = lane % PIECES;
        const double2 *src = reinterpret_cast<const double2 *>(stage + r0 * FK_PITCH + 2 * j);
#ifdef FK_EXP_G
        // TIMING EXPERIMENT ONLY, never defined in the product build (the results are wrong): the same flush, but the
        // 64-byte pieces of FK_EXP_G (2, 4, 8 or 32) adjacent output rows are written next to each other.  It measured
        // what the piece length of the transposed stores costs (profiles/README.md) and led to the tiled layout.
        static_assert(FK_EXP_G == 2 || FK_EXP_G == 4 || FK_EXP_G == 8 || FK_EXP_G == 32, "experiment sizes");
        const long long wr = w0 + r0;
        const long long off = (wr / FK_EXP_G) * (FK_EXP_G * A.pitch_out) + (mb >> 3) * (8 * FK_EXP_G) + (wr % FK_EXP_G) * 8 + 2 * j;
        const long long step = (ROWS >= FK_EXP_G) ? (ROWS / FK_EXP_G) * (FK_EXP_G * A.pitch_out) : ROWS * 8;
#else
        const long long off = (w0 + r0 + A.g) * A.pitch_out + (mb + A.g) + 2 * j;
        const long long step = ROWS * A.pitch_out;
#endif
#pragma unroll 1
        for (int v = 0; v < 4; v++) {
            double *dst = A.out[v] + off;
#pragma unroll
            for (int it = 0; it < 32 / ROWS; it++) {
                const double2 val = src[(v * 32 + it * ROWS) * FK_PITCH / 2];
                fk_store2(dst + it * step, val);
            }
        }
    } else {
        const int rsub = lane / FK_K, col = lane % FK_K;
#pragma unroll 1
        for (int v = 0; v < 4; v++) {
#pragma unroll 1
            for (int it = 0; it < 32 / (32 / FK_K); it++) {
                const int r = it * (32 / FK_K) + rsub;
                const double val = stage[(v * 32 + r) * FK_PITCH + col];
                const long long w = w0 + r, m = mb + col;
                if (w < A.nw && m >= m_lo && m < m1) A.out[v][(w + A.g) * A.pitch_out + (m + A.g)] = val;
            }
        }
    }
    __syncwarp();
}

// ---- mbarrier / bulk tensor copy ----------------------------------------------------------------------------------
__device__ __forceinline__ unsigned fk_smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void fk_mbar_init(unsigned bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fk_mbar_expect_tx(unsigned bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// Wait for the phase with parity `parity` of an mbarrier.  A copy that never completes (a malformed tensor map) must not
// hang the GPU: after ~2^24 expired try_wait periods the kernel traps instead.
__device__ __forceinline__ void fk_mbar_wait(unsigned bar, unsigned parity)
{
    unsigned done, spins = 0u;
    do {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        if (!done && ++spins > (1u << 24)) __trap();
    } while (!done);
}
// Non-blocking test of the same phase.  The phase check is a long-latency instruction even when the copies landed long
// ago; issued two steps before the group is needed, its result is there when the next iteration starts and the blocking
// wait is skipped (the profile showed 2.7 % of the warp samples sitting on the check at the top of the iteration).
__device__ __forceinline__ bool fk_mbar_test(unsigned bar, unsigned parity)
{
    unsigned done;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    return done != 0u;
}
// One lane of the (converged) warp: elect.sync tells the compiler that exactly one thread runs the guarded code, so the
// TMA instructions inside are issued straight from uniform registers (a `lane == 0` test makes it wrap each of them in
// a loop over the active lanes).
__device__ __forceinline__ bool fk_elect_one()
{
    unsigned pred = 0u;
    asm volatile(
        "{\n"
        ".reg .b32 rx;\n"
        ".reg .pred px;\n"
        "elect.sync rx|px, 0xffffffff;\n"
        "@px mov.u32 %0, 1;\n"
        "}\n" : "+r"(pred));
    return pred != 0u;
}

// [4 rows x 32 columns] box of one variable at (column c0, array row r0) -> shared memory, completion on `bar`
__device__ __forceinline__ void fk_tma_load_2d(unsigned dst, const CUtensorMap *map, int c0, int r0, unsigned bar)
{
#if FK_LOAD_HINT
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
                 " [%0], [%1, {%3, %4}], [%2], %5;"
                 ::"r"(dst), "l"(reinterpret_cast<unsigned long long>(map)), "r"(bar), "r"(c0), "r"(r0), "l"(FK_EVICT_FIRST)
                 : "memory");
#else
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(dst), "l"(reinterpret_cast<unsigned long long>(map)), "r"(bar), "r"(c0), "r"(r0)
                 : "memory");
#endif
}


__device__ __forceinline__ void async_copy8(unsigned dst, const double *src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
}

#ifndef FAST_MIN_BLOCKS
#define FAST_MIN_BLOCKS (256 / ASYNC_TPB_VALUE)   // 8 warps per SM
#endif

// CONS = 1: also accumulates the conservation sums of the cells it stores (per-cycle diagnostics fused into the last
// sweep of a cycle): per thread in march order, per warp by a fixed butterfly, one partial per warp for k_diag_final.
// MATH = MATH_STRICT: the reference's operation order with correctly rounded divisions (strict_step, section 6.) and the
// chunk-granular hand-over of out-of-range operands to sweep_fixup_kernel (chunk_end, sweep_staged_common.cuh).
// DXP (strict arithmetic only): 1 = the cell size is a power of two, the host checked it (SweepArgs::dx_pow2).
template <int STG, int RL, int PROJ, int EOS, int TR, int CONS = 0, int LAY = LAY_ROWS, int MATH = MATH_FAST, int DXP = 0>
__global__ void __launch_bounds__(ASYNC_TPB, FAST_MIN_BLOCKS)
sweep_fast_kernel(const SweepArgs A, const __grid_constant__ SweepTmaMaps M)
{
    static_assert(LAY == LAY_ROWS || STG == STG_TMA, "the tiled layout is staged by TMA only");
    static_assert(MATH == MATH_FAST || (LAY == LAY_ROWS && CONS == 0 && STG == STG_TMA), "strict arithmetic: row-major layouts, TMA staging");
    extern __shared__ __align__(128) unsigned char fast_smem_raw[];
    // warp index through a shuffle: the compiler then knows it (and every address derived from it: the warp's ring, its
    // barriers, its first column) is warp-uniform and keeps it in uniform registers, which the TMA instructions take
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
    FastWarpShared &S = reinterpret_cast<FastWarpShared *>(fast_smem_raw)[warp];

    // tiled: lane 0 <-> array column 32 k (whole tiles / bands), i.e. cell 32 k - g
    const long long w0 = (long long)blockIdx.x * ASYNC_TPB + warp * 32 - (LAY == LAY_TILED ? A.g : 0);
    const long long w = w0 + lane;
    // march segment k = [k seg - 4, (k+1) seg - 4): the first one starts with 4 virtual cells (masked), the last one
    // runs to the end of the domain
    const long long kseg = sweep_segment_index(A);
    const long long m0 = kseg * A.seg - 4;
    const long long m_lo = m0 < 0 ? 0 : m0;
    const long long m1 = (kseg == A.nseg - 1) ? A.nm : m0 + A.seg;

    SweepThread T;
    T.valid = w >= 0 && w < A.nw;
    T.col = (T.valid ? w : A.nw - 1) + A.g;
    T.amax = 0ULL; T.tmax = 0ULL;
    T.cm = 0.0; T.ce = 0.0;
    const long long cons_slot = (kseg * gridDim.x + blockIdx.x) * (ASYNC_TPB / 32) + warp;

    const DeviceTimeState *ts = A.ts;
    if (ts->done) {   // see sweep_kernel: copy the state through so that the host's buffer rotation stays valid
        if (CONS && lane == 0) { A.cons_m[cons_slot] = 0.0; A.cons_e[cons_slot] = 0.0; }   // the log line is dropped anyway
        if (T.valid) {
            for (long long m = m_lo; m < m1; m++) {
                const long long i = layout_index(LAY == LAY_TILED, m + A.g, T.col, A.pitch_in);
                const long long o = A.transpose_out ? layout_index(LAY == LAY_TILED, T.col, m + A.g, A.pitch_out)
                                                    : layout_index(LAY == LAY_TILED, m + A.g, T.col, A.pitch_out);
#pragma unroll
                for (int k = 0; k < 4; k++) A.out[k][o] = A.in[k][i];
            }
        }
        return;
    }
    if (w0 >= A.nw) {   // warp entirely outside the domain (warps are independent: no CTA barrier below)
        if (CONS && lane == 0) { A.cons_m[cons_slot] = 0.0; A.cons_e[cons_slot] = 0.0; }
        return;
    }

    const double dt = xmul(ts->current_dt, A.dt_factor);   // update_solver_state!, src/solver_state.jl:339-345
    const int len = (int)(m1 - m0);
    const bool first_seg = m0 < 0;
    const int nchunks = (len + FK_K - 1) / FK_K;
    const long long a_begin = m0 - 4;
    // Step t consumes array row a_begin + t and emits cell m0 + t - 11: 11 warm-up steps, FK_K per chunk of outputs, run
    // in groups of 4 (the last group needs 3 of its steps).
    const int n_groups = (FK_K / 4) * nchunks + 3;

    // benign finite state in the whole ring: the lagging stages read rows "before" the first one during warm-up, and
    // columns past the end of a row are never copied by the cp.async variants (rho = E = 1, u = v = 0, c = 1)
    for (int k = lane; k < FK_NG * FK_GS; k += 32) {
        const int v = (k / FK_VS) & 3;
        (&S.ring[0][0][0][0])[k] = (v == 0 || v == 3) ? 1.0 : 0.0;
    }
    for (int k = lane; k < FK_CS * 32; k += 32) (&S.cring[0][0])[k] = 1.0;

    // ---- staging ----
    // Group G holds array rows a_begin + 4G .. + 3 in ring slot G & 3.  Iteration `it` consumes group `it` (every value
    // a later step needs again travels in registers), so its slot is refilled with group it+4 at the end of the
    // iteration: 12 rows (12 KB per warp, ~14 MB per GPU) are in flight at any time, the copies run 12 steps ahead of
    // their use -- with 8 the profile showed 6 % of the warp time waiting on the barrier of the next group.
    // Prologue: groups 0 .. 3.
    const unsigned ring_u32 = fk_smem_u32(&S.ring[0][0][0][0]);
    const unsigned bar_u32 = fk_smem_u32(&S.full[0]);
    const int col0 = (int)(w0 + A.g), row0_arr = (int)(a_begin + A.g);
    AsyncLane L;   // cp.async variants: per-thread copy plan (see sweep_staged_common.cuh)
    L.active = false;
    if (STG == STG_TMA) {
        if (lane == 0) {
#pragma unroll
            for (int k = 0; k < FK_NG; k++) fk_mbar_init(bar_u32 + 8u * k, 1u);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the ring initialisation above vs the copy engine
        __syncwarp();
    } else {
        const int h = lane >> 4, piece = lane & 15;
        const long long cols = A.nw - w0 < 32 ? A.nw - w0 : 32;
        if (STG == STG_CPA16) {
            L.active = 2 * piece < cols;
#pragma unroll
            for (int j = 0; j < 2; j++) {
                L.src[j] = (h ? A.in[2 * j + 1] : A.in[2 * j]) + w0 + A.g + 2 * piece + (a_begin + A.g) * A.pitch_in;
                L.dst[j] = fk_smem_u32(&S.ring[0][2 * j + h][0][2 * piece]);
            }
        } else {
            L.active = lane < cols;
        }
        __syncwarp();
    }
    // issue the copies of group G (rows a_begin + 4G ..); rows past the last array row are skipped (cp.async; the
    // commit group is still formed, empty, so that the wait_group accounting holds) or zero-filled (TMA): they only
    // feed cells that are never stored
    auto issue_group = [&](int G) {
        // never start a copy that the warp will not wait for: it could land after the CTA has exited
        if (STG == STG_TMA) {
            if (G < n_groups && fk_elect_one()) {
                const unsigned bar = bar_u32 + 8u * (unsigned)fk_slot<MATH>(G);
                const unsigned dst = ring_u32 + (unsigned)fk_slot<MATH>(G) * (FK_GS * 8u);
                fk_mbar_expect_tx(bar, 4u * FK_VS * 8u);
#pragma unroll
                for (int v = 0; v < 4; v++) {
                    if (LAY == LAY_TILED)   // element 4 col0 of band row0_arr / 4 + G: the 4 tiles of columns col0 .. col0 + 31
                        fk_tma_load_2d(dst + v * (FK_VS * 8u), &M.m[v], 4 * col0, row0_arr / FK_GROUP + G, bar);
                    else
                        fk_tma_load_2d(dst + v * (FK_VS * 8u), &M.m[v], col0, row0_arr + FK_GROUP * G, bar);
                }
            }
        } else {
            const unsigned gbase = (unsigned)(G & (FK_NG - 1)) * (FK_GS * 8u);
#pragma unroll
            for (int r = 0; r < FK_GROUP; r++) {
                const long long row = a_begin + (long long)FK_GROUP * G + r;
                const bool row_ok = row >= -(long long)A.g && row <= A.nm + A.g - 1 && G < n_groups;
                const long long off = ((long long)FK_GROUP * G + r) * A.pitch_in;
                if (STG == STG_CPA16) {
                    if (L.active && row_ok) {
                        async_copy16(L.dst[0] + gbase + r * 256u, L.src[0] + off);
                        async_copy16(L.dst[1] + gbase + r * 256u, L.src[1] + off);
                    }
                } else if (L.active && row_ok) {
#pragma unroll
                    for (int v = 0; v < 4; v++)
                        async_copy8(ring_u32 + gbase + (unsigned)(v * FK_VS + r * 32 + lane) * 8u,
                                    A.in[v] + (a_begin + A.g) * A.pitch_in + off + w0 + A.g + lane);
                }
            }
            async_commit();
        }
    };
    // strict arithmetic: three staging groups, of which the one consumed by the previous iteration stays in the ring (the
    // chains re-read the inputs of the cells a-1 .. a-4 from it instead of carrying them in registers): groups it and
    // it + 1 are issued ahead
    constexpr int LEAD = MATH == MATH_STRICT ? 2 : FK_NG;
    issue_group(0);
    issue_group(1);
    if (LEAD == FK_NG) { issue_group(2); issue_group(3); }

    typename std::conditional<MATH == MATH_STRICT, PipeS, PipeF>::type P;
    fk_pipe_init(P);
    // strict arithmetic: range bookkeeping of the branch-free divisions, evaluated once per chunk of FK_K emitted cells
    typename Div<sd, DIV_FLAGGED>::Rcp inv_dx;
    inv_dx.b = A.dx; inv_dx.r = A.inv_dx;
    ChunkFix C;
    C.tot_a = 0ULL; C.tot_t = 0ULL; C.taint = 0; C.always = false;
    if (MATH == MATH_STRICT) {
        if (!DXP) inv_dx = Div<sd, DIV_FLAGGED>::prepare(sd(A.dx), T.flag);
        range_check_dividend(dt, T.flag);
        C.always = T.flag.bad();
    }

    const double *ring = &S.ring[0][0][0][0] + (LAY == LAY_TILED ? (lane >> 3) * 32 + (lane & 7) : lane);
    double *cring = &S.cring[0][lane];
    double *sbase = S.stage + lane * FK_PITCH;
    // tiled transposed stores: output row w + g = band (w0 + g) / 4 + lane / 4, row lane & 3 of its tiles
    const long long q_lane = (((w0 + A.g) >> 2) + (lane >> 2)) * (4 * A.pitch_out) + (lane & 3) * 8;
    FastIter I;
    I.q_off = 0; I.qok = false; I.row_mask = 0xfu;
    I.lr = const_cast<double *>(ring) + (FK_NG - 1) * FK_GS;

    // Iteration `it` = steps 4 it .. 4 it + 3 (J = 0 .. 3).
#define FK_STEP(Jv, EMITv, OK)                                                                              \
    if constexpr (MATH == MATH_STRICT) strict_step<RL, PROJ, EOS, Jv, TR, EMITv, DXP>(A, T, P, I, sd(dt), inv_dx, OK);  \
    else fast_step<RL, PROJ, EOS, Jv, TR, EMITv, CONS, LAY>(A, T, P, I, dt, OK);
#define FK_BEGIN(it)                                                                                        \
    {                                                                                                       \
        const int p_ = (it) & 1;                                                                            \
        I.gb0 = ring + fk_slot<MATH>(it) * FK_GS;                                                           \
        I.gbm = ring + fk_slot<MATH>((it) + (MATH == MATH_STRICT ? 2 : 3)) * FK_GS;   /* group it - 1 */    \
        I.cw = cring + p_ * 4 * 32;                                                                         \
        I.cr = cring + (p_ ^ 1) * 4 * 32;                                                                   \
        if (MATH == MATH_STRICT) {   /* rows a_begin + 4 it + J inside [-g, nm + g): bits J of the mask */          \
            const long long lo_ = -(long long)A.g - (a_begin + 4 * (it)), hi_ = A.nm + A.g - 1 - (a_begin + 4 * (it)); \
            const unsigned m_hi_ = hi_ >= 3 ? 0xfu : hi_ < 0 ? 0u : ((2u << (int)hi_) - 1u);                \
            const unsigned m_lo_ = lo_ <= 0 ? 0xfu : lo_ > 3 ? 0u : ((0xfu << (int)lo_) & 0xfu);            \
            I.row_mask = m_hi_ & m_lo_;                                                                     \
        }                                                                                                   \
        if (STG == STG_TMA) {                                                                               \
            if (!landed) fk_mbar_wait(bar_u32 + 8u * (unsigned)fk_slot<MATH>(it), fk_phase<MATH>(it));      \
        } else { async_wait<3>(); __syncwarp(); }                                                           \
        if (MATH == MATH_STRICT && I.row_mask != 0xfu) {                                                    \
            /* rows outside the array were zero-filled by the copy engine: give them a benign state (rho = 1e4, */ \
            /* E = 1: in range for both EOS) before anything reads them; they only feed cells never stored */ \
            double *g_ = const_cast<double *>(I.gb0);                                                       \
            _Pragma("unroll")                                                                               \
            for (int j_ = 0; j_ < 4; j_++)                                                                  \
                if (!((I.row_mask >> j_) & 1u)) { g_[j_ * 32] = 1e4; g_[j_ * 32 + 3 * FK_VS] = 1.0; }       \
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   /* ordered before the slot's next bulk copy */ \
            __syncwarp();                                                                                   \
        }                                                                                                   \
    }
    /* early look at the barrier of the next iteration's group */
#define FK_PEEK(it)                                                                                         \
    if (STG == STG_TMA) landed = fk_mbar_test(bar_u32 + 8u * (unsigned)fk_slot<MATH>((it) + 1), fk_phase<MATH>((it) + 1));
#define FK_END(it)                                                                                          \
    {                                                                                                       \
        __syncwarp();   /* every lane has read the last row of group it: refill its slot */                 \
        issue_group((it) + LEAD);                                                                           \
    }

    // warm-up: steps 0 .. 7 fill the head of the dependency cone of the first output, nothing is emitted
    int it = 0;
    bool landed = false;
#pragma unroll 1
    for (; it < 2; it++) {
        FK_BEGIN(it)
        FK_STEP(0, 0, false)
        FK_STEP(1, 0, false)
        FK_PEEK(it)
        FK_STEP(2, 0, false)
        FK_STEP(3, 0, false)
        FK_END(it)
    }
    // Steady state.  The cell emitted at step t is m0 + t - 11, its slot in the FK_K-cell transposed staging tile
    // (t - 11) & (FK_K - 1): the first three steps of an iteration fill slots (4 it + 5 ..+ 7) & (FK_K - 1), the last one
    // slot (4 it + 8) & (FK_K - 1); when that is 0 the tile was complete after the third step and is flushed there.
    // Steps 8 .. 10, the first three of the first iteration here, are the last warm-up steps: they run the emitting
    // code with stores and CFL maxima masked, and their tile slots are overwritten before the first flush.
#pragma unroll 1
    for (; it < n_groups; it++) {
        const bool live = it != 2;
        const int k3 = (4 * it + 8) & (FK_K - 1);
        // cells m0 .. m0 + 3 of the first segment are virtual (m < 0): emitted at the last step of iteration 2 and the
        // first three of iteration 3
        const bool ok012 = T.valid && live && !(first_seg && it == 3), ok3 = T.valid && !(first_seg && it == 2);
        FK_BEGIN(it)
        I.s0 = sbase + ((4 * it + 5) & (FK_K - 1));
        I.s3 = sbase + k3;
        I.rem = len - (4 * it - 11);
        if (TR == 1 && LAY == LAY_TILED) {
            // the unit completed at J = 2: cells m0 + 4 it - 12 .. - 9, i.e. columns mq .. mq + 3 of the output rows,
            // mq = m0 + 4 it - 12 + g a multiple of 4; the ends of a segment are multiples of 4 as well
            const long long mq = m0 + (4 * it - 12);
            I.qok = T.valid && mq >= m_lo && mq + 4 <= m1;
            I.q_off = q_lane + ((mq + A.g) >> 3) * 32 + ((mq + A.g) & 7);
        }
        if (TR == 0) {
            if (LAY == LAY_TILED)   // m0 + 4 it - 11 + g = 4 (band) + 1
                I.o_it = ((m0 + (4 * it - 11) + A.g) >> 2) * (4 * A.pitch_out) + ((w0 + A.g) >> 3) * 32 + (lane >> 3) * 32 + 8 + (lane & 7);
            else
                I.o_it = (m0 + (4 * it - 11) + A.g) * A.pitch_out + T.col;
        }
        FK_STEP(0, 1, ok012)
        FK_STEP(1, 1, ok012)
        FK_PEEK(it)
        FK_STEP(2, 1, ok012)
        if (TR == 1 && LAY == LAY_ROWS && k3 == 0 && live) fast_flush(A, S.stage, w0, m0 + (4 * it - 8 - FK_K), m_lo, m1);
        if (MATH == MATH_STRICT && k3 == 0 && live) {
            // the chunk [mb, mb + FK_K) is complete; an out-of-range operand met since the last check reaches cells
            // emitted at most 11 steps later, i.e. chunks k .. k+2 = the FIX_CHUNKS * SWEEP_CHUNK rows of a work-list entry
            const long long mb = m0 + (4 * it - 8 - FK_K);
            chunk_end<DIV_FLAGGED>(A, T, C, mb < 0 ? 0 : mb, w);
        }
        FK_STEP(3, 1, ok3)
        FK_END(it)
    }
#undef FK_STEP
#undef FK_PEEK
#undef FK_BEGIN
#undef FK_END
    if (STG != STG_TMA) async_wait<0>();

    unsigned long long am = T.valid ? (MATH == MATH_STRICT ? C.tot_a : T.amax) : 0ULL;
    unsigned long long tm = T.valid ? (MATH == MATH_STRICT ? C.tot_t : T.tmax) : 0ULL;
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
    if (CONS) {
        double cm = T.cm, ce = T.ce;   // 0 in the lanes without a real column
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {   // fixed butterfly: both lanes of a pair form the same sum
            cm = xadd(cm, __shfl_xor_sync(0xffffffffu, cm, off));
            ce = xadd(ce, __shfl_xor_sync(0xffffffffu, ce, off));
        }
        if (lane == 0) { A.cons_m[cons_slot] = cm; A.cons_e[cons_slot] = ce; }
    }
}
