// sweep_dispatch.h -- tables of the fused sweep kernel instantiations (one translation unit per kernel family x
// arithmetic mode x EOS so that they compile in parallel).
//
// Families on the product path:
//   sweep_kernel        (register prefetch)  : every math mode; the fallback for odd input pitches and math_mode ieee
//   sweep_fast_kernel   (TMA / cp.async)     : four chains per step; math_mode fast (explicit arithmetic, band-tiled layout)
//                                              and math_mode strict (MATH_STRICT, bit-exact, + sweep_fixup_kernel)
#pragma once
#include "sweep_kernel.cuh"
#include "sweep_fixup_kernel.cuh"
#include "sweep_staged_common.cuh"

typedef void (*sweep_fn_t)(const SweepArgs);
typedef void (*sweep_fixup_fn_t)(const SweepArgs, const FixupArgs);

// rl: 0 = Godunov (acoustic!), 1 + limiter code = GAD with that limiter; proj: ARMON_PROJ_*
sweep_fn_t sweep_table_ieee_pg(int rl, int proj);
sweep_fn_t sweep_table_ieee_biz(int rl, int proj);
sweep_fn_t sweep_table_strict_pg(int rl, int proj);
sweep_fn_t sweep_table_strict_biz(int rl, int proj);
sweep_fn_t sweep_table_fast_pg(int rl, int proj);
sweep_fn_t sweep_table_fast_biz(int rl, int proj);

#define ARMON_DEFINE_SWEEP_TABLE(NAME, R, DIV, EOS)                                          \
    sweep_fn_t NAME(int rl, int proj)                                                       \
    {                                                                                       \
        static const sweep_fn_t table[4][2] = {                                             \
            {sweep_kernel<R, DIV, 0, ARMON_PROJ_EULER, EOS>, sweep_kernel<R, DIV, 0, ARMON_PROJ_EULER_2ND, EOS>}, \
            {sweep_kernel<R, DIV, 1, ARMON_PROJ_EULER, EOS>, sweep_kernel<R, DIV, 1, ARMON_PROJ_EULER_2ND, EOS>}, \
            {sweep_kernel<R, DIV, 2, ARMON_PROJ_EULER, EOS>, sweep_kernel<R, DIV, 2, ARMON_PROJ_EULER_2ND, EOS>}, \
            {sweep_kernel<R, DIV, 3, ARMON_PROJ_EULER, EOS>, sweep_kernel<R, DIV, 3, ARMON_PROJ_EULER_2ND, EOS>}, \
        };                                                                                  \
        if (rl < 0 || rl > 3 || proj < 0 || proj > 1) return nullptr;                       \
        return table[rl][proj];                                                             \
    }

// IEEE fix-up kernels of the strict sweep (sweep_fast_kernel<..., MATH_STRICT>)
sweep_fixup_fn_t sweep_fixup_table_pg(int rl, int proj);
sweep_fixup_fn_t sweep_fixup_table_biz(int rl, int proj);

#define ARMON_DEFINE_FIXUP_TABLE(NAME, EOS)                                                  \
    sweep_fixup_fn_t NAME(int rl, int proj)                                                 \
    {                                                                                       \
        static const sweep_fixup_fn_t table[4][2] = {                                       \
            {sweep_fixup_kernel<sd, 0, ARMON_PROJ_EULER, EOS>, sweep_fixup_kernel<sd, 0, ARMON_PROJ_EULER_2ND, EOS>}, \
            {sweep_fixup_kernel<sd, 1, ARMON_PROJ_EULER, EOS>, sweep_fixup_kernel<sd, 1, ARMON_PROJ_EULER_2ND, EOS>}, \
            {sweep_fixup_kernel<sd, 2, ARMON_PROJ_EULER, EOS>, sweep_fixup_kernel<sd, 2, ARMON_PROJ_EULER_2ND, EOS>}, \
            {sweep_fixup_kernel<sd, 3, ARMON_PROJ_EULER, EOS>, sweep_fixup_kernel<sd, 3, ARMON_PROJ_EULER_2ND, EOS>}, \
        };                                                                                  \
        if (rl < 0 || rl > 3 || proj < 0 || proj > 1) return nullptr;                       \
        return table[rl][proj];                                                             \
    }

// Fast-mode marching kernels (sweep_fast_kernel.cuh): explicit arithmetic, staged by TMA / cp.async; tr = 1: transposed
// output.  One table per staging variant and EOS.
#include "sweep_fast_kernel.cuh"
typedef void (*sweep_fast_fn_t)(const SweepArgs, const SweepTmaMaps);
sweep_fast_fn_t sweep_fast_table_tma_pg(int rl, int proj, int tr);
sweep_fast_fn_t sweep_fast_table_tma_biz(int rl, int proj, int tr);
sweep_fast_fn_t sweep_fast_table_cpa16_pg(int rl, int proj, int tr);
sweep_fast_fn_t sweep_fast_table_cpa16_biz(int rl, int proj, int tr);
sweep_fast_fn_t sweep_fast_table_cpa8_pg(int rl, int proj, int tr);
sweep_fast_fn_t sweep_fast_table_cpa8_biz(int rl, int proj, int tr);
// TMA staging + conservation sums (the last sweep of a cycle when the per-cycle diagnostics are on)
sweep_fast_fn_t sweep_fast_table_tma_cons_pg(int rl, int proj, int tr);
sweep_fast_fn_t sweep_fast_table_tma_cons_biz(int rl, int proj, int tr);
// band-tiled layout between the sweeps (LAY_TILED), TMA staging, without / with the conservation sums
sweep_fast_fn_t sweep_fast_table_tiled_pg(int rl, int proj, int tr);
sweep_fast_fn_t sweep_fast_table_tiled_biz(int rl, int proj, int tr);
sweep_fast_fn_t sweep_fast_table_tiled_cons_pg(int rl, int proj, int tr);
sweep_fast_fn_t sweep_fast_table_tiled_cons_biz(int rl, int proj, int tr);

// strict arithmetic on the schedule of the fast kernel (MATH_STRICT: TMA staging, row-major layouts, four chains per step)
// (_dxp: cell size a power of two -- x / dx is a multiplication, decided at compile time)
sweep_fast_fn_t sweep_fast_table_strict_pg(int rl, int proj, int tr);
sweep_fast_fn_t sweep_fast_table_strict_biz(int rl, int proj, int tr);
sweep_fast_fn_t sweep_fast_table_strict_dxp_pg(int rl, int proj, int tr);
sweep_fast_fn_t sweep_fast_table_strict_dxp_biz(int rl, int proj, int tr);

#define ARMON_FAST_ROW_M(STG, RLV, EOS, CONS, LAY, MATH, DXP)                                \
    {{sweep_fast_kernel<STG, RLV, ARMON_PROJ_EULER, EOS, 0, CONS, LAY, MATH, DXP>, sweep_fast_kernel<STG, RLV, ARMON_PROJ_EULER, EOS, 1, CONS, LAY, MATH, DXP>}, \
     {sweep_fast_kernel<STG, RLV, ARMON_PROJ_EULER_2ND, EOS, 0, CONS, LAY, MATH, DXP>, sweep_fast_kernel<STG, RLV, ARMON_PROJ_EULER_2ND, EOS, 1, CONS, LAY, MATH, DXP>}}

#define ARMON_DEFINE_FAST_TABLE_M(NAME, STG, EOS, CONS, LAY, MATH, DXP)                      \
    sweep_fast_fn_t NAME(int rl, int proj, int tr)                                          \
    {                                                                                       \
        static const sweep_fast_fn_t table[4][2][2] = {                                     \
            ARMON_FAST_ROW_M(STG, 0, EOS, CONS, LAY, MATH, DXP), ARMON_FAST_ROW_M(STG, 1, EOS, CONS, LAY, MATH, DXP), \
            ARMON_FAST_ROW_M(STG, 2, EOS, CONS, LAY, MATH, DXP), ARMON_FAST_ROW_M(STG, 3, EOS, CONS, LAY, MATH, DXP), \
        };                                                                                  \
        if (rl < 0 || rl > 3 || proj < 0 || proj > 1 || tr < 0 || tr > 1) return nullptr;   \
        return table[rl][proj][tr];                                                         \
    }
#define ARMON_DEFINE_FAST_TABLE(NAME, STG, EOS, CONS, LAY) ARMON_DEFINE_FAST_TABLE_M(NAME, STG, EOS, CONS, LAY, MATH_FAST, 0)
