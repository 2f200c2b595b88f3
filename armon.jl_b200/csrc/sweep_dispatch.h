// sweep_dispatch.h -- tables of the fused sweep kernel instantiations (one translation unit per kernel family x
// arithmetic mode x EOS so that they compile in parallel).
//
// Families on the product path:
//   sweep_kernel        (register prefetch)  : every math mode; the fallback for odd input pitches and math_mode ieee
//   sweep_async_kernel  (cp.async staging)   : math_mode strict (bit-exact), + sweep_fixup_kernel
//   sweep_fast_kernel   (TMA / cp.async)     : math_mode fast, explicit arithmetic, 4 chains per step
#pragma once
#include "sweep_kernel.cuh"
#include "sweep_fixup_kernel.cuh"
#include "sweep_async_kernel.cuh"

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

// IEEE fix-up kernels of the strict cp.async sweep
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

// cp.async-staged marching kernels (sweep_async_kernel.cuh); tr = 1: transposed output.
sweep_fn_t sweep_async_table_strict_pg(int rl, int proj, int tr);
sweep_fn_t sweep_async_table_strict_biz(int rl, int proj, int tr);

#define ARMON_ASYNC_ROW(R, DIV, RLV, EOS)                                                    \
    {{sweep_async_kernel<R, DIV, RLV, ARMON_PROJ_EULER, EOS, 0>, sweep_async_kernel<R, DIV, RLV, ARMON_PROJ_EULER, EOS, 1>}, \
     {sweep_async_kernel<R, DIV, RLV, ARMON_PROJ_EULER_2ND, EOS, 0>, sweep_async_kernel<R, DIV, RLV, ARMON_PROJ_EULER_2ND, EOS, 1>}}

#define ARMON_DEFINE_ASYNC_TABLE(NAME, R, DIV, EOS)                                          \
    sweep_fn_t NAME(int rl, int proj, int tr)                                               \
    {                                                                                       \
        static const sweep_fn_t table[4][2][2] = {                                          \
            ARMON_ASYNC_ROW(R, DIV, 0, EOS), ARMON_ASYNC_ROW(R, DIV, 1, EOS),               \
            ARMON_ASYNC_ROW(R, DIV, 2, EOS), ARMON_ASYNC_ROW(R, DIV, 3, EOS),               \
        };                                                                                  \
        if (rl < 0 || rl > 3 || proj < 0 || proj > 1 || tr < 0 || tr > 1) return nullptr;   \
        return table[rl][proj][tr];                                                         \
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
sweep_fast_fn_t sweep_fast_table_strict_pg(int rl, int proj, int tr);
sweep_fast_fn_t sweep_fast_table_strict_biz(int rl, int proj, int tr);

#define ARMON_FAST_ROW_M(STG, RLV, EOS, CONS, LAY, MATH)                                     \
    {{sweep_fast_kernel<STG, RLV, ARMON_PROJ_EULER, EOS, 0, CONS, LAY, MATH>, sweep_fast_kernel<STG, RLV, ARMON_PROJ_EULER, EOS, 1, CONS, LAY, MATH>}, \
     {sweep_fast_kernel<STG, RLV, ARMON_PROJ_EULER_2ND, EOS, 0, CONS, LAY, MATH>, sweep_fast_kernel<STG, RLV, ARMON_PROJ_EULER_2ND, EOS, 1, CONS, LAY, MATH>}}

#define ARMON_DEFINE_FAST_TABLE_M(NAME, STG, EOS, CONS, LAY, MATH)                           \
    sweep_fast_fn_t NAME(int rl, int proj, int tr)                                          \
    {                                                                                       \
        static const sweep_fast_fn_t table[4][2][2] = {                                     \
            ARMON_FAST_ROW_M(STG, 0, EOS, CONS, LAY, MATH), ARMON_FAST_ROW_M(STG, 1, EOS, CONS, LAY, MATH), \
            ARMON_FAST_ROW_M(STG, 2, EOS, CONS, LAY, MATH), ARMON_FAST_ROW_M(STG, 3, EOS, CONS, LAY, MATH), \
        };                                                                                  \
        if (rl < 0 || rl > 3 || proj < 0 || proj > 1 || tr < 0 || tr > 1) return nullptr;   \
        return table[rl][proj][tr];                                                         \
    }
#define ARMON_DEFINE_FAST_TABLE(NAME, STG, EOS, CONS, LAY) ARMON_DEFINE_FAST_TABLE_M(NAME, STG, EOS, CONS, LAY, MATH_FAST)
