// Instantiations of the fast kernel's schedule with strict (bit-exact) arithmetic: staging STG_TMA, row-major layouts,
// EOS ARMON_EOS_PERFECT_GAS, cell sizes that are powers of two (x / dx == x * (1/dx) bit for bit).
#include "sweep_dispatch.h"
ARMON_DEFINE_FAST_TABLE_M(sweep_fast_table_strict_dxp_pg, STG_TMA, ARMON_EOS_PERFECT_GAS, 0, LAY_ROWS, MATH_STRICT, 1)
