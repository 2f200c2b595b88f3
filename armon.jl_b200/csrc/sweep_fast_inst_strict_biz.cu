// Instantiations of the fast kernel's schedule with strict (bit-exact) arithmetic: staging STG_TMA, row-major layouts,
// EOS ARMON_EOS_BIZARRIUM, any cell size (x / dx is a division).
#include "sweep_dispatch.h"
ARMON_DEFINE_FAST_TABLE_M(sweep_fast_table_strict_biz, STG_TMA, ARMON_EOS_BIZARRIUM, 0, LAY_ROWS, MATH_STRICT, 0)
