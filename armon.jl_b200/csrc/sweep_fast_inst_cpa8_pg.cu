// Instantiations of the fast-mode sweep kernel: staging STG_CPA8, EOS ARMON_EOS_PERFECT_GAS.
#include "sweep_dispatch.h"
ARMON_DEFINE_FAST_TABLE(sweep_fast_table_cpa8_pg, STG_CPA8, ARMON_EOS_PERFECT_GAS, 0, LAY_ROWS)
