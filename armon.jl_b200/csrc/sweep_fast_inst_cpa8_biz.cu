// Instantiations of the fast-mode sweep kernel: staging STG_CPA8, EOS ARMON_EOS_BIZARRIUM.
#include "sweep_dispatch.h"
ARMON_DEFINE_FAST_TABLE(sweep_fast_table_cpa8_biz, STG_CPA8, ARMON_EOS_BIZARRIUM, 0, LAY_ROWS)
