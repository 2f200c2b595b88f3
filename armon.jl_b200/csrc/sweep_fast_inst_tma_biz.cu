// Instantiations of the fast-mode sweep kernel: staging STG_TMA, EOS ARMON_EOS_BIZARRIUM.
#include "sweep_dispatch.h"
ARMON_DEFINE_FAST_TABLE(sweep_fast_table_tma_biz, STG_TMA, ARMON_EOS_BIZARRIUM, 0, LAY_ROWS)
