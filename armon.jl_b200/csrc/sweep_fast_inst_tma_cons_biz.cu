// Instantiations of the fast-mode sweep kernel: staging STG_TMA, EOS ARMON_EOS_BIZARRIUM, with the conservation sums of the per-cycle
// diagnostics accumulated in the sweep.
#include "sweep_dispatch.h"
ARMON_DEFINE_FAST_TABLE(sweep_fast_table_tma_cons_biz, STG_TMA, ARMON_EOS_BIZARRIUM, 1, LAY_ROWS)
