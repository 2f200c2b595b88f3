// Instantiations of the fast-mode sweep kernel: band-tiled layout, staging STG_TMA, EOS ARMON_EOS_PERFECT_GAS, with the conservation
// sums of the per-cycle diagnostics accumulated in the sweep.
#include "sweep_dispatch.h"
ARMON_DEFINE_FAST_TABLE(sweep_fast_table_tiled_cons_pg, STG_TMA, ARMON_EOS_PERFECT_GAS, 1, LAY_TILED)
