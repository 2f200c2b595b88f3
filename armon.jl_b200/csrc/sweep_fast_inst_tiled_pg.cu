// Instantiations of the fast-mode sweep kernel: band-tiled layout, staging STG_TMA, EOS ARMON_EOS_PERFECT_GAS.
#include "sweep_dispatch.h"
ARMON_DEFINE_FAST_TABLE(sweep_fast_table_tiled_pg, STG_TMA, ARMON_EOS_PERFECT_GAS, 0, LAY_TILED)
