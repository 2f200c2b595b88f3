// Instantiations of the fast-mode sweep kernel: band-tiled layout, staging STG_TMA, EOS ARMON_EOS_BIZARRIUM.
#include "sweep_dispatch.h"
ARMON_DEFINE_FAST_TABLE(sweep_fast_table_tiled_biz, STG_TMA, ARMON_EOS_BIZARRIUM, 0, LAY_TILED)
