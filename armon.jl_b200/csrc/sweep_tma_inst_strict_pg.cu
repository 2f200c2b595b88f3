// Instantiations of the TMA-staged sweep kernel: number type sd, division policy DIV_FLAGGED, EOS ARMON_EOS_PERFECT_GAS.
#include "sweep_dispatch.h"
ARMON_DEFINE_TMA_TABLE(sweep_tma_table_strict_pg, sd, DIV_FLAGGED, ARMON_EOS_PERFECT_GAS)
