// Instantiations of the fused sweep kernel: number type sd, division policy DIV_FLAGGED, EOS ARMON_EOS_BIZARRIUM.
#include "sweep_dispatch.h"
ARMON_DEFINE_SWEEP_TABLE(sweep_table_strict_biz, sd, DIV_FLAGGED, ARMON_EOS_BIZARRIUM)
