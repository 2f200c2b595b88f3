// Instantiations of the warp-specialised sweep kernel: number type sd, division policy DIV_FLAGGED, EOS ARMON_EOS_BIZARRIUM.
#include "sweep_dispatch.h"
ARMON_DEFINE_WS_TABLE(sweep_ws_table_strict_biz, sd, DIV_FLAGGED, ARMON_EOS_BIZARRIUM)
