// Instantiations of the software-pipelined cp.async-staged sweep kernel: number type sd, division policy DIV_FLAGGED, EOS ARMON_EOS_BIZARRIUM.
#include "sweep_dispatch.h"
ARMON_DEFINE_ASYNC2_TABLE(sweep_async2_table_strict_biz, sd, DIV_FLAGGED, ARMON_EOS_BIZARRIUM)
