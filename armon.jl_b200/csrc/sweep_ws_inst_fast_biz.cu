// Instantiations of the warp-specialised sweep kernel: number type fd, division policy DIV_FAST, EOS ARMON_EOS_BIZARRIUM.
#include "sweep_dispatch.h"
ARMON_DEFINE_WS_TABLE(sweep_ws_table_fast_biz, fd, DIV_FAST, ARMON_EOS_BIZARRIUM)
