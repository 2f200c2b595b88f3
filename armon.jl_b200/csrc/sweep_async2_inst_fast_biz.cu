// Instantiations of the software-pipelined cp.async-staged sweep kernel: number type fd, division policy DIV_FAST, EOS ARMON_EOS_BIZARRIUM.
#include "sweep_dispatch.h"
ARMON_DEFINE_ASYNC2_TABLE(sweep_async2_table_fast_biz, fd, DIV_FAST, ARMON_EOS_BIZARRIUM)
