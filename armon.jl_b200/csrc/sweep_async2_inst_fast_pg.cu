// Instantiations of the software-pipelined cp.async-staged sweep kernel: number type fd, division policy DIV_FAST, EOS ARMON_EOS_PERFECT_GAS.
#include "sweep_dispatch.h"
ARMON_DEFINE_ASYNC2_TABLE(sweep_async2_table_fast_pg, fd, DIV_FAST, ARMON_EOS_PERFECT_GAS)
