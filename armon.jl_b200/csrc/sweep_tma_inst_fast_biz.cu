// Instantiations of the TMA-staged sweep kernel: number type fd, division policy DIV_FAST, EOS ARMON_EOS_BIZARRIUM.
#include "sweep_dispatch.h"
ARMON_DEFINE_TMA_TABLE(sweep_tma_table_fast_biz, fd, DIV_FAST, ARMON_EOS_BIZARRIUM)
