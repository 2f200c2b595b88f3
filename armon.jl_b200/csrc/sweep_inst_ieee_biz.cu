// Instantiations of the fused sweep kernel: number type sd, division policy DIV_IEEE, EOS ARMON_EOS_BIZARRIUM.
#include "sweep_dispatch.h"
ARMON_DEFINE_SWEEP_TABLE(sweep_table_ieee_biz, sd, DIV_IEEE, ARMON_EOS_BIZARRIUM)
