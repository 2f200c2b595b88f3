// Instantiations of the fused sweep kernel: number type sd, division policy DIV_IEEE, EOS ARMON_EOS_PERFECT_GAS.
#include "sweep_dispatch.h"
ARMON_DEFINE_SWEEP_TABLE(sweep_table_ieee_pg, sd, DIV_IEEE, ARMON_EOS_PERFECT_GAS)
