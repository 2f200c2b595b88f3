// Instantiations of the fused sweep kernel: arithmetic sd, EOS ARMON_EOS_PERFECT_GAS.
#include "sweep_dispatch.h"
ARMON_DEFINE_SWEEP_TABLE(sweep_table_strict_pg, sd, ARMON_EOS_PERFECT_GAS)
