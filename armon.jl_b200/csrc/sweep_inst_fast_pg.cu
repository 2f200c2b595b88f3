// Instantiations of the fused sweep kernel: arithmetic fd, EOS ARMON_EOS_PERFECT_GAS.
#include "sweep_dispatch.h"
ARMON_DEFINE_SWEEP_TABLE(sweep_table_fast_pg, fd, ARMON_EOS_PERFECT_GAS)
