// Instantiations of the fused sweep kernel: arithmetic sd, EOS ARMON_EOS_BIZARRIUM.
#include "sweep_dispatch.h"
ARMON_DEFINE_SWEEP_TABLE(sweep_table_strict_biz, sd, ARMON_EOS_BIZARRIUM)
