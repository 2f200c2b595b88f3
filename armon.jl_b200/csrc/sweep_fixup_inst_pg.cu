// Instantiations of the IEEE fix-up kernel of the warp-specialised sweep, EOS ARMON_EOS_PERFECT_GAS.
#include "sweep_dispatch.h"
ARMON_DEFINE_FIXUP_TABLE(sweep_fixup_table_pg, ARMON_EOS_PERFECT_GAS)
