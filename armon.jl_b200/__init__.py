"""armon.jl_b200 -- B200-native backend for Armon.jl's axis-split Lagrange+remap time step.

Host-side mirror of the reference's public API for this path (`ArmonParameters`, `armon`, `BlockGrid`,
`init_test`, `time_loop`, `solver_cycle`, `conservation_vars`, ...) above the C-ABI library
`libarmon_b200.so` (include/armon_b200.h).  There is no CPU fallback: every compute entry point raises if
the CUDA library cannot be loaded or no B200 is visible.
"""
from .utils import Axis, Side, SolverException, solver_error
from .test_cases import (TestCase, Sod, Sod_y, Sod_circ, Bizarrium, Sedov, DebugIndexes,
                         test_from_name, create_test)
from .schemes import split_axes, stencil_width
from .parameters import ArmonParameters, StepsRanges, block_domain_range
from .solver_state import GlobalTimeStep, SolverState
from .backend import B200Device, B200Array, load_library, device_count, LIB_PATH
from .blocks import BlockGrid, BlockData, BLOCK_VARS, MAIN_VARS, SAVED_VARS, COMM_VARS
from .solver import (SolverStats, armon, time_loop, solver_cycle, init_test, update_EOS, boundary_conditions,
                     block_ghost_exchange, numerical_fluxes, cell_update, advection_fluxes, euler_projection,
                     projection_remap, local_time_step, next_time_step, conservation_vars)

__all__ = [
    "Axis", "Side", "SolverException", "solver_error",
    "TestCase", "Sod", "Sod_y", "Sod_circ", "Bizarrium", "Sedov", "DebugIndexes", "test_from_name", "create_test",
    "split_axes", "stencil_width", "ArmonParameters", "StepsRanges", "block_domain_range",
    "GlobalTimeStep", "SolverState", "B200Device", "B200Array", "load_library", "device_count", "LIB_PATH",
    "BlockGrid", "BlockData", "BLOCK_VARS", "MAIN_VARS", "SAVED_VARS", "COMM_VARS",
    "SolverStats", "armon", "time_loop", "solver_cycle", "init_test", "update_EOS", "boundary_conditions",
    "block_ghost_exchange", "numerical_fluxes", "cell_update", "advection_fluxes", "euler_projection",
    "projection_remap", "local_time_step", "next_time_step", "conservation_vars",
]
