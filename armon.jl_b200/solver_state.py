"""Host-side time-step state machine used by the per-step (unfused) path.

Mirrors `GlobalTimeStep`, `update_dt!`, `next_cycle!` (src/solver_state.jl:26-166) and `SolverState`
(src/solver_state.jl:275-345) for the synchronous, single-block case.  The fused path keeps the same state on
the device (csrc/solver.cu: DeviceTimeState / k_cycle_step).
"""
import math

from .utils import Axis, solver_error


class GlobalTimeStep:
    def __init__(self):
        self.cycle = 0
        self.time = 0.0
        self.current_dt = 0.0
        self.next_cycle_dt = math.inf

    def reset(self, params):                       # reset!, src/solver_state.jl:58-67
        self.cycle = 0
        self.time = 0.0
        self.current_dt = params.Dt if params.cst_dt else 0.0
        self.next_cycle_dt = math.inf

    def update_dt(self, params, new_dt):           # update_dt!, src/solver_state.jl:102-142 (no MPI)
        previous_dt = self.current_dt
        if not math.isfinite(new_dt) or new_dt <= 0:
            solver_error("time", f"Invalid time step for cycle {self.cycle}: {new_dt}")
        elif previous_dt == 0:
            new_dt = params.cfl * new_dt
        else:
            new_dt = min(params.cfl * new_dt, 1.05 * previous_dt)
        self.next_cycle_dt = new_dt
        if self.current_dt == 0:
            self.current_dt = self.next_cycle_dt

    def next_cycle(self, params):                  # next_cycle!, src/solver_state.jl:145-166
        self.cycle += 1
        self.time += self.current_dt
        if params.cst_dt:
            self.current_dt = self.next_cycle_dt = params.Dt
            return
        self.current_dt = self.next_cycle_dt
        self.next_cycle_dt = math.inf


class SolverState:
    """Non-constant parameters of the solver for one block (src/solver_state.jl:275-305)."""

    def __init__(self, params, global_dt):
        self.dx = 0.0
        self.dt = 0.0
        self.axis = Axis.X
        self.splitting = params.axis_splitting
        self.riemann_scheme = params.riemann_scheme
        self.riemann_limiter = params.riemann_limiter
        self.projection_scheme = params.projection_scheme
        self.test_case = params.test
        self.global_dt = global_dt
        self.steps_ranges = params.steps_ranges[0]

    def update(self, params, axis, dt_factor):     # update_solver_state!, src/solver_state.jl:339-345
        i_ax = int(axis)
        self.dx = params.domain_size[i_ax] / params.global_grid[i_ax]
        self.dt = self.global_dt.current_dt * dt_factor
        self.axis = Axis(axis)
        self.steps_ranges = params.steps_ranges[i_ax]

    def reset(self):
        self.dx = 0.0
        self.dt = 0.0
        self.axis = Axis.X
