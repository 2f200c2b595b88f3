"""Step orchestration: the host-side mirror of src/solver.jl, src/kernels.jl (wrappers), src/riemann_schemes.jl
(wrappers), src/projection_schemes.jl (wrappers), src/halo_exchange.jl (drivers) and src/reductions.jl (drivers)
for `ArmonParameters{T, <:B200Device}`.

Two paths, both through the C ABI only:
  * fused (default): `solver_cycle` enqueues one marching kernel per axis sweep, the time step lives on the device
    (`armon_solver_run` / `armon_solver_time_loop`);
  * per-step (`fused=False`): one library call per reference kernel, host-side dt state machine -- the overloads of
    update_EOS!, block_ghost_exchange, numerical_fluxes!, cell_update!, projection_remap!, used for step-by-step
    comparison with the oracle (the reference's `compare=true` checkpoints, src/io.jl:185-227).
"""
import ctypes as C
import math
import time as _time
from dataclasses import dataclass, field

from .backend import armon_domain, check
from .blocks import BlockGrid, fill_test_case
from . import backend
from .parameters import block_domain_range
from .schemes import limiter_code, split_axes
from .utils import Axis, Side, first_side, last_side, solver_error


@dataclass
class SolverStats:
    """src/solver.jl:13-23"""
    final_time: float
    last_dt: float
    cycles: int
    solve_time: float            # seconds (host wall clock around time_loop, like the reference)
    cell_count: int
    giga_cells_per_sec: float
    data: object = None
    device_ms: float = 0.0       # device time of the cycles (CUDA events on the solver's stream)
    timer: object = None
    grid_log: object = field(default=None, repr=False)


def _dom(params, corners):
    return armon_domain(*block_domain_range(params.N, corners))


# --------------------------------------------------------------------------------------------------------
# Per-step wrappers (kernel seam)
# --------------------------------------------------------------------------------------------------------
def init_test(params, grid):
    """init_test(params, grid), src/kernels.jl:176-214"""
    if params.fused:
        grid.init_fused()
        return
    d = grid.device_data
    tc = fill_test_case(backend.armon_test_case(), params.test)
    ds = (C.c_double * 2)(*params.domain_size)
    org = (C.c_double * 2)(*params.origin)
    check(grid.lib.armon_init_test(
        grid.device.ctx, grid.dims, params.N_origin[0], params.N_origin[1], params.global_grid[0],
        params.global_grid[1], ds, org, C.byref(tc),
        d.x.ptr, d.y.ptr, d.mask.ptr, d.rho.ptr, d.E.ptr, d.u.ptr, d.v.ptr, d.p.ptr, d.c.ptr, d.g.ptr,
        d.us.ptr, d.ps.ptr, d.work_1.ptr, d.work_2.ptr, d.work_3.ptr, d.work_4.ptr), "armon_init_test")


def update_EOS(params, state, grid):
    """update_EOS!, src/kernels.jl:151-173"""
    d = grid.device_data
    dom = _dom(params, state.steps_ranges.EOS)
    if params.test.bizarrium_eos:
        check(grid.lib.armon_bizarrium_EOS(grid.device.ctx, grid.dims, dom, d.rho.ptr, d.u.ptr, d.v.ptr, d.E.ptr,
                                           d.p.ptr, d.c.ptr, d.g.ptr), "armon_bizarrium_EOS")
    else:
        check(grid.lib.armon_perfect_gas_EOS(grid.device.ctx, grid.dims, dom, params.test.specific_heat_ratio(),
                                             d.rho.ptr, d.E.ptr, d.u.ptr, d.v.ptr, d.p.ptr, d.c.ptr, d.g.ptr),
              "armon_perfect_gas_EOS")


def boundary_conditions(params, state, grid, side):
    """boundary_conditions!(params, state, blk, side), src/halo_exchange.jl:32-36"""
    d = grid.device_data
    u_factor, v_factor = state.test_case.boundary_condition(side)
    check(grid.lib.armon_boundary_conditions(grid.device.ctx, grid.dims, int(side), float(u_factor), float(v_factor),
                                             d.rho.ptr, d.u.ptr, d.v.ptr, d.p.ptr, d.c.ptr, d.g.ptr, d.E.ptr),
          "armon_boundary_conditions")


def block_ghost_exchange(params, state, grid):
    """block_ghost_exchange, src/halo_exchange.jl:286-368: the two sides along the current axis."""
    for side in (first_side(state.axis), last_side(state.axis)):
        if params.neighbours[side] >= 0:
            solver_error("config", "the per-step path runs on one sub-domain; use the fused path with use_MPI=true")
        boundary_conditions(params, state, grid, side)


def numerical_fluxes(params, state, grid):
    """numerical_fluxes!, src/riemann_schemes.jl:46-52,107-123"""
    d = grid.device_data
    dom = _dom(params, state.steps_ranges.fluxes)
    ua = d.u if state.axis == Axis.X else d.v
    if state.riemann_scheme == "GAD":
        check(grid.lib.armon_acoustic_GAD(grid.device.ctx, grid.dims, dom, int(state.axis), state.dt, state.dx,
                                          limiter_code(state.riemann_limiter), d.us.ptr, d.ps.ptr, d.rho.ptr, ua.ptr,
                                          d.p.ptr, d.c.ptr), "armon_acoustic_GAD")
    else:
        check(grid.lib.armon_acoustic(grid.device.ctx, grid.dims, dom, int(state.axis), d.us.ptr, d.ps.ptr, d.rho.ptr,
                                      ua.ptr, d.p.ptr, d.c.ptr), "armon_acoustic")


def cell_update(params, state, grid):
    """cell_update!, src/kernels.jl:217-230"""
    d = grid.device_data
    dom = _dom(params, state.steps_ranges.cell_update)
    ua = d.u if state.axis == Axis.X else d.v
    check(grid.lib.armon_cell_update(grid.device.ctx, grid.dims, dom, int(state.axis), state.dx, state.dt,
                                     d.us.ptr, d.ps.ptr, d.rho.ptr, ua.ptr, d.E.ptr), "armon_cell_update")


def advection_fluxes(params, state, grid):
    """advection_fluxes!, src/projection_schemes.jl:81-145"""
    d = grid.device_data
    dom = _dom(params, state.steps_ranges.advection)
    if state.projection_scheme == "euler_2nd":
        check(grid.lib.armon_advection_second_order(
            grid.device.ctx, grid.dims, dom, int(state.axis), state.dx, state.dt, d.us.ptr, d.rho.ptr, d.u.ptr,
            d.v.ptr, d.E.ptr, d.work_1.ptr, d.work_2.ptr, d.work_3.ptr, d.work_4.ptr), "armon_advection_second_order")
    else:
        check(grid.lib.armon_advection_first_order(
            grid.device.ctx, grid.dims, dom, int(state.axis), state.dt, d.us.ptr, d.rho.ptr, d.u.ptr, d.v.ptr,
            d.E.ptr, d.work_1.ptr, d.work_2.ptr, d.work_3.ptr, d.work_4.ptr), "armon_advection_first_order")


def euler_projection(params, state, grid):
    """euler_projection!, src/projection_schemes.jl:44-59"""
    d = grid.device_data
    dom = _dom(params, state.steps_ranges.projection)
    check(grid.lib.armon_euler_projection(
        grid.device.ctx, grid.dims, dom, int(state.axis), state.dx, state.dt, d.us.ptr, d.rho.ptr, d.u.ptr, d.v.ptr,
        d.E.ptr, d.work_1.ptr, d.work_2.ptr, d.work_3.ptr, d.work_4.ptr), "armon_euler_projection")


def projection_remap(params, state, grid):
    """projection_remap!, src/projection_schemes.jl:148-157"""
    advection_fluxes(params, state, grid)
    euler_projection(params, state, grid)


def local_time_step(params, state, grid):
    """local_time_step / dtCFL_kernel, src/reductions.jl:56-110"""
    d = grid.device_data
    dX = params.cell_size()
    res = C.c_double()
    check(grid.lib.armon_dtCFL(grid.device.ctx, grid.dims, d.u.ptr, d.v.ptr, d.c.ptr, dX[0], dX[1], C.byref(res)),
          "armon_dtCFL")
    return res.value


def next_time_step(params, state, grid):
    """next_time_step(params, state, grid), src/reductions.jl:164-199 (per-step path)"""
    gdt = state.global_dt
    if params.cst_dt:
        state.dt = params.Dt
        return False
    local_dt = local_time_step(params, state, grid)
    gdt.update_dt(params, local_dt)
    state.dt = gdt.current_dt
    return False


def conservation_vars(params, grid):
    """conservation_vars, src/reductions.jl:202-323 -> (mass, energy): blocks of this process summed in block order,
    then over the ranks."""
    grid.finalize()
    dX = params.cell_size()
    mass = energy = 0.0
    for blk in grid.blocks:
        d = blk.device_data
        m, e = C.c_double(), C.c_double()
        check(grid.lib.armon_conservation_vars(grid.device.ctx, blk.dims, d.rho.ptr, d.E.ptr, dX[0] * dX[1],
                                               C.byref(m), C.byref(e)), "armon_conservation_vars")
        mass += m.value
        energy += e.value
    if params.use_MPI and params.proc_size > 1:
        from .distributed import allreduce_sum
        mass, energy = allreduce_sum((mass, energy))
    return mass, energy


# --------------------------------------------------------------------------------------------------------
# solver_cycle / time_loop / armon
# --------------------------------------------------------------------------------------------------------
def solver_cycle(params, grid):
    """solver_cycle(params, grid), src/solver.jl:288-320.  Returns True to stop (never, on this backend)."""
    if params.fused:
        grid.run(1)
        return False
    state = grid.state
    from .io import step_checkpoint as _cp

    def checkpoint(label):                         # @checkpoint, src/solver.jl:41-43
        return _cp(params, state, grid, label)

    if state.global_dt.cycle == 0:
        state.update(params, Axis.X, 1.0)
        if checkpoint("init_test"):
            return True
        update_EOS(params, state, grid)            # "EOS_init"
        if checkpoint("EOS_init"):
            return True
    if next_time_step(params, state, grid):
        return True
    if checkpoint("time_step"):
        return True
    for axis, dt_factor in split_axes(state.splitting, state.global_dt.cycle):
        state.update(params, axis, dt_factor)
        for step, label in ((update_EOS, "EOS"), (block_ghost_exchange, "boundary_conditions"),
                            (numerical_fluxes, "numerical_fluxes"), (cell_update, "cell_update"),
                            (projection_remap, "projection_remap")):
            step(params, state, grid)
            if checkpoint(label):
                return True
    return False


def _sync_fused_state(params, grid):
    st = grid.time_state()
    gdt = grid.global_dt
    gdt.cycle, gdt.time, gdt.current_dt, gdt.next_cycle_dt = st.cycle, st.time, st.current_dt, st.next_cycle_dt
    if st.error == backend.ARMON_ERR_RANGE:
        solver_error("cpp", f"cycle {st.error_cycle}: the work list of the strict mode's IEEE fix-up overflowed; "
                            "rerun with math_mode='ieee'")
    if st.error:
        solver_error("time", f"Invalid time step for cycle {st.error_cycle}")
    return st


def _print_cycle_line(params, cycle, dt, t, mass, energy):
    """The `silent <= 1` line of time_loop (src/solver.jl:359-371)."""
    if not params.is_root:
        return
    dM = abs(params.initial_mass - mass) / params.initial_mass * 100 if params.initial_mass else 0.0
    dE = abs(params.initial_energy - energy) / params.initial_energy * 100 if params.initial_energy else 0.0
    print(f"Cycle {cycle:4d}: dt = {dt:.18f}, t = {t:.18f}, |ΔM| = {dM:#8.6g}%, |ΔE| = {dE:#8.6g}%")


def time_loop(params, grid):
    """time_loop(params, grid), src/solver.jl:323-403 -> (time, current_dt, cycle, cells_per_sec, solve_time)"""
    grid.reset()
    gdt = grid.global_dt
    t1 = _time.perf_counter()
    single_rank = not (params.use_MPI and params.proc_size > 1)
    if params.fused and params.animation_step == 0 and (params.silent > 1 or single_rank):
        # whole loop on the device: exact `while time < maxtime && cycle < maxcycle`.  The per-cycle log of
        # `silent <= 1` is produced there too (a reduction over the state each cycle leaves, appended to a device
        # ring): no finalize, no blocking read per cycle; the lines are printed when the loop returns.
        verbose = params.silent <= 1
        if verbose:
            grid.diagnostics(max(16, min(params.maxcycle + 1, 1 << 20)))
        grid.run_time_loop()
        _sync_fused_state(params, grid)
        if verbose:
            grid.cycle_log = grid.read_diagnostics()
            for cycle, t, dt, mass, energy in grid.cycle_log:
                _print_cycle_line(params, cycle, dt, t, mass, energy)
            grid.diagnostics(0)
    else:
        while gdt.time < params.maxtime and gdt.cycle < params.maxcycle:
            if solver_cycle(params, grid):
                break
            if params.fused:
                _sync_fused_state(params, grid)
            else:
                gdt.next_cycle(params)
            if params.silent <= 1:
                mass, energy = conservation_vars(params, grid)
                _print_cycle_line(params, gdt.cycle, gdt.current_dt, gdt.time, mass, energy)
    params.backend_options.wait()                  # "Last fence"
    t2 = _time.perf_counter()
    solve_time = t2 - t1
    cells = params.N[0] * params.N[1]
    grind_time = solve_time / (gdt.cycle * cells) if gdt.cycle else math.inf
    if params.is_root and params.silent < 3:
        print(" ")
        print(f"Total time:  {solve_time:.5f} sec")
        print(f"Grind time:  {grind_time * 1e6:.5f} µs/cell/cycle")
        print(f"Cells/sec:   {1 / grind_time / 1e6:.5f} Mega cells/sec")
        print(f"Cycles:      {gdt.cycle}")
        print(f"Last cycle:  {gdt.time:.18f} sec, Δt={gdt.current_dt:.18f} sec")
    return gdt.time, gdt.current_dt, gdt.cycle, 1 / grind_time if gdt.cycle else 0.0, solve_time


def armon(params):
    """armon(params) -> SolverStats, src/solver.jl:411-516"""
    grid = BlockGrid(params)
    init_test(params, grid)
    if params.check_result or params.silent <= 1:
        params.initial_mass, params.initial_energy = conservation_vars(params, grid)
    final_time, dt, cycles, cells_per_sec, solve_time = time_loop(params, grid)
    device_ms = 0.0
    if params.fused and cycles > 0:
        device_ms = grid.elapsed_ms()
    if params.check_result and params.test.is_conservative:
        mass, energy = conservation_vars(params, grid)
        dM = abs(params.initial_mass - mass)
        dE = abs(params.initial_energy - energy)
        tol = params.comparison_tolerance
        if dM > tol * abs(params.initial_mass) or dE > tol * abs(params.initial_energy):
            print(f"WARNING: mass and energy are not constant: |ΔM| = {dM:.3g}, |ΔE| = {dE:.3g}")
    if params.write_output:
        from .io import write_sub_domain_file
        write_sub_domain_file(params, grid, params.output_file)
    stats = SolverStats(final_time, dt, cycles, solve_time, params.N[0] * params.N[1], cells_per_sec / 1e9,
                        grid if params.return_data else None, device_ms)
    if not params.return_data:
        grid.close()
    return stats
