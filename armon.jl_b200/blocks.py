"""Device storage: `BlockData` (src/blocking/blocks.jl:18-50) and a one-block `BlockGrid`
(src/blocking/block_grid.jl:46-183 with `block_size = N .+ 2nghost`).

The 16 variables of the reference exist by name; those the fused path never touches (x, y, mask, us, ps and,
until asked for, p, c, g) are allocated on first access so that production runs only hold the 8 arrays of the
ping-pong (rho, u, v, E + work_1..4).
"""
import ctypes as C
import weakref

import numpy as np

from . import backend
from .backend import B200Device, armon_dims, check
from .schemes import limiter_code, projection_code, riemann_code, splitting_code
from .solver_state import GlobalTimeStep, SolverState
from .utils import Side, solver_error

BLOCK_VARS = ("x", "y", "rho", "u", "v", "E", "p", "c", "g", "us", "ps",
              "work_1", "work_2", "work_3", "work_4", "mask")          # block_vars()
MAIN_VARS = ("x", "y", "rho", "u", "v", "E", "p", "c", "g", "us", "ps")  # main_vars(): host <-> device
SAVED_VARS = ("x", "y", "rho", "u", "v", "p")                          # saved_vars(): I/O
COMM_VARS = ("rho", "u", "v", "E", "p", "c", "g")                      # comm_vars(): ghost exchange


class BlockData:
    """Lazily allocated struct of device arrays."""

    def __init__(self, device, size):
        self._device = device
        self._size = int(size)
        self._arrays = {}

    def __getattr__(self, name):
        if name.startswith("_") or name not in BLOCK_VARS:
            raise AttributeError(name)
        arr = self._arrays.get(name)
        if arr is None:
            arr = self._arrays[name] = self._device.array(self._size)
        return arr

    def allocated(self):
        return tuple(self._arrays)

    def ptr(self, name, allocate=True):
        if not allocate and name not in self._arrays:
            return C.c_void_p(None)
        return getattr(self, name).ptr


def fill_test_case(tc, test):
    tc.test = test.code
    tc.eos = 1 if test.bizarrium_eos else 0
    for k, v in test.init_test_params().items():
        setattr(tc, k, v)
    tc.sedov_r = getattr(test, "r", 0.0)
    tc.gamma = test.specific_heat_ratio()
    for s in Side:
        uf, vf = test.boundary_condition(s)
        tc.bc_u[int(s)] = uf
        tc.bc_v[int(s)] = vf
    return tc


def solver_desc(params, N=None, N_origin=None, neighbours=None):
    """armon_solver_desc of one block: the whole sub-domain of this process by default."""
    N = params.N if N is None else N
    N_origin = params.N_origin if N_origin is None else N_origin
    neighbours = params.neighbours if neighbours is None else neighbours
    d = backend.armon_solver_desc()
    d.dims = armon_dims(N[0], N[1], params.nghost)
    d.global_nx, d.global_ny = params.global_grid
    d.origin_ix, d.origin_iy = N_origin
    d.domain_size[:] = params.domain_size
    d.origin[:] = params.origin
    d.riemann = riemann_code(params.riemann_scheme)
    d.limiter = limiter_code(params.riemann_limiter)
    d.projection = projection_code(params.projection_scheme)
    d.splitting = splitting_code(params.axis_splitting)
    d.cfl, d.maxtime, d.maxcycle = params.cfl, params.maxtime, params.maxcycle
    d.cst_dt, d.Dt = int(params.cst_dt), params.Dt
    for s in Side:
        d.neighbours[int(s)] = neighbours[s]
    d.math_mode = backend.MATH_MODES[params.math_mode]
    d.march_segment = params.march_segment
    d.kernel_variant = backend.KERNEL_VARIANTS[params.kernel_variant]
    d.cuda_graph = backend.CUDA_GRAPH_MODES[params.cuda_graph]
    fill_test_case(d.tc, params.test)
    return d


class LocalBlock:
    """One LocalTaskBlock (src/blocking/blocks.jl:57-103): geometry, device arrays and the fused solver of a block."""

    def __init__(self, grid, pos, N, N_origin):
        params = grid.params
        self.grid = grid
        self.pos = pos
        self.N = N
        self.N_origin = N_origin
        g = params.nghost
        self.dims = armon_dims(N[0], N[1], g)
        self.shape = (N[1] + 2 * g, N[0] + 2 * g)
        self.cell_count = self.shape[0] * self.shape[1]
        self.device_data = BlockData(grid.device, self.cell_count)
        self._solver = None
        # offset of the block's first real cell inside the sub-domain of this process (0-based)
        self.offset = (N_origin[0] - params.N_origin[0], N_origin[1] - params.N_origin[1])

    @property
    def solver(self):
        if self._solver is None:
            grid, params = self.grid, self.grid.params
            B = params.block_grid
            # sides facing another block of the group carry no rank: the group wires them (armon_group_create)
            nb = {s: params.neighbours[s] for s in Side}
            if self.pos[0] > 0: nb[Side.Left] = -1
            if self.pos[0] < B[0] - 1: nb[Side.Right] = -1
            if self.pos[1] > 0: nb[Side.Bottom] = -1
            if self.pos[1] < B[1] - 1: nb[Side.Top] = -1
            self._desc = solver_desc(params, self.N, self.N_origin, nb)
            s = C.c_void_p()
            check(grid.lib.armon_solver_create(grid.device.ctx, C.byref(self._desc), C.byref(s)), "armon_solver_create")
            self._solver = s
            # the finaliser keeps the context handle alive until the solver is destroyed
            self._solver_finalizer = weakref.finalize(self, backend._destroy_solver, grid.device.handle, s)
            d = self.device_data
            main = (C.c_void_p * 4)(d.rho.ptr, d.u.ptr, d.v.ptr, d.E.ptr)
            work = (C.c_void_p * 4)(d.work_1.ptr, d.work_2.ptr, d.work_3.ptr, d.work_4.ptr)
            pcg = (C.c_void_p * 3)(d.p.ptr, d.c.ptr, d.g.ptr) if params.bind_pcg else (C.c_void_p * 3)(None, None, None)
            check(grid.lib.armon_solver_bind(s, C.byref(main), C.byref(work), C.byref(pcg)), "armon_solver_bind")
        return self._solver

    def destroy(self):
        if self._solver is not None:
            self._solver_finalizer()
            self._solver = None


class BlockGrid:
    """The blocks of this process, resident on one B200: one block covering the whole sub-domain by default
    (`block_size = N .+ 2nghost`), or `params.block_grid` blocks advanced in lock step by a
    device-side block group (include/armon_b200.h, armon_group_*)."""

    def __init__(self, params, device=None):
        self.params = params
        self.device = device or params.backend_options or B200Device(params.device_id)
        params.backend_options = self.device
        self._owns_comm = False
        if params.use_MPI and params.proc_size > 1 and self.device.nranks == 1:
            from .distributed import setup_device_comm
            setup_device_comm(params, self.device)
            self._owns_comm = True
        self.lib = self.device.lib
        nx, ny = params.N
        g = params.nghost
        self.dims = armon_dims(nx, ny, g)
        self.shape = (ny + 2 * g, nx + 2 * g)
        self.cell_count = self.shape[0] * self.shape[1]
        self.blocks = [LocalBlock(self, (bx, by), n, o) for bx, by, n, o in params.block_layout()]
        self.multi = len(self.blocks) > 1
        self.host_data = {}
        self.global_dt = GlobalTimeStep()
        self.state = SolverState(params, self.global_dt)
        self._group = None
        self._fused_dirty = False     # device state lives in the solvers' rotating buffers
        self.reset()

    # -- fused solver objects ----------------------------------------------------------------------------
    @property
    def device_data(self):
        if self.multi:
            solver_error("config", "device_data names the arrays of one block: use grid.blocks[k].device_data")
        return self.blocks[0].device_data

    @property
    def solver(self):
        """The fused solver of a one-block grid (armon_solver_*)."""
        if self.multi:
            solver_error("config", "a grid of several blocks is driven through its group (grid.run, grid.time_state, ...)")
        return self.blocks[0].solver

    @property
    def group(self):
        """The block group of a grid of several blocks (armon_group_*)."""
        if self._group is None:
            B = self.params.block_grid
            handles = (C.c_void_p * len(self.blocks))(*[b.solver for b in self.blocks])
            gptr = C.c_void_p()
            check(self.lib.armon_group_create(self.device.ctx, B[0], B[1], handles, C.byref(gptr)), "armon_group_create")
            self._group = gptr
            self._group_finalizer = weakref.finalize(self, backend._destroy_group, self.device.handle, gptr)
        return self._group

    def _call(self, name, *args):
        """armon_group_<name> for several blocks, armon_solver_<name> for one."""
        if self.multi:
            check(getattr(self.lib, "armon_group_" + name)(self.group, *args), "armon_group_" + name)
        else:
            check(getattr(self.lib, "armon_solver_" + name)(self.solver, *args), "armon_solver_" + name)

    @property
    def _have_solver(self):
        return self._group is not None if self.multi else self.blocks[0]._solver is not None

    def init_fused(self):
        """init_test on the fused path: rho, u, v, E of every block + reset!(global_dt)."""
        self._call("init")
        self._fused_dirty = False

    def run(self, n_cycles):
        """Enqueue `n_cycles` solver cycles (asynchronous)."""
        self._call("run", int(n_cycles))
        self._fused_dirty = True

    def run_time_loop(self):
        self._call("time_loop")
        self._fused_dirty = True

    def elapsed_ms(self):
        ms = C.c_float()
        self._call("elapsed_ms", C.byref(ms))
        return ms.value

    def diagnostics(self, capacity):
        """Per-cycle (cycle, time, dt, mass, energy) lines produced on the device (0 disables)."""
        self._call("diagnostics", int(capacity))

    def read_diagnostics(self, max_lines=4096):
        lines = (backend.armon_cycle_diag * max_lines)()
        n = C.c_int64()
        self._call("read_diagnostics", lines, max_lines, C.byref(n))
        return [(ln.cycle, ln.time, ln.dt, ln.mass, ln.energy) for ln in lines[:n.value]]

    def fused_layout_is_tiled(self):
        """Layout of the state between the sweeps of the fused loop (armon_solver_tiled): 1 band-tiled, 0 row-major /
        transposed pair, -1 not decided yet (before the first cycle).  The arrays this object hands out are canonical
        either way."""
        tiled = C.c_int32(-1)
        check(self.lib.armon_solver_tiled(self.blocks[0].solver, C.byref(tiled)), "armon_solver_tiled")
        return tiled.value

    def strict_kernel_is_chains(self):
        """1 when the strict (bit-exact) sweeps run on the four-chain schedule of the fast kernel
        (armon_solver_strict_chains), 0 for the register-prefetch kernel or another arithmetic mode."""
        chains = C.c_int32(0)
        check(self.lib.armon_solver_strict_chains(self.blocks[0].solver, C.byref(chains)), "armon_solver_strict_chains")
        return chains.value

    def reset(self):                                # reset!(grid, params), src/blocking/block_grid.jl:555-561
        self.global_dt.reset(self.params)
        self.state.reset()
        if self._have_solver:
            self.finalize()
            self._call("reset")

    def finalize(self):
        """Bring the bound arrays back to the canonical layout (+ stale p, c, g) after fused cycles."""
        if self._have_solver and self._fused_dirty:
            self._call("finalize")
            self._fused_dirty = False

    def time_state(self):
        st = backend.armon_time_state()
        self._call("state", C.byref(st))
        return st

    def close(self):
        """Release the solvers and, for a multi-rank grid, the NCCL communicator.  Collective when `use_MPI`: every
        rank must call it at the same point (the communicator teardown waits for the peers)."""
        if self._group is not None:
            self._group_finalizer()
            self._group = None
        for b in self.blocks:
            b.destroy()
        if self._owns_comm:
            self.device.comm_destroy()
            self._owns_comm = False

    # -- host <-> device -----------------------------------------------------------------------------------
    def device_to_host(self, vars=MAIN_VARS):       # device_to_host!, src/blocking/blocks.jl:121-143
        self.finalize()
        for name in vars:
            if all(name in b.device_data.allocated() for b in self.blocks):
                self.host_data[name] = self.host_array(name)      # in the caller's type (params.dtype)
        return self.host_data

    def host_to_device(self, vars=None):
        for name, arr in self.host_data.items():
            if vars is None or name in vars:
                self.set_array(name, arr)

    def host_array(self, name):
        """Full [ny+2g, nx+2g] host copy of one variable (fresh read-back).  With several blocks the frame of ghost
        cells comes from the blocks on the edge of the sub-domain; real cells always from the block that owns them."""
        self.finalize()
        T = self.params.dtype
        if not self.multi:
            return getattr(self.blocks[0].device_data, name).copy_to_host(dtype=T).reshape(self.shape)
        g = self.params.nghost
        out = np.empty(self.shape, dtype=T)
        parts = [(b, getattr(b.device_data, name).copy_to_host(dtype=T).reshape(b.shape)) for b in self.blocks]
        for b, a in parts:      # ghosts included first ...
            out[b.offset[1]:b.offset[1] + b.shape[0], b.offset[0]:b.offset[0] + b.shape[1]] = a
        for b, a in parts:      # ... then every real cell from its owner
            out[g + b.offset[1]:g + b.offset[1] + b.N[1], g + b.offset[0]:g + b.offset[0] + b.N[0]] = a[g:-g, g:-g]
        return out

    def real(self, name):
        g = self.params.nghost
        return self.host_array(name)[g:-g, g:-g]

    def set_array(self, name, values):
        self.finalize()
        values = np.asarray(values)
        values = values.astype(np.float64 if values.dtype != np.float32 else np.float32, copy=False).reshape(self.shape)
        for b in self.blocks:   # every block receives its window of the array, ghosts included
            win = values[b.offset[1]:b.offset[1] + b.shape[0], b.offset[0]:b.offset[0] + b.shape[1]]
            getattr(b.device_data, name).copy_from_host(np.ascontiguousarray(win))

    def fill_ghosts(self, name, value):
        self.finalize()
        for b in self.blocks:
            check(self.lib.armon_fill_ghosts(self.device.ctx, b.dims, getattr(b.device_data, name).ptr, float(value)))
