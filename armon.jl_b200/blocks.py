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
from .utils import Side

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


def solver_desc(params):
    d = backend.armon_solver_desc()
    d.dims = armon_dims(params.N[0], params.N[1], params.nghost)
    d.global_nx, d.global_ny = params.global_grid
    d.origin_ix, d.origin_iy = params.N_origin
    d.domain_size[:] = params.domain_size
    d.origin[:] = params.origin
    d.riemann = riemann_code(params.riemann_scheme)
    d.limiter = limiter_code(params.riemann_limiter)
    d.projection = projection_code(params.projection_scheme)
    d.splitting = splitting_code(params.axis_splitting)
    d.cfl, d.maxtime, d.maxcycle = params.cfl, params.maxtime, params.maxcycle
    d.cst_dt, d.Dt = int(params.cst_dt), params.Dt
    for s in Side:
        d.neighbours[int(s)] = params.neighbours[s]
    d.math_mode = backend.MATH_MODES[params.math_mode]
    d.march_segment = params.march_segment
    d.kernel_variant = {"auto": 0, "single": 1, "ws": 2, "tma": 3, "async": 4, "async2": 5}[params.kernel_variant]
    fill_test_case(d.tc, params.test)
    return d


class BlockGrid:
    """One block covering the whole sub-domain of this process, resident on one B200."""

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
        self.device_data = BlockData(self.device, self.cell_count)
        self.host_data = {}
        self.global_dt = GlobalTimeStep()
        self.state = SolverState(params, self.global_dt)
        self._solver = None
        self._fused_dirty = False     # device state lives in the solver's rotating buffers
        self.reset()

    # -- fused solver object -----------------------------------------------------------------------------
    @property
    def solver(self):
        if self._solver is None:
            self._desc = solver_desc(self.params)
            s = C.c_void_p()
            check(self.lib.armon_solver_create(self.device.ctx, C.byref(self._desc), C.byref(s)), "armon_solver_create")
            self._solver = s
            # the finaliser keeps the context handle alive until the solver is destroyed
            self._solver_finalizer = weakref.finalize(self, backend._destroy_solver, self.device.handle, s)
            d = self.device_data
            main = (C.c_void_p * 4)(d.rho.ptr, d.u.ptr, d.v.ptr, d.E.ptr)
            work = (C.c_void_p * 4)(d.work_1.ptr, d.work_2.ptr, d.work_3.ptr, d.work_4.ptr)
            pcg = (C.c_void_p * 3)(d.p.ptr, d.c.ptr, d.g.ptr) if self.params.bind_pcg else (C.c_void_p * 3)(None, None, None)
            check(self.lib.armon_solver_bind(s, C.byref(main), C.byref(work), C.byref(pcg)), "armon_solver_bind")
        return self._solver

    def reset(self):                                # reset!(grid, params), src/blocking/block_grid.jl:555-561
        self.global_dt.reset(self.params)
        self.state.reset()
        if self._solver is not None:
            self.finalize()
            check(self.lib.armon_solver_reset(self._solver), "armon_solver_reset")

    def finalize(self):
        """Bring the bound arrays back to the canonical layout (+ stale p, c, g) after fused cycles."""
        if self._solver is not None and self._fused_dirty:
            check(self.lib.armon_solver_finalize(self._solver), "armon_solver_finalize")
            self._fused_dirty = False

    def time_state(self):
        st = backend.armon_time_state()
        check(self.lib.armon_solver_state(self.solver, C.byref(st)), "armon_solver_state")
        return st

    def close(self):
        """Release the solver and, for a multi-rank grid, the NCCL communicator.  Collective when `use_MPI`: every
        rank must call it at the same point (the communicator teardown waits for the peers)."""
        if self._solver is not None:
            self._solver_finalizer()
            self._solver = None
        if self._owns_comm:
            self.device.comm_destroy()
            self._owns_comm = False

    # -- host <-> device -----------------------------------------------------------------------------------
    def device_to_host(self, vars=MAIN_VARS):       # device_to_host!, src/blocking/blocks.jl:121-143
        self.finalize()
        for name in vars:
            if name in self.device_data.allocated():
                self.host_data[name] = getattr(self.device_data, name).copy_to_host().reshape(self.shape)
        return self.host_data

    def host_to_device(self, vars=None):
        for name, arr in self.host_data.items():
            if vars is None or name in vars:
                getattr(self.device_data, name).copy_from_host(arr)

    def host_array(self, name):
        """Full [ny+2g, nx+2g] host copy of one variable (fresh read-back)."""
        self.finalize()
        return getattr(self.device_data, name).copy_to_host().reshape(self.shape)

    def real(self, name):
        g = self.params.nghost
        return self.host_array(name)[g:-g, g:-g]

    def set_array(self, name, values):
        self.finalize()
        getattr(self.device_data, name).copy_from_host(np.asarray(values, dtype=np.float64))

    def fill_ghosts(self, name, value):
        self.finalize()
        check(self.lib.armon_fill_ghosts(self.device.ctx, self.dims, getattr(self.device_data, name).ptr, float(value)))
