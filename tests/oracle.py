"""ctypes wrapper of the CPU oracle (oracle/armon_oracle.c).  TEST INFRASTRUCTURE ONLY.

Used by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` legs.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
BUILD_DIR = os.path.join(ORACLE_DIR, "_build")


class orc_domain(C.Structure):
    _fields_ = [("ix0", C.c_int), ("ix1", C.c_int), ("iy0", C.c_int), ("iy1", C.c_int)]


class orc_test_case(C.Structure):
    _fields_ = [("test", C.c_int),
                ("high_rho", C.c_double), ("low_rho", C.c_double), ("high_E", C.c_double), ("low_E", C.c_double),
                ("high_u", C.c_double), ("low_u", C.c_double), ("high_v", C.c_double), ("low_v", C.c_double),
                ("sedov_r", C.c_double), ("gamma", C.c_double), ("eos", C.c_int),
                ("bc_u", C.c_double * 4), ("bc_v", C.c_double * 4)]


class orc_params(C.Structure):
    _fields_ = [("nx", C.c_int), ("ny", C.c_int), ("g", C.c_int),
                ("global_nx", C.c_int), ("global_ny", C.c_int),
                ("origin_ix", C.c_int), ("origin_iy", C.c_int),
                ("domain_size", C.c_double * 2), ("origin", C.c_double * 2),
                ("riemann", C.c_int), ("limiter", C.c_int), ("projection", C.c_int), ("splitting", C.c_int),
                ("cfl", C.c_double), ("maxtime", C.c_double), ("maxcycle", C.c_int),
                ("cst_dt", C.c_int), ("Dt", C.c_double),
                ("has_neighbour", C.c_int * 4),
                ("tc", orc_test_case),
                ("nthreads", C.c_int)]


PD = C.POINTER(C.c_double)


class orc_data(C.Structure):
    _names = ("x", "y", "rho", "u", "v", "E", "p", "c", "g", "us", "ps",
              "work_1", "work_2", "work_3", "work_4", "mask")
    _fields_ = [(n, PD) for n in _names]


class orc_dt_state(C.Structure):
    _fields_ = [("cycle", C.c_int), ("time", C.c_double), ("current_dt", C.c_double), ("next_cycle_dt", C.c_double)]


HALO_FN = C.CFUNCTYPE(None, C.c_void_p, C.c_int)
MIN_FN = C.CFUNCTYPE(C.c_double, C.c_void_p, C.c_double)


class orc_solver(C.Structure):
    _fields_ = [("p", orc_params), ("d", orc_data), ("t", orc_dt_state),
                ("halo_exchange", HALO_FN), ("halo_user", C.c_void_p),
                ("allreduce_min", MIN_FN), ("min_user", C.c_void_p),
                ("error", C.c_int)]


_LIBS = {}


def _cpu_stamp():
    """Identity of the host CPU's instruction set: the `fast` flavour is built -march=native and must be rebuilt on
    the machine that runs it (the built library travels with the repo snapshot)."""
    try:
        with open("/proc/cpuinfo") as f:
            for ln in f:
                if ln.startswith("flags"):
                    import hashlib
                    return hashlib.sha1(ln.encode()).hexdigest()
    except Exception:
        pass
    return "unknown"


def build(force=False):
    """Compile the three oracle flavours with oracle/Makefile (gcc)."""
    targets = [os.path.join(BUILD_DIR, f"liboracle_{f}.so") for f in ("strict", "fma", "fast")]
    src_mtime = max(os.path.getmtime(os.path.join(ORACLE_DIR, f)) for f in ("armon_oracle.c", "armon_oracle.h", "Makefile"))
    stamp_path = os.path.join(BUILD_DIR, "fast.cpustamp")
    stamp = _cpu_stamp()
    try:
        with open(stamp_path) as f:
            same_cpu = f.read().strip() == stamp
    except Exception:
        same_cpu = False
    if not same_cpu and os.path.exists(targets[2]):
        os.remove(targets[2])
    if force or not all(os.path.exists(t) and os.path.getmtime(t) >= src_mtime for t in targets):
        subprocess.run(["make", "-C", ORACLE_DIR, "CC=gcc"], check=True, capture_output=True)
        with open(stamp_path, "w") as f:
            f.write(stamp)
    return targets


def load(flavour="strict"):
    """flavour: 'strict' (parity oracle), 'fma' (contracted), 'fast' (CPU-baseline build)."""
    if flavour in _LIBS:
        return _LIBS[flavour]
    path = os.path.join(BUILD_DIR, f"liboracle_{flavour}.so")
    if not os.path.exists(path) or flavour == "fast":
        build()
    lib = C.CDLL(path)
    lib.orc_solver_create.restype = C.POINTER(orc_solver)
    lib.orc_solver_create.argtypes = [C.POINTER(orc_params)]
    lib.orc_solver_destroy.argtypes = [C.POINTER(orc_solver)]
    lib.orc_solver_init.argtypes = [C.POINTER(orc_solver)]
    lib.orc_solver_cycle.argtypes = [C.POINTER(orc_solver)]
    lib.orc_time_loop.argtypes = [C.POINTER(orc_solver)]
    lib.orc_next_time_step.argtypes = [C.POINTER(orc_solver)]
    lib.orc_local_time_step.argtypes = [C.POINTER(orc_solver)]
    lib.orc_local_time_step.restype = C.c_double
    for fn in ("orc_step_EOS", "orc_step_BC"):
        getattr(lib, fn).argtypes = [C.POINTER(orc_solver), C.c_int]
        getattr(lib, fn).restype = None
    for fn in ("orc_step_fluxes", "orc_step_cell_update", "orc_step_remap", "orc_sweep"):
        getattr(lib, fn).argtypes = [C.POINTER(orc_solver), C.c_int, C.c_double]
        getattr(lib, fn).restype = None
    lib.orc_conservation_vars.argtypes = [C.c_int, C.c_int, C.c_int, PD, PD, C.c_double, PD, PD]
    lib.orc_conservation_vars.restype = None
    lib.orc_num_threads.restype = C.c_int
    _LIBS[flavour] = lib
    return lib


def params_to_orc(params, nthreads=0):
    """Translate an `ArmonParameters` (armon.jl_b200/parameters.py) into the oracle's parameter struct."""
    from armon_jl_b200 import Side
    from armon_jl_b200.schemes import limiter_code, projection_code, riemann_code, splitting_code
    p = orc_params()
    p.nx, p.ny, p.g = params.N[0], params.N[1], params.nghost
    p.global_nx, p.global_ny = params.global_grid
    p.origin_ix, p.origin_iy = params.N_origin
    p.domain_size[:] = params.domain_size
    p.origin[:] = params.origin
    p.riemann = riemann_code(params.riemann_scheme)
    p.limiter = limiter_code(params.riemann_limiter)
    p.projection = projection_code(params.projection_scheme)
    p.splitting = splitting_code(params.axis_splitting)
    p.cfl, p.maxtime, p.maxcycle = params.cfl, params.maxtime, min(params.maxcycle, 2**31 - 1)
    p.cst_dt, p.Dt = int(params.cst_dt), params.Dt
    for s in Side:
        p.has_neighbour[int(s)] = int(params.neighbours[s] >= 0)
    fill_test_case(p.tc, params.test)
    p.nthreads = nthreads
    return p


def fill_test_case(tc, test):
    from armon_jl_b200 import Side
    tc.test = test.code
    for k, v in test.init_test_params().items():
        setattr(tc, k, v)
    tc.sedov_r = getattr(test, "r", 0.0)
    tc.gamma = test.specific_heat_ratio()
    tc.eos = 1 if test.bizarrium_eos else 0
    for s in Side:
        uf, vf = test.boundary_condition(s)
        tc.bc_u[int(s)] = uf
        tc.bc_v[int(s)] = vf


class OracleSolver:
    """The reference CPU path restated: init_test, solver_cycle, time_loop on 16 host arrays."""

    VARS = orc_data._names

    def __init__(self, params, flavour="strict", nthreads=0):
        self.lib = load(flavour)
        self.params = params
        self._p = params_to_orc(params, nthreads)
        self._s = self.lib.orc_solver_create(C.byref(self._p))
        if not self._s:
            raise MemoryError("oracle allocation failed")
        self.nx, self.ny, self.g = params.N[0], params.N[1], params.nghost
        self.shape = (self.ny + 2 * self.g, self.nx + 2 * self.g)
        self._cb = []
        self.lib.orc_solver_init(self._s)

    def close(self):
        if self._s:
            self.lib.orc_solver_destroy(self._s)
            self._s = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def array(self, name):
        """Full array (with ghosts) as a [ny+2g, nx+2g] numpy view of the oracle's memory."""
        ptr = getattr(self._s.contents.d, name)
        return np.ctypeslib.as_array(ptr, shape=self.shape)

    def real(self, name):
        g = self.g
        return self.array(name)[g:g + self.ny, g:g + self.nx]

    @property
    def state(self):
        return self._s.contents.t

    @property
    def error(self):
        return self._s.contents.error

    def set_hooks(self, halo=None, allreduce_min=None):
        if halo is not None:
            cb = HALO_FN(lambda user, axis: halo(axis))
            self._cb.append(cb)
            self._s.contents.halo_exchange = cb
        if allreduce_min is not None:
            cb = MIN_FN(lambda user, x: allreduce_min(x))
            self._cb.append(cb)
            self._s.contents.allreduce_min = cb

    def solver_cycle(self):
        return self.lib.orc_solver_cycle(self._s)

    def time_loop(self):
        err = self.lib.orc_time_loop(self._s)
        t = self.state
        return t.time, t.current_dt, t.cycle, err

    def conservation_vars(self):
        m, e = C.c_double(), C.c_double()
        dX = self.params.cell_size()
        d = self._s.contents.d
        self.lib.orc_conservation_vars(self.nx, self.ny, self.g, d.rho, d.E, dX[0] * dX[1], C.byref(m), C.byref(e))
        return m.value, e.value

    # single steps (for step-by-step comparisons)
    def step_EOS(self, axis):
        self.lib.orc_step_EOS(self._s, int(axis))

    def step_BC(self, axis):
        self.lib.orc_step_BC(self._s, int(axis))

    def step_fluxes(self, axis, dt):
        self.lib.orc_step_fluxes(self._s, int(axis), dt)

    def step_cell_update(self, axis, dt):
        self.lib.orc_step_cell_update(self._s, int(axis), dt)

    def step_remap(self, axis, dt):
        self.lib.orc_step_remap(self._s, int(axis), dt)

    def sweep(self, axis, dt):
        self.lib.orc_sweep(self._s, int(axis), dt)

    def next_time_step(self):
        return self.lib.orc_next_time_step(self._s)

    def local_time_step(self):
        return self.lib.orc_local_time_step(self._s)
