"""ctypes binding of libarmon_b200.so (include/armon_b200.h) -- the Python twin of the Julia `ccall` stub shown in
INTEGRATION.md.  There is no fallback: if the library is missing or no B200 is visible every entry point raises
`SolverException(:cpp | :config)`.
"""
import ctypes as C
import os
import weakref

import numpy as np

from .utils import SolverException, solver_error

_PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG_DIR, "lib", "libarmon_b200.so")
HEADER_PATH = os.path.join(os.path.dirname(_PKG_DIR), "include", "armon_b200.h")

ARMON_OK, ARMON_ERR_INVALID, ARMON_ERR_CUDA, ARMON_ERR_NCCL, ARMON_ERR_TIME, ARMON_ERR_NO_DEVICE, ARMON_ERR_RANGE = range(7)
MATH_MODES = {"strict": 0, "fast": 1, "ieee": 2}
KERNEL_VARIANTS = {"auto": 0, "single": 1, "async": 4, "async2": 5, "tma": 6}      # ARMON_KERNEL_*
CUDA_GRAPH_MODES = {"auto": 0, "on": 1, "off": 2}

PD = C.POINTER(C.c_double)


class armon_dims(C.Structure):
    _fields_ = [("nx", C.c_int64), ("ny", C.c_int64), ("g", C.c_int64)]


class armon_domain(C.Structure):
    _fields_ = [("ix0", C.c_int64), ("ix1", C.c_int64), ("iy0", C.c_int64), ("iy1", C.c_int64)]


class armon_test_case(C.Structure):
    _fields_ = [("test", C.c_int32), ("eos", C.c_int32),
                ("high_rho", C.c_double), ("low_rho", C.c_double), ("high_E", C.c_double), ("low_E", C.c_double),
                ("high_u", C.c_double), ("low_u", C.c_double), ("high_v", C.c_double), ("low_v", C.c_double),
                ("sedov_r", C.c_double), ("gamma", C.c_double),
                ("bc_u", C.c_double * 4), ("bc_v", C.c_double * 4)]


class armon_solver_desc(C.Structure):
    _fields_ = [("dims", armon_dims),
                ("global_nx", C.c_int64), ("global_ny", C.c_int64),
                ("origin_ix", C.c_int64), ("origin_iy", C.c_int64),
                ("domain_size", C.c_double * 2), ("origin", C.c_double * 2),
                ("riemann", C.c_int32), ("limiter", C.c_int32), ("projection", C.c_int32), ("splitting", C.c_int32),
                ("cfl", C.c_double), ("maxtime", C.c_double), ("maxcycle", C.c_int64),
                ("cst_dt", C.c_int32), ("Dt", C.c_double),
                ("neighbours", C.c_int32 * 4),
                ("math_mode", C.c_int32), ("march_segment", C.c_int32), ("kernel_variant", C.c_int32),
                ("cuda_graph", C.c_int32),
                ("tc", armon_test_case)]


class armon_time_state(C.Structure):
    _fields_ = [("cycle", C.c_int64), ("time", C.c_double), ("current_dt", C.c_double),
                ("next_cycle_dt", C.c_double), ("error", C.c_int32), ("done", C.c_int32),
                ("error_cycle", C.c_int64)]


class armon_cycle_diag(C.Structure):
    _fields_ = [("cycle", C.c_int64), ("time", C.c_double), ("dt", C.c_double), ("mass", C.c_double),
                ("energy", C.c_double)]


_VP = C.c_void_p
_DIMS_DOM = [_VP, armon_dims, armon_domain]

#: name -> argtypes of every int-returning entry point declared in include/armon_b200.h
SIGNATURES = {
    "armon_device_count": [C.POINTER(C.c_int)],
    "armon_ctx_create": [C.c_int, C.POINTER(_VP)],
    "armon_ctx_destroy": [_VP],
    "armon_ctx_sync": [_VP],
    "armon_device_memory_info": [_VP, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)],
    "armon_device_name": [_VP, C.c_char_p, C.c_int],
    "armon_ctx_launch_count": [_VP, C.POINTER(C.c_uint64)],
    "armon_alloc": [_VP, C.c_uint64, C.POINTER(_VP)],
    "armon_free": [_VP, _VP],
    "armon_copy_h2d": [_VP, _VP, _VP, C.c_uint64],
    "armon_copy_d2h": [_VP, _VP, _VP, C.c_uint64],
    "armon_copy_d2d": [_VP, _VP, _VP, C.c_uint64],
    "armon_copy_h2d_f32": [_VP, _VP, _VP, C.c_uint64],
    "armon_copy_d2h_f32": [_VP, _VP, _VP, C.c_uint64],
    "armon_fill": [_VP, _VP, C.c_double, C.c_uint64],
    "armon_fill_ghosts": [_VP, armon_dims, _VP, C.c_double],
    "armon_perfect_gas_EOS": _DIMS_DOM + [C.c_double] + [_VP] * 7,
    "armon_bizarrium_EOS": _DIMS_DOM + [_VP] * 7,
    "armon_boundary_conditions": [_VP, armon_dims, C.c_int, C.c_double, C.c_double] + [_VP] * 7,
    "armon_acoustic": _DIMS_DOM + [C.c_int] + [_VP] * 6,
    "armon_acoustic_GAD": _DIMS_DOM + [C.c_int, C.c_double, C.c_double, C.c_int] + [_VP] * 6,
    "armon_cell_update": _DIMS_DOM + [C.c_int, C.c_double, C.c_double] + [_VP] * 5,
    "armon_advection_first_order": _DIMS_DOM + [C.c_int, C.c_double] + [_VP] * 9,
    "armon_advection_second_order": _DIMS_DOM + [C.c_int, C.c_double, C.c_double] + [_VP] * 9,
    "armon_euler_projection": _DIMS_DOM + [C.c_int, C.c_double, C.c_double] + [_VP] * 9,
    "armon_dtCFL": [_VP, armon_dims, _VP, _VP, _VP, C.c_double, C.c_double, PD],
    "armon_conservation_vars": [_VP, armon_dims, _VP, _VP, C.c_double, PD, PD],
    "armon_init_test": [_VP, armon_dims, C.c_int64, C.c_int64, C.c_int64, C.c_int64, PD, PD,
                        C.POINTER(armon_test_case)] + [_VP] * 16,
    "armon_solver_create": [_VP, C.POINTER(armon_solver_desc), C.POINTER(_VP)],
    "armon_solver_destroy": [_VP],
    "armon_solver_bind": [_VP, C.POINTER(_VP * 4), C.POINTER(_VP * 4), C.POINTER(_VP * 3)],
    "armon_solver_init": [_VP],
    "armon_solver_reset": [_VP],
    "armon_solver_run": [_VP, C.c_int64],
    "armon_solver_time_loop": [_VP],
    "armon_solver_state": [_VP, C.POINTER(armon_time_state)],
    "armon_solver_finalize": [_VP],
    "armon_solver_halo_exchange": [_VP, C.c_int],
    "armon_solver_elapsed_ms": [_VP, C.POINTER(C.c_float)],
    "armon_solver_sweep_launches": [_VP, C.POINTER(C.c_uint64)],
    "armon_solver_tiled": [_VP, C.POINTER(C.c_int32)],
    "armon_solver_strict_chains": [_VP, C.POINTER(C.c_int32)],
    "armon_solver_profile": [_VP, C.c_int],
    "armon_solver_sweep_time_ms": [_VP, PD, C.POINTER(C.c_uint64)],
    "armon_solver_diagnostics": [_VP, C.c_int32],
    "armon_solver_read_diagnostics": [_VP, C.POINTER(armon_cycle_diag), C.c_int64, C.POINTER(C.c_int64)],
    "armon_group_create": [_VP, C.c_int32, C.c_int32, C.POINTER(_VP), C.POINTER(_VP)],
    "armon_group_destroy": [_VP],
    "armon_group_init": [_VP],
    "armon_group_reset": [_VP],
    "armon_group_run": [_VP, C.c_int64],
    "armon_group_time_loop": [_VP],
    "armon_group_state": [_VP, C.POINTER(armon_time_state)],
    "armon_group_finalize": [_VP],
    "armon_group_elapsed_ms": [_VP, C.POINTER(C.c_float)],
    "armon_group_diagnostics": [_VP, C.c_int32],
    "armon_group_read_diagnostics": [_VP, C.POINTER(armon_cycle_diag), C.c_int64, C.POINTER(C.c_int64)],
    "armon_selftest_math": [_VP, C.c_uint64, C.c_uint64, C.POINTER(C.c_uint64 * 5)],
    "armon_comm_unique_id": [C.c_char * 128],
    "armon_ctx_comm_init": [_VP, C.c_char * 128, C.c_int, C.c_int],
    "armon_ctx_comm_destroy": [_VP],
}
#: entry points with a non-status return type
PLAIN = {"armon_b200_abi_version": C.c_int, "armon_flt_size": C.c_int, "armon_idx_size": C.c_int,
         "armon_last_error": C.c_char_p}

_lib = None


def _preload_nccl():
    """libarmon_b200.so needs `libnccl.so.2`.  When PyTorch is installed, load ITS bundled NCCL first so that a
    later `import torch` in the same process (torch.distributed is the rank bootstrap) finds the version it was
    built against instead of the older system library.  Without PyTorch the system libnccl.so.2 is used."""
    import importlib.util
    try:
        spec = importlib.util.find_spec("nvidia.nccl")
        if spec is not None and spec.submodule_search_locations:
            for loc in spec.submodule_search_locations:
                cand = os.path.join(loc, "lib", "libnccl.so.2")
                if os.path.exists(cand):
                    C.CDLL(cand, mode=C.RTLD_GLOBAL)
                    return cand
    except Exception:
        pass
    return None


def load_library(path=None):
    """dlopen the C-ABI library and declare its prototypes.  Raises if it is missing (no fallback)."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    path = path or os.environ.get("ARMON_B200_LIB") or LIB_PATH   # ARMON_B200_LIB: kernel-variant experiments
    if not os.path.exists(path):
        solver_error("cpp", f"{path} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                            "(make -C armon.jl_b200/csrc); the B200 backend has no CPU fallback")
    _preload_nccl()
    lib = C.CDLL(path)
    for name, restype in PLAIN.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = restype, []
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = C.c_int, argtypes
    if lib.armon_flt_size() != 8 or lib.armon_idx_size() != 8:   # cf. ext/ArmonKokkos.jl:122-140
        solver_error("config", "eltype / index size mismatch with libarmon_b200.so")
    _lib = lib
    return lib


def check(status, what=""):
    """Status -> exception, like raise_cpp_exception (ext/ArmonKokkos.jl:72-76)."""
    if status == ARMON_OK:
        return
    msg = load_library().armon_last_error().decode(errors="replace")
    category = {ARMON_ERR_INVALID: "config", ARMON_ERR_TIME: "time", ARMON_ERR_NO_DEVICE: "config"}.get(status, "cpp")
    raise SolverException(category, f"{what}: {msg}" if what else msg)


def device_count():
    n = C.c_int(0)
    status = load_library().armon_device_count(C.byref(n))
    return n.value if status == ARMON_OK else 0


class _CtxHandle:
    """Owns the library context.  Device arrays and solvers keep a strong reference to the handle inside their own
    finalisers, so the context is destroyed only after everything allocated on it has been released, whatever the
    order in which the garbage collector finalises a dropped object graph."""

    def __init__(self, lib, ctx):
        self.lib, self.ctx = lib, ctx
        self._finalizer = weakref.finalize(self, lib.armon_ctx_destroy, ctx)


def _free_array(handle, ptr):
    handle.lib.armon_free(handle.ctx, ptr)


def _destroy_solver(handle, solver):
    handle.lib.armon_solver_destroy(solver)   # `handle` is only kept alive until here


def _destroy_group(handle, group):
    handle.lib.armon_group_destroy(group)


class B200Device:
    """`create_device(::Val{:B200})` (src/parameters.jl:738-755): a CUDA context + streams on one B200."""

    def __init__(self, device_id=0):
        self.lib = load_library()
        self._ctx = _VP()
        check(self.lib.armon_ctx_create(int(device_id), C.byref(self._ctx)), "armon_ctx_create")
        self.device_id = int(device_id)
        self.rank, self.nranks = 0, 1
        self.handle = _CtxHandle(self.lib, self._ctx)

    @property
    def ctx(self):
        return self._ctx

    def wait(self):
        """Base.wait(params) (src/parameters.jl:1031-1038)"""
        check(self.lib.armon_ctx_sync(self._ctx), "armon_ctx_sync")

    def memory_info(self):
        """device_memory_info (src/parameters.jl:916-926) -> (total, free)"""
        free, total = C.c_uint64(), C.c_uint64()
        check(self.lib.armon_device_memory_info(self._ctx, C.byref(free), C.byref(total)))
        return total.value, free.value

    def name(self):
        buf = C.create_string_buffer(256)
        check(self.lib.armon_device_name(self._ctx, buf, 256))
        return buf.value.decode()

    def launch_count(self):
        n = C.c_uint64()
        check(self.lib.armon_ctx_launch_count(self._ctx, C.byref(n)))
        return n.value

    def comm_init(self, unique_id, rank, nranks):
        """NCCL communicator from a 128-byte id broadcast by the launcher (stands for MPI.Cart_create)."""
        buf = (C.c_char * 128).from_buffer_copy(bytes(unique_id))
        check(self.lib.armon_ctx_comm_init(self._ctx, buf, int(rank), int(nranks)), "armon_ctx_comm_init")
        self.rank, self.nranks = int(rank), int(nranks)

    def comm_destroy(self):
        """Collective teardown of the communicator: every rank must call it at the same point of the program."""
        if self.nranks > 1:
            check(self.lib.armon_ctx_comm_destroy(self._ctx), "armon_ctx_comm_destroy")
            self.rank, self.nranks = 0, 1

    def unique_id(self):
        buf = (C.c_char * 128)()
        check(self.lib.armon_comm_unique_id(buf), "armon_comm_unique_id")
        return bytes(buf)

    def array(self, n):
        return B200Array(self, n)

    def selftest_math(self, n_samples=1 << 26, seed=1):
        """(division, sqrt, shared-reciprocal, spurious-flag, missed-flag) mismatch counts of the strict-mode
        division / sqrt against nvcc's IEEE instructions; all zero when healthy."""
        out = (C.c_uint64 * 5)()
        check(self.lib.armon_selftest_math(self._ctx, int(seed), int(n_samples), C.byref(out)), "armon_selftest_math")
        return tuple(out)


class B200Array:
    """`device_array_type(::B200Device)`: a 1-D Float64 device array owned by the host object
    (finaliser frees it; the library never frees caller memory)."""

    def __init__(self, device, n):
        self.device = device
        self.n = int(n)
        self._ptr = _VP()
        check(device.lib.armon_alloc(device.ctx, self.n, C.byref(self._ptr)), "armon_alloc")
        self._finalizer = weakref.finalize(self, _free_array, device.handle, self._ptr)

    @property
    def ptr(self):
        return self._ptr

    def __len__(self):
        return self.n

    def copy_from_host(self, host):
        """copyto!(device, host): Float64 arrays as they are, Float32 arrays (`ArmonParameters{Float32}`) widened on
        the device -- the device side is always Float64."""
        host = np.asarray(host)
        f32 = host.dtype == np.float32
        host = np.ascontiguousarray(host, dtype=np.float32 if f32 else np.float64).reshape(-1)
        if host.size != self.n:
            raise ValueError(f"size mismatch: {host.size} != {self.n}")
        fn = self.device.lib.armon_copy_h2d_f32 if f32 else self.device.lib.armon_copy_h2d
        check(fn(self.device.ctx, self._ptr, host.ctypes.data_as(_VP), self.n))

    def copy_to_host(self, out=None, dtype=np.float64):
        """copyto!(host, device); a Float32 destination receives the values rounded to nearest."""
        if out is None:
            out = np.empty(self.n, dtype=dtype)
        assert out.dtype in (np.float64, np.float32) and out.size == self.n and out.flags["C_CONTIGUOUS"]
        fn = self.device.lib.armon_copy_d2h_f32 if out.dtype == np.float32 else self.device.lib.armon_copy_d2h
        check(fn(self.device.ctx, out.ctypes.data_as(_VP), self._ptr, self.n))
        return out

    def fill(self, value):
        check(self.device.lib.armon_fill(self.device.ctx, self._ptr, float(value), self.n))

    def free(self):
        self._finalizer()
