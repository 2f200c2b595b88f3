"""`ArmonParameters(**options)` -- host-side mirror of src/parameters.jl.

The Julia host keeps this struct as is (the native backend only adds `device=:B200`); here, where no Julia
toolchain exists, the same keyword names, defaults, staged option consumption and errors are restated in
Python so that the drivers and tests read like the reference's (`ArmonParameters(; test=:Sod, ...)`).
Only what the hot path and its callers need is kept; CPU-only machinery (threading, cache blocking, NUMA,
Scotch workload distribution, profiling callbacks) is accepted for compatibility and ignored.
"""
import os

from .schemes import (limiter_from_name, projection_from_name, scheme_from_name, splitting_from_name,
                      stencil_width)
from .test_cases import TestCase, create_test, test_from_name
from .utils import Axis, Side, solver_error


class StepsRanges:
    """Offsets of each step's iteration domain from the corners of the real domain, for one axis.

    Mirrors `StepsRanges` (src/domain_ranges.jl:96-105) as filled by `compute_steps_ranges`
    (src/parameters.jl:984-1025).  Each entry is ((bl_x, bl_y), (tr_x, tr_y)).
    """

    def __init__(self, axis, ghosts, projection):
        extra_flx = stencil_width(projection)
        extra_up = stencil_width(projection)
        zero = (0, 0)
        self.axis = Axis(axis)
        self.real_domain = (zero, zero)
        self.full_domain = ((-ghosts, -ghosts), (ghosts, ghosts))
        self.EOS = (zero, zero)
        if self.axis == Axis.X:
            fl_bl, fl_tr = (extra_flx, 0), (extra_flx + 1, 0)
            cu_bl, cu_tr = (extra_up, 0), (extra_up, 0)
            ad_bl, ad_tr = (0, 0), (1, 0)
        else:
            fl_bl, fl_tr = (0, extra_flx), (0, extra_flx + 1)
            cu_bl, cu_tr = (0, extra_up), (0, extra_up)
            ad_bl, ad_tr = (0, 0), (0, 1)
        neg = lambda t: (-t[0], -t[1])
        self.fluxes = (neg(fl_bl), fl_tr)
        self.cell_update = (neg(cu_bl), cu_tr)
        self.advection = (neg(ad_bl), ad_tr)
        self.projection = (zero, zero)


def block_domain_range(N, corners):
    """Inclusive (ix0, ix1, iy0, iy1) in 1-based real-cell coordinates of a StepsRanges entry.

    Same cells as `block_domain_range(bsize, bottom_left, top_right)` (src/blocking/blocking.jl:71-85).
    """
    (blx, bly), (trx, try_) = corners
    return (1 + blx, N[0] + trx, 1 + bly, N[1] + try_)


_IGNORED_DEFAULTS = dict(
    # init_device (src/parameters.jl:470-529): CPU-side machinery with no meaning on this backend
    use_threading=False, use_simd=True, use_kokkos=False, use_cache_blocking=True, async_cycle=False,
    use_two_step_reduction=False, workload_distribution="simple", distrib_params=None, numa_aware=False,
    lock_memory=False, busy_wait_limit=100, block_size=None,
    # init_MPI
    reorder_grid=True, global_comm=None, gpu_aware=True,
    # init_profiling (src/parameters.jl:532-574)
    profiling=(), measure_time=False, time_async=True, log_blocks=False, estimated_blk_log_size=0,
)


class ArmonParameters:
    """The parameters of the solver (src/parameters.jl:267-389).

    Backend selection follows `get_device` (src/parameters.jl:392-405): `use_gpu=True, device="B200"` picks this
    backend; any other device is rejected because the library has no CPU or multi-vendor fallback.
    """

    def __init__(self, *, data_type="Float64", N=(10, 10), **options):
        # ArmonParameters{T}: T is the type of the caller's arrays.  The device arrays and all arithmetic are Float64
        # for both; with Float32 the conversion happens in the host <-> device copies (armon_copy_*_f32).
        name = str(data_type).lstrip(":")
        if name in ("Float64", "float64", "<class 'float'>", "<class 'numpy.float64'>"):
            self.data_type = "Float64"
        elif name in ("Float32", "float32", "<class 'numpy.float32'>"):
            self.data_type = "Float32"
        else:
            solver_error("config", f"the B200 backend supports Float64 and Float32 (computed in Float64), got data_type={data_type}")
        self.N = tuple(int(n) for n in N)
        if len(self.N) != 2:
            solver_error("config", f"Expected a 2D domain, got N={N}")
        options = self._get_device(**options)
        options = self._init_scheme(**options)
        options = self._init_test(**options)
        options = self._init_MPI(**options)
        options = self._init_device(**options)
        options = self._init_indexing(**options)
        options = self._init_output(**options)
        options = self._init_backend(**options)
        if options:
            names = " and ".join(f"'{k}'" for k in options)
            raise ValueError(f"{len(options)} unconsumed options:\n{names}")   # src/parameters.jl:369-372

    # -- get_device, src/parameters.jl:392-405 ----------------------------------------------------
    def _get_device(self, device="B200", use_gpu=True, **options):
        dev = str(device).lstrip(":")
        if not use_gpu or dev != "B200":
            solver_error("config", "this package only provides the B200 backend (use_gpu=true, device=:B200); "
                                   f"got use_gpu={use_gpu}, device={dev} -- there is no CPU fallback")
        self.use_gpu = True
        self.device = "B200"
        return options

    # -- init_scheme, src/parameters.jl:577-629 ---------------------------------------------------
    def _init_scheme(self, scheme="GAD", projection="euler_2nd", riemann_limiter="minmod",
                     axis_splitting="Sequential", nghost=4, cst_dt=False, Dt=0.0, dt_on_even_cycles=False,
                     **options):
        self.riemann_scheme = scheme_from_name(scheme)
        self.projection_scheme = projection_from_name(projection)
        self.riemann_limiter = limiter_from_name(riemann_limiter)
        self.axis_splitting = splitting_from_name(axis_splitting)
        min_nghost = stencil_width(self.riemann_scheme) * stencil_width(self.projection_scheme)
        if nghost < min_nghost:
            solver_error("config", "Not enough ghost cells for the riemann solver and projection, "
                                   f"at least {min_nghost} are needed, got {nghost}")
        if cst_dt and Dt == 0:
            solver_error("config", "Dt == 0 with constant step enabled")
        if dt_on_even_cycles:
            solver_error("config", "dt_on_even_cycles is not supported by the B200 backend")
        self.nghost = int(nghost)
        self.cst_dt = bool(cst_dt)
        self.Dt = float(Dt)
        self.dt_on_even_cycles = False
        return options

    # -- init_test, src/parameters.jl:632-670 -----------------------------------------------------
    def _init_test(self, test="Sod", domain_size=None, origin=None, cfl=0.0, maxtime=0.0, maxcycle=500_000,
                   **options):
        if isinstance(test, TestCase):
            test_type, test_obj = type(test), test
        elif isinstance(test, str):
            test_type, test_obj = test_from_name(test), None
        else:
            solver_error("config", f"Expected a TestCase type or a symbol, got: {test}")
        self.domain_size = tuple(float(d) for d in (domain_size or test_type.default_domain_size))
        self.origin = tuple(float(o) for o in (origin or test_type.default_domain_origin))
        if test_obj is None:
            dX = tuple(d / n for d, n in zip(self.domain_size, self.N))
            test_obj = create_test(dX, test_type)
        self.test = test_obj
        self.maxcycle = int(maxcycle)
        self.cfl = float(cfl) if cfl != 0 else test_obj.default_CFL
        self.maxtime = float(maxtime) if maxtime != 0 else test_obj.default_max_time
        return options

    # -- init_MPI, src/parameters.jl:408-467: the Cartesian grid of one-process-per-GPU ranks -----
    def _init_MPI(self, use_MPI=False, P=(1, 1), rank=None, proc_size=None, **options):
        P = tuple(int(p) for p in P)
        if len(P) != len(self.N):
            solver_error("config", f"Mismatched dimensions: expected a grid of {len(self.N)} processes, got: {len(P)}")
        self.use_MPI = bool(use_MPI)
        for k in ("reorder_grid", "global_comm", "gpu_aware"):
            options.pop(k, None)
        if self.use_MPI:
            # ranks come from the launcher (torchrun: RANK / WORLD_SIZE), like MPI.Comm_rank / Comm_size
            self.rank = int(os.environ.get("RANK", 0)) if rank is None else int(rank)
            self.proc_size = int(os.environ.get("WORLD_SIZE", 1)) if proc_size is None else int(proc_size)
            self.proc_dims = P
            if P[0] * P[1] != self.proc_size:
                solver_error("config", f"could not create a {P[0]}x{P[1]} cartesian topology using "
                                       f"{self.proc_size} processes")
            # MPI_Cart_create is row-major: the last dimension varies fastest
            self.cart_coords = (self.rank // P[1], self.rank % P[1])
            cx, cy = self.cart_coords

            def rank_of(x, y):
                return x * P[1] + y if (0 <= x < P[0] and 0 <= y < P[1]) else -1   # MPI.PROC_NULL
            self.neighbours = {Side.Left: rank_of(cx - 1, cy), Side.Right: rank_of(cx + 1, cy),
                               Side.Bottom: rank_of(cx, cy - 1), Side.Top: rank_of(cx, cy + 1)}
        else:
            self.rank, self.proc_size, self.proc_dims, self.cart_coords = 0, 1, (1, 1), (0, 0)
            self.neighbours = {s: -1 for s in Side}
        self.root_rank = 0
        self.is_root = self.rank == self.root_rank
        return options

    # -- init_device, src/parameters.jl:470-529 ---------------------------------------------------
    def _init_device(self, **options):
        for k, default in _IGNORED_DEFAULTS.items():
            setattr(self, k, options.pop(k, default))
        if self.async_cycle:
            solver_error("config", "async_cycle=true is a CPU cache-blocking scheduler; unsupported on the B200 backend")
        return options

    # -- init_indexing, src/parameters.jl:673-697 -------------------------------------------------
    def _init_indexing(self, **options):
        self.global_grid = self.N
        P, C = self.proc_dims, self.cart_coords
        self.N = tuple(self.global_grid[d] // P[d] + (self.global_grid[d] % P[d] if C[d] == P[d] - 1 else 0)
                       for d in range(2))
        if any(P[d] > 1 and self.N[d] < self.nghost for d in range(2)):
            solver_error("config", f"domain {self.global_grid} is too small to be split by {P} processes while "
                                   f"keeping more than {self.nghost} cells along each axis")
        self.N_origin = tuple(C[d] * (self.global_grid[d] // P[d]) + 1 for d in range(2))
        self.compute_steps_ranges()
        return options

    def compute_steps_ranges(self):
        self.steps_ranges = [StepsRanges(axis, self.nghost, self.projection_scheme) for axis in (Axis.X, Axis.Y)]

    # -- init_output (subset), src/parameters.jl:700-735 ------------------------------------------
    def _init_output(self, silent=0, output_dir=".", output_file="output", write_output=False, write_ghosts=False,
                     write_slices=False, output_precision=None, animation_step=0, compare=False, is_ref=False,
                     comparison_tolerance=1e-10, check_result=False, return_data=False, **options):
        self.silent = int(silent)
        self.output_dir, self.output_file = output_dir, output_file
        self.write_output, self.write_ghosts, self.write_slices = bool(write_output), bool(write_ghosts), bool(write_slices)
        # exact decimal output by default (src/parameters.jl:708-710)
        self.output_precision = (17 if self.data_type == "Float64" else 9) if output_precision is None else int(output_precision)
        self.animation_step = int(animation_step)
        self.compare, self.is_ref = bool(compare), bool(is_ref)
        self.comparison_tolerance = float(comparison_tolerance)
        self.check_result = bool(check_result)
        self.return_data = bool(return_data)
        self.initial_mass = 0.0
        self.initial_energy = 0.0
        return options

    # -- init_backend(params, ::B200Device; options...), src/parameters.jl:758-778 ----------------
    def _init_backend(self, math_mode="strict", march_segment=0, fused=True, device_id=None, bind_pcg=True,
                      kernel_variant="auto", cuda_graph="auto", block_grid=(1, 1), **options):
        """Backend-specific options (like `armon_cpp_lib_src`/`use_md_iter` for Kokkos, ext/ArmonKokkos.jl:83-89).

        math_mode     "strict": IEEE order of the reference source, bit-exact against the oracle (branch-free correctly
                      rounded division; column chunks whose operands leave its proven range -- divisors in
                      [2^-120, 2^120], dividends 0 or in [2^-900, 2^900] -- are recomputed with the full IEEE division);
                      "ieee": nvcc's full IEEE division everywhere (slower);
                      "fast": FMA contraction + reciprocal division (the reference's own @fastmath latitude).
        march_segment cells per marching segment along the swept axis (0 = auto).
        fused         True: one marching kernel per sweep (`solver_cycle` overload);
                      False: one kernel per reference kernel (the per-step overloads / `compare` path).
        device_id     CUDA ordinal; default LOCAL_RANK (one process per GPU).
        bind_pcg      keep p, c, g arrays so that the stale `p` the reference saves can be produced.
        kernel_variant "auto" | "single" | "async" | "async2" | "tma" (include/armon_b200.h, ARMON_KERNEL_*).
        cuda_graph    "auto" (grids of <= 512x512 cells on one rank) | "on" | "off": replay captured cycle pairs.
        block_grid    (bx, by) blocks per GPU: the sub-domain of this process is cut like `init_indexing` cuts the
                      global domain (N // B, remainder on the last block) into LocalTaskBlocks that exchange their
                      ghost rows on the device (the BlockGrid of src/blocking/block_grid.jl:46-183 with
                      `block_size` = sub-domain / B); one process only for now.
        """
        if math_mode not in ("strict", "fast", "ieee"):
            solver_error("config", f"unknown math_mode '{math_mode}'")
        self.math_mode = math_mode
        self.march_segment = int(march_segment)
        # the step checkpoints of `compare` need every intermediate array of the reference: per-step path
        self.fused = bool(fused) and not self.compare
        self.bind_pcg = bool(bind_pcg)
        if kernel_variant not in ("auto", "single", "async", "async2", "tma"):
            solver_error("config", f"unknown kernel_variant '{kernel_variant}'")
        self.kernel_variant = kernel_variant
        if cuda_graph in (True, False):
            cuda_graph = "on" if cuda_graph else "off"
        if cuda_graph not in ("auto", "on", "off"):
            solver_error("config", f"unknown cuda_graph mode '{cuda_graph}'")
        self.cuda_graph = cuda_graph
        self.block_grid = tuple(int(b) for b in block_grid)
        if len(self.block_grid) != 2 or min(self.block_grid) < 1:
            solver_error("config", f"block_grid must be two positive integers, got {block_grid}")
        if self.block_grid != (1, 1):
            if self.use_MPI and self.proc_size > 1:
                solver_error("config", "several blocks per GPU are supported on one process only")
            if not self.fused:
                solver_error("config", "the per-step path runs on one block")
            if any(self.N[d] // self.block_grid[d] < self.nghost for d in range(2)):
                solver_error("config", f"sub-domain {self.N} is too small to be cut into {self.block_grid} blocks of "
                                       f"at least {self.nghost} cells along each axis")
        self.device_id = int(os.environ.get("LOCAL_RANK", 0)) if device_id is None else int(device_id)
        self.backend_options = None   # set by BlockGrid / armon(): the library context
        return options

    def block_layout(self):
        """[(bx, by, N_block, N_origin_block)] of the blocks of this process, x fastest (`block_grid`)."""
        B = self.block_grid
        out = []
        for by in range(B[1]):
            for bx in range(B[0]):
                pos = (bx, by)
                n = tuple(self.N[d] // B[d] + (self.N[d] % B[d] if pos[d] == B[d] - 1 else 0) for d in range(2))
                o = tuple(self.N_origin[d] + pos[d] * (self.N[d] // B[d]) for d in range(2))
                out.append((bx, by, n, o))
        return out

    # -- helpers ---------------------------------------------------------------------------------
    @property
    def dtype(self):
        """numpy dtype of the host-side arrays: `data_type(params)`"""
        import numpy as np
        return np.float32 if self.data_type == "Float32" else np.float64

    def cell_size(self):
        """ΔX = domain_size ./ global_grid (src/kernels.jl:184, src/reductions.jl:92)"""
        return tuple(d / n for d, n in zip(self.domain_size, self.global_grid))

    def __repr__(self):
        return (f"ArmonParameters(test={self.test}, N={self.global_grid}, local N={self.N}, scheme={self.riemann_scheme}, "
                f"limiter={self.riemann_limiter}, projection={self.projection_scheme}, splitting={self.axis_splitting}, "
                f"nghost={self.nghost}, cfl={self.cfl}, maxtime={self.maxtime}, maxcycle={self.maxcycle}, "
                f"P={self.proc_dims}, coords={self.cart_coords}, math={self.math_mode})")
