"""The drop-in boundary without a GPU: libarmon_b200.so loads, exports every symbol include/armon_b200.h declares,
describes its ABI, and refuses to compute when no CUDA device is visible (there is no CPU fallback).  Host-side
mirror: option handling and errors of `ArmonParameters` (src/parameters.jl:349-388), axis splitting, steps ranges."""
import ctypes as C
import os
import re

import pytest

import armon_jl_b200 as armon
from armon_jl_b200 import backend

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "armon_b200.h")


def declared_functions():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = re.findall(r"^\s*(?:int|const char \*)\s*(armon_\w+)\s*\(", text, flags=re.M)
    assert len(names) >= 45, names
    return names


def test_header_symbols_are_exported_and_bound():
    lib = C.CDLL(backend.LIB_PATH) if os.path.exists(backend.LIB_PATH) else None
    assert lib is not None, f"{backend.LIB_PATH} missing: run __graft_entry__.build()"
    bound = set(backend.SIGNATURES) | set(backend.PLAIN)
    for name in declared_functions():
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
        assert name in bound, f"{name} has no ctypes prototype in backend.py"
    for name in bound:
        assert name in declared_functions(), f"{name} bound in backend.py but not declared in the header"


def test_abi_self_description():
    lib = armon.load_library()
    assert lib.armon_b200_abi_version() == 2
    assert lib.armon_flt_size() == 8 and lib.armon_idx_size() == 8      # cf. ext/ArmonKokkos.jl:122-140
    assert C.sizeof(backend.armon_dims) == 24 and C.sizeof(backend.armon_domain) == 32
    assert C.sizeof(backend.armon_time_state) == 48 and C.sizeof(backend.armon_cycle_diag) == 40


def test_no_cpu_fallback_without_a_device():
    if armon.device_count() > 0:
        pytest.skip("a CUDA device is visible")
    with pytest.raises(armon.SolverException) as e:
        armon.B200Device(0)
    assert e.value.category in ("config", "cpp")
    with pytest.raises(armon.SolverException):
        armon.armon(armon.ArmonParameters(test="Sod", N=(16, 16), silent=5))


def test_parameters_reject_what_the_backend_does_not_provide():
    for kw in (dict(use_gpu=False), dict(device="CUDA"), dict(data_type="Float16"), dict(async_cycle=True),
               dict(dt_on_even_cycles=True), dict(nghost=2), dict(cst_dt=True, Dt=0.0), dict(math_mode="sloppy"),
               dict(P=(1, 1, 1))):
        with pytest.raises(armon.SolverException) as e:
            armon.ArmonParameters(test="Sod", N=(16, 16), **kw)
        assert e.value.category == "config", kw
    with pytest.raises(ValueError):                      # leftover options, src/parameters.jl:369-372
        armon.ArmonParameters(test="Sod", N=(16, 16), no_such_option=1)


def test_axis_splitting_and_steps_ranges():
    from armon_jl_b200 import Axis, split_axes
    X, Y = Axis.X, Axis.Y
    assert list(split_axes("Sequential", 3)) == [(X, 1.0), (Y, 1.0)]                       # src/axis_splitting.jl:24-46
    assert list(split_axes("Godunov", 0)) == [(X, 1.0), (Y, 1.0)] and list(split_axes("Godunov", 1)) == [(Y, 1.0), (X, 1.0)]
    assert list(split_axes("Strang", 0)) == [(X, 0.5), (Y, 1.0), (X, 0.5)]
    assert list(split_axes("Strang", 1)) == [(Y, 0.5), (X, 1.0), (Y, 0.5)]
    assert list(split_axes("X_only", 5)) == [(X, 1.0)] and list(split_axes("Y_only", 5)) == [(Y, 1.0)]
    p = armon.ArmonParameters(test="Sod", N=(20, 10), projection="euler_2nd", silent=5)
    sx = p.steps_ranges[0]                                                                  # src/parameters.jl:984-1025
    assert armon.block_domain_range(p.N, sx.fluxes) == (1 - 2, 20 + 3, 1, 10)
    assert armon.block_domain_range(p.N, sx.cell_update) == (1 - 2, 20 + 2, 1, 10)
    assert armon.block_domain_range(p.N, sx.advection) == (1, 21, 1, 10)
    assert armon.block_domain_range(p.N, p.steps_ranges[1].fluxes) == (1, 20, 1 - 2, 10 + 3)


def test_host_time_step_state_machine_matches_oracle():
    """GlobalTimeStep (the per-step path's mirror of update_dt! / next_cycle!, src/solver_state.jl:102-166) driven with
    the oracle's local time steps reproduces the oracle's own dt sequence bit for bit, including the one-cycle lag and
    the 5 % growth cap."""
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from helpers import reference_params
    from oracle import OracleSolver
    from armon_jl_b200 import GlobalTimeStep

    params = reference_params("Sod_circ", N=(48, 40), maxcycle=25)
    a = OracleSolver(params, "strict", nthreads=1)       # runs its own state machine
    b = OracleSolver(params, "strict", nthreads=1)       # stepped by hand with the Python state machine
    gdt = GlobalTimeStep()
    gdt.reset(params)
    for cycle in range(25):
        a.solver_cycle()
        if cycle == 0:
            for axis in (0,):
                b.step_EOS(axis)                              # EOS_init (src/solver.jl:291-295)
        gdt.update_dt(params, b.local_time_step())            # next_time_step (src/reductions.jl:164-199)
        dt = gdt.current_dt
        for axis in (0, 1):                                   # Sequential splitting
            b.sweep(axis, dt)
        gdt.next_cycle(params)
        assert gdt.cycle == a.state.cycle and gdt.time == a.state.time and gdt.current_dt == a.state.current_dt, cycle
    assert np_equal(a.real("rho"), b.real("rho")) and np_equal(a.real("E"), b.real("E"))


def np_equal(x, y):
    import numpy as np
    return bool(np.array_equal(x, y))


def test_missing_library_fails_loudly(tmp_path):
    with pytest.raises(armon.SolverException) as e:
        backend.load_library(str(tmp_path / "libarmon_b200.so"))
    assert e.value.category == "cpp" and "no CPU fallback" in str(e.value)
