"""Numerical scheme tags.  Mirrors the singleton types used for dispatch in the reference:
src/riemann_schemes.jl:1-18, src/limiters.jl, src/projection_schemes.jl:1-12, src/axis_splitting.jl.
Integer codes match include/armon_b200.h.
"""
from .utils import Axis, solver_error


def _strip(name):
    return str(name).lstrip(":")


_RIEMANN = {"Godunov": 0, "GAD": 1}
_LIMITER = {"no_limiter": 0, "minmod": 1, "superbee": 2}
_PROJECTION = {"euler": 0, "euler_2nd": 1}
_SPLITTING = {"Sequential": 0, "Godunov": 1, "SequentialSym": 1, "Strang": 2, "X_only": 3, "Y_only": 4}


def scheme_from_name(name):          # src/riemann_schemes.jl:5-9
    n = _strip(name)
    if n not in _RIEMANN:
        solver_error("config", f"Unknown scheme: '{n}'")
    return n


def limiter_from_name(name):         # src/limiters.jl:10-15
    n = _strip(name)
    if n not in _LIMITER:
        solver_error("config", f"Unknown limiter name: '{n}'")
    return n


def projection_from_name(name):      # src/projection_schemes.jl:5-6
    n = _strip(name)
    if n not in _PROJECTION:
        solver_error("config", f"Unknown scheme: '{n}'")
    return n


def splitting_from_name(name):       # src/axis_splitting.jl:7-16
    n = _strip(name)
    if n not in _SPLITTING:
        solver_error("config", f"Unknown splitting method: '{n}'")
    return "Godunov" if n == "SequentialSym" else n


def riemann_code(n):
    return _RIEMANN[n]


def limiter_code(n):
    return _LIMITER[n]


def projection_code(n):
    return _PROJECTION[n]


def splitting_code(n):
    return _SPLITTING[n]


def stencil_width(name):
    """src/riemann_schemes.jl:17-18 and src/projection_schemes.jl:11-12"""
    return {"Godunov": 1, "GAD": 2, "euler": 1, "euler_2nd": 2}[name]


def uses_limiter(scheme):            # src/riemann_schemes.jl:14-15
    return scheme == "GAD"


def split_axes(splitting, cycle):
    """((axis, dt_factor), ...) of one cycle -- src/axis_splitting.jl:22-46."""
    even = cycle % 2 == 0
    if splitting == "Sequential":
        return ((Axis.X, 1.0), (Axis.Y, 1.0))
    if splitting == "Godunov":
        return ((Axis.X, 1.0), (Axis.Y, 1.0)) if even else ((Axis.Y, 1.0), (Axis.X, 1.0))
    if splitting == "Strang":
        if even:
            return ((Axis.X, 0.5), (Axis.Y, 1.0), (Axis.X, 0.5))
        return ((Axis.Y, 0.5), (Axis.X, 1.0), (Axis.Y, 0.5))
    if splitting == "X_only":
        return ((Axis.X, 1.0),)
    if splitting == "Y_only":
        return ((Axis.Y, 1.0),)
    solver_error("config", f"Unknown splitting method: '{splitting}'")
