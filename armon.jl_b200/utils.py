"""Axis / Side enumerations and the solver error convention.

Mirrors src/utils.jl:15-78 (`Axis`, `Side`, `first_side`, `opposite_of`, `axis_of`, ...) and
src/utils.jl:89-117 (`SolverException`, `solver_error`).  Integer codes are 0-based and identical to
include/armon_b200.h (ARMON_AXIS_*, ARMON_SIDE_*).
"""
from enum import IntEnum


class Axis(IntEnum):
    X = 0
    Y = 1


class Side(IntEnum):
    Left = 0
    Right = 1
    Bottom = 2
    Top = 3


def first_sides():
    return (Side.Left, Side.Bottom)


def last_sides():
    return (Side.Right, Side.Top)


def first_side(axis):
    return Side.Left if Axis(axis) == Axis.X else Side.Bottom


def last_side(axis):
    return Side.Right if Axis(axis) == Axis.X else Side.Top


def sides_along(axis):
    return (first_side(axis), last_side(axis))


def axis_of(side):
    return Axis.X if Side(side) in (Side.Left, Side.Right) else Axis.Y


def opposite_of(side):
    return {Side.Left: Side.Right, Side.Right: Side.Left, Side.Bottom: Side.Top, Side.Top: Side.Bottom}[Side(side)]


def next_axis(axis):
    return Axis.Y if Axis(axis) == Axis.X else Axis.X


#: categories accepted by `solver_error` (src/utils.jl:89-100)
ERROR_CATEGORIES = ("config", "cpp", "time", "timeout", "other")


class SolverException(Exception):
    """Thrown when the solver encounters an invalid state (src/utils.jl:102-106)."""

    def __init__(self, category, msg):
        if category not in ERROR_CATEGORIES:
            raise ValueError(f"unknown SolverException category: {category}")
        super().__init__(f"SolverException({category}): {msg}")
        self.category = category
        self.msg = msg


def solver_error(category, msg):
    raise SolverException(category, msg)
