"""Test cases = initial and boundary conditions.  Mirrors src/tests.jl.

Everything here is host-side scalar work; the values are handed to the CUDA library (and, in tests/, to the
CPU oracle) through `armon_test_case` so that both start from bit-identical states.
"""
import math

from .utils import Side, solver_error

FreeFlow, Dirichlet = 0, 1   # src/tests.jl:124


class TestCase:
    __test__ = False          # not a pytest class
    code = -1                 # ARMON_TEST_*
    default_domain_size = (1.0, 1.0)     # src/tests.jl:32-33
    default_domain_origin = (0.0, 0.0)   # src/tests.jl:35-36
    default_CFL = 0.0
    default_max_time = 0.0
    is_conservative = True               # src/tests.jl:48-49
    bizarrium_eos = False
    boundaries = {}                      # Side -> FreeFlow | Dirichlet (src/tests.jl:163-211)

    @property
    def name(self):
        return type(self).__name__

    def specific_heat_ratio(self):       # src/tests.jl:46
        return 7 / 5

    def init_test_params(self):          # src/tests.jl:84-121 -> dict(high_rho, low_rho, high_E, ...)
        raise NotImplementedError

    def boundary_condition(self, side):
        """(u_factor, v_factor) of `side` -- src/tests.jl:150-160."""
        side = Side(side)
        if self.boundaries[side] == FreeFlow:
            return (1, 1)
        if side in (Side.Left, Side.Right):
            return (-1, 1)
        return (1, -1)

    def __repr__(self):
        return self.name


_SOD_PARAMS = dict(high_rho=1.0, low_rho=0.125, high_E=2.5, low_E=2.0,
                   high_u=0.0, low_u=0.0, high_v=0.0, low_v=0.0)


class Sod(TestCase):
    code = 0
    default_CFL = 0.95
    default_max_time = 0.20
    boundaries = {Side.Left: Dirichlet, Side.Right: Dirichlet, Side.Bottom: FreeFlow, Side.Top: FreeFlow}

    def init_test_params(self):
        return dict(_SOD_PARAMS)


class Sod_y(Sod):
    code = 1
    boundaries = {Side.Left: FreeFlow, Side.Right: FreeFlow, Side.Bottom: Dirichlet, Side.Top: Dirichlet}


class Sod_circ(Sod):
    code = 2
    boundaries = {Side.Left: Dirichlet, Side.Right: Dirichlet, Side.Bottom: Dirichlet, Side.Top: Dirichlet}


class Bizarrium(TestCase):
    code = 3
    default_CFL = 0.6
    default_max_time = 80e-6
    is_conservative = False
    bizarrium_eos = True
    boundaries = {Side.Left: Dirichlet, Side.Right: FreeFlow, Side.Bottom: Dirichlet, Side.Top: Dirichlet}

    def init_test_params(self):   # src/tests.jl:97-108
        return dict(high_rho=1.42857142857e+4, low_rho=10000., high_E=4.48657821135e+6,
                    low_E=0.5 * 250 ** 2, high_u=0.0, low_u=250., high_v=0.0, low_v=0.0)


class Sedov(TestCase):
    code = 4
    default_domain_size = (2.0, 2.0)
    default_domain_origin = (-1.0, -1.0)
    default_CFL = 0.7
    default_max_time = 1.0
    boundaries = {Side.Left: FreeFlow, Side.Right: FreeFlow, Side.Bottom: FreeFlow, Side.Top: FreeFlow}

    def __init__(self, r=0.0):
        self.r = float(r)

    def init_test_params(self):   # src/tests.jl:110-121
        return dict(high_rho=1.0, low_rho=1.0,
                    high_E=(1 / 1.033) ** 5 / (math.pi * (self.r * self.r)), low_E=2.5e-14,
                    high_u=0.0, low_u=0.0, high_v=0.0, low_v=0.0)


class DebugIndexes(TestCase):
    """Sets every variable to its index in the global domain (src/tests.jl:217-233)."""
    code = 5
    boundaries = {Side.Left: Dirichlet, Side.Right: Dirichlet, Side.Bottom: Dirichlet, Side.Top: Dirichlet}

    def init_test_params(self):
        return dict(high_rho=0.0, low_rho=0.0, high_E=0.0, low_E=0.0, high_u=0.0, low_u=0.0, high_v=0.0, low_v=0.0)


_TESTS = {"Sod": Sod, "Sod_y": Sod_y, "Sod_circ": Sod_circ, "Bizarrium": Bizarrium, "Sedov": Sedov,
          "DebugIndexes": DebugIndexes}


def test_from_name(name):
    """src/tests.jl:21-28"""
    name = str(name).lstrip(":")
    if name not in _TESTS:
        solver_error("config", f"Unknown test case: '{name}'")
    return _TESTS[name]


test_from_name.__test__ = False


def create_test(dX, test_type):
    """src/tests.jl:13-19: Sedov's radius is hypot(dx, dy)/sqrt(2) of the GLOBAL cell size."""
    if test_type is Sedov:
        return Sedov(math.hypot(*dX) / math.sqrt(2))
    return test_type()
