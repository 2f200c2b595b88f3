"""Shared helpers of the test-suite."""
import numpy as np

import armon_jl_b200 as armon

GOLDEN_TESTS = ("Sod", "Sod_y", "Sod_circ", "Bizarrium", "Sedov")
SAVED_VARS = ("x", "y", "rho", "u", "v", "p")          # saved_vars(), src/blocking/blocks.jl:49
EPS = np.finfo(np.float64).eps


def reference_params(test, **overrides):
    """get_reference_params (test/reference_data/reference_functions.jl:7-19) for the B200 backend."""
    opts = dict(test=test, scheme="GAD", projection="euler_2nd", riemann_limiter="minmod",
                nghost=4, N=(100, 100), cfl=0, maxcycle=1000, maxtime=0, silent=5,
                write_output=False, measure_time=False, use_MPI=False)
    opts.update(overrides)
    return armon.ArmonParameters(**opts)


def count_differences(a, b, atol=1e-13, rtol=4 * EPS):
    """Cells failing the reference's own acceptance (reference_functions.jl:55-58, isapprox)."""
    return int((~np.isclose(a, b, atol=atol, rtol=rtol, equal_nan=False)).sum())


def scaled_max_diff(a, b):
    """max|a-b| / max|b| : the field-scale normalised distance of SURVEY.md section 9."""
    scale = float(np.abs(b).max())
    if scale == 0.0:
        return float(np.abs(a - b).max())
    return float(np.abs(a - b).max() / scale)
