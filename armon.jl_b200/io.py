"""Output format of the reference: `write_blocks_to_file` / `write_sub_domain_file` (src/io.jl:4-27,46-74).

One line per real cell, X fastest, the `saved_vars()` (x, y, rho, u, v, p) as "%#24.17e" (p = output_precision = 17,
width p+7), a blank line after each grid row (gnuplot pm3d), per-rank file suffix "_<cx>×<cy>" under MPI.
"""
import os

from .blocks import SAVED_VARS
from .solver import init_test   # noqa: F401  (kept for API parity with Armon.jl's module layout)


def build_file_path(params, file_name):
    """src/io.jl:46-59"""
    path = os.path.join(params.output_dir, file_name)
    if params.is_root and not os.path.isdir(params.output_dir):
        os.makedirs(params.output_dir, exist_ok=True)
    if params.use_MPI:
        path += "_" + "×".join(str(c) for c in params.cart_coords)
    return path


def write_blocks_to_file(params, grid, file, vars=SAVED_VARS, for_3D=True):
    """src/io.jl:4-27.  x, y come from the init kernel (debug path) or are recomputed exactly as it does."""
    p = params.output_precision
    fmt = ", ".join([f"%#{p + 7}.{p}e"] * len(vars)) + "\n"
    g = params.nghost
    cols = []
    for name in vars:
        if name in ("x", "y") and name not in grid.device_data.allocated():
            cols.append(_coordinate(params, name))
        else:
            cols.append(grid.host_array(name)[g:-g, g:-g])
    ny, nx = cols[0].shape
    for iy in range(ny):
        if iy > 0 and for_3D:
            file.write("\n")
        for ix in range(nx):
            file.write(fmt % tuple(float(c[iy, ix]) for c in cols))


def _coordinate(params, name):
    """(x, y) = gI .* ΔX .+ origin (src/kernels.jl:119-125), strict IEEE, real cells only."""
    import numpy as np
    nx, ny = params.N
    dX = params.cell_size()
    if name == "x":
        gi = np.arange(nx, dtype=np.float64) + (params.N_origin[0] - 1)
        return np.broadcast_to(gi * dX[0] + params.origin[0], (ny, nx))
    gi = np.arange(ny, dtype=np.float64) + (params.N_origin[1] - 1)
    return np.broadcast_to((gi * dX[1] + params.origin[1])[:, None], (ny, nx))


def write_sub_domain_file(params, grid, file_name):
    """src/io.jl:62-74"""
    path = build_file_path(params, file_name)
    with open(path, "w") as f:
        write_blocks_to_file(params, grid, f)
    return path


def read_data_from_file(params, file, vars=SAVED_VARS):
    """src/io.jl:30-43 -> dict of [ny, nx] arrays"""
    import numpy as np
    nx, ny = params.N
    rows = [ln for ln in file if ln.strip()]
    data = np.array([[float(t) for t in ln.split(",")] for ln in rows], dtype=np.float64)
    return {name: data[:, k].reshape(ny, nx) for k, name in enumerate(vars)}
