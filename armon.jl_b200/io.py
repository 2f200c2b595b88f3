"""Output and verification formats of the reference (src/io.jl).

* `write_blocks_to_file` / `write_sub_domain_file` / `read_data_from_file` (src/io.jl:4-74): one line per real cell, X
  fastest, the `saved_vars()` (x, y, rho, u, v, p) as "%#24.17e" (p = output_precision = 17, width p+7), a blank line
  after each grid row (gnuplot pm3d), per-rank file suffix "_<cx>×<cy>" under MPI.
* the `compare` / `is_ref` step-checkpoint protocol (src/io.jl:77-227): with `compare=true, is_ref=true` every step of
  `solver_cycle` (src/solver.jl:288-320) leaves `<output_file>_<cycle:03d>_<step>_<axis letter>` (the time step for the
  "time_step" checkpoints, the saved variables otherwise); with `is_ref=false` the same files are read back and compared
  (`isapprox`, rtol = comparison_tolerance), the first differing step stops the run and leaves a `_diff` file.  This is
  what lets the reference's own GPU-vs-CPU debugging workflow (test/gpu.jl-style comparisons) point at `device=:B200`.
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


def read_sub_domain_file(params, file_name, vars=SAVED_VARS):
    """read_sub_domain_file!, src/io.jl:77-85 -> dict of [ny, nx] arrays"""
    with open(build_file_path(params, file_name)) as f:
        return read_data_from_file(params, f, vars)


def write_time_step_file(params, dt, file_name):
    """src/io.jl:88-97: the current time step as "%#24.17e" """
    p = params.output_precision
    with open(build_file_path(params, file_name), "w") as f:
        f.write(f"%#{p + 7}.{p}e\n" % dt)


def read_time_step_file(params, file_name):
    """src/io.jl:100-106"""
    with open(build_file_path(params, file_name)) as f:
        return float(f.read().strip())


def compare_data(params, ref, ours, label, vars=SAVED_VARS, out=print):
    """compare_block / compare_data, src/io.jl:111-168: cells where `!isapprox(ref, ours; rtol=comparison_tolerance)`
    (isapprox: |a - b| <= rtol * max(|a|, |b|), NaN never approximately equal).  `ref`, `ours`: dicts of [ny, nx] arrays of
    the real cells.  Prints the reference's report and returns True when something differs."""
    import numpy as np
    different = False
    gx, gy = params.N_origin[0] - 1, params.N_origin[1] - 1
    for var in vars:
        a, b = np.asarray(ref[var]), np.asarray(ours[var])
        with np.errstate(invalid="ignore"):
            ok = np.abs(a - b) <= params.comparison_tolerance * np.maximum(np.abs(a), np.abs(b))
        ok |= (a == b)                       # equal infinities
        bad = np.argwhere(~ok)
        if len(bad) == 0:
            continue
        if not different:
            out(f"At {label}, in block (1, 1):")
        different = True
        line = f"  {len(bad)} differences found in {var}"
        if len(bad) <= 200:
            out(line + " (ref ≢ current)")
            for iy, ix in bad:
                r, o = float(a[iy, ix]), float(b[iy, ix])
                ulp = (r - o) / np.spacing(abs(r)) if np.isfinite(r) and r != 0 else float("inf")
                if abs(ulp) > 1e10:
                    ulp = float("inf")
                out("   - %5d (%3d,%3d | %3d,%3d): %12.5g ≢ %12.5g (%12.5g, ulp: %8g)"
                    % (iy * a.shape[1] + ix + 1, ix + 1, iy + 1, ix + 1 + gx, iy + 1 + gy, r, o, r - o, ulp))
        else:
            out(line)
    return different


def _saved_arrays(params, grid, vars=SAVED_VARS):
    g = params.nghost
    out = {}
    for name in vars:
        if name in ("x", "y") and name not in grid.device_data.allocated():
            out[name] = _coordinate(params, name)
        else:
            out[name] = grid.host_array(name)[g:-g, g:-g]
    return out


def compare_with_file(params, grid, file_name, label, out=print):
    """src/io.jl:171-182 (the MPI `|` reduction of the flag is done by the caller's ranks independently here)"""
    ref = read_sub_domain_file(params, file_name)
    different = compare_data(params, ref, _saved_arrays(params, grid), label, out=out)
    if params.use_MPI and params.proc_size > 1:
        from .distributed import allreduce_max
        different = allreduce_max(1.0 if different else 0.0) > 0.0
    return different


def step_checkpoint(params, state, grid, step_label, out=print):
    """step_checkpoint, src/io.jl:185-227 -> True when the run must stop (a difference was found)."""
    if not params.compare:
        return False
    import math
    params.backend_options.wait()
    cycle = state.global_dt.cycle
    from .utils import Axis
    axis = Axis.X if (cycle == 0 and step_label == "time_step") else state.axis
    name = params.output_file + "_%03d_%s" % (cycle, step_label) + "_" + Axis(axis).name[0]
    if params.is_ref:
        if step_label == "time_step":
            write_time_step_file(params, state.global_dt.current_dt, name)
        else:
            write_sub_domain_file(params, grid, name)
        return False
    if step_label == "time_step":
        ref_dt = read_time_step_file(params, name)
        dt = state.dt
        different = not (abs(ref_dt - dt) <= params.comparison_tolerance * max(abs(ref_dt), abs(dt))) or math.isnan(dt)
        if different:
            out("Time step difference: ref Δt = %.18f, Δt = %.18f, diff = %.18f" % (ref_dt, dt, ref_dt - dt))
    else:
        different = compare_with_file(params, grid, name, step_label, out=out)
    if different:
        write_sub_domain_file(params, grid, name + "_diff")
        out(f"Difference file written to {name}_diff")
    return different
