"""Host-side logic that needs neither a GPU nor several processes: the Cartesian decomposition of init_MPI /
init_indexing (src/parameters.jl:408-467,673-697) over the process grids and uneven domains of the reference's own MPI
tests (test/mpi.jl:465-475,551-561; single-process construction like `non_mpi_params`, test/mpi.jl:122-130), and the
output format of src/io.jl:4-43 on a stand-in grid."""
import io as _io

import numpy as np
import pytest

import armon_jl_b200 as armon
from armon_jl_b200 import Side
from armon_jl_b200.io import read_data_from_file, write_blocks_to_file

PROC_GRIDS = [(1, 1), (1, 2), (1, 4), (4, 1), (2, 2), (4, 4), (5, 2), (2, 5), (5, 5)]
DOMAINS = [(100, 100), (107, 113), (20, 20), (37, 241)]
KW = dict(test="Sod", scheme="GAD", projection="euler_2nd", riemann_limiter="minmod", nghost=4, silent=5)


@pytest.mark.parametrize("P", PROC_GRIDS)
@pytest.mark.parametrize("N", DOMAINS)
def test_cartesian_decomposition_tiles_the_domain(P, N):
    world = P[0] * P[1]
    if any(N[d] // P[d] < 4 for d in range(2)):
        with pytest.raises(armon.SolverException) as e:      # src/parameters.jl:684-690
            for rank in range(world):
                armon.ArmonParameters(N=N, use_MPI=True, P=P, rank=rank, proc_size=world, **KW)
        assert e.value.category == "config"
        return
    cover = np.zeros((N[1], N[0]), dtype=np.int32)
    ranks = {}
    for rank in range(world):
        p = armon.ArmonParameters(N=N, use_MPI=True, P=P, rank=rank, proc_size=world, **KW)
        ranks[rank] = p
        (ox, oy), (nx, ny) = p.N_origin, p.N
        cover[oy - 1:oy - 1 + ny, ox - 1:ox - 1 + nx] += 1
        assert p.global_grid == N and p.cart_coords == (rank // P[1], rank % P[1])
        assert p.cell_size() == (1.0 / N[0], 1.0 / N[1])      # global cell size, whatever the rank
    assert (cover == 1).all()
    opposite = {Side.Left: Side.Right, Side.Right: Side.Left, Side.Bottom: Side.Top, Side.Top: Side.Bottom}
    for rank, p in ranks.items():
        for side, nb in p.neighbours.items():
            if nb < 0:      # global edge: the sub-domain touches the domain border on that side
                (ox, oy), (nx, ny) = p.N_origin, p.N
                assert {Side.Left: ox == 1, Side.Right: ox - 1 + nx == N[0], Side.Bottom: oy == 1,
                        Side.Top: oy - 1 + ny == N[1]}[side]
            else:           # neighbour relations are symmetric and the shared face has the same length
                q = ranks[nb]
                assert q.neighbours[opposite[side]] == rank
                axis = 1 if side in (Side.Left, Side.Right) else 0
                assert p.N[axis] == q.N[axis] and p.N_origin[axis] == q.N_origin[axis]


class _StubData:
    def allocated(self):
        return ()


class _StubGrid:
    """What write_blocks_to_file needs from a BlockGrid: host copies of the saved variables."""

    def __init__(self, params, fields):
        self.device_data = _StubData()
        g, (nx, ny) = params.nghost, params.N
        self._full = {}
        for name, real in fields.items():
            full = np.full((ny + 2 * g, nx + 2 * g), np.nan)
            full[g:-g, g:-g] = real
            self._full[name] = full

    def host_array(self, name):
        return self._full[name]


def test_output_format_round_trip_on_a_stub_grid(golden):
    ref = golden("Sod_circ")
    params = armon.ArmonParameters(N=(100, 100), **dict(KW, test="Sod_circ"))
    grid = _StubGrid(params, {v: ref[v] for v in ("rho", "u", "v", "p")})
    buf = _io.StringIO()
    write_blocks_to_file(params, grid, buf)
    lines = buf.getvalue().split("\n")
    assert len([ln for ln in lines if ln.strip()]) == 100 * 100 and lines[100] == ""
    assert all(len(tok) == 24 for tok in lines[0].split(", "))          # "%#24.17e"
    data = read_data_from_file(params, _io.StringIO(buf.getvalue()))
    for v in ("rho", "u", "v", "p"):
        assert np.array_equal(data[v], ref[v]), v                        # 17 significant digits round-trip exactly
    for v in ("x", "y"):                                                 # coordinates recomputed like the init kernel
        assert np.allclose(data[v], ref[v], rtol=0, atol=1e-15), v


def test_compare_data_and_time_step_files(tmp_path):
    """compare_block's isapprox rule and report, the time-step checkpoint file (src/io.jl:88-168), without a GPU."""
    import numpy as np
    import armon_jl_b200 as armon
    from armon_jl_b200.io import compare_data, read_time_step_file, write_time_step_file
    p = armon.ArmonParameters(test="Sod", N=(6, 4), output_dir=str(tmp_path), comparison_tolerance=1e-10, silent=5)
    write_time_step_file(p, 0.0043196268869671031, "dt_000")
    assert (tmp_path / "dt_000").read_text() == " 4.31962688696710308e-03\n"       # "%#24.17e"
    assert read_time_step_file(p, "dt_000") == 0.0043196268869671031
    base = {v: np.arange(24, dtype=np.float64).reshape(4, 6) + 1.0 for v in ("x", "y", "rho", "u", "v", "p")}
    other = {v: a.copy() for v, a in base.items()}
    lines = []
    assert not compare_data(p, base, other, "cell_update", out=lines.append) and not lines
    other["rho"][2, 3] *= 1 + 1e-9          # beyond rtol
    other["u"][0, 0] *= 1 + 1e-12           # within rtol
    other["p"][1, 1] = np.nan
    assert compare_data(p, base, other, "cell_update", out=lines.append)
    assert lines[0] == "At cell_update, in block (1, 1):"
    assert lines[1] == "  1 differences found in rho (ref ≢ current)" and "(  4,  3 |   4,  3)" in lines[2]
    assert any("1 differences found in p" in ln for ln in lines) and not any("found in u" in ln for ln in lines)
