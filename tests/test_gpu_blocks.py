"""Decomposition invariance on ONE GPU: the sub-domains of a Cartesian decomposition run as the blocks of a device-side
block group (several LocalTaskBlocks per GPU, include/armon_b200.h armon_group_*), their ghost rows moved by
device-to-device copies where the multi-rank path uses ncclSend/ncclRecv, the CFL maxima of all blocks meeting in one
device-resident time-step state where the multi-rank path all-reduces them.  Gathered fields must equal the
single-block fields bit for bit, in both arithmetic modes -- the analogue of test/mpi.jl:363-398 (sub-domains against
the global reference) and :551-561 (uneven domains) that runs on the driver's 1-GPU box; the NCCL twin of these tests
is tests/test_z_distributed.py.
"""
import numpy as np
import pytest

import armon_jl_b200 as armon
from helpers import reference_params
from oracle import OracleSolver

pytestmark = pytest.mark.gpu

GRIDS = [(1, 2), (2, 1), (2, 2), (2, 4), (3, 1), (1, 4), (4, 1)]


def run(test, blocks, **kw):
    params = reference_params(test, block_grid=blocks, return_data=True, **kw)
    stats = armon.armon(params)
    return stats, stats.data


def assert_same(a, b, what):
    if not np.array_equal(a, b):
        bad = np.argwhere(a != b)
        raise AssertionError(f"{what}: {len(bad)} cells differ, first at {bad[0]}: {a[tuple(bad[0])]!r} vs {b[tuple(bad[0])]!r}")


@pytest.mark.parametrize("mode", ["strict", "fast"])
@pytest.mark.parametrize("blocks", GRIDS)
def test_blocks_equal_single_block_bitwise(blocks, mode):
    """The reference's process grids (test/mpi.jl:363-398) as block grids, golden 100x100 Sod_circ case to its end."""
    kw = dict(math_mode=mode)
    s0, g0 = run("Sod_circ", (1, 1), **kw)
    s1, g1 = run("Sod_circ", blocks, **kw)
    assert s1.cycles == s0.cycles and s1.last_dt == s0.last_dt and s1.final_time == s0.final_time
    for var in ("rho", "u", "v", "E", "p"):
        assert_same(g1.real(var), g0.real(var), f"{blocks} {mode} {var}")
    g0.close(); g1.close()


@pytest.mark.parametrize("blocks", [(2, 2), (1, 3), (3, 2), (2, 4)])
@pytest.mark.parametrize("N", [(107, 113), (20, 20), (37, 241)])
def test_uneven_domains_bitwise_and_against_oracle(N, blocks):
    """test/mpi.jl:551-561: domains that do not divide evenly (the remainder goes to the last block of each axis,
    src/parameters.jl:678-682), 100 cycles at most; blocks == one block == oracle."""
    if any(n // b < 4 for n, b in zip(N, blocks)):
        with pytest.raises(armon.SolverException):     # "too small to be split", src/parameters.jl:684-690
            reference_params("Sod_circ", N=N, block_grid=blocks, maxcycle=100)
        return
    kw = dict(N=N, maxcycle=100)
    s0, g0 = run("Sod_circ", (1, 1), **kw)
    s1, g1 = run("Sod_circ", blocks, **kw)
    orc = OracleSolver(reference_params("Sod_circ", **kw), "strict", nthreads=1)
    _, dt, cycles, err = orc.time_loop()
    assert err == 0 and s1.cycles == s0.cycles == cycles and s1.last_dt == s0.last_dt == dt
    for var in ("rho", "u", "v", "E", "p"):
        assert_same(g1.real(var), g0.real(var), f"{N} {blocks} {var}")
        assert_same(g1.real(var), orc.real(var), f"{N} {blocks} {var} vs oracle")
    g0.close(); g1.close()


@pytest.mark.parametrize("test,splitting,scheme,projection", [
    ("Sedov", "Strang", "GAD", "euler_2nd"),
    ("Bizarrium", "Godunov", "GAD", "euler"),
    ("Sod", "Sequential", "Godunov", "euler_2nd"),
    ("Sod_y", "Y_only", "GAD", "euler_2nd"),
])
def test_blocks_other_cases_and_splittings(test, splitting, scheme, projection):
    """Boundary conditions differ per side and per case (src/tests.jl:150-211): only the blocks on the edge of the
    domain apply them; alternating and 3-sweep splittings keep the blocks in lock step."""
    kw = dict(N=(120, 88), axis_splitting=splitting, scheme=scheme, projection=projection, maxcycle=17)
    s0, g0 = run(test, (1, 1), **kw)
    s1, g1 = run(test, (3, 2), **kw)
    assert s1.cycles == s0.cycles and s1.last_dt == s0.last_dt
    for var in ("rho", "u", "v", "E", "p"):
        assert_same(g1.real(var), g0.real(var), f"{test} {var}")
    g0.close(); g1.close()


def test_blocks_large_grid_uneven_pitches():
    """Blocks larger than L2 with an odd/even pitch mix (the staged kernels need an even pitch and fall back to the
    register-prefetch kernel otherwise): 2050 x 1537 cells in 2 x 3 blocks, fast mode."""
    kw = dict(N=(2050, 1537), maxcycle=6, math_mode="fast")
    s0, g0 = run("Sod_circ", (1, 1), **kw)
    s1, g1 = run("Sod_circ", (2, 3), **kw)
    assert s1.last_dt == s0.last_dt
    for var in ("rho", "u", "v", "E"):
        assert_same(g1.real(var), g0.real(var), var)
    g0.close(); g1.close()


def test_ghost_poisoning_with_blocks(golden):
    """test/convergence.jl:67-102 on a block grid: internal ghost rows are overwritten by the exchange before use."""
    from helpers import count_differences
    params = reference_params("Sod_circ", block_grid=(2, 2), return_data=True)
    grid = armon.BlockGrid(params)
    armon.init_test(params, grid)
    for var in ("rho", "u", "v", "E", "work_1", "work_2", "work_3", "work_4"):
        grid.fill_ghosts(var, 1e100)
    _, dt, cycles, _, _ = armon.time_loop(params, grid)
    ref = golden("Sod_circ")
    assert cycles == int(ref["cycles"])
    for var in ("rho", "u", "v", "p"):
        assert count_differences(grid.real(var), ref[var]) == 0
    grid.close()


def test_group_rejects_foreign_use():
    params = reference_params("Sod", N=(64, 64), block_grid=(2, 1), maxcycle=3)
    grid = armon.BlockGrid(params)
    armon.init_test(params, grid)
    from armon_jl_b200.backend import check
    with pytest.raises(armon.SolverException):          # a grouped solver may only be driven through its group
        check(grid.lib.armon_solver_run(grid.blocks[0].solver, 1))
    armon.time_loop(params, grid)
    assert grid.time_state().cycle == 3
    grid.close()
