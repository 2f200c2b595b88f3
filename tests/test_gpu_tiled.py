"""The band-tiled layout between the sweeps (sweep_fast_kernel.cuh, 5.; common.cuh tiled_index) is a pure change of
addressing: on every grid where it applies (fast mode, TMA staging, extents multiples of 8) the results must be the same
BITS as with the row-major layouts (ARMON_B200_TILED=0), and within the fast-mode tolerance of the oracle.
"""
import os

import numpy as np
import pytest

import armon_jl_b200 as armon
from helpers import reference_params, scaled_max_diff
from oracle import OracleSolver

pytestmark = pytest.mark.gpu

VARS = ("rho", "u", "v", "E", "p", "c", "g")


def assert_same(a, b, what):
    if not np.array_equal(a, b):
        bad = np.argwhere(a != b)
        raise AssertionError(f"{what}: {len(bad)} cells differ, first at {bad[0]}: {a[tuple(bad[0])]!r} vs "
                             f"{b[tuple(bad[0])]!r}, max rel {scaled_max_diff(a, b):.3e}")


def run(params, tiled):
    """armon(params) with the layout forced; returns (stats, {var: real cells})"""
    old = os.environ.get("ARMON_B200_TILED")
    os.environ["ARMON_B200_TILED"] = "1" if tiled else "0"
    try:
        params.return_data = True
        stats = armon.armon(params)
        grid = stats.data
        assert grid.fused_layout_is_tiled() == (1 if tiled else 0)      # the layout asked for is the layout that ran
        out = {v: grid.real(v).copy() for v in VARS}
        grid.close()
    finally:
        if old is None:
            del os.environ["ARMON_B200_TILED"]
        else:
            os.environ["ARMON_B200_TILED"] = old
    return stats, out


CASES = [
    # test, N, scheme, limiter, projection, splitting, cycles, march_segment
    ("Sod_circ", (64, 64), "GAD", "minmod", "euler_2nd", "Sequential", 12, 0),
    ("Sod_circ", (128, 40), "GAD", "minmod", "euler_2nd", "Sequential", 11, 16),      # 40 columns: a warp of 4 + 28 + 8
    ("Sod_circ", (40, 136), "GAD", "superbee", "euler_2nd", "Godunov", 10, 32),
    ("Sod_circ", (256, 96), "Godunov", "minmod", "euler", "Strang", 9, 64),
    ("Sod_circ", (96, 256), "GAD", "no_limiter", "euler", "X_only", 8, 0),
    ("Sod_circ", (96, 256), "GAD", "minmod", "euler_2nd", "Y_only", 8, 48),
    ("Sod", (200, 8), "GAD", "minmod", "euler_2nd", "Sequential", 20, 0),             # one band of real rows
    ("Sod_y", (8, 200), "GAD", "minmod", "euler_2nd", "Sequential", 20, 0),
    ("Sedov", (120, 120), "GAD", "minmod", "euler_2nd", "Strang", 15, 0),
    ("Bizarrium", (152, 40), "GAD", "superbee", "euler", "Godunov", 11, 0),
    ("Bizarrium", (304, 264), "GAD", "minmod", "euler_2nd", "Sequential", 20, 128),
    ("Sod_circ", (1024, 520), "GAD", "minmod", "euler_2nd", "Sequential", 6, 0),
]


@pytest.mark.parametrize("test,N,scheme,limiter,projection,splitting,cycles,seg", CASES)
def test_tiled_layout_gives_the_same_bits(test, N, scheme, limiter, projection, splitting, cycles, seg):
    kw = dict(N=N, scheme=scheme, riemann_limiter=limiter, projection=projection, axis_splitting=splitting,
              maxcycle=cycles, math_mode="fast", march_segment=seg)
    s_t, t = run(reference_params(test, **kw), True)
    s_r, r = run(reference_params(test, **kw), False)
    assert s_t.cycles == s_r.cycles == cycles
    assert s_t.last_dt == s_r.last_dt and s_t.final_time == s_r.final_time
    for v in VARS:
        assert_same(t[v], r[v], f"{test} {N} {v}: tiled vs row-major")
    orc = OracleSolver(reference_params(test, **{k: v for k, v in kw.items() if k not in ("math_mode", "march_segment")}),
                       "strict", nthreads=2)
    _, dt, ncyc, err = orc.time_loop()
    assert err == 0 and ncyc == cycles and abs(s_t.last_dt - dt) <= 1e-12 * dt
    for v in ("rho", "u", "v", "E", "p"):
        assert scaled_max_diff(t[v], orc.real(v)) <= 1e-12, v
    orc.close()


@pytest.mark.parametrize("blocks", [(2, 1), (1, 2), (2, 2), (4, 2)])
def test_tiled_layout_block_grids(blocks):
    """Ghost bands copied between local blocks are layout-blind: a block grid in the tiled layout = one block, bitwise."""
    kw = dict(N=(256, 128), maxcycle=10, math_mode="fast")
    _, one = run(reference_params("Sod_circ", **kw), True)
    _, many = run(reference_params("Sod_circ", block_grid=blocks, **kw), True)
    _, rows = run(reference_params("Sod_circ", block_grid=blocks, **kw), False)
    for v in ("rho", "u", "v", "E"):
        assert_same(many[v], one[v], f"{v}: {blocks} blocks vs one (tiled)")
        assert_same(many[v], rows[v], f"{v}: {blocks} blocks, tiled vs row-major")


def test_tiled_layout_is_what_runs_and_is_refused_on_ragged_grids(capfd):
    """The library says which layout a group runs (ARMON_B200_VERBOSE): on for 64x64, off for 100x100 (not multiples of 8)."""
    os.environ["ARMON_B200_VERBOSE"] = "1"
    try:
        for n, word in ((64, "on"), (100, "off")):
            s = armon.armon(reference_params("Sod", N=(n, n), maxcycle=2, math_mode="fast", return_data=True))
            s.data.close()
            err = capfd.readouterr().err
            assert f"band-tiled layout: {word}" in err, err
    finally:
        del os.environ["ARMON_B200_VERBOSE"]


def test_tiled_layout_diagnostics_and_restart():
    """Per-cycle conservation sums accumulated by the tiled sweep = the row-major ones; a reset re-enters the tiled layout
    from the canonical one."""
    kw = dict(N=(128, 64), maxcycle=7, math_mode="fast")
    lines = {}
    for tiled in (True, False):
        os.environ["ARMON_B200_TILED"] = "1" if tiled else "0"
        try:
            p = reference_params("Sod_circ", **kw)
            g = armon.BlockGrid(p)
            armon.init_test(p, g)
            g.diagnostics(16)
            g.run(7)
            first = [(d[0], d[3], d[4]) for d in g.read_diagnostics()]
            rho1 = g.real("rho").copy()          # finalize: back to the canonical layout
            armon.init_test(p, g)                # same initial state, reset
            g.run(7)
            again = [(d[0], d[3], d[4]) for d in g.read_diagnostics()]
            assert_same(g.real("rho"), rho1, "second run after reset")
            assert first == again and len(first) == 7
            lines[tiled] = (first, rho1)
            g.close()
        finally:
            del os.environ["ARMON_B200_TILED"]
    # the columns are grouped into warps 4 cells earlier: another summation order, not another sum
    for (c1, m1, e1), (c0, m0, e0) in zip(lines[True][0], lines[False][0]):
        assert c1 == c0 and abs(m1 - m0) <= 1e-14 * abs(m0) and abs(e1 - e0) <= 1e-14 * abs(e0)
    assert_same(lines[True][1], lines[False][1], "rho")


def test_tiled_layout_cycles_past_the_end_and_graph_replay():
    """The copy-through of a finished run (cycles enqueued past maxcycle) and the CUDA-graph replay of cycle pairs in the
    tiled layout: same bits as the plain run, stale p and c included."""
    kw = dict(N=(128, 64), maxcycle=9, math_mode="fast")
    os.environ["ARMON_B200_TILED"] = "1"
    try:
        s0 = armon.armon(reference_params("Sod_circ", cuda_graph="off", return_data=True, **kw))
        s1 = armon.armon(reference_params("Sod_circ", cuda_graph="on", return_data=True, **kw))
        p2 = reference_params("Sod_circ", cuda_graph="off", **kw)
        g2 = armon.BlockGrid(p2)
        armon.init_test(p2, g2)
        g2.run(14)                                  # 5 cycles past the end
        st = g2.time_state()
        assert st.done and st.cycle == 9 and st.time == s0.final_time
        assert s1.cycles == s0.cycles == 9 and s1.last_dt == s0.last_dt
        for v in VARS:
            assert_same(s1.data.real(v), s0.data.real(v), f"{v}: graph replay")
            if v != "g":
                assert_same(g2.real(v), s0.data.real(v), f"{v}: past the end")
        s0.data.close(); s1.data.close(); g2.close()
    finally:
        del os.environ["ARMON_B200_TILED"]


def run_strict(params, kernel):
    """armon(params) in math_mode strict with the sweep kernel forced (ARMON_B200_STRICT = chains | single)."""
    old = os.environ.get("ARMON_B200_STRICT")
    os.environ["ARMON_B200_STRICT"] = kernel
    try:
        params.return_data = True
        stats = armon.armon(params)
        grid = stats.data
        assert grid.strict_kernel_is_chains() == (1 if kernel == "chains" else 0)   # the kernel asked for is the kernel that ran
        out = {v: grid.real(v).copy() for v in VARS}
        grid.close()
    finally:
        if old is None:
            del os.environ["ARMON_B200_STRICT"]
        else:
            os.environ["ARMON_B200_STRICT"] = old
    return stats, out


@pytest.mark.parametrize("test,N,scheme,limiter,projection,splitting,cycles,segment", CASES)
def test_strict_chains_kernel_gives_the_same_bits(test, N, scheme, limiter, projection, splitting, cycles, segment):
    """The two kernels of the bit-exact mode -- the strict arithmetic on the four-chain schedule of the fast kernel
    (sweep_fast_kernel<..., MATH_STRICT>, the default) and the register-prefetch kernel (one chain per step, what rows
    that are not 16-byte aligned still take) -- and the strict CPU oracle agree bit for bit: every test case, scheme, limiter, projection and splitting, forced march segments, extents from
    one band of rows to 1024 x 520 (non-power-of-two cell sizes included: the x / dx division path)."""
    kw = dict(N=N, maxcycle=cycles, scheme=scheme, riemann_limiter=limiter, projection=projection,
              axis_splitting=splitting, march_segment=segment, math_mode="strict")
    s_c, chains = run_strict(reference_params(test, **kw), "chains")
    s_a, unskewed = run_strict(reference_params(test, **kw), "single")
    assert (s_c.cycles, s_c.last_dt, s_c.final_time) == (s_a.cycles, s_a.last_dt, s_a.final_time)
    for v in VARS:
        assert_same(chains[v], unskewed[v], f"{test} {N} {v}: chains vs single")
    okw = {k: v for k, v in kw.items() if k not in ("march_segment", "math_mode")}
    orc = OracleSolver(reference_params(test, **okw), "strict", nthreads=os.cpu_count() or 1)
    _, dt, ncyc, err = orc.time_loop()
    assert err == 0 and (ncyc, dt) == (s_c.cycles, s_c.last_dt)
    for v in ("rho", "u", "v", "E", "p"):
        assert_same(chains[v], orc.real(v), f"{test} {N} {v}: chains vs oracle")
    orc.close()
