"""GPU parity tests proper: the CUDA path, called through the C ABI, against the CPU oracle on the same inputs,
and against the reference's golden vectors (tests/golden/).

Bar (north_star): bit-exact in strict mode (the operation order is kept); within 1e-12 relative (of field scale) in
fast mode and against the @fastmath-generated golden CSVs.
"""
import numpy as np
import pytest

import armon_jl_b200 as armon
from helpers import GOLDEN_TESTS, SAVED_VARS, count_differences, reference_params, scaled_max_diff
from oracle import OracleSolver

pytestmark = pytest.mark.gpu

EXEMPT = ("Bizarrium", "Sedov")


def run_gpu(params):
    params.return_data = True
    stats = armon.armon(params)
    return stats, stats.data


def assert_same(a, b, what):
    """bit-exact up to the sign of zero"""
    if not np.array_equal(a, b):
        bad = np.argwhere(a != b)
        raise AssertionError(f"{what}: {len(bad)} cells differ, first at {bad[0]}: {a[tuple(bad[0])]!r} vs "
                             f"{b[tuple(bad[0])]!r}, max rel {scaled_max_diff(a, b):.3e}")


# ---- 1. per-step kernels vs oracle, step by step (the reference's compare=true protocol) ----------------
@pytest.mark.parametrize("test,scheme,limiter,projection", [
    ("Sod_circ", "GAD", "minmod", "euler_2nd"),
    ("Sod", "Godunov", "minmod", "euler"),
    ("Bizarrium", "GAD", "superbee", "euler_2nd"),
    ("Sedov", "GAD", "no_limiter", "euler"),
])
def test_per_step_kernels_match_oracle_step_by_step(test, scheme, limiter, projection):
    kw = dict(N=(60, 44), scheme=scheme, riemann_limiter=limiter, projection=projection, maxcycle=3)
    params = reference_params(test, fused=False, **kw)
    grid = armon.BlockGrid(params)
    armon.init_test(params, grid)
    orc = OracleSolver(reference_params(test, **kw), "strict", nthreads=1)
    for v in armon.BLOCK_VARS:
        assert_same(grid.host_array(v), orc.array(v), f"init {v}")
    state, gdt = grid.state, grid.global_dt
    for cycle in range(3):
        if cycle == 0:
            state.update(params, armon.Axis.X, 1.0)
            armon.update_EOS(params, state, grid)
            orc.step_EOS(0)
        armon.next_time_step(params, state, grid)
        orc.next_time_step()
        assert gdt.current_dt == orc.state.current_dt and gdt.next_cycle_dt == orc.state.next_cycle_dt
        for axis, factor in armon.split_axes(state.splitting, gdt.cycle):
            state.update(params, axis, factor)
            dt = orc.state.current_dt * factor
            steps = [(armon.update_EOS, lambda: orc.step_EOS(axis), ("p", "c", "g")),
                     (armon.block_ghost_exchange, lambda: orc.step_BC(axis), armon.COMM_VARS),
                     (armon.numerical_fluxes, lambda: orc.step_fluxes(axis, dt), ("us", "ps")),
                     (armon.cell_update, lambda: orc.step_cell_update(axis, dt), ("rho", "u", "v", "E")),
                     (armon.projection_remap, lambda: orc.step_remap(axis, dt),
                      ("rho", "u", "v", "E", "work_1", "work_2", "work_3", "work_4"))]
            for gpu_step, orc_step, vars_ in steps:
                gpu_step(params, state, grid)
                orc_step()
                for v in vars_:
                    assert_same(grid.host_array(v), orc.array(v), f"cycle {cycle} axis {axis} {gpu_step.__name__} {v}")
        gdt.next_cycle(params)
        orc.state.cycle += 1
        orc.state.time += orc.state.current_dt
        orc.state.current_dt = orc.state.next_cycle_dt
    grid.close()


# ---- 2. golden vectors, both paths ---------------------------------------------------------------------
@pytest.mark.parametrize("fused", [True, False])
@pytest.mark.parametrize("test", GOLDEN_TESTS)
def test_golden_vectors(test, fused, golden):
    ref = golden(test)
    stats, grid = run_gpu(reference_params(test, fused=fused))
    assert stats.cycles == int(ref["cycles"])
    assert np.isclose(stats.last_dt, float(ref["dt"]), atol=1e-13, rtol=1e-13)
    for var in ("rho", "u", "v", "p"):
        got, want = grid.real(var), ref[var]
        assert scaled_max_diff(got, want) <= 1e-12, var
        if test not in EXEMPT:
            assert count_differences(got, want) == 0, var
    grid.close()


# ---- 3. fused strict path == oracle, bit for bit --------------------------------------------------------
@pytest.mark.parametrize("variant", ["single", "async"])
@pytest.mark.parametrize("test", GOLDEN_TESTS)
def test_fused_strict_bit_exact_on_golden_cases(test, variant):
    stats, grid = run_gpu(reference_params(test, kernel_variant=variant))
    orc = OracleSolver(reference_params(test), "strict", nthreads=1)
    _, dt, cycles, err = orc.time_loop()
    assert err == 0 and stats.cycles == cycles
    assert stats.last_dt == dt and stats.final_time == orc.state.time
    for var in ("rho", "u", "v", "E", "p", "c"):
        assert_same(grid.real(var), orc.real(var), f"{test} {var}")
    grid.close()


VARIANTS = [
    # test, N, scheme, limiter, projection, splitting, cycles
    ("Sod_circ", (96, 72), "Godunov", "minmod", "euler", "Sequential", 12),
    ("Sod_circ", (107, 113), "GAD", "superbee", "euler_2nd", "Godunov", 13),
    ("Sod_circ", (131, 37), "GAD", "no_limiter", "euler", "Strang", 9),
    ("Sod", (64, 200), "Godunov", "minmod", "euler_2nd", "X_only", 10),
    ("Sod_y", (200, 64), "GAD", "minmod", "euler_2nd", "Y_only", 10),
    ("Sedov", (129, 129), "GAD", "minmod", "euler_2nd", "Strang", 15),
    ("Bizarrium", (150, 40), "GAD", "superbee", "euler", "Godunov", 11),
    ("Sod_circ", (300, 260), "GAD", "minmod", "euler_2nd", "Sequential", 20),
]


@pytest.mark.parametrize("variant", ["single", "async"])
@pytest.mark.parametrize("test,N,scheme,limiter,projection,splitting,cycles", VARIANTS)
def test_fused_strict_bit_exact_variants(test, N, scheme, limiter, projection, splitting, cycles, variant):
    kw = dict(N=N, scheme=scheme, riemann_limiter=limiter, projection=projection, axis_splitting=splitting,
              maxcycle=cycles)
    stats, grid = run_gpu(reference_params(test, kernel_variant=variant, **kw))
    orc = OracleSolver(reference_params(test, **kw), "strict", nthreads=1)
    _, dt, ncyc, err = orc.time_loop()
    assert err == 0 and stats.cycles == ncyc == cycles
    assert stats.last_dt == dt and stats.final_time == orc.state.time
    for var in ("rho", "u", "v", "E", "p"):
        assert_same(grid.real(var), orc.real(var), f"{test} {var}")
    grid.close()


@pytest.mark.parametrize("variant,mode", [("single", "strict"), ("async", "strict"), ("tma", "fast"), ("async2", "fast")])
@pytest.mark.parametrize("seg", [8, 16, 40, 1000])
def test_march_segment_does_not_change_results(seg, variant, mode):
    """Also in fast mode: the arithmetic of a cell does not depend on where the march segments are cut."""
    kw = dict(N=(90, 75), maxcycle=8, math_mode=mode)
    _, g0 = run_gpu(reference_params("Sod_circ", kernel_variant=variant, **kw))
    _, g1 = run_gpu(reference_params("Sod_circ", march_segment=seg, kernel_variant=variant, **kw))
    for var in ("rho", "u", "v", "E"):
        assert_same(g1.real(var), g0.real(var), var)
    g0.close(); g1.close()


@pytest.mark.parametrize("N", [(4, 4), (5, 33), (8, 12), (33, 6), (130, 4)])
def test_tiny_and_ragged_grids_bit_exact(N):
    """Edge cases of the marching kernels: sub-domains as small as the ghost width (src/parameters.jl:684-690),
    fewer cells than a warp / a staging chunk / the prefetch lead, odd and even pitches."""
    kw = dict(N=N, maxcycle=7)
    stats, grid = run_gpu(reference_params("Sod_circ", **kw))
    orc = OracleSolver(reference_params("Sod_circ", **kw), "strict", nthreads=1)
    _, dt, ncyc, err = orc.time_loop()
    assert err == 0 and stats.cycles == ncyc and stats.last_dt == dt   # tiny grids reach maxtime before 7 cycles
    for var in ("rho", "u", "v", "E", "p"):
        assert_same(grid.real(var), orc.real(var), f"{N} {var}")
    grid.close()


def test_fast_mode_large_grid_tracks_strict_mode():
    """Size-independent property at a grid larger than L2: the fast arithmetic mode stays within 1e-12 (of field
    scale) of the bit-exact mode, and conserves mass and energy like it."""
    kw = dict(N=(2048, 1536), maxcycle=25)
    ps, pf = reference_params("Sod_circ", **kw), reference_params("Sod_circ", math_mode="fast", **kw)
    gs, gf = armon.BlockGrid(ps), armon.BlockGrid(pf)
    armon.init_test(ps, gs); armon.init_test(pf, gf)
    m0, e0 = armon.conservation_vars(pf, gf)
    armon.time_loop(ps, gs); armon.time_loop(pf, gf)
    m1, e1 = armon.conservation_vars(pf, gf)
    assert abs(m0 - m1) <= 1e-11 and abs(e0 - e1) <= 1e-11
    assert gs.global_dt.cycle == gf.global_dt.cycle == 25
    assert abs(gs.global_dt.current_dt - gf.global_dt.current_dt) <= 1e-12 * gs.global_dt.current_dt
    for var in ("rho", "u", "v", "E"):
        assert scaled_max_diff(gf.real(var), gs.real(var)) <= 1e-12, var
    gs.close(); gf.close()


def test_cst_dt():
    kw = dict(N=(80, 80), cst_dt=True, Dt=1e-3, maxcycle=10)
    stats, grid = run_gpu(reference_params("Sod", **kw))
    orc = OracleSolver(reference_params("Sod", **kw), "strict", nthreads=1)
    orc.time_loop()
    assert stats.cycles == 10 and stats.final_time == orc.state.time
    for var in ("rho", "u", "v", "E"):
        assert_same(grid.real(var), orc.real(var), var)
    grid.close()


# ---- 4. fast arithmetic mode: 1e-12 of field scale -------------------------------------------------------
@pytest.mark.parametrize("variant", ["single", "async2", "tma"])
@pytest.mark.parametrize("test", GOLDEN_TESTS)
def test_fused_fast_mode_within_tolerance(test, variant, golden):
    ref = golden(test)
    stats, grid = run_gpu(reference_params(test, math_mode="fast", kernel_variant=variant))
    orc = OracleSolver(reference_params(test), "strict", nthreads=1)
    orc.time_loop()
    assert stats.cycles == int(ref["cycles"]) == orc.state.cycle
    assert abs(stats.last_dt - orc.state.current_dt) <= 1e-12 * orc.state.current_dt
    for var in ("rho", "u", "v", "E", "p"):
        assert scaled_max_diff(grid.real(var), orc.real(var)) <= 1e-12, var    # tolerance of north_star
    for var in ("rho", "u", "v", "p"):
        assert scaled_max_diff(grid.real(var), ref[var]) <= 1e-12, var
        if test not in EXEMPT:     # the reference's own acceptance test (atol=1e-13, rtol=4eps, 0 differing cells)
            assert count_differences(grid.real(var), ref[var]) == 0, var
    grid.close()


FAST_VARIANTS = [
    # test, N, scheme, limiter, projection, splitting, cycles
    ("Sod_circ", (96, 72), "Godunov", "minmod", "euler", "Sequential", 12),
    ("Sod_circ", (107, 113), "GAD", "superbee", "euler_2nd", "Godunov", 13),
    ("Sod_circ", (131, 37), "GAD", "no_limiter", "euler", "Strang", 9),
    ("Sedov", (129, 129), "GAD", "minmod", "euler_2nd", "Strang", 15),
    ("Bizarrium", (150, 40), "GAD", "superbee", "euler", "Godunov", 11),
    ("Bizarrium", (300, 260), "GAD", "minmod", "euler_2nd", "Sequential", 20),
]


@pytest.mark.parametrize("test,N,scheme,limiter,projection,splitting,cycles", FAST_VARIANTS)
def test_fast_kernel_schemes_and_staging_variants(test, N, scheme, limiter, projection, splitting, cycles):
    """Every scheme combination of the fast kernel: within 1e-12 (of field scale) of the oracle, and the three staging
    variants (TMA, 16-byte and 8-byte cp.async: odd pitches take the last one) give the same bits -- the arithmetic of
    the fast mode is explicit, not left to the compiler's contraction."""
    kw = dict(N=N, scheme=scheme, riemann_limiter=limiter, projection=projection, axis_splitting=splitting,
              maxcycle=cycles, math_mode="fast")
    orc = OracleSolver(reference_params(test, **{k: v for k, v in kw.items() if k != "math_mode"}), "strict", nthreads=1)
    _, dt, ncyc, err = orc.time_loop()
    s_t, g_t = run_gpu(reference_params(test, kernel_variant="tma", **kw))
    s_a, g_a = run_gpu(reference_params(test, kernel_variant="async2", **kw))
    assert err == 0 and s_t.cycles == s_a.cycles == ncyc == cycles
    assert s_t.last_dt == s_a.last_dt and abs(s_t.last_dt - dt) <= 1e-12 * dt
    for var in ("rho", "u", "v", "E", "p"):
        assert_same(g_t.real(var), g_a.real(var), f"{test} {var}: tma vs cp.async staging")
        assert scaled_max_diff(g_t.real(var), orc.real(var)) <= 1e-12, var
    g_t.close(); g_a.close()


# ---- 5. reference property tests -------------------------------------------------------------------------
def test_ghost_poisoning(golden):
    """test/convergence.jl:67-102"""
    test = "Sod_circ"
    params = reference_params(test, return_data=True)
    grid = armon.BlockGrid(params)
    armon.init_test(params, grid)
    for var in ("rho", "u", "v", "E", "p", "c", "g", "work_1", "work_2", "work_3", "work_4"):
        grid.fill_ghosts(var, 1e100)
    _, dt, cycles, _, _ = armon.time_loop(params, grid)
    ref = golden(test)
    assert cycles == int(ref["cycles"])
    for var in ("rho", "u", "v", "p"):
        assert count_differences(grid.real(var), ref[var]) == 0
    grid.close()


@pytest.mark.parametrize("mode", ["strict", "fast"])
@pytest.mark.parametrize("test", ["Sod", "Sod_y", "Sod_circ"])
def test_conservation(test, mode):
    """test/conservation.jl:1-15 as written: 10 000 cycles (maxtime = 10 000 is never reached), mass and energy
    constant to 1e-12 absolute.  The initial sums are also checked against the oracle's conservation_vars."""
    params = reference_params(test, maxcycle=10000, maxtime=10000, math_mode=mode)
    grid = armon.BlockGrid(params)
    armon.init_test(params, grid)
    m0, e0 = armon.conservation_vars(params, grid)
    orc = OracleSolver(reference_params(test), "strict", nthreads=1)
    om, oe = orc.conservation_vars()
    assert abs(m0 - om) <= 1e-13 * abs(om) and abs(e0 - oe) <= 1e-13 * abs(oe)
    _, _, cycles, _, _ = armon.time_loop(params, grid)
    assert cycles == 10000
    m1, e1 = armon.conservation_vars(params, grid)
    assert abs(m0 - m1) <= 1e-12 and abs(e0 - e1) <= 1e-12
    grid.close()


@pytest.mark.parametrize("test,N", [("Sod_circ", (300, 260)), ("Bizarrium", (150, 40)), ("Sedov", (129, 129))])
def test_conservation_vars_match_oracle(test, N):
    """armon_conservation_vars (fixed-tree device reduction) against the oracle's sequential sums of the same
    evolved state (src/reductions.jl:202-259): the fields are bit-equal, only the summation order differs."""
    kw = dict(N=N, maxcycle=12)
    params = reference_params(test, **kw)
    grid = armon.BlockGrid(params)
    armon.init_test(params, grid)
    armon.time_loop(params, grid)
    m, e = armon.conservation_vars(params, grid)
    orc = OracleSolver(reference_params(test, **kw), "strict", nthreads=1)
    orc.time_loop()
    om, oe = orc.conservation_vars()
    # Sedov's energy spans 20 orders of magnitude: the order of the sum shows at 2e-13
    tol = 1e-12 if test == "Sedov" else 1e-13
    assert abs(m - om) <= tol * abs(om) and abs(e - oe) <= tol * abs(oe)
    grid.close()


@pytest.mark.parametrize("N,blocks,mode", [((96, 80), (1, 1), "strict"), ((96, 80), (2, 3), "strict"),
                                           ((96, 80), (1, 1), "fast"), ((96, 80), (2, 3), "fast"),
                                           ((97, 81), (1, 1), "fast"), ((130, 75), (3, 1), "fast"),
                                           ((128, 96), (2, 2), "fast")])   # 64x48 blocks: tiled layout + fused sums
def test_per_cycle_diagnostics_ring(N, blocks, mode, capsys):
    """The `silent <= 1` log (src/solver.jl:359-371) produced on the device: one line per cycle, same cycle / time / dt
    as the time-step state, mass and energy equal to conservation_vars of that cycle's state.  In fast mode with even
    pitches the sums are accumulated by the last sweep itself (one partial per warp), otherwise -- strict mode, odd
    pitches -- by a reduction over the state that sweep left; both against the oracle's sequential sums."""
    kw = dict(N=N, maxcycle=9)
    params = reference_params("Sod_circ", silent=1, block_grid=blocks, math_mode=mode, **kw)
    grid = armon.BlockGrid(params)
    armon.init_test(params, grid)
    params.initial_mass, params.initial_energy = armon.conservation_vars(params, grid)
    armon.time_loop(params, grid)
    out = capsys.readouterr().out
    assert out.count("Cycle ") == 9 and "|ΔM| =" in out
    log = grid.cycle_log
    assert [ln[0] for ln in log] == list(range(1, 10))
    orc = OracleSolver(reference_params("Sod_circ", **kw), "strict", nthreads=1)
    for cycle, t, dt, mass, energy in log:
        orc.solver_cycle()                      # includes next_cycle!
        st = orc.state
        if mode == "strict":
            assert (cycle, t, dt) == (st.cycle, st.time, st.current_dt)
        else:
            assert cycle == st.cycle and abs(t - st.time) <= 1e-12 * st.time and abs(dt - st.current_dt) <= 1e-12 * dt
        om, oe = orc.conservation_vars()
        assert abs(mass - om) <= 1e-13 * abs(om) and abs(energy - oe) <= 1e-13 * abs(oe), cycle
    grid.close()


@pytest.mark.parametrize("test,splitting", [("Sod_circ", "Sequential"), ("Sedov", "Strang"), ("Bizarrium", "Godunov"),
                                            ("Sod", "X_only")])
def test_cuda_graph_replay_equals_plain_launches(test, splitting):
    """Cycle pairs replayed from a CUDA graph (launch-bound grids) give the bits of the plain launch sequence."""
    kw = dict(N=(100, 100), axis_splitting=splitting, maxcycle=31)
    s0, g0 = run_gpu(reference_params(test, cuda_graph="off", **kw))
    s1, g1 = run_gpu(reference_params(test, cuda_graph="on", **kw))
    assert s0.cycles == s1.cycles and s0.last_dt == s1.last_dt and s0.final_time == s1.final_time
    for var in ("rho", "u", "v", "E", "p"):
        assert_same(g1.real(var), g0.real(var), var)
    g0.close(); g1.close()


def test_cycles_enqueued_past_the_end_keep_the_stale_pressure():
    """armon_solver_run beyond maxcycle: the extra cycles are no-ops, and the p the reference would hold (EOS at the
    start of the last real sweep) survives them."""
    kw = dict(N=(64, 48), maxcycle=6)
    s0, g0 = run_gpu(reference_params("Sod_circ", **kw))
    p1 = reference_params("Sod_circ", **kw)
    g1 = armon.BlockGrid(p1)
    armon.init_test(p1, g1)
    g1.run(11)                                  # 5 cycles past the end
    st = g1.time_state()
    assert st.done and st.cycle == 6 and st.time == s0.final_time
    for var in ("rho", "u", "v", "E", "p", "c"):
        assert_same(g1.real(var), g0.real(var), var)
    g0.close(); g1.close()


@pytest.mark.parametrize("test", ["Sod", "Sod_y", "Bizarrium"])
def test_symmetry(test):
    """test/convergence.jl:31-64"""
    _, grid = run_gpu(reference_params(test))
    for var in ("rho", "u", "v", "E", "p"):
        a = grid.real(var)
        if test == "Sod_y":
            assert np.array_equal(a, np.repeat(a[:, :1], a.shape[1], axis=1)), var
        else:
            assert np.array_equal(a, np.repeat(a[:1, :], a.shape[0], axis=0)), var
    grid.close()


def test_invalid_time_step_raises():
    params = reference_params("Sod", N=(32, 32), maxcycle=5)
    grid = armon.BlockGrid(params)
    armon.init_test(params, grid)
    grid.set_array("rho", np.full(grid.shape, np.nan))
    with pytest.raises(armon.SolverException) as exc:
        armon.time_loop(params, grid)
    assert exc.value.category == "time"
    grid.close()


def test_large_grid_properties():
    """BASELINE-size-independent properties on a grid larger than L2: symmetry of Sod along Y, conservation, and
    bit-equality of the canonical and transposed code paths (Sod on NxM vs Sod_y on MxN)."""
    kw = dict(N=(2048, 1024), maxcycle=6)
    pa = reference_params("Sod", **kw)
    ga = armon.BlockGrid(pa)
    armon.init_test(pa, ga)
    m0, e0 = armon.conservation_vars(pa, ga)
    armon.time_loop(pa, ga)
    m1, e1 = armon.conservation_vars(pa, ga)
    assert abs(m0 - m1) <= 1e-11 and abs(e0 - e1) <= 1e-11
    rho = ga.real("rho")
    assert np.array_equal(rho, np.repeat(rho[:1, :], rho.shape[0], axis=0))
    ga.close()


# ---- 6. arithmetic building blocks ------------------------------------------------------------------------
def test_branch_free_division_and_sqrt_match_ieee():
    """The strict mode's division / sqrt return the bits of nvcc's div.rn.f64 / sqrt.rn.f64 (2^27 samples)."""
    dev = armon.B200Device(0)
    assert dev.selftest_math(n_samples=1 << 27, seed=12345) == (0, 0, 0, 0, 0)


@pytest.mark.parametrize("test", ["Sod_circ", "Bizarrium"])
def test_ieee_mode_equals_strict_mode(test):
    kw = dict(N=(120, 90), maxcycle=10)
    _, g0 = run_gpu(reference_params(test, math_mode="strict", **kw))
    _, g1 = run_gpu(reference_params(test, math_mode="ieee", **kw))
    for var in ("rho", "u", "v", "E", "p"):
        assert_same(g1.real(var), g0.real(var), var)
    g0.close(); g1.close()


def test_strict_mode_handles_tiny_operands_like_ieee():
    """Operands below 2^-900 are outside the proven range of the branch-free division: the affected threads must
    recompute with the full IEEE division, so strict == ieee == oracle even there."""
    kw = dict(N=(48, 40), maxcycle=4)
    scale = 1e-290

    def run(mode, variant="single"):
        params = reference_params("Sod_circ", math_mode=mode, kernel_variant=variant, **kw)
        grid = armon.BlockGrid(params)
        armon.init_test(params, grid)
        grid.set_array("rho", grid.host_array("rho") * scale)     # tiny densities: impedances ~1e-290 as divisors
        armon.time_loop(params, grid)
        return grid

    g_strict, g_async, g_ieee = run("strict"), run("strict", "async"), run("ieee")
    for var in ("rho", "u", "v", "E"):
        assert_same(g_strict.real(var), g_ieee.real(var), var)
        assert_same(g_async.real(var), g_ieee.real(var), var + " (four-chain strict kernel + fix-up kernel)")
    assert g_strict.time_state().current_dt == g_ieee.time_state().current_dt == g_async.time_state().current_dt
    g_strict.close(); g_async.close(); g_ieee.close()


# ---- 7. output format (src/io.jl:4-43): the file the reference's own comparison tooling reads -------------------
def test_output_file_round_trip(tmp_path, golden):
    from armon_jl_b200.io import read_data_from_file
    test = "Sod"
    params = reference_params(test, write_output=True, output_dir=str(tmp_path), output_file="out", return_data=True)
    stats = armon.armon(params)
    path = tmp_path / "out"
    text = path.read_text()
    lines = text.split("\n")
    assert len([ln for ln in lines if ln.strip()]) == 100 * 100
    assert lines[100] == "" and lines[99] != ""                       # blank line after each grid row (gnuplot pm3d)
    first = lines[0].split(", ")
    assert len(first) == 6 and all(len(tok) == 24 for tok in first)   # "%#24.17e" x 6 saved_vars
    with open(path) as f:
        data = read_data_from_file(params, f)
    ref = golden(test)
    for var in ("rho", "u", "v", "p"):
        assert np.array_equal(data[var], stats.data.real(var)), var   # 17 significant digits round-trip exactly
        assert count_differences(data[var], ref[var]) == 0, var
    for var in ("x", "y"):
        assert np.allclose(data[var], ref[var], rtol=0, atol=1e-15), var
    stats.data.close()


def test_compare_is_ref_checkpoint_protocol(tmp_path, capsys):
    """`compare=true` step checkpoints (src/io.jl:185-227): a reference run leaves one file per step, an identical run
    compares clean against them, a run with another limiter stops at the first checkpoint whose saved variables differ
    and leaves the `_diff` file."""
    from armon_jl_b200.io import step_checkpoint, write_sub_domain_file
    kw = dict(N=(40, 30), maxcycle=3, output_dir=str(tmp_path), output_file="chk", compare=True, silent=5)
    ref = armon.armon(reference_params("Sod_circ", is_ref=True, return_data=True, **kw))
    assert ref.cycles == 3
    files = sorted(p.name for p in tmp_path.iterdir())
    assert "chk_000_init_test_X" in files and "chk_000_time_step_X" in files and "chk_002_projection_remap_Y" in files
    assert len(files) == 2 + 3 * (1 + 2 * 5)        # init_test, EOS_init + per cycle: time_step + 2 axes x 5 steps
    # the axis letter of a time_step checkpoint is the axis of the previous sweep (X only before the first cycle)
    assert len((tmp_path / "chk_001_time_step_Y").read_text().strip()) == 23
    ref.data.close()

    same = armon.armon(reference_params("Sod_circ", is_ref=False, return_data=True, **kw))
    assert same.cycles == 3 and "differences found" not in capsys.readouterr().out
    same.data.close()

    other = armon.armon(reference_params("Sod_circ", is_ref=False, return_data=True, riemann_limiter="superbee", **kw))
    out = capsys.readouterr().out
    # the saved variables (x, y, rho, u, v, p) first feel the limiter in a cell_update of cycle 0: the run stops there
    assert other.cycles == 0 and "At cell_update, in block (1, 1):" in out and "differences found in rho (ref ≢ current)" in out
    diff = [p.name for p in tmp_path.iterdir() if p.name.endswith("_diff")]
    assert len(diff) == 1 and diff[0].startswith("chk_000_cell_update_") and f"Difference file written to {diff[0]}" in out
    other.data.close()


@pytest.mark.parametrize("test", GOLDEN_TESTS)
def test_float32_boundary(test, golden):
    """`ArmonParameters{Float32}`: the caller's arrays are Float32, the device computes in Float64 (conversion in the
    host <-> device copies).  Against the reference's golden `ref_*_32bits.csv` (test/reference_data, tolerances
    reference_functions.jl:56,58: atol 1e-5, rtol 20 eps(Float32)): same cycle count; fields equal to the Float64 golden
    data rounded to Float32, i.e. as far from the 32-bit golden data as the reference's own two precisions are from each
    other -- its Float32 arithmetic carries rounding noise above its own tolerance for Sedov and Bizarrium, which a
    Float64 computation cannot and should not reproduce.  Sod, Sod_y and Sod_circ pass the reference's acceptance test
    outright."""
    import os
    g32 = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", f"ref_{test}_32bits.npz"))
    g64 = golden(test)
    stats, grid = run_gpu(reference_params(test, data_type="Float32"))
    assert grid.real("rho").dtype == np.float32
    assert stats.cycles == int(g32["cycles"]) == int(g64["cycles"])
    eps32 = float(np.finfo(np.float32).eps)
    report = {}
    for var in ("rho", "u", "v", "p"):
        got = grid.real(var)
        # our Float32 output is the Float64 result rounded once
        assert np.array_equal(got, g64[var].astype(np.float32)) or scaled_max_diff(got.astype(np.float64), g64[var]) <= eps32
        ours = float(np.abs(got.astype(np.float64) - g32[var].astype(np.float64)).max())
        theirs = float(np.abs(g64[var] - g32[var].astype(np.float64)).max())
        assert ours <= theirs + 2 * eps32 * float(np.abs(g64[var]).max()), (var, ours, theirs)
        report[var] = count_differences(got.astype(np.float64), g32[var].astype(np.float64), atol=1e-5, rtol=20 * eps32)
    assert abs(stats.last_dt - float(g32["dt"])) <= abs(float(g64["dt"]) - float(g32["dt"])) + 1e-13
    print(f"float32 {test}: cells failing the reference's Float32 acceptance vs ref_{test}_32bits: {report}")
    if test in ("Sod", "Sod_y", "Sod_circ"):
        assert sum(report.values()) == 0, report      # the reference's own Float32 acceptance test
    grid.close()
