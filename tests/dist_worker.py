"""Multi-rank worker, launched by tests/test_distributed.py under `python -m torch.distributed.run`.

  --mode cpu : world_size-N gloo run.  Host-side logic of the one-process-per-GPU path (Cartesian process grid,
               sub-domain sizes / origins / neighbours of init_MPI + init_indexing, src/parameters.jl:408-467,673-697;
               unique-id broadcast; scalar all-reduces) and the decomposition semantics themselves: every rank runs the
               CPU oracle on its sub-domain with the halo exchange (4 strips of rho,u,v,E,p,c,g along the swept axis,
               no corners, src/halo_exchange.jl:187-310) and the dt all-reduce(min) (src/solver_state.jl:107-111) done
               over gloo; the gathered result must equal the single-domain oracle bit for bit
               (the analogue of test/mpi.jl:363-398 "sub-domain vs global reference").
  --mode gpu : world_size-N NCCL run on N B200s: the fused CUDA path with NCCL halo exchange + dt all-reduce must
               equal the 1-GPU fused result bit for bit; DebugIndexes halo test (test/mpi.jl:272-360).
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch                      # noqa: E402
import torch.distributed as dist  # noqa: E402

import armon_jl_b200 as armon                      # noqa: E402
from armon_jl_b200 import Axis, Side               # noqa: E402
from armon_jl_b200 import distributed as adist     # noqa: E402

COMM_VARS = ("rho", "u", "v", "E", "p", "c", "g")   # comm_vars(), src/blocking/blocks.jl:50
SCHEME = dict(scheme="GAD", riemann_limiter="minmod", projection="euler_2nd", nghost=4, silent=5)


def expected_decomposition(global_n, P, rank):
    """Independent restatement of init_MPI / init_indexing for the checks below."""
    cx, cy = rank // P[1], rank % P[1]
    coords = (cx, cy)
    n = tuple(global_n[d] // P[d] + (global_n[d] % P[d] if coords[d] == P[d] - 1 else 0) for d in range(2))
    origin = tuple(coords[d] * (global_n[d] // P[d]) + 1 for d in range(2))

    def rk(x, y):
        return x * P[1] + y if 0 <= x < P[0] and 0 <= y < P[1] else -1
    nb = {Side.Left: rk(cx - 1, cy), Side.Right: rk(cx + 1, cy), Side.Bottom: rk(cx, cy - 1), Side.Top: rk(cx, cy + 1)}
    return coords, n, origin, nb


def check_decomposition(params, global_n, P, rank):
    coords, n, origin, nb = expected_decomposition(global_n, P, rank)
    assert params.cart_coords == coords, (params.cart_coords, coords)
    assert params.N == n and params.N_origin == origin and params.global_grid == tuple(global_n)
    assert dict(params.neighbours) == nb, (params.neighbours, nb)
    # the sub-domains tile the global grid exactly
    t = torch.tensor([n[0] * n[1]], dtype=torch.float64)
    if dist.get_backend() == "nccl":
        t = t.cuda()
    dist.all_reduce(t)
    assert int(t.item()) == global_n[0] * global_n[1]


def exchange_rows(send_lo, send_hi, lo, hi):
    """Two-sided exchange of numpy blocks with the low / high neighbour (ranks or -1) over torch.distributed."""
    ops, recv = [], {}
    for name, peer, payload in (("lo", lo, send_lo), ("hi", hi, send_hi)):
        if peer < 0:
            continue
        s = torch.from_numpy(np.ascontiguousarray(payload))
        r = torch.empty_like(s)
        recv[name] = r
        ops.append(dist.P2POp(dist.isend, s, peer))
        ops.append(dist.P2POp(dist.irecv, r, peer))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    return recv.get("lo"), recv.get("hi")


def oracle_halo(orc, params):
    """block_ghost_exchange with remote neighbours for the oracle's arrays: the g innermost real strips of each side
    along `axis` go to the neighbour's ghost strips, same orientation; rows/columns of REAL cells only (no corners)."""
    g, nx, ny = params.nghost, params.N[0], params.N[1]

    def halo(axis):
        x = int(axis) == int(Axis.X)
        lo = params.neighbours[Side.Left if x else Side.Bottom]
        hi = params.neighbours[Side.Right if x else Side.Top]
        if lo < 0 and hi < 0:
            return
        arrs = [orc.array(v) for v in COMM_VARS]
        if x:
            send_lo = np.stack([a[g:g + ny, g:2 * g] for a in arrs])
            send_hi = np.stack([a[g:g + ny, nx:nx + g] for a in arrs])
        else:
            send_lo = np.stack([a[g:2 * g, g:g + nx] for a in arrs])
            send_hi = np.stack([a[ny:ny + g, g:g + nx] for a in arrs])
        r_lo, r_hi = exchange_rows(send_lo, send_hi, lo, hi)
        for k, a in enumerate(arrs):
            if x:
                if r_lo is not None:
                    a[g:g + ny, 0:g] = r_lo[k].numpy()
                if r_hi is not None:
                    a[g:g + ny, nx + g:nx + 2 * g] = r_hi[k].numpy()
            else:
                if r_lo is not None:
                    a[0:g, g:g + nx] = r_lo[k].numpy()
                if r_hi is not None:
                    a[ny + g:ny + 2 * g, g:g + nx] = r_hi[k].numpy()
    return halo


def allreduce_min(x):
    t = torch.tensor([x], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    return float(t[0])


def gather_global(local, params, P, global_n):
    """Gather the real cells of every rank on rank 0 into one [global_ny, global_nx] array."""
    world = dist.get_world_size()
    objs = [None] * world if dist.get_rank() == 0 else None
    dist.gather_object((params.N_origin, params.N, np.ascontiguousarray(local)), objs, dst=0)
    if dist.get_rank() != 0:
        return None
    out = np.full((global_n[1], global_n[0]), np.nan)
    for (ox, oy), (nx, ny), block in objs:
        out[oy - 1:oy - 1 + ny, ox - 1:ox - 1 + nx] = block
    assert not np.isnan(out).any()
    return out


def run_cpu(args):
    from oracle import OracleSolver
    rank, world = adist.init_process_group("gloo")
    assert world == args.world
    # unique-id broadcast plumbing (stands for MPI.bcast of the NCCL id)
    payload = bytes(range(128)) if rank == 0 else None
    assert adist.broadcast_bytes(payload, src=0) == bytes(range(128))
    assert adist.allreduce_max(float(rank)) == float(world - 1)
    assert adist.allreduce_sum((1.0, float(rank))) == (float(world), float(sum(range(world))))

    cases = [("Sod_circ", (40, 36), 12), ("Sod", (37, 24), 8), ("Sedov", (30, 30), 10)]
    grids = [(1, world), (world, 1)] + ([(2, world // 2)] if world >= 4 and world % 2 == 0 else [])
    for test, global_n, cycles in cases:
        ref = None
        if rank == 0:
            single = OracleSolver(armon.ArmonParameters(test=test, N=global_n, maxcycle=cycles, **SCHEME), "strict", nthreads=1)
            _, ref_dt, ref_cycles, err = single.time_loop()
            assert err == 0
            ref = {v: single.real(v).copy() for v in ("rho", "u", "v", "E", "p")}
        for P in grids:
            params = armon.ArmonParameters(test=test, N=global_n, maxcycle=cycles, use_MPI=True, P=P, rank=rank,
                                           proc_size=world, **SCHEME)
            check_decomposition(params, global_n, P, rank)
            orc = OracleSolver(params, "strict", nthreads=1)
            orc.set_hooks(halo=oracle_halo(orc, params), allreduce_min=allreduce_min)
            _, dt, ncyc, err = orc.time_loop()
            assert err == 0
            for v in ("rho", "u", "v", "E", "p"):
                glob = gather_global(orc.real(v), params, P, global_n)
                if rank == 0:
                    assert np.array_equal(glob, ref[v]), f"{test} P={P} {v}: max diff {np.abs(glob - ref[v]).max()}"
            if rank == 0:
                assert (dt, ncyc) == (ref_dt, ref_cycles), (test, P, dt, ref_dt)
            orc.close()
            dist.barrier()
    # configuration errors of init_MPI / init_indexing
    for bad in (dict(P=(3, world)), dict(N=(4, 2 * world), P=(1, world), nghost=4)):
        kw = dict(test="Sod", N=(40, 40), use_MPI=True, P=(1, world), rank=rank, proc_size=world, **SCHEME)
        kw.update(bad)
        try:
            armon.ArmonParameters(**kw)
        except armon.SolverException as e:
            assert e.category == "config"
        else:
            raise AssertionError(f"{bad} accepted")
    if rank == 0:
        print("dist_worker cpu ok")
    dist.barrier()
    dist.destroy_process_group()


def log(*a):
    print(f"[rank {os.environ.get('RANK', '?')}]", *a, file=sys.stderr, flush=True)


def run_gpu(args):
    rank, world = adist.init_process_group("nccl", timeout_s=90)
    assert world == args.world
    grids = [(1, world), (world, 1)] + ([(2, world // 2)] if world >= 4 and world % 2 == 0 else [])
    cases = [("Sod_circ", (192, 160), 12, "strict", "Sequential", 0), ("Sedov", (128, 128), 10, "strict", "Godunov", 0),
             ("Bizarrium", (160, 96), 8, "strict", "Strang", 0), ("Sod_circ", (2048, 1024), 4, "fast", "Sequential", 0),
             # uneven decompositions (remainder on the last rank, test/mpi.jl:551-561) and local extents of 50 cells with
             # 16-row march segments: the last segment is 2 rows (< nghost), so the segment before it also reads ghost rows
             # and must run with the edge launches of the overlapped sweep
             ("Sod_circ", (107, 113), 12, "strict", "Sequential", 16),
             ("Sod_circ", (50 * world, 50 * world), 10, "strict", "Sequential", 16),
             ("Sod_circ", (50 * world, 50 * world), 10, "fast", "Godunov", 16),
             # band-tiled layout of the fast mode (local extents multiples of 8): ghost bands through NCCL, sweeps that
             # keep their axis (Godunov splitting); and a decomposition where only SOME ranks could tile (64 + 65 columns
             # on 2 ranks): the ranks must agree not to, the ghost bands travel as raw memory
             ("Sod_circ", (64 * world, 128), 11, "fast", "Godunov", 0),
             ("Sod_circ", (64 * world + 1, 64), 10, "fast", "Sequential", 0)]
    if args.quick:
        cases = [cases[0], cases[5], cases[7]]
    for test, global_n, cycles, math, splitting, segment in cases:
        kw = dict(test=test, N=global_n, maxcycle=cycles, math_mode=math, axis_splitting=splitting,
                  march_segment=segment, return_data=True, **SCHEME)
        ref = None
        log("case", test, global_n, math, splitting)
        if rank == 0:
            stats = armon.armon(armon.ArmonParameters(**kw))
            ref = {v: stats.data.real(v).copy() for v in ("rho", "u", "v", "E", "p")}
            ref_dt, ref_cycles = stats.last_dt, stats.cycles
            stats.data.close()
        dist.barrier()
        for P in grids:
            params = armon.ArmonParameters(use_MPI=True, P=P, rank=rank, proc_size=world, **kw)
            check_decomposition(params, global_n, P, rank)
            log("  P", P, "running")
            stats = armon.armon(params)
            log("  P", P, "done", stats.cycles, stats.last_dt)
            for v in ("rho", "u", "v", "E", "p"):
                glob = gather_global(stats.data.real(v), params, P, global_n)
                if rank == 0:
                    assert np.array_equal(glob, ref[v]), f"{test} P={P} {v}: max diff {np.abs(glob - ref[v]).max()}"
            if rank == 0:
                assert (stats.last_dt, stats.cycles) == (ref_dt, ref_cycles), (test, P, stats.last_dt, ref_dt)
            # conservation through the all-reduce path (src/reductions.jl:317-320)
            mass, energy = armon.conservation_vars(params, stats.data)
            assert np.isfinite(mass) and np.isfinite(energy)
            stats.data.close()
            dist.barrier()

    # DebugIndexes-style halo test (test/mpi.jl:272-360): payload = rank*1e6 + iy*1000 + ix of the sender
    for P in grids:
        params = armon.ArmonParameters(test="Sod", N=(64, 48), use_MPI=True, P=P, rank=rank, proc_size=world,
                                       return_data=True, **SCHEME)
        log("halo test P", P)
        grid = armon.BlockGrid(params)
        armon.init_test(params, grid)
        g, (nx, ny) = params.nghost, params.N
        iy, ix = np.meshgrid(np.arange(1, ny + 1), np.arange(1, nx + 1), indexing="ij")
        for k, v in enumerate(("rho", "u", "v", "E")):
            full = np.full(grid.shape, -1.0)
            full[g:g + ny, g:g + nx] = rank * 1e6 + iy * 1e3 + ix + 0.25 * k
            grid.set_array(v, full)
        for axis in (Axis.X, Axis.Y):
            from armon_jl_b200.backend import check
            check(grid.lib.armon_solver_halo_exchange(grid.solver, int(axis)))
            grid._fused_dirty = True
        for k, v in enumerate(("rho", "u", "v", "E")):
            full = grid.host_array(v)
            for side, sl_ghost, own_slice in (
                    (Side.Left, (slice(g, g + ny), slice(0, g)), "x_hi"), (Side.Right, (slice(g, g + ny), slice(nx + g, nx + 2 * g)), "x_lo"),
                    (Side.Bottom, (slice(0, g), slice(g, g + nx)), "y_hi"), (Side.Top, (slice(ny + g, ny + 2 * g), slice(g, g + nx)), "y_lo")):
                nb = params.neighbours[side]
                got = full[sl_ghost]
                if nb < 0:
                    assert (got == -1.0).all(), f"{v} {side}: ghost of a global edge was touched"
                    continue
                _, nb_n, _, _ = expected_decomposition(params.global_grid, P, nb)
                nnx, nny = nb_n
                if own_slice == "x_hi":      # neighbour's g right-most real columns
                    jy, jx = np.meshgrid(np.arange(1, nny + 1), np.arange(nnx - g + 1, nnx + 1), indexing="ij")
                elif own_slice == "x_lo":
                    jy, jx = np.meshgrid(np.arange(1, nny + 1), np.arange(1, g + 1), indexing="ij")
                elif own_slice == "y_hi":
                    jy, jx = np.meshgrid(np.arange(nny - g + 1, nny + 1), np.arange(1, nnx + 1), indexing="ij")
                else:
                    jy, jx = np.meshgrid(np.arange(1, g + 1), np.arange(1, nnx + 1), indexing="ij")
                want = nb * 1e6 + jy * 1e3 + jx + 0.25 * k
                assert np.array_equal(got, want), f"rank {rank} {v} {side}: halo payload mismatch"
        grid.close()
        dist.barrier()
    if rank == 0:
        print("dist_worker gpu ok")
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--mode", choices=["cpu", "gpu"], required=True)
    ap.add_argument("--world", type=int, required=True)
    ap.add_argument("--quick", action="store_true")
    a = ap.parse_args()
    (run_cpu if a.mode == "cpu" else run_gpu)(a)
