#!/usr/bin/env python
"""bench.py -- giga cell-updates/s of the fused Lagrange+remap time step on B200, one JSON line.

  python bench.py --gpus N --steps K --warmup W [--workload NAME] [--math strict|fast] [--impl reference]

A "step" is one solver cycle (all axis sweeps of the splitting + the CFL time-step update) over the whole grid.
Default workload: BASELINE.json configs[4], Sod_circ weak scaling with 16384 x 16384 cells per GPU (GAD + minmod +
euler_2nd, Float64) -- the largest single-GPU configuration at N=1, 65536 x 32768 on 8 GPUs -- with Cartesian process
grid, NCCL halo exchange and dt all-reduce for N>1.  At N=1 the same JSON line carries secondary figures for
configs[1] (Sod_circ 8192^2), configs[2] (Bizarrium 16384^2) and the bit-exact strict mode.  `value` is timed with the
state resident in HBM; `e2e` is the same job through the public API with host buffers (h2d of the initial state, a
blocking read of the time-step state every cycle, d2h of the final fields); `parity` compares the benched
configuration with the CPU oracle.  `--impl reference` times the restated reference CPU path (oracle/, C + OpenMP, all
host cores) on the same grid -- the reference itself is Julia and cannot run here.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

METRIC = "giga cell-updates/s per time step (Float64)"
UNIT = "Gcell-updates/s"
BYTES_PER_CELL_SWEEP = 64   # read rho,u,v,E + write rho,u,v,E
FP64_LANES_PER_CLK_SM = 64  # DFMA per clock and SM on B200 (ncu sm__sass_thread_inst_executed_op_dfma_pred_on.peak_sustained)

WORKLOADS = {
    "sod_circ_16384": dict(test="Sod_circ", n=(16384, 16384), scaling="weak", prefer="x",
                           desc="Sod_circ 16384x16384 per GPU weak scaling, GAD+minmod+euler_2nd, Float64 (BASELINE configs[4])"),
    "sod_circ_8192": dict(test="Sod_circ", n=(8192, 8192), scaling="weak", prefer="y",
                          desc="Sod_circ 8192x8192 per GPU, GAD+minmod+euler_2nd, Float64 (BASELINE configs[1])"),
    "bizarrium_16384": dict(test="Bizarrium", n=(16384, 16384), scaling="weak", prefer="y",
                            desc="Bizarrium 16384x16384 per GPU, GAD+minmod+euler_2nd, Float64 (BASELINE configs[2])"),
    "sedov_32768_strong": dict(test="Sedov", n=(32768, 32768), scaling="strong", prefer="y",
                               desc="Sedov 32768x32768 global, strong scaling, GAD+minmod+euler_2nd, Float64 (BASELINE configs[3])"),
    "sod_circ_1024": dict(test="Sod_circ", n=(1024, 1024), scaling="weak", prefer="y", desc="small smoke workload"),
}
DEFAULT_WORKLOAD = "sod_circ_16384"
SECONDARY = ("sod_circ_8192", "bizarrium_16384")     # extra figures in the N=1 line


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def load_profile_json(name):
    try:
        with open(os.path.join(ROOT, "profiles", name)) as f:
            return json.load(f)
    except Exception:
        return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device_index):
        self.idx = device_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.idx)], stdout=subprocess.PIPE, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons, power = [], [], set(), []
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


def scheme_kwargs():
    return dict(scheme="GAD", riemann_limiter="minmod", projection="euler_2nd", axis_splitting="Sequential",
                nghost=4, silent=5)


def process_grid(world, w, override=None):
    from armon_jl_b200 import distributed as adist
    return tuple(override) if override else adist.process_grid_for(world, prefer=w["prefer"])


def workload_config(w, world, P, cfl):
    """The `config` object of the JSON line: identical for the B200 arm and the reference arm."""
    if w["scaling"] == "weak":
        global_n = (w["n"][0] * P[0], w["n"][1] * P[1])
    else:
        global_n = w["n"]
    per_gpu = (global_n[0] // P[0], global_n[1] // P[1])
    cells = (per_gpu[0] + 8) * (per_gpu[1] + 8)
    return {"workload": w["desc"], "test": w["test"], "global_grid": list(global_n), "per_gpu_grid": list(per_gpu),
            "process_grid": list(P), "scheme": "GAD + minmod + euler_2nd", "axis_splitting": "Sequential",
            "nghost": 4, "cfl": cfl, "dtype": "f64",
            "l2_policy": "inputs larger than L2 (8 arrays x %.0f MB per GPU vs 126 MB L2)" % (cells * 8 / 1e6)
            if cells * 8 * 8 > 4 * 126e6 else "inputs fit in L2 (small workload)"}, global_n


def pin_to_gpu_numa_node(local_rank):
    """Run this process (and first-touch its pinned host buffers) on the CPUs next to its GPU."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = {64 * i + b for i, wd in enumerate(words) for b in range(64) if (wd >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
        return len(cpus)
    except Exception:
        return 0


def host_memory_available():
    """Bytes of host memory this process may still take (MemAvailable, capped by the container's cgroup limit)."""
    try:
        with open("/proc/meminfo") as f:
            avail = next(int(ln.split()[1]) * 1024 for ln in f if ln.startswith("MemAvailable"))
    except Exception:
        return 1 << 62
    try:   # a container may be capped below what /proc/meminfo shows
        with open("/sys/fs/cgroup/memory.max") as f, open("/sys/fs/cgroup/memory.current") as g:
            cap = f.read().strip()
            if cap != "max":
                avail = min(avail, int(cap) - int(g.read().strip()))
    except Exception:
        pass
    return avail


def oracle_fits_in_host_memory(n, reserve_gb=8.0):
    """The oracle holds the reference's 16 arrays of (nx+8)(ny+8) doubles; never start one that would push the box
    into swap / the OOM killer (malloc succeeds under overcommit)."""
    need = 16 * (n[0] + 8) * (n[1] + 8) * 8
    return need + reserve_gb * 2**30 < host_memory_available()


# -------------------------------------------------------------------------------------------------------------
# CPU arm: the restated reference CPU path (oracle/), all host threads
# -------------------------------------------------------------------------------------------------------------
def cpu_run(test, n, steps, warmup, nthreads=0):
    """(Gcell-updates/s, seconds, threads, grid actually run) of the oracle's `fast` flavour.  Falls back to smaller
    grids when the 16 host arrays of the reference do not fit in memory."""
    import armon_jl_b200 as armon
    import oracle
    oracle.build()
    threads = nthreads or (os.cpu_count() or 1)
    n = tuple(n)
    while not oracle_fits_in_host_memory(n) and min(n) > 2048:
        n = (n[0] // 2, n[1] // 2)
    params = armon.ArmonParameters(test=test, N=n, maxcycle=10**9, **scheme_kwargs())
    s = oracle.OracleSolver(params, "fast", nthreads=threads)
    for _ in range(warmup):
        s.solver_cycle()
    t0 = time.perf_counter()
    for _ in range(steps):
        s.solver_cycle()
    dt = time.perf_counter() - t0
    s.close()
    return n[0] * n[1] * steps / dt / 1e9, dt, threads, n


def cpu_sample_text(test, n, ran, steps, warmup):
    where = "the workload's per-GPU grid" if tuple(n) == tuple(ran) else f"a {ran[0]}x{ran[1]} sub-grid (host memory)"
    return (f"{test} {ran[0]}x{ran[1]} ({where}), {steps} cycles after {warmup} warm-up, same scheme; restated reference "
            "CPU path (oracle/, C + OpenMP over rows, -O3 -ffast-math -march=native, unfused 7-pass structure like "
            "the reference), not the Julia code (no julia in this image)")


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    if rank != 0:
        return 0
    import armon_jl_b200 as armon
    w = WORKLOADS[args.workload]
    P = process_grid(max(world, args.gpus), w, args.proc_grid)
    cfl = armon.ArmonParameters(test=w["test"], N=(64, 64), **scheme_kwargs()).cfl
    config, global_n = workload_config(w, max(world, args.gpus), P, cfl)
    # bounded sample: one rank's sub-domain of the workload (the whole grid at N=1)
    n = tuple(config["per_gpu_grid"])
    value, secs, threads, ran = cpu_run(w["test"], n, args.steps, args.warmup)
    sample = cpu_sample_text(w["test"], n, ran, args.steps, args.warmup)
    out = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": secs / args.steps * 1e3, "higher_is_better": True, "scaling": w["scaling"],
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(out)
    return 0


# -------------------------------------------------------------------------------------------------------------
# GPU arm
# -------------------------------------------------------------------------------------------------------------
class GpuJob:
    """One workload on this rank's GPU: parameters, grid, and the measurements the JSON line is built from."""

    def __init__(self, args, w, math, world, rank, local_rank, P):
        import armon_jl_b200 as armon
        self.armon, self.args, self.w, self.math = armon, args, w, math
        self.world, self.rank, self.multi = world, rank, world > 1
        cfl = armon.ArmonParameters(test=w["test"], N=(64, 64), **scheme_kwargs()).cfl
        self.config, global_n = workload_config(w, world, P, cfl)
        self.params = armon.ArmonParameters(test=w["test"], N=global_n, use_MPI=self.multi, P=P, rank=rank,
                                            proc_size=world, maxcycle=10**9, math_mode=math,
                                            march_segment=args.segment, bind_pcg=False, device_id=local_rank,
                                            return_data=True, **scheme_kwargs())
        self.grid = armon.BlockGrid(self.params)
        self.lib, self.dev = self.grid.lib, self.grid.device
        self.local_cells = self.params.N[0] * self.params.N[1]
        self.global_cells = self.params.global_grid[0] * self.params.global_grid[1]

    def sync_all(self):
        from armon_jl_b200 import distributed as adist
        self.dev.wait()
        if self.multi:
            adist.barrier()

    def timed_run(self, steps, warmup):
        """W warm-up cycles, then K cycles timed with CUDA events on the solver stream (max over ranks) with the sweep
        launches timed individually (roofline of the dominant kernel)."""
        from armon_jl_b200 import distributed as adist
        from armon_jl_b200.backend import check
        armon, grid, lib = self.armon, self.grid, self.lib
        armon.init_test(self.params, grid)
        solver = grid.solver
        grid.run(warmup)
        self.sync_all()
        launches0 = self.dev.launch_count()
        check(lib.armon_solver_profile(solver, 1))
        t0 = time.perf_counter()
        grid.run(steps)
        ms = grid.elapsed_ms()
        self.sync_all()
        host_s = time.perf_counter() - t0
        launches = self.dev.launch_count() - launches0
        sweep_ms, sweep_n = C.c_double(), C.c_uint64()
        check(lib.armon_solver_sweep_time_ms(solver, C.byref(sweep_ms), C.byref(sweep_n)))
        check(lib.armon_solver_profile(solver, 0))
        elapsed_s = ms / 1e3
        if self.multi:
            elapsed_s = adist.allreduce_max(elapsed_s)
        st = grid.time_state()
        if st.error or st.done:
            raise SystemExit(f"bench invalid: solver stopped early (error={st.error}, done={st.done}, cycle={st.cycle})")
        return {"elapsed_s": elapsed_s, "host_s": host_s, "launches": int(launches), "sweep_ms": sweep_ms.value,
                "sweep_n": int(sweep_n.value), "device_ms": ms, "state": st,
                "value": self.global_cells * steps / elapsed_s / 1e9}

    def roofline(self, run, clocks_mhz):
        from armon_jl_b200.backend import check
        peak, peak_src = measured_peak()
        avg_s = run["sweep_ms"] / max(run["sweep_n"], 1) / 1e3
        achieved = BYTES_PER_CELL_SWEEP * self.local_cells / avg_s / 1e9
        biz = self.w["test"] == "Bizarrium"
        variant = os.environ.get("ARMON_B200_KERNEL") or {"fast": "tma", "strict": "single", "ieee": "single"}[self.math]
        kname = {"tma": "sweep_fast_kernel<STG_TMA", "async2": "sweep_fast_kernel<STG_CPA16",
                 "single": "sweep_kernel<%s" % {"fast": "fd, DIV_FAST", "strict": "sd, DIV_FLAGGED", "ieee": "sd, DIV_IEEE"}[self.math]}.get(variant, variant)
        # which layout the fused path actually ran in (sweep_fast_kernel.cuh 5.): asked from the library, not assumed
        if self.grid.fused_layout_is_tiled() == 1:
            variant, kname = "tiled", "sweep_fast_kernel<STG_TMA, LAY_TILED"
        elif self.math == "strict" and self.grid.strict_kernel_is_chains() == 1 and self.params.N[0] % 2 == 0 and self.params.N[1] % 2 == 0:
            variant, kname = "chains", "sweep_fast_kernel<STG_TMA, MATH_STRICT"
        key = f"{variant}_{self.math}_{'biz' if biz else 'pg'}"
        traffic = load_profile_json("sweep_traffic.json") or {}
        t = traffic.get(key)
        budget = (load_profile_json("sass_budget.json") or {}).get(key)
        out = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
               # ncu dram__bytes_read.sum + dram__bytes_write.sum of one launch of this kernel at this grid
               # (profiles/sweep_traffic.json, captured once per kernel change -- not measured in this run)
               "traffic": t.get("bytes_per_launch") if t and t.get("cells_per_launch") == self.local_cells else None,
               "peak_source": peak_src,
               "kernel": f"{kname}, GAD+minmod, euler_2nd, {'bizarrium' if biz else 'perfect gas'}>",
               "avg_launch_ms": avg_s * 1e3, "launches_timed": run["sweep_n"],
               "algorithmic_bytes_per_launch": BYTES_PER_CELL_SWEEP * self.local_cells,
               "sweep_share_of_step": run["sweep_ms"] / run["device_ms"]}
        if budget and clocks_mhz:
            # second roof: FP64 pipe.  Static count of FP64 instructions per march step (= per cell and
            # thread) of the steady-state loop (profiles/sass_budget.json, cuobjdump) x cells per launch / time, against
            # 64 FP64 lanes per clock and SM at the SM clock sampled during the timed region.
            fp64_peak = FP64_LANES_PER_CLK_SM * self.dev_sm_count() * clocks_mhz * 1e6
            fp64_rate = budget["fp64_per_step"] * self.local_cells / avg_s
            issue_rate = budget["instr_per_step"] * self.local_cells / 32 / avg_s
            out["fp64"] = {"achieved": fp64_rate / 1e12, "peak": fp64_peak / 1e12, "unit": "T FP64 instr/s (thread level)",
                           "frac": fp64_rate / fp64_peak, "fp64_instr_per_cell": budget["fp64_per_step"],
                           "warp_instr_per_32_cells": budget["instr_per_step"],
                           "issue_slot_frac": issue_rate / (4 * self.dev_sm_count() * clocks_mhz * 1e6),
                           "sm_clock_mhz": clocks_mhz, "source": "profiles/sass_budget.json (static SASS count of the march loop)"}
        return out

    def dev_sm_count(self):
        return 148

    def e2e(self, steps):
        """Same job through the public API with HOST buffers: h2d of rho,u,v,E from pinned memory, K x (solver_cycle +
        blocking read of the time-step state), finalize, d2h of rho,u,v,E; wall clock, max over ranks."""
        import torch
        from armon_jl_b200 import distributed as adist
        from armon_jl_b200.backend import check
        armon, grid, params = self.armon, self.grid, self.params
        names = ("rho", "u", "v", "E")
        armon.init_test(params, grid)          # fresh initial state on the device -> host copy (untimed set-up)
        # One set of host buffers per rank, used in both directions (the result overwrites the input, as a caller that
        # advances a state in place would do).  Pinned when every rank's set fits comfortably in host memory
        # (8 ranks x 8.6 GB at 16384^2 per GPU), pageable otherwise -- slower copies, but never a box out of memory.
        need = 4 * grid.cell_count * 8
        pin = host_memory_available() > 1.5 * need * self.world + (16 << 30)
        host0 = {}
        for k in names:
            buf = torch.empty(grid.cell_count, dtype=torch.float64, pin_memory=pin)
            getattr(grid.device_data, k).copy_to_host(buf.numpy())
            host0[k] = buf.numpy()
        out_pinned = {k: torch.from_numpy(host0[k]) for k in names}
        self.sync_all()
        t0 = time.perf_counter()
        for k in names:                        # h2d of the job's input
            getattr(grid.device_data, k).copy_from_host(host0[k])
        check(self.lib.armon_solver_reset(grid.solver))
        grid._fused_dirty = False
        last = None
        for _ in range(steps):                 # one public-API call + a blocking read of the step's result per cycle
            armon.solver_cycle(params, grid)
            last = grid.time_state()
        grid.finalize()
        for k in names:                        # d2h of the job's result
            getattr(grid.device_data, k).copy_to_host(out_pinned[k].numpy())
        self.dev.wait()
        e2e_s = time.perf_counter() - t0
        if self.multi:
            adist.barrier()
            e2e_s = adist.allreduce_max(e2e_s)
        import numpy as np
        checksum = float(out_pinned["rho"].numpy()[::97].sum())
        assert np.isfinite(checksum) and last.cycle == steps
        h2d = 4 * grid.cell_count * 8 / steps
        d2h = (4 * grid.cell_count * 8 + steps * C.sizeof(armon.backend.armon_time_state)) / steps
        return {"value": self.global_cells * steps / e2e_s / 1e9, "unit": UNIT, "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h, "steps": steps,
                "host_buffers": "pinned" if pin else "pageable (not enough host memory to pin every rank's buffers)",
                "note": "h2d of rho,u,v,E from pinned host memory + K x (solver_cycle + blocking read of the time-step "
                        "state) + finalize + d2h of rho,u,v,E, wall clock, max over ranks; the two field transfers are "
                        "one-off per job (a real run to maxtime is thousands of cycles)"}

    def parity(self, cycles):
        """The benched configuration (this grid, this arithmetic mode, automatic march segments) against the CPU oracle
        (strict flavour, all host cores) after `cycles` cycles from the initial state; N=1 only."""
        import numpy as np
        import oracle
        armon, grid, params = self.armon, self.grid, self.params
        n = tuple(params.N)
        note = "the bench grid itself"
        t0 = time.perf_counter()
        if not oracle_fits_in_host_memory(n, reserve_gb=16.0):
            return {"skipped": "the oracle's 16 host arrays do not fit in host memory at this grid"}
        op = armon.ArmonParameters(test=self.w["test"], N=n, maxcycle=cycles, **scheme_kwargs())
        orc = oracle.OracleSolver(op, "strict", nthreads=os.cpu_count() or 1)
        _, odt, ocyc, err = orc.time_loop()
        armon.init_test(params, grid)
        grid.run(cycles)
        st = grid.time_state()
        worst = {}
        for v in ("rho", "u", "v", "E"):
            want = orc.real(v)
            got = grid.real(v)
            scale = float(np.abs(want).max()) or 1.0
            worst[v] = float(np.abs(got - want).max() / scale)
            del got
        orc.close()
        return {"cycles": int(cycles), "max_scaled_diff": max(worst.values()), "per_field": worst,
                "dt_rel_diff": abs(st.current_dt - odt) / odt, "cycles_match": bool(st.cycle == ocyc and err == 0),
                "grid": list(n), "math_mode": self.math, "oracle": "oracle/armon_oracle.c strict flavour, OpenMP "
                f"{os.cpu_count()} threads", "against": note, "tolerance": 0.0 if self.math != "fast" else 1e-12,
                "seconds": time.perf_counter() - t0}

    def close(self):
        self.grid.close()


def run_gpu_arm(args):
    from armon_jl_b200 import distributed as adist

    world = int(os.environ.get("WORLD_SIZE", 1))
    rank = int(os.environ.get("RANK", 0))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torchrun for --gpus > 1 (one process per GPU)")
    multi = world > 1
    all_cpus = os.sched_getaffinity(0)
    numa_cpus = pin_to_gpu_numa_node(local_rank)   # undone before the CPU legs, which use every host core
    if multi:
        import torch  # noqa: F401  (torch's NCCL is loaded before the library links against libnccl.so.2)
        adist.init_process_group("nccl")

    w = WORKLOADS[args.workload]
    P = process_grid(world, w, args.proc_grid)
    job = GpuJob(args, w, args.math, world, rank, local_rank, P)

    # ---- timed region: K cycles, state resident in HBM ----
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
    run = job.timed_run(args.steps, args.warmup)
    clocks = sampler.stop() if rank == 0 else None
    sm_mhz = (clocks or {}).get("sm_mhz")
    roofline = job.roofline(run, sm_mhz)

    # ---- e2e: same job through the public API with HOST buffers ----
    e2e = job.e2e(args.steps)

    # ---- parity of the benched configuration against the CPU oracle (N=1) ----
    os.sched_setaffinity(0, all_cpus)
    parity = None
    if world == 1 and args.parity_cycles > 0 and not args.no_cpu:
        parity = job.parity(args.parity_cycles)
    config = job.config
    job.close()
    del job
    import gc
    gc.collect()

    # ---- secondary figures: the bit-exact (strict) mode on the same workload, and at N=1 the other single-GPU configs ----
    def short_figure(wname, math, steps):
        sj = GpuJob(args, WORKLOADS[wname], math, world, rank, local_rank, process_grid(world, WORKLOADS[wname], args.proc_grid))
        r = sj.timed_run(steps, 3)
        rf = sj.roofline(r, sm_mhz)
        sj.close()
        del sj
        gc.collect()
        return {"workload": WORKLOADS[wname]["desc"], "math_mode": math, "value": r["value"], "unit": UNIT, "steps": steps,
                "ms_per_step": r["elapsed_s"] / steps * 1e3, "roofline_frac": rf["frac"], "avg_launch_ms": rf["avg_launch_ms"],
                "fp64_frac": (rf.get("fp64") or {}).get("frac"), "kernel": rf["kernel"]}

    strict, other = None, []
    if not args.no_secondary:
        n_short = min(args.steps, 10)
        if args.math != "strict":
            strict = short_figure(args.workload, "strict", n_short)
            strict["note"] = "math_mode strict: bit-identical to the CPU oracle"
        if world == 1 and args.workload == DEFAULT_WORKLOAD:
            for wname in SECONDARY:
                other.append(short_figure(wname, args.math, n_short))
            other.append(short_figure("bizarrium_16384", "strict", min(n_short, 5)))
    if rank != 0:
        return 0

    # ---- CPU baseline beside it (rank 0, N=1 only, bounded sample) ----
    cpu = None
    if world == 1 and not args.no_cpu:
        n = tuple(config["per_gpu_grid"])
        cv, csecs, cthreads, ran = cpu_run(w["test"], n, 3, 1)
        cpu = {"value": cv, "unit": UNIT, "cores": cthreads, "kind": "port", "sample": cpu_sample_text(w["test"], n, ran, 3, 1)}

    st = run["state"]
    out = {
        "metric": METRIC, "value": run["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": run["elapsed_s"] / args.steps * 1e3, "higher_is_better": True, "scaling": w["scaling"],
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config,
        "math_mode": args.math,
        "timing": "CUDA events on the solver stream, max over ranks; W >= 3 warm-up cycles",
        "roofline": roofline,
        "cpu_baseline": cpu,
        "parity": parity,
        "strict_mode": strict,
        "other_configs": other,
        "e2e": e2e,
        "gpu_launches": run["launches"],
        "clocks": clocks,
        "host_wall_ms_per_step": run["host_s"] / args.steps * 1e3,
        "numa_local_cpus": numa_cpus,
        "final_state": {"cycle": int(st.cycle), "time": st.time, "dt": st.current_dt},
    }
    emit(out)
    return 0


_REAL_STDOUT = None


def _quiet_stdout():
    """Library chatter on fd 1 (NCCL prints its version banner there) goes to stderr: stdout carries the JSON line only."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(obj):
    line = json.dumps(obj)
    out = _REAL_STDOUT or sys.stdout
    out.write(line + "\n")
    out.flush()


def main():
    _quiet_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50,
                    help="cycles in the timed region (a real run is thousands of cycles; the one-off h2d/d2h of the fields in the "
                         "e2e leg is amortised over them)")
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS))
    ap.add_argument("--math", default="fast", choices=["strict", "fast", "ieee"],
                    help="fast: FMA + reciprocal division, the reference's own @fastmath latitude (default, within 1e-12 of "
                         "the oracle and of the golden data); strict: bit-exact vs the oracle; ieee: strict with nvcc's full division")
    ap.add_argument("--segment", type=int, default=0, help="march segment length (0 = auto)")
    ap.add_argument("--proc-grid", type=int, nargs=2, default=None)
    ap.add_argument("--parity-cycles", type=int, default=3, help="cycles of the in-bench oracle comparison (0 = skip)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline and parity legs")
    ap.add_argument("--no-secondary", "--no-strict", dest="no_secondary", action="store_true",
                    help="skip the secondary figures (strict mode, other single-GPU configurations)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3   # timing rule: W >= 3
    if args.impl == "reference":
        return run_reference_arm(args)
    return run_gpu_arm(args)


if __name__ == "__main__":
    sys.exit(main())
