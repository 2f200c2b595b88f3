#!/usr/bin/env python
"""bench.py -- giga cell-updates/s of the fused Lagrange+remap time step on B200, one JSON line.

  python bench.py --gpus N --steps K --warmup W [--workload NAME] [--math strict|fast] [--impl reference]

A "step" is one solver cycle (all axis sweeps of the splitting + the CFL time-step update) over the whole grid.
N=1 runs BASELINE.json's configs[1] (Sod_circ 8192x8192, GAD + minmod + euler_2nd, Float64); N>1 keeps 8192x8192 cells
per GPU (weak scaling, Cartesian process grid, NCCL halo exchange + dt all-reduce).  `value` is timed with the state
resident in HBM; `e2e` is the same job through the public API with host buffers (h2d of the initial state, a
blocking read of the time-step state every cycle, d2h of the final fields).  `--impl reference` times the restated
reference CPU path (oracle/, C + OpenMP, all host cores) -- the reference itself is Julia and cannot run here.
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

WORKLOADS = {
    # name: (test, per-GPU or global N, scaling, description)
    "sod_circ_8192": dict(test="Sod_circ", n=(8192, 8192), scaling="weak",
                          desc="Sod_circ 8192x8192 per GPU, GAD+minmod+euler_2nd, Float64 (BASELINE configs[1])"),
    "bizarrium_16384": dict(test="Bizarrium", n=(16384, 16384), scaling="weak",
                            desc="Bizarrium 16384x16384 per GPU, GAD+minmod+euler_2nd (BASELINE configs[2])"),
    "sedov_32768_strong": dict(test="Sedov", n=(32768, 32768), scaling="strong",
                               desc="Sedov 32768x32768 global, strong scaling (BASELINE configs[3])"),
    "sod_circ_weak_16384": dict(test="Sod_circ", n=(16384, 16384), scaling="weak",
                                desc="Sod_circ 16384x16384 per GPU weak scaling (BASELINE configs[4])"),
    "sod_circ_1024": dict(test="Sod_circ", n=(1024, 1024), scaling="weak", desc="small smoke workload"),
}
CPU_SAMPLE_N = (4096, 4096)


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum per sweep launch from the committed ncu capture, if any."""
    try:
        with open(os.path.join(ROOT, "profiles", "sweep_traffic.json")) as f:
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


# -------------------------------------------------------------------------------------------------------------
# CPU arm: the restated reference CPU path (oracle/), all host threads
# -------------------------------------------------------------------------------------------------------------
def cpu_run(test, n, steps, warmup, nthreads=0):
    import armon_jl_b200 as armon
    import oracle
    oracle.build()
    lib = oracle.load("fast")
    threads = nthreads or (os.cpu_count() or 1)
    params = armon.ArmonParameters(test=test, N=n, maxcycle=10**9, **scheme_kwargs())
    s = oracle.OracleSolver(params, "fast", nthreads=threads)
    for _ in range(warmup):
        s.solver_cycle()
    t0 = time.perf_counter()
    for _ in range(steps):
        s.solver_cycle()
    dt = time.perf_counter() - t0
    s.close()
    return n[0] * n[1] * steps / dt / 1e9, dt, threads, lib.orc_num_threads()


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return 0
    w = WORKLOADS[args.workload]
    value, secs, threads, _ = cpu_run(w["test"], CPU_SAMPLE_N, args.steps, args.warmup)
    sample = (f"{w['test']} {CPU_SAMPLE_N[0]}x{CPU_SAMPLE_N[1]} sub-sample of the workload grid, {args.steps} cycles "
              f"after {args.warmup} warm-up, same scheme; restated reference CPU path (C + OpenMP, -O3 -ffast-math), "
              "not the Julia code (no julia in this image)")
    out = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": secs / args.steps * 1e3, "higher_is_better": True, "scaling": w["scaling"],
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": w["desc"], "sample": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(out)
    return 0


# -------------------------------------------------------------------------------------------------------------
# GPU arm
# -------------------------------------------------------------------------------------------------------------
def run_gpu_arm(args):
    import numpy as np
    import armon_jl_b200 as armon
    from armon_jl_b200 import distributed as adist
    from armon_jl_b200.backend import check

    world = int(os.environ.get("WORLD_SIZE", 1))
    rank = int(os.environ.get("RANK", 0))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torchrun for --gpus > 1 (one process per GPU)")
    multi = world > 1
    if multi:
        import torch  # noqa: F401  (torch's NCCL is loaded before the library links against libnccl.so.2)
        adist.init_process_group("nccl")

    w = WORKLOADS[args.workload]
    P = adist.process_grid_for(world) if not args.proc_grid else tuple(args.proc_grid)
    if w["scaling"] == "weak":
        global_n = (w["n"][0] * P[0], w["n"][1] * P[1])
    else:
        global_n = w["n"]
    params = armon.ArmonParameters(test=w["test"], N=global_n, use_MPI=multi, P=P, rank=rank, proc_size=world,
                                   maxcycle=10**9, math_mode=args.math, march_segment=args.segment, bind_pcg=False,
                                   device_id=local_rank, return_data=True, **scheme_kwargs())
    grid = armon.BlockGrid(params)
    lib, solver, dev = grid.lib, None, grid.device
    armon.init_test(params, grid)
    solver = grid.solver
    local_cells = params.N[0] * params.N[1]
    global_cells = params.global_grid[0] * params.global_grid[1]

    def sync_all():
        dev.wait()
        if multi:
            adist.barrier()

    # ---- warm-up ----
    check(lib.armon_solver_run(solver, args.warmup))
    sync_all()

    # ---- timed region: K cycles, state resident in HBM ----
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
    launches0 = dev.launch_count()
    sync_all()
    check(lib.armon_solver_profile(solver, 1))
    t_host0 = time.perf_counter()
    check(lib.armon_solver_run(solver, args.steps))
    ms = C.c_float()
    check(lib.armon_solver_elapsed_ms(solver, C.byref(ms)))
    sync_all()
    t_host = time.perf_counter() - t_host0
    clocks = sampler.stop() if rank == 0 else None
    launches = dev.launch_count() - launches0
    sweep_ms, sweep_n = C.c_double(), C.c_uint64()
    check(lib.armon_solver_sweep_time_ms(solver, C.byref(sweep_ms), C.byref(sweep_n)))
    check(lib.armon_solver_profile(solver, 0))
    elapsed_s = ms.value / 1e3
    if multi:
        elapsed_s = adist.allreduce_max(elapsed_s)
    st = grid.time_state()
    if st.error or st.done:
        raise SystemExit(f"bench invalid: solver stopped early (error={st.error}, done={st.done}, cycle={st.cycle})")
    value = global_cells * args.steps / elapsed_s / 1e9

    # ---- e2e: same job through the public API with HOST buffers ----
    import torch
    names = ("rho", "u", "v", "E")
    host0 = {}
    armon.init_test(params, grid)          # fresh initial state on the device -> host copy (untimed set-up)
    pinned = {k: torch.empty(grid.cell_count, dtype=torch.float64, pin_memory=True) for k in names}
    for k in names:
        getattr(grid.device_data, k).copy_to_host(pinned[k].numpy())
        host0[k] = pinned[k].numpy()
    out_pinned = {k: torch.empty(grid.cell_count, dtype=torch.float64, pin_memory=True) for k in names}
    e2e_steps = args.steps
    sync_all()
    t0 = time.perf_counter()
    for k in names:                        # h2d of the job's input
        getattr(grid.device_data, k).copy_from_host(host0[k])
    check(lib.armon_solver_reset(solver))
    grid._fused_dirty = False
    last = None
    for _ in range(e2e_steps):             # one public-API call + a blocking read of the step's result per cycle
        armon.solver_cycle(params, grid)
        last = grid.time_state()
    grid.finalize()
    for k in names:                        # d2h of the job's result
        getattr(grid.device_data, k).copy_to_host(out_pinned[k].numpy())
    dev.wait()
    e2e_s = time.perf_counter() - t0
    if multi:
        adist.barrier()
        e2e_s = adist.allreduce_max(e2e_s)
    e2e_value = global_cells * e2e_steps / e2e_s / 1e9
    h2d = 4 * grid.cell_count * 8 / e2e_steps
    d2h = (4 * grid.cell_count * 8 + e2e_steps * C.sizeof(armon.backend.armon_time_state)) / e2e_steps
    checksum = float(out_pinned["rho"].numpy().sum())
    assert np.isfinite(checksum) and last.cycle == e2e_steps

    # ---- roofline of the dominant kernel (the sweep) ----
    peak, peak_src = measured_peak()
    avg_sweep_s = sweep_ms.value / max(sweep_n.value, 1) / 1e3
    achieved = BYTES_PER_CELL_SWEEP * local_cells / avg_sweep_s / 1e9
    traffic = ncu_traffic()
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                # ncu dram__bytes_read.sum + dram__bytes_write.sum of one launch (profiles/sweep_traffic.json), only quoted
                # for the workload it was captured on
                "traffic": traffic.get("bytes_per_launch") if traffic and traffic.get("cells_per_launch") == local_cells
                and w["test"] == "Sod_circ" else None,
                "peak_source": peak_src, "kernel": f"sweep_{os.environ.get('ARMON_B200_KERNEL', 'async2')}_kernel<{args.math}, GAD+minmod, euler_2nd, "
                          f"{'bizarrium' if w['test'] == 'Bizarrium' else 'perfect gas'}>",
                "avg_launch_ms": avg_sweep_s * 1e3, "launches_timed": int(sweep_n.value),
                "algorithmic_bytes_per_launch": BYTES_PER_CELL_SWEEP * local_cells,
                "sweep_share_of_step": sweep_ms.value / 1e3 / (ms.value / 1e3)}

    grid.close()

    # ---- secondary figure: the bit-exact (strict) arithmetic mode, same workload, short run ----
    strict = None
    if args.math != "strict" and not args.no_strict:
        sp = armon.ArmonParameters(test=w["test"], N=global_n, use_MPI=multi, P=P, rank=rank, proc_size=world,
                                   maxcycle=10**9, math_mode="strict", march_segment=args.segment, bind_pcg=False,
                                   device_id=local_rank, return_data=True, **scheme_kwargs())
        sg = armon.BlockGrid(sp)
        armon.init_test(sp, sg)
        n_strict = min(args.steps, 10)
        check(lib.armon_solver_run(sg.solver, 3))
        dev2 = sg.device
        dev2.wait()
        if multi:
            adist.barrier()
        check(lib.armon_solver_profile(sg.solver, 1))
        check(lib.armon_solver_run(sg.solver, n_strict))
        sms = C.c_float()
        check(lib.armon_solver_elapsed_ms(sg.solver, C.byref(sms)))
        s_sweep_ms, s_sweep_n = C.c_double(), C.c_uint64()
        check(lib.armon_solver_sweep_time_ms(sg.solver, C.byref(s_sweep_ms), C.byref(s_sweep_n)))
        s_el = sms.value / 1e3
        if multi:
            s_el = adist.allreduce_max(s_el)
        s_avg = s_sweep_ms.value / max(s_sweep_n.value, 1) / 1e3
        strict = {"value": global_cells * n_strict / s_el / 1e9, "unit": UNIT, "steps": n_strict,
                  "roofline_frac": BYTES_PER_CELL_SWEEP * local_cells / s_avg / 1e9 / peak,
                  "avg_launch_ms": s_avg * 1e3, "note": "math_mode strict: bit-identical to the CPU oracle"}
        sg.close()
    if rank != 0:
        return 0

    # ---- CPU baseline beside it (rank 0, N=1 only, bounded sample) ----
    cpu = None
    if world == 1 and not args.no_cpu:
        cv, csecs, cthreads, _ = cpu_run(w["test"], CPU_SAMPLE_N, 3, 1)
        cpu = {"value": cv, "unit": UNIT, "cores": cthreads, "kind": "port",
               "sample": f"{w['test']} {CPU_SAMPLE_N[0]}x{CPU_SAMPLE_N[1]}, 3 cycles after 1 warm-up, same scheme; "
                         "restated reference CPU path (oracle/, C + OpenMP -O3 -ffast-math); the Julia reference "
                         "cannot run here"}

    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": elapsed_s / args.steps * 1e3, "higher_is_better": True, "scaling": w["scaling"],
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": w["desc"], "global_grid": list(params.global_grid), "per_gpu_grid": list(params.N),
                   "process_grid": list(P), "math_mode": args.math, "nghost": 4, "cfl": params.cfl,
                   "l2_policy": "inputs larger than L2 (8 arrays x %.0f MB per GPU vs 126 MB L2)" % (grid.cell_count * 8 / 1e6)
                   if grid.cell_count * 8 * 8 > 4 * 126e6 else "inputs fit in L2 (small workload)",
                   "timing": "CUDA events on the solver stream, max over ranks"},
        "roofline": roofline,
        "cpu_baseline": cpu,
        "strict_mode": strict,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "steps": e2e_steps, "note": "h2d of rho,u,v,E from pinned host memory + K x (solver_cycle + blocking read "
                "of the time-step state) + finalize + d2h of rho,u,v,E, wall clock, max over ranks"},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "host_wall_ms_per_step": t_host / args.steps * 1e3,
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
    ap.add_argument("--steps", type=int, default=200,
                    help="cycles in the timed region (a real run is thousands of cycles; the one-off h2d/d2h of the fields in the "
                         "e2e leg is amortised over them)")
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="sod_circ_8192", choices=sorted(WORKLOADS))
    ap.add_argument("--math", default="fast", choices=["strict", "fast", "ieee"],
                    help="fast: FMA + reciprocal division, the reference's own @fastmath latitude (default, within 1e-12 of "
                         "the golden data); strict: bit-exact vs the oracle; ieee: strict with nvcc's full division")
    ap.add_argument("--segment", type=int, default=0, help="march segment length (0 = auto)")
    ap.add_argument("--proc-grid", type=int, nargs=2, default=None)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-strict", action="store_true", help="skip the secondary strict-mode measurement")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3   # timing rule: W >= 3
    if args.impl == "reference":
        return run_reference_arm(args)
    return run_gpu_arm(args)


if __name__ == "__main__":
    sys.exit(main())
