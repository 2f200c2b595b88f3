"""How far do two correct implementations of the same scheme drift apart over a full-length run?  (one-off measurement,
not collected by pytest; output committed as profiles/r2_drift.txt)

  python tests/drift_study.py [N ...]

For each grid the case runs to its own maxtime four times: the B200 backend in `strict` and `fast` arithmetic, and the
CPU oracle in its `strict` (IEEE, reference operation order) and `fma` (contracted, what a @fastmath build is allowed
to do) flavours.  Reported: max|a - b| / max|b| per field for
  * GPU fast   vs GPU strict   -- the drift of the benched mode,
  * CPU fma    vs CPU strict   -- the drift the reference's own CPU path has between an IEEE and a fastmath build,
  * GPU strict vs CPU strict   -- must be exactly 0 (bit parity).
The two drifts being of the same order shows that the distance between the fast and the strict mode after thousands of
cycles is the sensitivity of the flow (rounding noise amplified along the contact discontinuity), not an error of the
fast kernel.
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]

import numpy as np

import armon_jl_b200 as armon
from helpers import reference_params, scaled_max_diff
from oracle import OracleSolver

FIELDS = ("rho", "u", "v", "E")


def gpu(test, n, mode):
    p = reference_params(test, N=(n, n), maxcycle=10**6, math_mode=mode, bind_pcg=False)
    g = armon.BlockGrid(p)
    armon.init_test(p, g)
    armon.time_loop(p, g)
    out = {v: g.real(v).copy() for v in FIELDS}
    cyc = g.time_state().cycle
    g.close()
    return out, cyc


def cpu(test, n, flavour):
    o = OracleSolver(reference_params(test, N=(n, n), maxcycle=10**6), flavour, nthreads=os.cpu_count() or 1)
    o.time_loop()
    out = {v: o.real(v).copy() for v in FIELDS}
    cyc = o.state.cycle
    o.close()
    return out, cyc


def dist(a, b):
    return ", ".join(f"{v}={scaled_max_diff(a[v], b[v]):.2e}" for v in FIELDS)


def main():
    test = "Sod_circ"
    for n in [int(x) for x in sys.argv[1:]] or [512, 1024, 2048]:
        t0 = time.time()
        gs, c1 = gpu(test, n, "strict")
        gf, c2 = gpu(test, n, "fast")
        cs, c3 = cpu(test, n, "strict")
        cf, c4 = cpu(test, n, "fma")
        print(f"{test} {n}x{n}: cycles gpu strict/fast {c1}/{c2}, cpu strict/fma {c3}/{c4}  ({time.time() - t0:.0f} s)")
        print(f"   GPU fast   vs GPU strict: {dist(gf, gs)}")
        print(f"   CPU fma    vs CPU strict: {dist(cf, cs)}")
        print(f"   GPU strict vs CPU strict: {dist(gs, cs)}   bit-equal: {all(np.array_equal(gs[v], cs[v]) for v in FIELDS)}")
        print(f"   GPU fast   vs CPU fma   : {dist(gf, cf)}", flush=True)


if __name__ == "__main__":
    main()
