"""Convert the reference's golden CSVs into compact fixtures (run once, in the build container).

Source: /root/reference/test/reference_data/ref_<Test>_64bits.csv -- written by
test/reference_data/create_references.jl:8-15 with the parameters of
test/reference_data/reference_functions.jl:7-19 (GAD + minmod + euler_2nd, nghost=4, N=(100,100),
Sequential splitting, default cfl/maxtime, maxcycle=1000, use_threading=false, use_simd=false).
Format (src/io.jl:4-27, reference_functions.jl:41): line 1 = "%#.15g, %d" -> final dt, cycles; then one
line per real cell, X fastest, "x, y, rho, u, v, p" in "%#24.17e" (exact decimal round trip), blank line
after each grid row.

The fixtures hold DATA only (no reference source).  /root/reference does not exist on the GPU box, so the
tests read these .npz files instead.  Float parsing with Python's float() is exact for 17 significant digits.
"""
import os
import sys
import numpy as np

REF_DIR = "/root/reference/test/reference_data"
OUT_DIR = os.path.dirname(os.path.abspath(__file__))
TESTS = ("Sod", "Sod_y", "Sod_circ", "Bizarrium", "Sedov")
N = 100


def convert(test, bits=64):
    path = os.path.join(REF_DIR, f"ref_{test}_{bits}bits.csv")
    with open(path) as f:
        header = f.readline().strip()
        dt_str, cycles_str = [s.strip() for s in header.split(",")]
        rows = [ln for ln in f if ln.strip()]
    assert len(rows) == N * N, (test, len(rows))
    data = np.array([[float(tok) for tok in ln.split(",")] for ln in rows], dtype=np.float64)
    assert data.shape == (N * N, 6)
    fields = {name: data[:, k].reshape(N, N) for k, name in enumerate(("x", "y", "rho", "u", "v", "p"))}
    if bits == 32:   # "%#16.9e": exact decimal round trip of Float32 values (src/parameters.jl:708-710)
        fields = {k: v.astype(np.float32) for k, v in fields.items()}
    np.savez_compressed(
        os.path.join(OUT_DIR, f"ref_{test}_{bits}bits.npz"),
        dt=np.float64(float(dt_str)), dt_str=np.array(dt_str), cycles=np.int64(int(cycles_str)), **fields)
    print(test, dt_str, cycles_str, {k: float(np.abs(v).max()) for k, v in fields.items()})


if __name__ == "__main__":
    if not os.path.isdir(REF_DIR):
        sys.exit("reference data not found (this script only runs in the build container)")
    for t in TESTS:
        convert(t, 64)
        convert(t, 32)
