#!/usr/bin/env python
"""Static instruction budget of the steady-state loop of a sweep kernel (no GPU needed).

  python profiles/sass_histogram.py armon.jl_b200/build/sweep_fast_inst_tma_pg.o sweep_fast_kernelILi0ELi2ELi1ELi0ELi1E

Finds the innermost-but-largest backward branch of the function (the 4-step unrolled march loop), stops at the first
forward branch out of it (the staging flush that runs every second iteration) and prints the opcode histogram per march
step (= per row of 32 cells and warp)."""
import collections
import re
import subprocess
import sys


def histogram(obj, pat):
    """Opcode histogram of the steady-state march loop of the kernel whose mangled name contains `pat`: the loop is the
    backward branch with the largest span; the transposed-store flush inside it (runs once per chunk of 8 steps, found
    by its 128-bit shared loads) is left out and reported separately."""
    out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True, check=True).stdout
    lines, on = [], False
    for ln in out.splitlines():
        if "Function :" in ln:
            on = pat in ln
        if on:
            m = re.match(r"\s*/\*([0-9a-f]{4,5})\*/\s+(.*?);", ln)
            if m:
                lines.append((int(m.group(1), 16), m.group(2).strip()))
    bra = re.compile(r"BRA(?:\.U)?(?:\.ANY)?\s+(?:!?U?P\w+,\s*)?0x([0-9a-f]+)")
    back = [(a, int(t.group(1), 16)) for a, i in lines if (t := bra.search(i)) and int(t.group(1), 16) < a]
    end, start = max(back, key=lambda x: x[0] - x[1])
    body = [(a, i) for a, i in lines if start <= a <= end]
    # flush region: from the forward branch that skips it (the last one before the first LDS.128) to that branch's target
    lds128 = [a for a, i in body if "LDS.128" in i]
    skip = (0, 0)
    if lds128:
        fwd = [(a, int(t.group(1), 16)) for a, i in body if (t := bra.search(i)) and max(lds128) < int(t.group(1), 16) <= end and a < min(lds128)]
        if fwd:
            skip = max(fwd, key=lambda x: x[1])   # the branch that skips the whole flush, slow path included
    c = collections.Counter()
    for a, i in body:
        if skip[0] < a < skip[1]:
            continue
        op = re.sub(r"^@!?U?P\w+\s+", "", i).split()[0]
        op = op if op.startswith(("IMAD.MOV", "MUFU")) else op.split(".")[0]
        c[op] += 1
    stop = end
    tot = sum(c.values())
    fp64 = sum(v for k, v in c.items() if k in ("DFMA", "DMUL", "DADD", "DSETP"))
    return start, stop, tot, fp64, c


# kernels of the product path (GAD + minmod + euler_2nd, transposed output): key -> (object, mangled-name pattern)
PRODUCT = {
    "tma_fast_pg": ("sweep_fast_inst_tma_pg.o", "sweep_fast_kernelILi0ELi2ELi1ELi0ELi1E"),
    "tma_fast_biz": ("sweep_fast_inst_tma_biz.o", "sweep_fast_kernelILi0ELi2ELi1ELi1ELi1E"),
    "tiled_fast_pg": ("sweep_fast_inst_tiled_pg.o", "sweep_fast_kernelILi0ELi2ELi1ELi0ELi1ELi0ELi1E"),
    "tiled_fast_biz": ("sweep_fast_inst_tiled_biz.o", "sweep_fast_kernelILi0ELi2ELi1ELi1ELi1ELi0ELi1E"),
    "async2_fast_pg": ("sweep_fast_inst_cpa16_pg.o", "sweep_fast_kernelILi1ELi2ELi1ELi0ELi1E"),
    "async2_fast_biz": ("sweep_fast_inst_cpa16_biz.o", "sweep_fast_kernelILi1ELi2ELi1ELi1ELi1E"),
    # strict arithmetic on the four-chain schedule; the benched grids have power-of-two cell sizes (the _dxp instantiations)
    "chains_strict_pg": ("sweep_fast_inst_strict_dxp_pg.o", "sweep_fast_kernelILi0ELi2ELi1ELi0ELi1E"),
    "chains_strict_biz": ("sweep_fast_inst_strict_dxp_biz.o", "sweep_fast_kernelILi0ELi2ELi1ELi1ELi1E"),
}


def write_budget(build_dir, path):
    """profiles/sass_budget.json: FP64 and total warp instructions per march step of the product kernels (read by
    bench.py for the FP64 roof)."""
    import json
    import os
    out = {"_doc": "static SASS count of the steady-state march loop per march step (= per cell and thread; per row of 32 "
                   "cells and warp), python profiles/sass_histogram.py --budget armon.jl_b200/build profiles/sass_budget.json"}
    for key, (obj, pat) in PRODUCT.items():
        start, stop, tot, fp64, c = histogram(os.path.join(build_dir, obj), pat)
        out[key] = {"instr_per_step": tot / 4, "fp64_per_step": fp64 / 4, "mufu_per_step": sum(v for k, v in c.items() if k.startswith("MUFU")) / 4,
                    "kernel": pat}
    with open(path, "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out, indent=1))


def main():
    if sys.argv[1] == "--budget":
        return write_budget(sys.argv[2], sys.argv[3])
    obj, pat = sys.argv[1], sys.argv[2]
    start, stop, tot, fp64, c = histogram(obj, pat)
    print(f"loop 0x{start:x}..0x{stop:x}: {tot} instructions per 4 steps = {tot / 4:.1f} per step, FP64 {fp64 / 4:.1f} per step")
    for k, v in c.most_common():
        print(f"  {k:14s} {v / 4:6.2f}")


if __name__ == "__main__":
    main()
