#!/usr/bin/env python
"""Reduce an ncu report (.ncu-rep, `--set full`) to the metrics quoted in DESIGN.md / profiles/README.md.

  python profiles/summarize.py gpurun_out/prof.ncu-rep profiles/out_summary.csv \
         [--traffic profiles/sweep_traffic.json --key tiled_fast_pg --cells 268435456]
"""
import csv
import json
import subprocess
import sys

KEYS = [
    "Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.per_cycle_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second", "dram__bytes_write.sum.per_second",
    "smsp__inst_executed.sum", "sm__issue_active.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    stall = [h for h in hdr if "issue_stalled" in h and h.endswith("_per_issue_active.ratio")]
    cols = [k for k in KEYS if k in hdr] + stall
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(cols)
        w.writerow([units[hdr.index(c)] for c in cols])
        for r in rows[2:]:
            w.writerow([r[hdr.index(c)] for c in cols])
    if "--traffic" in sys.argv:
        d = dict(zip(hdr, rows[2]))
        u = dict(zip(hdr, units))
        scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
        rd = float(d["dram__bytes_read.sum"]) * scale[u["dram__bytes_read.sum"]]
        wr = float(d["dram__bytes_write.sum"]) * scale[u["dram__bytes_write.sum"]]
        path = sys.argv[sys.argv.index("--traffic") + 1]
        key = sys.argv[sys.argv.index("--key") + 1]        # <kernel>_<math>_<eos>, as bench.py looks it up
        cells = int(sys.argv[sys.argv.index("--cells") + 1])
        try:
            with open(path) as f:
                allk = json.load(f)
        except FileNotFoundError:
            allk = {}
        allk[key] = {"kernel": d["Kernel Name"], "grid": d["Grid Size"],
                     "duration_ms_under_ncu": float(d["gpu__time_duration.sum"]),
                     "dram_bytes_read": rd, "dram_bytes_write": wr, "bytes_per_launch": rd + wr, "cells_per_launch": cells,
                     "algorithmic_bytes_per_launch": 64 * cells, "traffic_over_algorithmic": (rd + wr) / (64 * cells),
                     "source": rep.split("/")[-1]}
        with open(path, "w") as f:
            json.dump(allk, f, indent=1)

if __name__ == "__main__":
    main()
