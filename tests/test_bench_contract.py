"""bench.py contract checks that need no GPU: the reference arm (`--impl reference`) prints exactly one JSON line on
stdout with the keys the driver reads, and non-zero ranks of a multi-rank launch stay silent."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_bench(*args, env=None):
    e = dict(os.environ, OMP_NUM_THREADS=str(min(8, os.cpu_count() or 1)))
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True,
                          timeout=600, env=e, cwd=ROOT)


def test_reference_arm_prints_one_json_line():
    res = run_bench("--impl", "reference", "--steps", "1", "--warmup", "0")
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [ln for ln in res.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, res.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "Gcell-updates/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["steps"] == 1 and d["dtype"] == "f64" and d["data"] == "synthetic"
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]


def test_reference_arm_other_ranks_are_silent():
    res = run_bench("--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0", env={"RANK": "1", "WORLD_SIZE": "2"})
    assert res.returncode == 0 and res.stdout.strip() == ""


def test_every_kernel_translation_unit_is_built_and_profiled():
    """csrc/Makefile lists exactly the .cu files of csrc/ (a translation unit left out would drop its kernel table from the
    library; a stale entry breaks the build), and profiles/sass_budget.json / sweep_traffic.json carry the keys bench.py
    looks up for the kernels of the product path (roofline.fp64, roofline.traffic)."""
    import json
    import re
    csrc = os.path.join(ROOT, "armon.jl_b200", "csrc")
    with open(os.path.join(csrc, "Makefile")) as f:
        mk = f.read()
    srcs = set(re.findall(r"[\w]+\.cu\b", mk.split("SRCS")[1].split("OBJS")[0]))
    on_disk = {n for n in os.listdir(csrc) if n.endswith(".cu")}
    assert srcs == on_disk, (sorted(srcs - on_disk), sorted(on_disk - srcs))
    with open(os.path.join(ROOT, "profiles", "sass_budget.json")) as f:
        budget = json.load(f)
    for key in ("tiled_fast_pg", "tiled_fast_biz", "chains_strict_pg", "chains_strict_biz"):
        assert budget[key]["fp64_per_step"] > 0 and budget[key]["instr_per_step"] > budget[key]["fp64_per_step"], key
    with open(os.path.join(ROOT, "profiles", "sweep_traffic.json")) as f:
        traffic = json.load(f)
    for key in ("tiled_fast_pg", "chains_strict_pg"):
        t = traffic[key]
        assert t["cells_per_launch"] == 16384 * 16384 and 1.0 <= t["traffic_over_algorithmic"] < 1.05, (key, t)
