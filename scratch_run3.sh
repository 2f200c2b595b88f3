timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r1d_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r1d_pytest_gpu.log
for v in single async tma; do for m in fast strict; do
ARMON_B200_KERNEL=$v timeout 300 python bench.py --steps 10 --warmup 3 --math $m --no-cpu > gpurun_out/r1d_bench_${v}_${m}.json 2> gpurun_out/r1d_bench_${v}_${m}.err; echo "bench $v $m rc=$?"
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r1d_bench_${v}_${m}.json"))
    print("$v $m", round(d["value"],2), "Gc/s", round(d["roofline"]["avg_launch_ms"],3), "ms/sweep frac", round(d["roofline"]["frac"],3), d["gpu_launches"])
except Exception as e: print("$v $m failed", e)
PY
done; done
