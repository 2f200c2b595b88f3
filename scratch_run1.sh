set -x
nvidia-smi -L
python -m pytest tests -m gpu -x -q > gpurun_out/r1b_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r1b_pytest_gpu.log
python bench.py --steps 20 --warmup 3 --math fast > gpurun_out/r1b_bench_fast_n1.json 2> gpurun_out/r1b_bench_fast_n1.err; echo "bench rc=$?"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 3 --math fast > gpurun_out/r1b_bench_fast_n2.json 2> gpurun_out/r1b_bench_fast_n2.err; echo "bench2 rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r1b_bench_ref.json 2> gpurun_out/r1b_bench_ref.err; echo "ref rc=$?"
cat gpurun_out/r1b_bench_fast_n1.json gpurun_out/r1b_bench_fast_n2.json gpurun_out/r1b_bench_ref.json | cut -c1-600
nproc
