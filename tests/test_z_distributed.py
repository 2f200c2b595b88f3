"""One process per GPU: multi-rank tests (tests/dist_worker.py under torch.distributed.run).

CPU (gloo, world_size 2 and 4): host-side decomposition logic + decomposition invariance of the oracle.
GPU (NCCL): N-GPU fused result == 1-GPU fused result bit for bit, DebugIndexes halo test; skipped when the box has
fewer than 2 GPUs.
"""
import os
import socket
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WORKER = os.path.join(ROOT, "tests", "dist_worker.py")


def free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def launch(mode, world, extra=(), timeout=600):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(free_port()), WORKER, "--mode", mode,
           "--world", str(world), *extra]
    env = dict(os.environ, OMP_NUM_THREADS="1")
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, env=env, cwd=ROOT)
    if res.returncode != 0:
        lines = [ln for ln in res.stderr.splitlines() if "frame #" not in ln and "libtorch" not in ln]
        raise AssertionError("worker failed:\n" + res.stdout[-2000:] + "\n" + "\n".join(lines)[-8000:])
    assert f"dist_worker {mode} ok" in res.stdout
    return res


@pytest.mark.parametrize("world", [2, 4])
def test_gloo_decomposition_and_oracle_invariance(world):
    launch("cpu", world)


def test_process_grid_for():
    from armon_jl_b200.distributed import process_grid_for
    assert [process_grid_for(n) for n in (1, 2, 4, 8)] == [(1, 1), (1, 2), (2, 2), (2, 4)]
    with pytest.raises(ValueError):
        process_grid_for(6)


def _gpu_count():
    import armon_jl_b200 as armon
    return armon.device_count()


@pytest.mark.gpu
def test_nccl_two_gpus_equal_one_gpu_bitwise():
    n = _gpu_count()
    if n < 2:
        pytest.skip(f"needs >= 2 GPUs, found {n}")
    launch("gpu", 2, timeout=400)


@pytest.mark.gpu
def test_nccl_four_gpus_equal_one_gpu_bitwise():
    n = _gpu_count()
    if n < 4:
        pytest.skip(f"needs >= 4 GPUs, found {n}")
    launch("gpu", 4, extra=("--quick",), timeout=400)


@pytest.mark.gpu
def test_nccl_eight_gpus_equal_one_gpu_bitwise():
    n = _gpu_count()
    if n < 8:
        pytest.skip(f"needs 8 GPUs, found {n}")
    launch("gpu", 8, extra=("--quick",), timeout=400)
