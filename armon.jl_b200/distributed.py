"""One process per GPU: the plumbing that replaces `init_MPI` (src/parameters.jl:408-467).

`torch.distributed` (NCCL on the GPU box, gloo in the CPU tests) bootstraps the ranks -- it plays the role of
`MPI.Init` / `MPI.bcast`: rank 0 creates the NCCL unique id through the library, every rank receives it and hands it
to `armon_ctx_comm_init`.  The data path itself (halo send/recv, dt all-reduce) is NCCL inside libarmon_b200.so.
"""
import os

_pg = None


def is_initialized():
    import torch.distributed as dist
    return dist.is_available() and dist.is_initialized()


def init_process_group(backend=None, timeout_s=None):
    """Rendezvous from the torchrun environment (RANK, WORLD_SIZE, MASTER_ADDR, MASTER_PORT)."""
    import torch
    import torch.distributed as dist
    if dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    if backend is None:
        backend = "nccl" if torch.cuda.is_available() else "gloo"
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29511")
    kwargs = {}
    if timeout_s is not None:
        import datetime
        kwargs["timeout"] = datetime.timedelta(seconds=float(timeout_s))
    if backend == "nccl":
        local = int(os.environ.get("LOCAL_RANK", 0))
        torch.cuda.set_device(local)
        kwargs["device_id"] = torch.device("cuda", local)
    dist.init_process_group(backend=backend, **kwargs)
    return dist.get_rank(), dist.get_world_size()


def broadcast_bytes(payload, src=0):
    """Broadcast a bytes object from `src` (stands for MPI.bcast of the NCCL unique id)."""
    import torch.distributed as dist
    obj = [payload if dist.get_rank() == src else None]
    dist.broadcast_object_list(obj, src=src)
    return obj[0]


def setup_device_comm(params, device):
    """Create the NCCL communicator of the library context for the process grid of `params`."""
    if not params.use_MPI or params.proc_size == 1:
        return
    init_process_group()
    uid = device.unique_id() if params.rank == 0 else None
    uid = broadcast_bytes(uid, src=0)
    device.comm_init(uid, params.rank, params.proc_size)


def process_grid_for(n_ranks, prefer="y"):
    """Default Cartesian grid P=(px, py) for n ranks: cuts along Y first (contiguous halo rows in the canonical
    layout), SURVEY.md section 8e.  1 -> (1,1), 2 -> (1,2), 4 -> (2,2), 8 -> (2,4)."""
    px, py = 1, 1
    n = int(n_ranks)
    turn = 0
    while n > 1:
        if n % 2:
            solver_error_msg = f"cannot build a default process grid for {n_ranks} ranks"
            raise ValueError(solver_error_msg)
        if turn % 2 == 0:
            py *= 2
        else:
            px *= 2
        n //= 2
        turn += 1
    return (px, py) if prefer == "y" else (py, px)


def allreduce_sum(values):
    """MPI.Allreduce(SUM) of a few host scalars (conservation_vars, src/reductions.jl:317-320)."""
    import torch
    import torch.distributed as dist
    t = torch.tensor(list(values), dtype=torch.float64)
    if dist.get_backend() == "nccl":
        t = t.cuda()
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return tuple(t.cpu().tolist())


def allreduce_max(value):
    import torch
    import torch.distributed as dist
    t = torch.tensor([float(value)], dtype=torch.float64)
    if dist.get_backend() == "nccl":
        t = t.cuda()
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.cpu()[0])


def barrier():
    import torch.distributed as dist
    if dist.is_initialized():
        dist.barrier()
