"""Data parallelism the B200 way: one process per GPU, batch sharded by image, per-GPU BN
statistics (exactly what the reference's single-process nn.DataParallel does, stack-hg.py:49 --
no SyncBN) and ONE all-reduce of the flat gradient buffer per step over NCCL / NVLink.  The
1/world_size factor is folded into the RMSprop kernel (`grad_scale`)."""
import os

import torch
import torch.distributed as dist_


def init_from_env(backend=None):
    """Initialise torch.distributed from the torchrun environment.  Returns (rank, world, local_rank)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist_.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist_.init_process_group(backend, rank=rank, world_size=world, device_id=torch.device("cuda", local))
        else:
            dist_.init_process_group(backend, rank=rank, world_size=world)
    return rank, world, local


def world_size():
    return dist_.get_world_size() if dist_.is_available() and dist_.is_initialized() else 1


def shard_range(n_total, rank, world):
    """Images [lo, hi) of a global batch owned by `rank` (equal shards; remainder to the first ranks)."""
    base, rem = divmod(n_total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def allreduce_flat_grads(flat_grad, async_op=False):
    """Sum the flat gradient buffer over all ranks (no-op for a single process)."""
    if world_size() == 1:
        return None
    return dist_.all_reduce(flat_grad, op=dist_.ReduceOp.SUM, async_op=async_op)


def broadcast_flat_params(flat, src=0):
    """Make every replica start from rank `src`'s parameters (DataParallel's per-forward replicate)."""
    if world_size() > 1:
        dist_.broadcast(flat, src=src)
