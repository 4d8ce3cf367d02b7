"""N>1 host logic on CPU (gloo, world_size 2): image sharding, the single flat-gradient all-reduce and the
parameter broadcast used by the data-parallel step (one process per GPU; reference: nn.DataParallel,
stack-hg.py:49).  The GPU kernels are not involved."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pose_adv_aug_b200 import dist as hdist


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    r, w, l = hdist.init_from_env(backend="gloo")
    assert (r, w) == (rank, world) and hdist.world_size() == world
    # sharding: ranks own disjoint, covering, (almost) equal image ranges
    lo, hi = hdist.shard_range(48, rank, world)
    assert hi - lo == 24 and lo == 24 * rank
    lo2, hi2 = hdist.shard_range(7, rank, world)
    assert (hi2 - lo2) == (4 if rank == 0 else 3)
    # flat gradient all-reduce (sum) + 1/world folded later into the optimiser
    g = torch.full((1000,), float(rank + 1))
    hdist.allreduce_flat_grads(g)
    assert torch.equal(g, torch.full((1000,), 3.0))
    # broadcast: every replica starts from rank 0's parameters
    p = torch.arange(10, dtype=torch.float32) * (rank + 1)
    hdist.broadcast_flat_params(p, src=0)
    assert torch.equal(p, torch.arange(10, dtype=torch.float32))
    # mean-of-local-means == global mean for equal shards (the data-parallel loss gradient identity)
    x = torch.arange(48, dtype=torch.float64)
    local = x[lo:hi].mean().reshape(1)
    dist.all_reduce(local)
    assert abs(float(local) / world - float(x.mean())) < 1e-12
    # bucketed all-reduce (trainer.py: the flat gradient buffer is reduced bucket by bucket, back to front, as the backward
    # pass completes the buckets): slices of the flat buffer reduced in any order == one all-reduce of the whole buffer
    from pose_adv_aug_b200.trainer import bucket_splits
    offsets = [0, 128, 1024, 1152, 4096, 4224, 9000, 9128, 20000]
    numel = 24064
    splits = bucket_splits(offsets, numel, 4)
    assert splits == sorted(splits) and all(s_ in offsets for s_ in splits) and 1 <= len(splits) <= 3
    edges = [0] + splits + [numel]
    gen = torch.Generator().manual_seed(100 + rank)
    gflat = torch.randn(numel, generator=gen)
    whole = gflat.clone()
    hdist.allreduce_flat_grads(whole)
    for lo, hi in reversed(list(zip(edges[:-1], edges[1:]))):
        hdist.allreduce_flat_grads(gflat[lo:hi])
    assert torch.equal(gflat, whole)
    dist.barrier()
    dist.destroy_process_group()
    ret[rank] = 1


def test_two_rank_gloo_allreduce_and_sharding():
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, port, ret), nprocs=2, join=True)
    assert dict(ret) == {0: 1, 1: 1}


def test_single_process_is_noop():
    g = torch.ones(8)
    assert hdist.world_size() == 1
    assert hdist.allreduce_flat_grads(g) is None
    assert torch.equal(g, torch.ones(8))
    assert hdist.shard_range(24, 0, 1) == (0, 24)
