"""world_size-2 gloo test of the row-sharded recall plumbing on CPU: each rank scores its shard with the
oracle (global row ids via row_base), ONE all-gather exchanges the per-shard top-k keys, every rank merges with the
same total order -> replica-identical result equal to the unsharded top-k."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle
    rng = np.random.default_rng(9)
    n, d, B, k = 6001, 64, 5, 37
    E = rng.standard_normal((n, d)).astype(np.float32)
    E[3000:3050] = E[:50]           # ties across the shard boundary
    Q = rng.standard_normal((B, d)).astype(np.float32)
    shard = n // world
    lo = rank * shard
    hi = n if rank == world - 1 else lo + shard
    local = oracle.recall_topk(E[lo:hi], Q, k, row_base=lo)
    t_local = torch.from_numpy(local.view(np.int64))
    gathered = [torch.empty_like(t_local) for _ in range(world)]
    dist.all_gather(gathered, t_local)
    keys = np.stack([g.numpy().view(np.uint64) for g in gathered])
    merged = oracle.merge_keys(keys)
    full = oracle.recall_topk(E, Q, k)
    ret[rank] = bool((merged == full).all())
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_recall_two_ranks_gloo():
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert all(ret[r] for r in range(world)) and len(ret) == world
