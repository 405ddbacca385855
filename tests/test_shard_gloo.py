"""world_size-2 gloo test of the row-sharded recall plumbing on CPU: each rank scores its shard with the
oracle (global row ids via row_base), ONE all-gather exchanges the per-shard top-k keys, every rank merges with the
same total order -> replica-identical result equal to the unsharded top-k."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    import socket
    with socket.socket() as sk:          # a port nobody holds right now (two tests, or two checkouts, on one machine)
        sk.bind(("127.0.0.1", 0))
        return sk.getsockname()[1]


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
    port = _free_port()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert all(ret[r] for r in range(world)) and len(ret) == world


# ---------------------------------------------------------------------------------------------------------------------
# The protocol bench.py runs at N > 1 (include/pairec_gpu.h "Row-sharded recall with ONE global threshold per query" and
# its all-to-all form), restated in numpy over the oracle's exact keys and driven through REAL exchanges between two gloo
# ranks: sample keys -> all-gather -> tau -> candidates that reach tau -> pack per owner -> ONE all-to-all -> owner check ->
# (agreed through one all-reduce) exact local lists when the check fails -> merge.  The owned queries' merged lists must
# equal the unsharded top-k — with a useful sample and with one the row order defeats (every rank then redoes the batch).
_STRIDE = 8          # the emulated sample: every 8th row of the shard (the kernels sample 1 tile in 128)


def _sample_len(k):
    # the r-th best of the gathered samples sits near rank r * _STRIDE of the whole catalog: 1.4 k rows are expected to reach
    # tau — at least k with a wide margin (else: redo), and per shard of a 2-rank run still below the list capacity k
    return (14 * k) // (10 * _STRIDE)


def _a2a_worker(rank, world, port, adversarial, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle
    rng = np.random.default_rng(21)
    n, d, B, k = 64000, 64, 3, 1000
    Bg, r = B * world, _sample_len(k)
    E = rng.standard_normal((n, d)).astype(np.float32)
    E[32000:32030] = E[:30]                                 # ties across the shard boundary
    Q = rng.standard_normal((Bg, d)).astype(np.float32)
    shard = n // world
    lo, hi = rank * shard, (n if rank == world - 1 else (rank + 1) * shard)
    if adversarial:                                          # the sampled rows score 0: tau says nothing about the rest
        E[np.arange(0, n, _STRIDE)] = 0.0
    local = oracle.recall_topk(E[lo:hi], Q, hi - lo, row_base=lo)           # [Bg, shard] every local key, sorted
    # 1. sample keys of this shard (sorted, 0-padded) -> all-gather #1
    samp = oracle.recall_topk(np.ascontiguousarray(E[lo:hi:_STRIDE]), Q, r)  # row ids are irrelevant for a threshold
    samp = np.where(samp != 0, (samp & np.uint64(0xFFFFFFFF00000000)), samp)  # keep the score bits only
    gathered = [torch.empty(Bg, r, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(gathered, torch.from_numpy(samp.view(np.int64).copy()))
    allsamp = np.concatenate([g.numpy().view(np.uint64) for g in gathered], axis=1)
    tau = np.sort(allsamp, axis=1)[:, ::-1][:, r - 1]        # r-th largest gathered sample key per query
    # 2. candidates: this shard's exact keys whose score bits reach tau; at most k, 0-padded; status = overflow
    cand = np.zeros((Bg, k), dtype=np.uint64)
    status = np.zeros(Bg, dtype=np.uint64)
    for q in range(Bg):
        reach = local[q][(local[q] | np.uint64(0xFFFFFFFF)) >= (tau[q] | np.uint64(0xFFFFFFFF))]
        status[q] = reach.size > k
        cand[q, :min(k, reach.size)] = reach[:k]
    # 3. pack per owner: chunk o = [B*k keys | B status words] of queries [o*B, o*B + B)  -> ONE all-to-all
    packed = np.concatenate([np.concatenate([cand[o * B:(o + 1) * B].ravel(), status[o * B:(o + 1) * B]]) for o in range(world)])
    recv = torch.empty(packed.size, dtype=torch.int64)
    dist.all_to_all_single(recv, torch.from_numpy(packed.view(np.int64).copy()))
    chunks = recv.numpy().view(np.uint64).reshape(world, B * k + B)
    # 4. owner check for the owned queries [rank*B, rank*B + B): a shard overflowed, or fewer than k keys reach tau
    retry = 0
    for ql in range(B):
        lists = chunks[:, ql * k:(ql + 1) * k]
        if chunks[:, B * k + ql].any() or np.count_nonzero(lists) < k:
            retry = 1
    flag = torch.tensor([retry])
    dist.all_reduce(flag, op=dist.ReduceOp.MAX)              # the ranks agree on a redo
    redone = bool(flag.item())
    if redone:                                               # exact protocol: per-shard top-k lists of every query
        exact = np.ascontiguousarray(local[:, :k])
        g2 = [torch.empty(Bg, k, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(g2, torch.from_numpy(exact.view(np.int64).copy()))
        keys = np.stack([g.numpy().view(np.uint64)[rank * B:(rank + 1) * B] for g in g2])
    else:
        keys = np.stack([chunks[g, :B * k].reshape(B, k) for g in range(world)])
    merged = oracle.merge_keys(keys)                         # [B, k]
    full = oracle.recall_topk(E, Q, k)
    ret[rank] = (bool((merged == full[rank * B:(rank + 1) * B]).all()), redone)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("adversarial", [False, True])
def test_global_threshold_all_to_all_protocol_two_ranks_gloo(adversarial):
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    port = _free_port()
    mp.spawn(_a2a_worker, args=(world, port, adversarial, ret), nprocs=world, join=True)
    assert len(ret) == world
    assert all(ret[r][0] for r in range(world)), "a rank's merged lists differ from the unsharded top-k"
    assert all(ret[r][1] == adversarial for r in range(world)), dict(ret)
