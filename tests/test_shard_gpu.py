"""Row-sharded path on ONE GPU: two engines hold the two shards (row_base 0 / N/2), their local top-k key
lists are concatenated exactly as the all-gather would lay them out, and prg_merge_keys / prg_recommend_from_keys must
give the unsharded result.  (bench.py --gpus N runs the same entry points with torch.distributed/NCCL in between.)"""
import ctypes as C

import numpy as np
import pytest

from tests import synth

pytestmark = pytest.mark.gpu


def test_two_shards_merge_to_unsharded_result(oracle_lib):
    import torch
    from pairec_b200 import DppParams, Engine
    from pairec_b200.binding import MEM_DEVICE, MODEL_FM
    n, d, B, k, T = 600_000, 64, 70, 300, 15     # B = 70: two query blocks of <= 64
    rng = np.random.default_rng(31)
    E = (rng.standard_normal((n, d)) / 8).astype(np.float32)
    E[n // 2:n // 2 + 500] = E[:500]              # ties across the shard boundary
    Q = (rng.standard_normal((B, d)) / 8).astype(np.float32)
    dev = torch.device("cuda:0")
    half = n // 2
    engs = [Engine(0), Engine(0)]
    try:
        engs[0].set_item_matrix(E[:half], row_base=0)
        engs[1].set_item_matrix(E[half:], row_base=half)
        q_dev = torch.from_numpy(Q).to(dev)
        keys_all = torch.zeros(2, B, k, dtype=torch.int64, device=dev)       # the all-gather output layout [G][B][k]
        for g in range(2):
            engs[g].recall_local_keys_dev(q_dev.data_ptr(), B, k, keys_all[g].data_ptr())
            engs[g].sync()
        full = oracle_lib.recall_topk(E, Q, k)
        # per-shard lists equal the oracle's per-shard lists
        for g, (lo, hi) in enumerate(((0, half), (half, n))):
            want = oracle_lib.recall_topk(E[lo:hi], Q, k, row_base=lo)
            assert (keys_all[g].cpu().numpy().view(np.uint64) == want).all()
        rows, scores, cnt = engs[0].merge_keys(keys_all.data_ptr(), 2, B, k)
        orows, oscores, on = oracle_lib.keys_split(full)
        assert (rows == orows).all() and (scores.view(np.uint32) == oscores.view(np.uint32)).all() and (cnt == on).all()

        # downstream of the merge on a slice of the requests (what each rank does for its own 64)
        fields, factors, linear = synth.rank_tables(n_items=n, n_fields=32)
        Dm = synth.diversity(n_items=n, dim=32)
        e = engs[1]
        e.set_item_fields(fields)
        for t, (f, l) in enumerate(zip(factors, linear)):
            e.set_feature_table(t, f, l)
        e.set_fm_bias(0.05)
        e.set_diversity_matrix(Dm)
        b0, nb = 6, 5                                                         # requests 6..10 of the global batch
        out_rows = torch.empty(nb, T, dtype=torch.int32, device=dev)
        out_sc = torch.empty(nb, T, dtype=torch.float64, device=dev)
        out_n = torch.empty(nb, dtype=torch.int32, device=dev)
        p = DppParams(top_n=T, alpha=1.0, window_size=10)
        e.recommend_from_keys_dev(keys_all.data_ptr() + b0 * k * 8, 2, B * k, nb, k, MODEL_FM, p, out_rows.data_ptr(),
                                  out_sc.data_ptr(), out_n.data_ptr())
        e.sync()
        got_rows = out_rows.cpu().numpy().view(np.uint32)
        for i in range(nb):
            r = orows[b0 + i]
            logit, _ = oracle_lib.gather_fm(fields, factors, linear, 0.05, r, want_x=False)
            sc = oracle_lib.sigmoid(logit).astype(np.float64)
            perm = oracle_lib.stable_sort_desc(sc)
            idx, st = oracle_lib.dpp_request(Dm[r[perm]].astype(np.float64), sc[perm], T, alpha=1.0, window_size=10)
            assert st == 0 and (got_rows[i, :len(idx)] == r[perm][idx]).all()
    finally:
        for g in engs:
            g.close()


def _global_protocol(engs, Q, k, dev):
    """The global-threshold protocol on one GPU: G engines hold the shards; the two all-gathers are emulated by writing
    every shard's output into its slot of the gathered tensor."""
    import torch
    G, Bg = len(engs), Q.shape[0]
    r = engs[0].shard_sample_len(k)
    q_dev = torch.from_numpy(Q).to(dev)
    samples = torch.zeros(G, Bg, r, dtype=torch.int64, device=dev)
    for g, e in enumerate(engs):
        e.shard_sample_dev(q_dev.data_ptr(), Bg, k, G, samples[g].data_ptr())
        e.sync()
    blk = Bg * k + Bg
    gathered = torch.zeros(G, blk, dtype=torch.int64, device=dev)
    for g, e in enumerate(engs):
        e.shard_candidates_dev(q_dev.data_ptr(), Bg, k, G, samples.data_ptr(), gathered[g].data_ptr())
        e.sync()
    retry = torch.zeros(2, dtype=torch.int32, device=dev)
    engs[0].shard_check_dev(gathered.data_ptr(), G, Bg, k, retry.data_ptr())
    engs[0].sync()
    return gathered, retry.cpu().numpy()


# (8, 64, 512, 1000) is the shape of the 8-GPU bench: 64 requests per rank, two filter passes of 256 queries per shard
@pytest.mark.parametrize("G,d,B,k", [(2, 64, 70, 300), (4, 64, 130, 1000), (2, 128, 9, 200), (8, 64, 512, 1000),
                                     (8, 128, 130, 1000)])
def test_global_threshold_protocol_matches_unsharded_oracle(oracle_lib, G, d, B, k):
    import torch
    from pairec_b200 import Engine
    n = max(1_200_000, 300_000 * G)   # every shard large enough to join the sampled protocol (>= 8 sample tiles)
    rng = np.random.default_rng(37 + G)
    E = (rng.standard_normal((n, d)) / np.sqrt(d)).astype(np.float32)
    E[n // G:n // G + 300] = E[:300]              # ties across a shard boundary
    Q = (rng.standard_normal((B, d)) / np.sqrt(d)).astype(np.float32)
    dev = torch.device("cuda:0")
    bounds = [g * (n // G) for g in range(G)] + [n]
    engs = [Engine(0) for _ in range(G)]
    try:
        for g, e in enumerate(engs):
            e.set_item_matrix(E[bounds[g]:bounds[g + 1]], row_base=bounds[g])
        gathered, retry = _global_protocol(engs, Q, k, dev)
        assert retry[0] == 0 and retry[1] == 0, "a query asked for the exact protocol on well-mixed data"
        lists = gathered[:, :B * k].reshape(G, B, k)
        status = gathered[:, B * k:]
        assert (status == 0).all()
        # a shard only emits rows that reach the global threshold: ~4*max(k,1024)/G of them (capped at k)
        per = (lists != 0).sum(dim=2).float().mean().item()
        assert per <= min(k, 1.5 * 4 * max(k, 1024) / G)
        compact = lists.contiguous()
        rows, scores, cnt = engs[0].merge_keys(compact.data_ptr(), G, B, k)
        full = oracle_lib.recall_topk(E, Q, k)
        orows, oscores, on = oracle_lib.keys_split(full)
        assert (cnt == on).all()
        assert (rows == orows).all(), "merged rows differ from the unsharded oracle"
        assert (scores.view(np.uint32) == oscores.view(np.uint32)).all()
    finally:
        for e in engs:
            e.close()


def test_global_threshold_protocol_flags_adversarial_order(oracle_lib):
    # every large score sits in tiles the strided sample never visits: tau is far too low, candidate lists overflow,
    # and the check must ask for the exact protocol (on every rank alike) instead of returning a wrong list
    import torch
    from pairec_b200 import Engine
    n, d, G, k = 800_000, 64, 2, 500
    rng = np.random.default_rng(43)
    E = (rng.standard_normal((n, d)) * 0.01).astype(np.float32)
    tile = np.arange(n) // 256
    E[:, 0] = np.where(tile % 128 == 0, 0.0, 1.0 + rng.random(n) * 0.5).astype(np.float32)
    Q = np.zeros((3, d), dtype=np.float32)
    Q[:, 0] = 1.0
    dev = torch.device("cuda:0")
    engs = [Engine(0) for _ in range(G)]
    try:
        half = n // 2
        engs[0].set_item_matrix(E[:half], row_base=0)
        engs[1].set_item_matrix(E[half:], row_base=half)
        _, retry = _global_protocol(engs, Q, k, dev)
        assert retry[0] == 1 and retry[1] == 3
    finally:
        for e in engs:
            e.close()


def test_all_to_all_exchange_form_equals_unsharded_path(oracle_lib):
    """Exchange #2 as ONE all-to-all (bench.py's default at N > 1): prg_shard_pack_owner regroups a shard's candidates into
    per-owner chunks, each owner checks and merges only its own B requests (prg_shard_check_owner, g_stride = B*k + B).
    Emulated on one GPU: chunk o of shard g lands in slot g of owner o's receive buffer."""
    import torch
    from pairec_b200 import DppParams, Engine
    from pairec_b200.binding import MODEL_FM
    G, d, B, k, T = 4, 64, 9, 300, 12
    Bg = G * B
    n = 1_200_000
    rng = np.random.default_rng(61)
    E = (rng.standard_normal((n, d)) / np.sqrt(d)).astype(np.float32)
    Q = (rng.standard_normal((Bg, d)) / np.sqrt(d)).astype(np.float32)
    fields, factors, linear = synth.rank_tables(n_items=n, n_fields=6)
    Dm = synth.diversity(n_items=n, dim=32)
    dev = torch.device("cuda:0")
    p = DppParams(top_n=T, alpha=1.0, window_size=10)

    def load_rank(e):
        e.set_item_fields(fields)
        for t, (f, l) in enumerate(zip(factors, linear)):
            e.set_feature_table(t, f, l)
        e.set_fm_bias(0.05)
        e.set_diversity_matrix(Dm)
    ref = Engine(0)
    ref.set_item_matrix(E)
    load_rank(ref)
    want = ref.recommend(Q, k, MODEL_FM, p)
    ref.close()
    bounds = [g * (n // G) for g in range(G)] + [n]
    engs = [Engine(0) for _ in range(G)]
    try:
        for g, e in enumerate(engs):
            e.set_item_matrix(E[bounds[g]:bounds[g + 1]], row_base=bounds[g])
            load_rank(e)
        r = engs[0].shard_sample_len(k)
        q_dev = torch.from_numpy(Q).to(dev)
        samples = torch.zeros(G, Bg, r, dtype=torch.int64, device=dev)
        for g, e in enumerate(engs):
            e.shard_sample_dev(q_dev.data_ptr(), Bg, k, G, samples[g].data_ptr())
            e.sync()
        chunk = B * k + B
        recv = torch.zeros(G, G, chunk, dtype=torch.int64, device=dev)          # [owner][source shard][chunk]
        for g, e in enumerate(engs):
            cand = torch.zeros(Bg * k + Bg, dtype=torch.int64, device=dev)
            packed = torch.zeros(G, chunk, dtype=torch.int64, device=dev)
            e.shard_candidates_dev(q_dev.data_ptr(), Bg, k, G, samples.data_ptr(), cand.data_ptr())
            e.shard_pack_owner_dev(cand.data_ptr(), Bg, B, k, packed.data_ptr())
            e.sync()
            # the pack is a pure regrouping of the lists and the status words
            assert torch.equal(packed[:, :B * k].reshape(-1), cand[:Bg * k])
            assert torch.equal(packed[:, B * k:].reshape(-1), cand[Bg * k:])
            recv[:, g] = packed                                                  # the all-to-all
        for o, e in enumerate(engs):
            retry = torch.zeros(2, dtype=torch.int32, device=dev)
            e.shard_check_owner_dev(recv[o].data_ptr(), G, B, k, o * B, retry.data_ptr())
            out_rows = torch.empty(B, T, dtype=torch.int32, device=dev)
            out_sc = torch.empty(B, T, dtype=torch.float64, device=dev)
            out_n = torch.empty(B, dtype=torch.int32, device=dev)
            e.recommend_from_keys_dev(recv[o].data_ptr(), G, chunk, B, k, MODEL_FM, p, out_rows.data_ptr(), out_sc.data_ptr(),
                                      out_n.data_ptr())
            e.sync()
            assert retry.cpu().numpy().tolist() == [0, 0]
            sl = slice(o * B, (o + 1) * B)
            assert (out_rows.cpu().numpy().view(np.uint32) == want[0][sl]).all(), f"owner {o}: rows differ from the unsharded path"
            assert (out_sc.cpu().numpy().view(np.uint64) == want[1][sl].view(np.uint64)).all()
            assert (out_n.cpu().numpy() == want[2][sl]).all()
    finally:
        for e in engs:
            e.close()
