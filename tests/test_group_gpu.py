"""prg_group: all GPUs of a box behind ONE C-ABI call — row-sharded recall with the exchanges done as peer
stores inside the library, then rank / sort / DPP of each request on its owner GPU.  The result must be bit-identical to
prg_recommend_ex over the unsharded matrix.  With one GPU the G members share the device (every line of group.cu runs;
the stores are local); with >= 2 GPUs (gpurun --gpus N) the members sit on different devices and the stores cross NVLink."""
import numpy as np
import pytest

from tests import synth

pytestmark = pytest.mark.gpu


def _build(G, n, d, F, U, nd, dims, devices, E, fields, factors, linear, D):
    from pairec_b200 import Engine
    engs = []
    bounds = [g * (n // G) for g in range(G)] + [n]
    for g in range(G):
        e = Engine(devices[g % len(devices)])
        e.set_item_matrix(E[bounds[g]:bounds[g + 1]], row_base=bounds[g])
        e.set_item_fields(fields)
        for t, (f, l) in enumerate(zip(factors, linear)):
            e.set_feature_table(t, f, l)
        e.set_fm_bias(0.04)
        e.set_user_fields(U, nd)
        W, b = synth.mlp_weights(dims)
        e.set_mlp(dims, W, b)
        e.set_diversity_matrix(D)
        engs.append(e)
    return engs


def _reference(n, d, F, U, nd, dims, E, fields, factors, linear, D):
    from pairec_b200 import Engine
    e = Engine(0)
    e.set_item_matrix(E)
    e.set_item_fields(fields)
    for t, (f, l) in enumerate(zip(factors, linear)):
        e.set_feature_table(t, f, l)
    e.set_fm_bias(0.04)
    e.set_user_fields(U, nd)
    W, b = synth.mlp_weights(dims)
    e.set_mlp(dims, W, b)
    e.set_diversity_matrix(D)
    return e


def _devices():
    import torch
    return list(range(torch.cuda.device_count()))


@pytest.mark.parametrize("G,d,n_req,k", [(2, 64, 24, 300), (4, 64, 70, 1000), (3, 128, 10, 200)])
def test_group_equals_unsharded_path(G, d, n_req, k):
    from pairec_b200 import DppParams, Group
    from pairec_b200.binding import MODEL_FM_MLP
    n = max(1_200_000, 300_000 * G)
    F, U, nd = 8, 2, 3
    dims = [(F + U) * 16 + nd, 128, 64, 1]
    rng = np.random.default_rng(50 + G)
    E = (rng.standard_normal((n, d)) / np.sqrt(d)).astype(np.float32)
    E[n // G:n // G + 300] = E[:300]                      # ties across a shard boundary
    fields, factors, linear = synth.rank_tables(n_items=n, n_fields=F + U)
    factors = [f * 8 for f in factors]
    fields = np.ascontiguousarray(fields[:, :F])
    D = synth.diversity(n_items=n, dim=32)
    Q = (rng.standard_normal((n_req, d)) / np.sqrt(d)).astype(np.float32)
    uids = np.stack([rng.integers(0, factors[F + u].shape[0], size=n_req) for u in range(U)], axis=1).astype(np.uint32)
    dense = (rng.standard_normal((n_req, nd)) * 0.5).astype(np.float32)
    p = DppParams(top_n=20, alpha=1.0, window_size=10)
    ref = _reference(n, d, F, U, nd, dims, E, fields, factors, linear, D)
    try:
        want = ref.recommend(Q, k, MODEL_FM_MLP, p, user_ids=uids, user_dense=dense)
    finally:
        ref.close()
    engs = _build(G, n, d, F, U, nd, dims, _devices(), E, fields, factors, linear, D)
    grp = Group(engs)
    try:
        for _ in range(2):                                 # a second batch reuses every buffer
            rows, scores, cnt, redone = grp.recommend(Q, k, MODEL_FM_MLP, p, user_ids=uids, user_dense=dense)
            assert not redone
            assert (cnt == want[2]).all()
            assert (rows == want[0]).all(), "group rows differ from the unsharded path"
            assert (scores.view(np.uint64) == want[1].view(np.uint64)).all()
        # a smaller batch that does not divide by G (zero-padded inside), without user features
        rows2, scores2, cnt2, _ = grp.recommend(Q[:G + 1], k, MODEL_FM_MLP, p)
        ref = _reference(n, d, F, U, nd, dims, E, fields, factors, linear, D)
        try:
            w2 = ref.recommend(Q[:G + 1], k, MODEL_FM_MLP, p)
        finally:
            ref.close()
        assert (rows2 == w2[0]).all() and (scores2.view(np.uint64) == w2[1].view(np.uint64)).all() and (cnt2 == w2[2]).all()
        # more requests than any call before but fewer results in total (3 x the requests, top_n 20 -> 5): the per-request
        # count staging grows although the result staging does not
        if 3 * n_req <= 100:
            p3 = DppParams(top_n=5, alpha=1.0, window_size=10)
            Q3 = np.ascontiguousarray(np.tile(Q, (3, 1)))
            rows3, scores3, cnt3, _ = grp.recommend(Q3, k, MODEL_FM_MLP, p3)
            ref = _reference(n, d, F, U, nd, dims, E, fields, factors, linear, D)
            try:
                w3 = ref.recommend(Q, k, MODEL_FM_MLP, p3)
            finally:
                ref.close()
            for r in range(3):
                s = slice(r * n_req, (r + 1) * n_req)
                assert (rows3[s] == w3[0]).all() and (scores3[s].view(np.uint64) == w3[1].view(np.uint64)).all()
                assert (cnt3[s] == w3[2]).all()
    finally:
        grp.close()
        for e in engs:
            e.close()


def test_group_redoes_an_adversarial_batch_exactly():
    """Every large score sits in tiles the strided sample never visits: the global threshold is useless, the check fails,
    and the group redoes the batch with exact per-shard lists — same answer as the unsharded path, redone flag set."""
    from pairec_b200 import DppParams, Engine, Group
    from pairec_b200.binding import MODEL_FM
    n, d, G, k = 800_000, 64, 2, 500
    rng = np.random.default_rng(43)
    E = (rng.standard_normal((n, d)) * 0.01).astype(np.float32)
    tile = np.arange(n) // 256
    E[:, 0] = np.where(tile % 128 == 0, 0.0, 1.0 + rng.random(n) * 0.5).astype(np.float32)
    Q = np.zeros((3, d), dtype=np.float32)
    Q[:, 0] = 1.0
    fields, factors, linear = synth.rank_tables(n_items=n, n_fields=4)
    D = synth.diversity(n_items=n, dim=32)
    p = DppParams(top_n=10, alpha=1.0, window_size=10)

    def load(e, lo, hi):
        e.set_item_matrix(E[lo:hi], row_base=lo)
        e.set_item_fields(fields)
        for t, (f, l) in enumerate(zip(factors, linear)):
            e.set_feature_table(t, f, l)
        e.set_diversity_matrix(D)
    ref = Engine(0)
    load(ref, 0, n)
    want = ref.recommend(Q, k, MODEL_FM, p)
    ref.close()
    devs = _devices()
    engs = [Engine(devs[g % len(devs)]) for g in range(G)]
    load(engs[0], 0, n // 2)
    load(engs[1], n // 2, n)
    grp = Group(engs)
    try:
        rows, scores, cnt, redone = grp.recommend(Q, k, MODEL_FM, p)
        assert redone
        assert (rows == want[0]).all() and (scores.view(np.uint64) == want[1].view(np.uint64)).all() and (cnt == want[2]).all()
    finally:
        grp.close()
        for e in engs:
            e.close()
