"""GPU parity of the user / context side of the rank features: the reference merges
features = userFeatures, then itemFeatures on top (service/rank/algo_data.go:104-118) — two users get different
scores for the same items.  FM scores bit-exact, tower scores 1e-5 relative, multi-head score maps
(algorithm/eas/easyrec_response.go:35-70) and the RankScore sum of products (service/rank/rank_service.go:339-363)."""
import threading

import numpy as np
import pytest

from tests import synth

pytestmark = pytest.mark.gpu

RTOL = 1e-5


def _setup(F, U, n_dense, dims=None, n_items=3000, w0=0.03, scale=8.0, seed=6):
    from pairec_b200 import Engine
    fields, factors, linear = synth.rank_tables(n_items=n_items, n_fields=F + U)
    factors = [f * scale for f in factors]
    fields = np.ascontiguousarray(fields[:, :F])      # the last U tables belong to the user fields
    eng = Engine(0)
    eng.set_item_fields(fields)
    for t, (f, l) in enumerate(zip(factors, linear)):
        eng.set_feature_table(t, f, l)
    eng.set_fm_bias(w0)
    eng.set_user_fields(U, n_dense)
    W = b = None
    if dims is not None:
        W, b = synth.mlp_weights(dims, seed=seed)
        eng.set_mlp(dims, W, b)
    return eng, fields, factors, linear, W, b


def _users(rng, B, U, factors, F, n_dense):
    ids = np.stack([rng.integers(0, factors[F + u].shape[0], size=B) for u in range(U)], axis=1).astype(np.uint32)
    dense = (rng.standard_normal((B, n_dense)) * 0.5).astype(np.float32) if n_dense else None
    return ids, dense


def test_fm_with_user_fields_bit_exact(oracle_lib):
    from pairec_b200.binding import MODEL_FM
    F, U = 6, 3
    eng, fields, factors, linear, _, _ = _setup(F, U, 0, scale=1.0)
    try:
        rng = np.random.default_rng(0)
        B, n = 5, 333
        rows = rng.integers(0, 3000, size=(B, n)).astype(np.uint32)
        rows[1] = rows[0]                              # same items, different users
        rows[2, -4:] = 0xFFFFFFFF
        ids, _ = _users(rng, B, U, factors, F, 0)
        ids[3, 1] = 0xFFFFFFFF                         # the user has no such feature
        ids[4] = 0xFFFFFFFF                            # a user without features at all
        got = eng.rank(MODEL_FM, rows, user_ids=ids)
        for b in range(B):
            logit, _ = oracle_lib.gather_fm(fields, factors, linear, 0.03, rows[b], want_x=False, user_ids=ids[b])
            want = oracle_lib.sigmoid(logit).astype(np.float64)
            want[rows[b] == 0xFFFFFFFF] = 0.0
            assert (got[b].view(np.uint64) == want.view(np.uint64)).all(), f"request {b}"
        assert (got[0] != got[1]).mean() > 0.9, "two users must get different scores for the same items"
        # no user block at all == every user feature absent == the item-only model
        none = eng.rank(MODEL_FM, rows)
        logit, _ = oracle_lib.gather_fm(fields, factors[:F], linear[:F], 0.03, rows[4], want_x=False)
        assert (none[4].view(np.uint64) == got[4].view(np.uint64)).all()
        assert (oracle_lib.sigmoid(logit).astype(np.float64).view(np.uint64) == got[4].view(np.uint64)).all()
    finally:
        eng.close()


@pytest.mark.parametrize("F,U,nd,hidden,shape", [(32, 8, 0, [512, 256, 128], (3, 333)), (4, 2, 8, [64], (2, 320)),
                                                 (16, 3, 5, [192, 64], (4, 100))])
def test_tower_with_user_and_context_features(oracle_lib, F, U, nd, hidden, shape):
    from pairec_b200.binding import MODEL_MLP, MODEL_FM_MLP
    dims = [(F + U) * 16 + nd] + hidden + [1]
    eng, fields, factors, linear, W, b = _setup(F, U, nd, dims)
    try:
        rng = np.random.default_rng(1)
        B, n = shape
        rows = rng.integers(0, 3000, size=shape).astype(np.uint32)
        rows[1] = rows[0]
        rows[-1, -3:] = 0xFFFFFFFF
        ids, dense = _users(rng, B, U, factors, F, nd)
        got = eng.rank(MODEL_MLP, rows, user_ids=ids, user_dense=dense)
        got2 = eng.rank(MODEL_FM_MLP, rows, user_ids=ids, user_dense=dense)
        spread = []
        for bi in range(B):
            fm, x = oracle_lib.gather_fm(fields, factors, linear, 0.03, rows[bi], user_ids=ids[bi],
                                         user_dense=None if dense is None else dense[bi])
            ml = oracle_lib.mlp_forward(x, dims, W, b)
            spread.append(np.ptp(ml))
            live = rows[bi] != 0xFFFFFFFF
            want = oracle_lib.sigmoid(ml).astype(np.float64)
            rel = np.abs(got[bi][live] - want[live]) / np.abs(want[live])
            assert rel.max() <= RTOL, f"request {bi}: max relative error {rel.max():.3e}"
            want2 = oracle_lib.sigmoid((fm + ml).astype(np.float32)).astype(np.float64)
            rel2 = np.abs(got2[bi][live] - want2[live]) / np.abs(want2[live])
            assert rel2.max() <= RTOL
            assert (got[bi][~live] == 0).all()
        assert max(spread) > 0.05
        assert np.abs(got[0] - got[1]).max() > 1e-4, "the user features must reach the tower"
    finally:
        eng.close()


def test_multi_head_score_map_and_rank_score(oracle_lib):
    """3 output heads (a multi-target EasyRec model): GetScoreMap values per head, Item.Score = 0.5*h0 + 2*h1 + 0.25*h2."""
    from pairec_b200.binding import MODEL_MLP
    F, U = 8, 2
    dims = [(F + U) * 16, 128, 64, 3]
    eng, fields, factors, linear, W, b = _setup(F, U, 0, dims)
    try:
        coef = [0.5, 2.0, 0.25]
        eng.set_rank_score(coef)
        rng = np.random.default_rng(2)
        rows = rng.integers(0, 3000, size=(3, 200)).astype(np.uint32)
        ids, _ = _users(rng, 3, U, factors, F, 0)
        score, smap = eng.rank(MODEL_MLP, rows, user_ids=ids, score_map=True)
        assert smap.shape == (3, 200, 3)
        for bi in range(3):
            _, x = oracle_lib.gather_fm(fields, factors, linear, 0.03, rows[bi], user_ids=ids[bi])
            ml = oracle_lib.mlp_forward(x, dims, W, b)                       # [n, 3]
            want = oracle_lib.sigmoid(ml).astype(np.float64)
            rel = np.abs(smap[bi] - want) / np.abs(want)
            assert rel.max() <= RTOL
            # Item.Score is the oracle's left-to-right sum of products over the GPU's own head scores, bit for bit
            for i in (0, 7, 199):
                assert score[bi, i] == oracle_lib.rank_score_expr(smap[bi, i], coef)
        assert np.ptp(smap[..., 0] - smap[..., 1]) > 1e-3
    finally:
        eng.close()


def test_fused_path_and_batcher_carry_user_features(oracle_lib):
    """prg_recommend_ex == recall -> prg_rank_ex -> sort -> DPP staged calls, and one-request-per-thread through the
    batcher gives each caller the answer for ITS user."""
    from pairec_b200 import Batcher, DppParams
    from pairec_b200.binding import MODEL_FM_MLP
    F, U, nd = 8, 2, 4
    dims = [(F + U) * 16 + nd, 128, 64, 1]
    n_items = 300_000
    eng, fields, factors, linear, W, b = _setup(F, U, nd, dims, n_items=n_items)
    try:
        rng = np.random.default_rng(3)
        E = (rng.standard_normal((n_items, 64)) / 8).astype(np.float32)
        eng.set_item_matrix(E)
        eng.set_diversity_matrix(synth.diversity(n_items=n_items, dim=32))
        B, k, T = 6, 200, 20
        Q = (rng.standard_normal((B, 64)) / 8).astype(np.float32)
        Q[1] = Q[0]                                                # same query, different users
        ids, dense = _users(rng, B, U, factors, F, nd)
        p = DppParams(top_n=T, alpha=1.0, window_size=10)
        rows, scores, n = eng.recommend(Q, k, MODEL_FM_MLP, p, user_ids=ids, user_dense=dense)
        rr, _, _ = eng.recall_topk(Q, k)
        rs = eng.rank(MODEL_FM_MLP, rr, user_ids=ids, user_dense=dense)
        perm = eng.sort_desc(rs)
        srows = np.take_along_axis(rr, perm, axis=1)
        sscore = np.take_along_axis(rs, perm, axis=1)
        idx, cnt, st = eng.dpp(srows, sscore, p)
        for bi in range(B):
            assert n[bi] == cnt[bi]
            assert (rows[bi, :n[bi]] == srows[bi][idx[bi, :cnt[bi]]]).all()
            assert (scores[bi, :n[bi]] == sscore[bi][idx[bi, :cnt[bi]]]).all()
        assert not (scores[0] == scores[1]).all()
        # oracle spot check of request 2's rank scores
        fm, x = oracle_lib.gather_fm(fields, factors, linear, 0.03, rr[2], user_ids=ids[2], user_dense=dense[2])
        want = oracle_lib.sigmoid((fm + oracle_lib.mlp_forward(x, dims, W, b)).astype(np.float32)).astype(np.float64)
        assert (np.abs(rs[2] - want) / want).max() <= RTOL

        bt = Batcher(eng, k, MODEL_FM_MLP, p, max_batch=4)
        try:
            out = [None] * B

            def call(i):
                out[i] = bt.recommend(Q[i], user_ids=ids[i], user_dense=dense[i])
            th = [threading.Thread(target=call, args=(i,)) for i in range(B)]
            for t in th:
                t.start()
            for t in th:
                t.join()
            for i in range(B):
                assert (out[i][0] == rows[i, :n[i]]).all() and (out[i][1] == scores[i, :n[i]]).all(), f"request {i}"
        finally:
            bt.close()
    finally:
        eng.close()


def test_device_resident_prerank_equals_staged_calls(oracle_lib):
    """general (pre-)rank with Action truncation (service/general_rank/base_general_rank.go:66-109,183) as a
    stage of the fused path: FM over the whole recall set, the best 150 go to the DeepFM-shaped rank, sort, DPP."""
    from pairec_b200 import DppParams
    from pairec_b200.binding import MODEL_FM, MODEL_FM_MLP
    F, U, nd = 8, 2, 0
    dims = [(F + U) * 16, 128, 64, 1]
    n_items = 300_000
    eng, fields, factors, linear, W, b = _setup(F, U, nd, dims, n_items=n_items)
    try:
        rng = np.random.default_rng(5)
        E = (rng.standard_normal((n_items, 64)) / 8).astype(np.float32)
        eng.set_item_matrix(E)
        eng.set_diversity_matrix(synth.diversity(n_items=n_items, dim=128))
        B, k, keep, T = 5, 600, 150, 20
        Q = (rng.standard_normal((B, 64)) / 8).astype(np.float32)
        ids, _ = _users(rng, B, U, factors, F, 0)
        p = DppParams(top_n=T, alpha=1.0, window_size=10)
        eng.set_prerank(MODEL_FM, keep)
        rows, scores, n = eng.recommend(Q, k, MODEL_FM_MLP, p, user_ids=ids)
        eng.set_prerank(MODEL_FM, 0)                                   # staged calls without the stage
        rr, _, _ = eng.recall_topk(Q, k)
        pre = eng.rank(MODEL_FM, rr, user_ids=ids)
        perm = eng.sort_desc(pre)
        kept = np.ascontiguousarray(np.take_along_axis(rr, perm, axis=1)[:, :keep])
        rs = eng.rank(MODEL_FM_MLP, kept, user_ids=ids)
        perm2 = eng.sort_desc(rs)
        srows = np.take_along_axis(kept, perm2, axis=1)
        sscore = np.take_along_axis(rs, perm2, axis=1)
        idx, cnt, st = eng.dpp(srows, sscore, p)
        for bi in range(B):
            assert n[bi] == cnt[bi] == T
            assert (rows[bi] == srows[bi][idx[bi]]).all() and (scores[bi] == sscore[bi][idx[bi]]).all()
        # the pre-rank order is the oracle's: FM scores bit-exact, stable order
        logit, _ = oracle_lib.gather_fm(fields, factors, linear, 0.03, rr[0], want_x=False, user_ids=ids[0])
        want = oracle_lib.sigmoid(logit).astype(np.float64)
        assert (pre[0].view(np.uint64) == want.view(np.uint64)).all()
        assert (perm[0] == oracle_lib.stable_sort_desc(want)).all()
    finally:
        eng.close()
