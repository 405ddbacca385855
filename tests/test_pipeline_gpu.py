"""Fused request path (prg_recommend) against the oracle run stage by stage."""
import numpy as np
import pytest

from tests import synth

pytestmark = pytest.mark.gpu


def test_recommend_fm_matches_staged_oracle(engine, oracle_lib):
    from pairec_b200 import DppParams
    from pairec_b200.binding import MODEL_FM
    n_items, d, B, k, T = 300_000, 64, 6, 500, 20
    rng = np.random.default_rng(42)
    E = (rng.standard_normal((n_items, d)) / 8).astype(np.float32)
    Q = (rng.standard_normal((B, d)) / 8).astype(np.float32)
    fields, factors, linear = synth.rank_tables(n_items=n_items, n_fields=32)
    D = synth.diversity(n_items=n_items, dim=32)
    engine.set_item_matrix(E)
    engine.set_item_fields(fields)
    for t, (f, l) in enumerate(zip(factors, linear)):
        engine.set_feature_table(t, f, l)
    engine.set_fm_bias(0.05)
    engine.set_diversity_matrix(D)
    p = DppParams(top_n=T, alpha=1.0, window_size=10)
    rows, scores, n = engine.recommend(Q, k, MODEL_FM, p)

    keys = oracle_lib.recall_topk(E, Q, k)
    rrows, _, _ = oracle_lib.keys_split(keys)
    for b in range(B):
        logit, _ = oracle_lib.gather_fm(fields, factors, linear, 0.05, rrows[b], want_x=False)
        sc = oracle_lib.sigmoid(logit).astype(np.float64)
        perm = oracle_lib.stable_sort_desc(sc)
        srows, ssc = rrows[b][perm], sc[perm]
        idx, st = oracle_lib.dpp_request(D[srows].astype(np.float64), ssc, T, alpha=1.0, window_size=10)
        assert st == 0 and n[b] == len(idx)
        assert (rows[b, :n[b]] == srows[idx]).all(), "final ordering differs"
        assert (scores[b, :n[b]].view(np.uint64) == ssc[idx].view(np.uint64)).all()


def test_recommend_repairs_failed_recall_queries(oracle_lib):
    """Adversarial row order: the sampled threshold fails for query 0, the deferred status check catches it after
    the downstream stages were already enqueued, the query is redone densely and rank/sort/DPP run again — through
    both the host-buffer call and the device-buffer call + prg_sync."""
    import torch
    from pairec_b200 import DppParams, Engine
    from pairec_b200.binding import MEM_DEVICE, MODEL_FM
    n, d, k, T = 400_000, 64, 500, 12
    rng = np.random.default_rng(11)
    E = (rng.standard_normal((n, d)) * 0.01).astype(np.float32)
    n_tiles = (n + 255) // 256
    stride = n_tiles // max(64, n_tiles // 128)
    tile = np.arange(n) // 256
    E[:, 0] = np.where(tile % stride == 0, 0.0, 1.0 + rng.random(n) * 0.5).astype(np.float32)
    Q = (rng.standard_normal((3, d)) * 0.01).astype(np.float32)
    Q[0] = 0
    Q[0, 0] = 1.0
    fields, factors, linear = synth.rank_tables(n_items=n, n_fields=32)
    D = synth.diversity(n_items=n, dim=32)
    eng = Engine(0)
    try:
        eng.set_item_matrix(E)
        eng.set_item_fields(fields)
        for t, (f, l) in enumerate(zip(factors, linear)):
            eng.set_feature_table(t, f, l)
        eng.set_fm_bias(0.05)
        eng.set_diversity_matrix(D)
        p = DppParams(top_n=T, alpha=1.0, window_size=10)
        rows_h, scores_h, n_h = eng.recommend(Q, k, MODEL_FM, p)
        assert eng.recall_stats()["fallback_queries"] >= 1
        dev = torch.device("cuda:0")
        q_dev = torch.from_numpy(Q).to(dev)
        rows_d = torch.zeros(3, T, dtype=torch.int32, device=dev)
        sc_d = torch.zeros(3, T, dtype=torch.float64, device=dev)
        n_d = torch.zeros(3, dtype=torch.int32, device=dev)
        eng.recommend_dev(q_dev.data_ptr(), 3, k, MODEL_FM, p, rows_d.data_ptr(), sc_d.data_ptr(), n_d.data_ptr())
        eng.sync()
        assert (rows_d.cpu().numpy().view(np.uint32) == rows_h).all()
        keys = oracle_lib.recall_topk(E, Q, k)
        rrows, _, _ = oracle_lib.keys_split(keys)
        for b in range(3):
            logit, _ = oracle_lib.gather_fm(fields, factors, linear, 0.05, rrows[b], want_x=False)
            sc = oracle_lib.sigmoid(logit).astype(np.float64)
            perm = oracle_lib.stable_sort_desc(sc)
            idx, st = oracle_lib.dpp_request(D[rrows[b][perm]].astype(np.float64), sc[perm], T, alpha=1.0, window_size=10)
            assert (rows_h[b, :len(idx)] == rrows[b][perm][idx]).all()
    finally:
        eng.close()
