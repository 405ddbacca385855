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
