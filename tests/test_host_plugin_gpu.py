"""The GPU-backed plugins behind the reference's interfaces, driven like /api/recommend: VectorRecall
(service/recall/vector_recall.go) -> algorithm.Run("gpu_recall") ; RankService -> algorithm.Run("gpu_rank") ;
SortService -> "gpu_dpp" (DPPSort).  One recconf JSON, string item ids in and out."""
import numpy as np
import pytest

from tests import synth

pytestmark = pytest.mark.gpu

RECCONF = {
    "RecallConfs": [{"Name": "u2i_vec", "RecallType": "VectorRecall", "RecallCount": 200, "RecallAlgo": "gpu_recall",
                     "ItemType": "video"}],
    "AlgoConfs": [],
    "SceneConfs": {"home_feed": {"default": {"RecallNames": ["u2i_vec"]}}},
    "FilterNames": {"default": ["UniqueFilter"]},
    "RankConf": {"home_feed": {"RankAlgoList": ["gpu_rank"], "RankScore": "${gpu_rank}", "BatchCount": 1000}},
    "SortNames": {"home_feed": ["ItemRankScore", "gpu_dpp"]},
    "SortConfs": [{"Name": "gpu_dpp", "SortType": "DPPSort", "DPPConf": {"Alpha": 1.0, "WindowSize": 10}}],
}


def test_recommend_through_reference_plugin_interfaces(engine, oracle_lib):
    from pairec_b200.binding import MODEL_FM
    from pairec_b200.plugin import HostServer
    n_items, d = 300_000, 64
    rng = np.random.default_rng(77)
    E = (rng.standard_normal((n_items, d)) / 8).astype(np.float32)
    fields, factors, linear = synth.rank_tables(n_items=n_items, n_fields=32)
    D = synth.diversity(n_items=n_items, dim=64)
    engine.set_item_matrix(E)
    engine.set_item_fields(fields)
    for t, (f, l) in enumerate(zip(factors, linear)):
        engine.set_feature_table(t, f, l)
    engine.set_fm_bias(0.05)
    engine.set_diversity_matrix(D)
    ids = ["item_%d" % i for i in range(n_items)]

    srv = HostServer(RECCONF)
    try:
        srv.attach_engine(engine, ids)
        srv.register_gpu_plugins(recall_algo="gpu_recall", rank_algo="gpu_rank", model=MODEL_FM, dpp_sort="gpu_dpp")
        q = (rng.standard_normal(d) / 8).astype(np.float32)
        srv.set_user_vector("u42", q)
        resp = srv.recommend(scene_id="home_feed", uid="u42", size=20)
        assert resp["code"] == 200 and resp["size"] == 20, resp
        got = [it["item_id"] for it in resp["items"]]
        assert all(it["retrieve_id"] == "u2i_vec" and it["item_type"] == "video" for it in resp["items"])

        # oracle, stage by stage (the user vector travels as the "i:v i:v" string of vector_recall.go:70-82)
        qv = np.array([np.float32(float(repr(float(v)))) for v in q], dtype=np.float32).reshape(1, -1)
        keys = oracle_lib.recall_topk(E, qv, 200)
        rows, _, _ = oracle_lib.keys_split(keys)
        logit, _ = oracle_lib.gather_fm(fields, factors, linear, 0.05, rows[0], want_x=False)
        sc = oracle_lib.sigmoid(logit).astype(np.float64)
        perm = oracle_lib.go_sort(sc)                      # ItemRankScore = Go sort.Sort order
        srows, ssc = rows[0][perm], sc[perm]
        idx, st = oracle_lib.dpp_request(D[srows].astype(np.float64), ssc, 20, alpha=1.0, window_size=10)
        assert st == 0
        assert got == [ids[r] for r in srows[idx]]
        assert [it["score"] for it in resp["items"]] == [float(s) for s in ssc[idx]]

        # unknown user -> VectoryEmptyError path: empty result, no abort (vector_recall.go:60-67)
        assert srv.recommend(scene_id="home_feed", uid="nobody", size=5)["size"] == 0
    finally:
        srv.close()
