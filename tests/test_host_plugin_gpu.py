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


def test_gpu_rank_as_custom_irank_with_user_features_and_hook_embeddings(oracle_lib):
    """The full-featured drop-in: rank.RegisterRank(scene, GpuRank) (service/rank/custom_rank.go:8-13) gets the Items and
    the User in ONE call per request — item ids map to rows without any injected feature, the user's categorical
    features (request `features`) reach FM and tower (service/rank/algo_data.go:104-118), and the DPP sort concatenates
    a registered embedding hook with the table embedding (sort/dpp_sort.go:362-370, :416-421)."""
    from pairec_b200 import Engine
    from pairec_b200.binding import MODEL_FM_MLP
    from pairec_b200.plugin import HostServer
    n_items, d, F, U = 200_000, 64, 8, 2
    rng = np.random.default_rng(78)
    E = (rng.standard_normal((n_items, d)) / 8).astype(np.float32)
    fields, factors, linear = synth.rank_tables(n_items=n_items, n_fields=F + U)
    factors = [f * 8 for f in factors]
    fields = np.ascontiguousarray(fields[:, :F])
    D = synth.diversity(n_items=n_items, dim=64)
    dims = [(F + U) * 16 + 1, 128, 64, 1]
    W, b = synth.mlp_weights(dims)
    eng = Engine(0)
    conf = dict(RECCONF)
    conf["RankConf"] = {"home_feed": {"RankAlgoList": [], "RankScore": "${gpu_rank}", "BatchCount": 100}}
    conf["SortConfs"] = [{"Name": "gpu_dpp", "SortType": "DPPSort",
                          "DPPConf": {"Alpha": 1.0, "WindowSize": 10, "TableName": "item_emb", "EmbeddingHookNames": ["cat_emb"]}}]
    srv = HostServer(conf)
    try:
        eng.set_item_matrix(E)
        eng.set_item_fields(fields)
        for t, (f, l) in enumerate(zip(factors, linear)):
            eng.set_feature_table(t, f, l)
        eng.set_fm_bias(0.05)
        eng.set_user_fields(U, 1)
        eng.set_mlp(dims, W, b)
        eng.set_diversity_matrix(D)
        ids = ["item_%d" % i for i in range(n_items)]
        srv.attach_engine(eng, ids)
        srv.register_gpu_plugins(recall_algo="gpu_recall", dpp_sort="gpu_dpp")
        vocab = ["bj", "hz", "sh", "sz"]
        srv.register_gpu_rank("home_feed", "gpu_rank", MODEL_FM_MLP,
                              user_fields=[{"column": "age_bucket", "id": True}, {"column": "city", "vocab": vocab}],
                              dense_columns=["ctx_hour"])
        hook = rng.standard_normal((n_items, 6))
        srv.register_embedding_hook("cat_emb", ids, hook)
        q = (rng.standard_normal(d) / 8).astype(np.float32)
        srv.set_user_vector("u1", q)
        srv.set_user_vector("u2", q)
        feats = {"u1": {"age_bucket": 7, "city": "hz", "ctx_hour": 0.5}, "u2": {"age_bucket": 31, "city": "paris", "ctx_hour": -1.25}}
        out = {}
        for uid in ("u1", "u2"):
            resp = srv.recommend(scene_id="home_feed", uid=uid, size=20, features=feats[uid])
            assert resp["code"] == 200 and resp["size"] == 20, resp
            out[uid] = resp
            # oracle, stage by stage
            qv = np.array([np.float32(float(repr(float(v)))) for v in q], dtype=np.float32).reshape(1, -1)
            rows, _, _ = oracle_lib.keys_split(oracle_lib.recall_topk(E, qv, 200))
            city = feats[uid]["city"]
            uid_ids = np.array([feats[uid]["age_bucket"], vocab.index(city) if city in vocab else 0xFFFFFFFF], dtype=np.uint32)
            fm, x = oracle_lib.gather_fm(fields, factors, linear, 0.05, rows[0], user_ids=uid_ids,
                                         user_dense=np.array([feats[uid]["ctx_hour"]], dtype=np.float32))
            want = oracle_lib.sigmoid((fm + oracle_lib.mlp_forward(x, dims, W, b)).astype(np.float32)).astype(np.float64)
            got_rows = np.array([int(it["item_id"][5:]) for it in resp["items"]])
            pos = {int(r): i for i, r in enumerate(rows[0])}
            got_sc = np.array([it["score"] for it in resp["items"]])
            ref_sc = np.array([want[pos[int(r)]] for r in got_rows])
            assert (np.abs(got_sc - ref_sc) / ref_sc).max() <= 1e-5
            # final order: the oracle's DPP (hook ++ table) over the GPU's own scores in Go sort order
            sc_gpu = eng.rank(MODEL_FM_MLP, rows, user_ids=uid_ids.reshape(1, -1),
                              user_dense=np.array([[feats[uid]["ctx_hour"]]], dtype=np.float32))[0]
            perm = oracle_lib.go_sort(sc_gpu)
            srows, ssc = rows[0][perm], sc_gpu[perm]
            idx, st = oracle_lib.dpp_request_ex(D[srows].astype(np.float64), ssc, 20, hook=hook[srows], alpha=1.0, window_size=10)
            assert st == 0 and got_rows.tolist() == srows[idx].tolist()
        assert [it["score"] for it in out["u1"]["items"]] != [it["score"] for it in out["u2"]["items"]]

        # The same requests through the OTHER drop-in route: RankConf.Processor "EasyRec".  The stock RankService then
        # builds one easyrec.PBRequest per batch — item ids + the request's user features (algo_data.go:173-325) — and
        # hands it to the IAlgorithm registered under the algo name: no IRank, no feature-load hook.  64 items per batch
        # (4 algorithm.Run calls per request): identical responses.
        conf2 = dict(conf)
        conf2["RankConf"] = {"home_feed": {"RankAlgoList": ["gpu_easyrec"], "RankScore": "${gpu_easyrec}",
                                           "Processor": "EasyRec", "BatchCount": 64}}
        srv2 = HostServer(conf2)
        try:
            srv2.attach_engine(eng, ids)
            srv2.register_gpu_plugins(recall_algo="gpu_recall", dpp_sort="gpu_dpp")
            srv2.register_gpu_easyrec("gpu_easyrec", MODEL_FM_MLP,
                                      user_fields=[{"column": "age_bucket", "id": True}, {"column": "city", "vocab": vocab}],
                                      dense_columns=["ctx_hour"])
            srv2.register_embedding_hook("cat_emb", ids, hook)
            for uid in ("u1", "u2"):
                srv2.set_user_vector(uid, q)
                resp2 = srv2.recommend(scene_id="home_feed", uid=uid, size=20, features=feats[uid])
                assert resp2["code"] == 200 and resp2["items"] == out[uid]["items"], uid
        finally:
            srv2.close()
    finally:
        srv.close()
        eng.close()


def test_short_user_vector_is_rejected_not_overread(engine):
    """service/recall/vector_recall.go:72-82 skips malformed 'i:v' pairs silently; the plugin must not hand a short
    vector to the C ABI (which takes no length)."""
    from pairec_b200.plugin import HostServer
    rng = np.random.default_rng(79)
    E = (rng.standard_normal((50_000, 64)) / 8).astype(np.float32)
    engine.set_item_matrix(E)
    srv = HostServer(RECCONF)
    try:
        srv.attach_engine(engine, ["i%d" % i for i in range(50_000)])
        srv.register_gpu_plugins(recall_algo="gpu_recall")
        srv.set_user_vector("short", rng.standard_normal(40).astype(np.float32))
        resp = srv.recommend(scene_id="home_feed", uid="short", size=5)
        assert resp["size"] == 0
        assert any("user vector has 40 elements" in l for l in resp["log"]), resp["log"]
    finally:
        srv.close()
