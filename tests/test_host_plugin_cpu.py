"""Config 1 (BASELINE.json configs[0]): single-scene recconf JSON, 1k-item in-memory catalog, LOOKUP scoring +
AlgoScoreSort / ItemRankScore ordering on the host (plumbing, no GPU) through the C++ mirror of the reference's
plugin interfaces.  Written in the style of the reference's own plugin tests (sort/custom_field_sort_test.go:56-80:
build items + context, run the plugin, assert order/count)."""
import json

import numpy as np
import pytest

# the sample config the reference's scaffolder emits (commands/commands/project_data.go:40-94), with the LOOKUP
# algorithm bound as the scene's rank model
RECCONF = {
    "RunMode": "product",
    "ListenConf": {"HttpAddr": "", "HttpPort": 8000},
    "RecallConfs": [],
    "SortNames": {"default": ["ItemRankScore"], "by_field": ["field_sort"]},
    "FilterNames": {"default": ["UniqueFilter"]},
    "AlgoConfs": [{"Name": "lookup_score", "Type": "LOOKUP", "LookupConf": {"FieldName": "score"}}],
    "SceneConfs": {"home_feed": {"default": {"RecallNames": ["mem_recall"]}},
                   "by_field": {"default": {"RecallNames": ["mem_recall"]}}},
    "RankConf": {"home_feed": {"RankAlgoList": ["lookup_score"], "RankScore": "${lookup_score}", "BatchCount": 100},
                 "by_field": {"RankAlgoList": ["lookup_score"], "RankScore": "${lookup_score} * 2 + ${bonus}"}},
    "SortConfs": [{"Name": "field_sort", "SortType": "AlgoScoreSort", "SortByField": "freshness", "SwitchThreshold": 10.0}],
}


@pytest.fixture()
def server():
    from pairec_b200.plugin import HostServer
    s = HostServer(RECCONF)
    yield s
    s.close()


def _catalog(server, n=1000, seed=1, ties=False):
    rng = np.random.default_rng(seed)
    score = rng.random(n)
    if ties:
        score = np.round(score, 2)
    fresh = rng.random(n)
    for i in range(n):
        props = {"score": float(score[i]), "freshness": float(fresh[i]), "bonus": 0.25}
        if i % 97 == 0:
            del props["score"]          # LOOKUP default 0.5 (algorithm/lookup.go:47-49)
            score[i] = 0.5
        server.add_context_item("mem_recall", "i%07d" % i, 0.0, props)
    server.add_context_item("mem_recall", "i%07d" % 3, 0.0, {"score": 0.9})   # duplicate id -> UniqueFilter drops it
    server.commit()
    return score, fresh


def test_config1_lookup_rank_item_rank_score_sort(server, oracle_lib):
    score, _ = _catalog(server)
    resp = server.recommend(scene_id="home_feed", uid="u1", size=50)
    assert resp["code"] == 200 and resp["size"] == 50
    got = [it["item_id"] for it in resp["items"]]
    want = oracle_lib.go_sort(oracle_lib.lookup(score, np.ones_like(score, dtype=np.uint8)))[:50]
    assert got == ["i%07d" % i for i in want]
    assert [it["score"] for it in resp["items"]] == [float(score[i]) for i in want]
    assert all(it["retrieve_id"] == "mem_recall" for it in resp["items"])


def test_config1_tie_order_follows_go_sort(server, oracle_lib):
    score, _ = _catalog(server, ties=True)
    resp = server.recommend(scene_id="home_feed", uid="u1", size=1000)
    got = [int(it["item_id"][1:]) for it in resp["items"]]
    assert got == oracle_lib.go_sort(score).tolist()


def test_config1_rank_score_expression_and_algo_score_sort(server, oracle_lib):
    score, fresh = _catalog(server, n=300)
    resp = server.recommend(scene_id="by_field", uid="u1", size=20)
    # RankScore "${lookup_score} * 2 + ${bonus}" -> Item.Score; max(Score) <= SwitchThreshold -> sort by the field
    final = score * 2 + 0.25
    want = oracle_lib.algo_score_sort(final, fresh, 10.0)[:20]
    assert [int(it["item_id"][1:]) for it in resp["items"]] == want.tolist()
    assert [it["score"] for it in resp["items"]] == [float(final[i]) for i in want]


def test_size_larger_than_catalog_is_code_299(server):
    _catalog(server, n=30)
    resp = server.recommend(scene_id="home_feed", uid="u1", size=100)
    assert resp["code"] == 299 and resp["size"] == 30    # web/recommend_controller.go:131-142


def test_unknown_scene_and_missing_algorithm_do_not_abort():
    from pairec_b200.plugin import HostServer
    conf = dict(RECCONF, RankConf={"home_feed": {"RankAlgoList": ["nope"], "RankScore": "${nope}"}})
    s = HostServer(conf)
    try:
        s.add_context_item("mem_recall", "a", 0.0, {"score": 0.3})
        s.commit()
        r = s.recommend(scene_id="home_feed", uid="u", size=1)
        assert r["size"] == 1 and any("not found algorithm, name:nope" in l for l in r["log"])
        assert s.recommend(scene_id="no_such_scene", uid="u", size=1)["size"] == 0
    finally:
        s.close()


def test_ast_known_answer_and_operators():
    from pairec_b200.plugin import HostError, eval_expr
    # utils/ast/ast_test.go:12-27
    assert eval_expr("${ctr} + ${click} + ${price}", {"ctr": 0.1, "click": 0.3, "price": 0.1}) == 0.5
    assert eval_expr("${a} * 2 + ${b} ^ 2", {"a": 3, "b": 4}) == 22.0
    assert eval_expr("(${a} + 1) * (${b} - 1) / 2", {"a": 3, "b": 4}) == 6.0
    assert eval_expr("${missing} # 7", {}) == 7.0          # '#': first non-zero operand
    assert eval_expr("7 % 4", {}) == 3.0
    with pytest.raises(HostError):
        eval_expr("1 / ${z}", {"z": 0.0})                  # the reference panics (utils/ast/ast.go:243-249)


def test_bad_recconf_is_an_error_not_a_crash():
    from pairec_b200.plugin import HostError, HostServer
    with pytest.raises(HostError):
        HostServer("{not json")


def test_general_rank_prerank_then_actions(oracle_lib):
    # service/general_rank: pre-rank the whole recall set with the scene's GeneralRankConfs.RankConf, then the actions
    # (sort by a registered ISort, truncate with AdjustCountFilter) before the main rank (user_recommend.go:116,137)
    from pairec_b200.plugin import HostServer
    conf = dict(RECCONF)
    conf["FilterConfs"] = [{"Name": "keep100", "FilterType": "AdjustCountFilter", "RetainNum": 100, "ShuffleItem": False}]
    conf["GeneralRankConfs"] = {"home_feed": {"RankConf": {"RankAlgoList": ["lookup_score"], "RankScore": "${lookup_score}"},
                                              "ActionConfs": [{"ActionType": "sort", "ActionName": "ItemRankScore"},
                                                              {"ActionType": "filter", "ActionName": "keep100"},
                                                              {"ActionType": "bogus", "ActionName": "x"}]}}
    conf["RankConf"] = {"home_feed": {"RankAlgoList": ["lookup_score"], "RankScore": "1 - ${lookup_score}"}}
    s = HostServer(conf)
    try:
        rng = np.random.default_rng(5)
        score = rng.random(400)
        for i in range(400):
            s.add_context_item("mem_recall", "i%07d" % i, 0.0, {"score": float(score[i])})
        s.commit()
        r = s.recommend(scene_id="home_feed", uid="u", size=100)
        # pre-rank keeps the 100 best by score; the main rank then inverts the score, so the final order is ascending
        keep = oracle_lib.go_sort(score)[:100]
        final = 1 - score[keep]
        want = keep[oracle_lib.go_sort(final)]
        assert [int(it["item_id"][1:]) for it in r["items"]] == want.tolist()
        assert any("error to find actionType:bogus" in l for l in r["log"])
    finally:
        s.close()


def test_ingest_formats():
    from pairec_b200.plugin import parse_embedding, recall_cache_roundtrip
    assert parse_embedding("{0.25,-1.5,3e-2}") == [0.25, -1.5, 0.03]
    assert parse_embedding("{1|2|x|4}", "|") == [1.0, 2.0, 0.0, 4.0]      # ParseFloat error -> 0 (dpp_sort.go:228-232)
    cache, back = recall_cache_roundtrip(["a", "b", "c"], [0.5, 0.123456789, 1e-7], "u2i")
    assert cache == "a:u2i:0.5,b:u2i:0.123456789,c:u2i:1e-07"             # fmt %v of float64 (vector_recall.go:107)
    assert [(d["item_id"], d["score"], d["retrieve_id"]) for d in back] == [("a", 0.5, "u2i"), ("b", 0.123456789, "u2i"),
                                                                           ("c", 1e-7, "u2i")]
    # %v = strconv 'g' with the shortest digits: exponent form below 1e-4 and from 1e6 on, whatever the digit count
    # (C's %g would print 100.0 as 1e+02 and 1234567.0 as 1234567)
    vals = [100.0, 20000.0, 999999.0, 1000000.0, 1234567.0, 0.0001, 0.00009, 12.5, -3.0, 0.0, 1e21, 123456789.125]
    want = ["100", "20000", "999999", "1e+06", "1.234567e+06", "0.0001", "9e-05", "12.5", "-3", "0", "1e+21",
            "1.23456789125e+08"]
    cache, back = recall_cache_roundtrip(["i%d" % i for i in range(len(vals))], vals, "m")
    assert cache == ",".join("i%d:m:%s" % (i, w) for i, w in enumerate(want))
    assert [d["score"] for d in back] == vals


def test_dosort_head_truncation_and_embedding_miss_threshold(oracle_lib):
    """DPPSort/SSDSort.doSort before the embeddings are needed (sort/dpp_sort.go:280-300, sort/ssd_sort.go:301-331) and
    loadEmbeddingCache's guard (:246-249): the list doSort holds — and returns unchanged on every error path."""
    from pairec_b200.plugin import dosort_head
    rng = np.random.default_rng(3)
    n, size = 300, 20
    scores = np.round(rng.random(n), 2)          # ties on purpose: the order must be Go's sort order
    all_emb = np.ones(n, dtype=np.uint8)

    def want(cc, msp, always):
        idx = np.arange(n)
        trunc = (cc > 0 or msp > 0) and n > size
        if always or trunc:
            idx = np.asarray(oracle_lib.go_sort(scores))
        if trunc:
            if cc > 0 and max(size, cc) < len(idx):
                idx = idx[:max(size, cc)]
            if msp > 0 and len(idx) > size:
                j = size
                while j < len(idx) and scores[idx[j]] / scores[idx[0]] >= msp:
                    j += 1
                idx = idx[:j]
        return idx

    for cc, msp, always in [(0, 0.0, False), (0, 0.0, True), (100, 0.0, False), (10, 0.0, False), (0, 0.8, False),
                            (150, 0.9, True), (1000, 0.5, False)]:
        got, missed = dosort_head(scores, all_emb, size, cc, msp, always_sort=always)
        assert (got == want(cc, msp, always)).all(), (cc, msp, always)
        assert not missed
    # the guard looks at the TRUNCATED list: 60 % of the best 100 have no embedding -> above the default 0.5
    best100 = want(100, 0.0, False)
    has = np.ones(n, dtype=np.uint8)
    has[best100[:60]] = 0
    got, missed = dosort_head(scores, has, size, 100, 0.0)
    assert missed and (got == best100).all()
    _, missed = dosort_head(scores, has, size, 100, 0.0, miss_threshold=0.7)      # EmbMissedThreshold raised
    assert not missed
    _, missed = dosort_head(scores, has, size, 0, 0.0)                            # 60 of 300 overall: below
    assert not missed
    has[:] = 0
    has[best100[:50]] = 1                                                         # exactly half missing: not ABOVE
    _, missed = dosort_head(scores, has, size, 100, 0.0)
    assert not missed
    got, missed = dosort_head(np.zeros(0), np.zeros(0, dtype=np.uint8), size)
    assert len(got) == 0 and not missed


def test_item_feature_column_sets_encode_to_field_ids():
    """what a FeatureDao fetches per item (feature_hologres_dao.go:644-675: one property per non-NULL
    column; values int / float / string) -> the id-encoded field matrix the gather reads."""
    from pairec_b200.plugin import encode_fields
    spec = [{"column": "category", "vocab": ["news", "sport", "music"]},
            {"column": "brand_id", "id": True},
            {"column": "price_bucket", "vocab": [0, 10, 20.5]},       # numeric vocabulary: matched by utils.ToString
            {"column": "city", "vocab": ["hz", "bj", "hz"]}]           # duplicate entry: the first row wins
    rows = [{"item_id": "i1", "category": "sport", "brand_id": 17, "price_bucket": 10, "city": "hz"},
            {"item_id": "i2", "category": "opera", "brand_id": "42", "price_bucket": 20.5, "city": None},
            {"item_id": "i3", "brand_id": -1, "price_bucket": 10.0},                       # NULL category and city
            {"item_id": "i4", "category": "news", "brand_id": 3.0, "price_bucket": 7, "city": "bj"},
            {"item_id": "i5", "category": "music", "brand_id": 2.5, "price_bucket": "10", "city": "sh"}]
    A = 0xFFFFFFFF
    got = encode_fields(spec, rows)
    want = np.array([[1, 17, 1, 0],
                     [A, 42, 2, A],       # unknown category; id given as a decimal string; float value; NULL city
                     [A, A, 1, A],        # negative id is invalid; 10.0 prints as "10" (FormatFloat 'f', -1)
                     [0, 3, A, 1],        # 3.0 is a valid id; 7 is not in the vocabulary
                     [2, A, 1, A]],       # 2.5 is not an id; the string "10" matches the numeric entry 10
                    dtype=np.uint32)
    assert got.shape == (5, 4)
    assert (got == want).all(), got
    assert encode_fields(spec, []).shape[0] == 0


def test_easyrec_generator_builds_the_columnar_request():
    """RankConf.Processor == "EasyRec" (service/rank/rank_service.go:173-213, algo_data.go:173-350): one PBRequest per
    batch with the user features once, the item ids, and a column per configured context / input item feature — the
    request a GPU IAlgorithm receives from the stock RankService (item ids and user features included)."""
    from pairec_b200.plugin import easyrec_requests
    items = [("a", {"brand": "x", "price": 3.5, "cnt": 7}), ("b", {"price": 1.25}), ("c", {"brand": "z", "cnt": 2})]
    user = {"age": 31, "city": "hz", "type": "vip", "level": "3"}
    # neither ContextFeatures nor ItemFeatures: item.GetFeatures() is not even fetched (:207-212)
    r = easyrec_requests(items, user)
    assert r == [{"user_features": user, "item_ids": ["a", "b", "c"], "context_features": {}, "item_features": {}}]
    # (MakeUserFeatures2 is a clone: "type" stays and "3" stays a string, unlike MakeUserFeatures, module/user.go:137-167)
    # configured context features are string typed: a missing value is "" (algo_data.go:159-176, :208-214)
    r = easyrec_requests(items, user, context_features=["brand", "nope"])
    assert r[0]["context_features"] == {"brand": ["x", "", "z"], "nope": ["", "", ""]} and r[0]["item_features"] == {}
    # ItemFeatures ["*"]: the FIRST item's features that are not context features become the input item features, with the
    # default of the type they had there; one request per BatchCount items, the user features in each
    r = easyrec_requests(items, {"age": 31}, context_features=["brand"], item_features=["*"], batch_count=2)
    assert [q["item_ids"] for q in r] == [["a", "b"], ["c"]]
    assert r[0]["context_features"] == {"brand": ["x", ""]} and r[1]["context_features"] == {"brand": ["z"]}
    assert r[0]["item_features"] == {"cnt": [7, 0], "price": [3.5, 1.25]}
    assert r[1]["item_features"] == {"cnt": [2], "price": [0]}
    assert all(q["user_features"] == {"age": 31} for q in r)
    # named input item features are string typed too
    r = easyrec_requests(items, {}, item_features=["price", "zzz"])
    assert r[0]["item_features"] == {"price": [3.5, 1.25, ""], "zzz": ["", "", ""]} and r[0]["context_features"] == {}
    assert easyrec_requests([], user) == []


def test_user_features_reach_the_generic_processor_as_make_user_features_builds_them(server):
    """module/user.go:137-159: a user property that is a numeric STRING reaches the algorithm as float64 and "type" is
    dropped; an item feature of the same name wins the merge (service/rank/algo_data.go:104-118).  LOOKUP reads the merged
    map: items without a "score" property take the user's."""
    _catalog(server)
    resp = server.recommend(scene_id="home_feed", uid="u1", size=1000, features={"score": "0.875", "type": "vip"})
    by_id = {it["item_id"]: it["score"] for it in resp["items"]}
    assert by_id["i%07d" % 97] == 0.875 and by_id["i%07d" % 194] == 0.875      # no item "score": the user's, parsed
    assert by_id["i%07d" % 1] != 0.875                                          # the item's own score wins
    # a non-numeric string stays a string: LOOKUP then fails the float64 assertion for those batches (the reference
    # panics there, algorithm/lookup.go:45; the mirror logs and leaves the scores)
    resp = server.recommend(scene_id="home_feed", uid="u1", size=5, features={"score": "n/a"})
    assert any("not float64" in l for l in resp["log"]), resp["log"]


def test_easyrec_processor_hands_the_algorithm_a_pbrequest_not_feature_maps():
    """RankConf.Processor == "EasyRec": algorithm.Run receives *easyrec.PBRequest (rank_service.go:273 over
    algo_data.go:84-86).  LOOKUP type-asserts []map[string]interface{} (algorithm/lookup.go:38; the reference panics), so
    behind this processor it fails per batch and the items keep their scores — the error is logged, nothing aborts."""
    from pairec_b200.plugin import HostServer
    conf = json.loads(json.dumps(RECCONF))
    conf["RankConf"]["home_feed"]["Processor"] = "EasyRec"
    conf["RankConf"]["home_feed"]["BatchCount"] = 400
    s = HostServer(conf)
    try:
        _catalog(s)
        resp = s.recommend(scene_id="home_feed", uid="u1", size=10)
        assert resp["size"] == 10
        errs = [l for l in resp["log"] if "algoData is not []map" in l]
        assert len(errs) == 3, resp["log"]          # 1000 unique items / 400 per batch
    finally:
        s.close()


def test_alink_fm_response_score_flip():
    """algorithm/eas/fm_response.go:28-53: the ALINK_FM processor answers with the predicted label and the probability of
    THAT label; GetScore() is result == 0 ? 1 - score : score.  (The in-process FM emits P(label 1) directly; note that the
    wire form does not round-trip: 1 - 0.7 is 0.30000000000000004.)"""
    from pairec_b200.plugin import HostError, alink_fm_scores
    body = json.dumps([{"prediction_result": 1, "prediction_score": 0.8},
                       {"prediction_result": 0.0, "prediction_score": 0.7, "prediction_detail": "{...}"},
                       {"prediction_result": 0, "prediction_score": 1.0}])
    assert alink_fm_scores(body) == [0.8, 1 - 0.7, 0.0]
    assert alink_fm_scores("[]") == []
    with pytest.raises(HostError, match="body:"):
        alink_fm_scores('[{"prediction_result": 1,')


def test_user_vector_elements_are_rounded_to_float32_once():
    """service/recall/vector_recall.go:78-79: strconv.ParseFloat(v, 32) rounds the decimal text to float32 in ONE step (and
    the ignored error leaves 0).  Going through a double first rounds twice: just below the midpoint of two floats the
    double lands ON the midpoint and the tie then goes to the even neighbour — the wrong one."""
    from pairec_b200.plugin import parse_float32
    s = "1.000000178813934326171874999"                     # a hair below 1 + 1.5 * 2^-23
    assert parse_float32(s) == float(np.float32(1 + 2.0 ** -23))
    assert float(np.float32(float(s))) == float(np.float32(1 + 2.0 ** -22))     # what double rounding would give
    assert parse_float32("0.25") == 0.25 and parse_float32("1e-3") == float(np.float32(1e-3))
    assert parse_float32("abc") == 0.0 and parse_float32("1.5x") == 0.0 and parse_float32("") == 0.0


def test_tfserving_response_flattens_row_major():
    """algorithm/tfserving/response.go:51-63: one single-score response per VALUE of Outputs, rows first — the layout of
    prg_rank_ex's score map, so a one-output tower yields one response per item in item order."""
    from pairec_b200.plugin import tfserving_scores
    assert tfserving_scores([[0.1], [0.7], [0.3]]) == [0.1, 0.7, 0.3]
    assert tfserving_scores([[0.1, 0.9], [0.7, 0.2]]) == [0.1, 0.9, 0.7, 0.2]
