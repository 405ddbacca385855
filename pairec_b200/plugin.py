"""ctypes front-end of libpairec_host.so — the C++ mirror of the reference's plugin interfaces (pairec_b200/host/).

`HostServer` plays the role of the pairec process for the hot path: it loads ONE recconf JSON, registers the
in-memory / GPU-backed plugins under the names the JSON binds, and answers /api/recommend-shaped requests.
"""
import ctypes as C
import json
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None


def load_host_library():
    global _lib
    if _lib is None:
        p = os.path.join(_HERE, "libpairec_host.so")
        if not os.path.exists(p):
            raise OSError(f"{p} is missing: run __graft_entry__.build() (make -C pairec_b200/host)")
        # the host library links libpairec_gpu.so ($ORIGIN rpath)
        _lib = C.CDLL(p)
        _lib.ph_last_error.restype = C.c_char_p
        _lib.ph_recommend.restype = C.c_longlong
        _lib.ph_destroy.restype = None
    return _lib


class HostError(RuntimeError):
    pass


class HostServer:
    def __init__(self, recconf):
        self._lib = load_host_library()
        self._h = C.c_void_p(0)
        js = recconf if isinstance(recconf, str) else json.dumps(recconf)
        if self._lib.ph_create(js.encode(), C.byref(self._h)) != 0:
            raise HostError(self._lib.ph_last_error().decode())

    def _ck(self, rc):
        if rc != 0:
            raise HostError(self._lib.ph_last_error().decode())

    def close(self):
        if self._h:
            self._lib.ph_destroy(self._h)
            self._h = C.c_void_p(0)

    def add_context_item(self, recall_name, item_id, score=0.0, properties=None):
        self._ck(self._lib.ph_add_context_item(self._h, recall_name.encode(), item_id.encode(), C.c_double(score),
                                               json.dumps(properties or {}).encode()))

    def commit(self):
        self._ck(self._lib.ph_commit(self._h))

    def attach_engine(self, engine, ids):
        """Binds a pairec_b200.Engine (GPU) and the row -> item-id catalog."""
        self._ck(self._lib.ph_attach_engine(self._h, engine._h))
        arr = (C.c_char_p * len(ids))(*[s.encode() for s in ids])
        self._ck(self._lib.ph_set_catalog_ids(self._h, arr, C.c_uint64(len(ids))))

    def register_gpu_plugins(self, recall_algo="", rank_algo="", model=0, dpp_sort=""):
        self._ck(self._lib.ph_register_gpu_plugins(self._h, recall_algo.encode(), rank_algo.encode(), C.c_int(model),
                                                   dpp_sort.encode()))

    def register_gpu_rank(self, scene, name, model, user_fields=(), dense_columns=(), heads=1):
        """The GPU rank as the scene's rank.IRank (service/rank/custom_rank.go:8-13): Items + User in one call per request."""
        self._ck(self._lib.ph_register_gpu_rank(self._h, scene.encode(), name.encode(), C.c_int(model),
                                                json.dumps(list(user_fields)).encode(),
                                                json.dumps(list(dense_columns)).encode(), C.c_int(heads)))

    def register_gpu_easyrec(self, algo_name, model, user_fields=(), dense_columns=(), outputs=()):
        """The GPU rank as an IAlgorithm for scenes with RankConf.Processor == "EasyRec": RankService hands it one
        easyrec.PBRequest per batch (item ids + the request's user features, service/rank/algo_data.go:292-325)."""
        self._ck(self._lib.ph_register_gpu_easyrec(self._h, algo_name.encode(), C.c_int(model),
                                                   json.dumps(list(user_fields)).encode(),
                                                   json.dumps(list(dense_columns)).encode(),
                                                   json.dumps(list(outputs)).encode()))

    def register_embedding_hook(self, hook_name, ids, emb):
        """sort.RegisterEmbeddingHook(hook_name, fn) with fn = lookup of the item id in emb [len(ids), dim] (f64)."""
        import numpy as np
        emb = np.ascontiguousarray(emb, dtype=np.float64)
        arr = (C.c_char_p * len(ids))(*[s.encode() for s in ids])
        self._ck(self._lib.ph_register_embedding_hook(self._h, hook_name.encode(), arr, emb.ctypes.data_as(C.c_void_p),
                                                      C.c_ulonglong(len(ids)), C.c_int(emb.shape[1])))

    def set_user_vector(self, uid, vector):
        s = " ".join(f"{i}:{float(v)!r}" for i, v in enumerate(vector))
        self._ck(self._lib.ph_set_user_vector(self._h, uid.encode(), s.encode()))

    def recommend(self, **param):
        """RecommendParam (scene_id, category, uid, size, debug, features) -> RecommendResponse dict."""
        req = json.dumps(param).encode()
        need = self._lib.ph_recommend(self._h, req, None, C.c_ulonglong(0))
        if need < 0:
            raise HostError(self._lib.ph_last_error().decode())
        buf = C.create_string_buffer(int(need))
        self._lib.ph_recommend(self._h, req, buf, C.c_ulonglong(need))
        return json.loads(buf.value.decode())


def eval_expr(expr, values):
    """utils/ast expression over named values."""
    lib = load_host_library()
    names = list(values)
    arr = (C.c_char_p * len(names))(*[n.encode() for n in names])
    vals = (C.c_double * len(names))(*[float(values[n]) for n in names])
    out = C.c_double(0)
    if lib.ph_eval_expr(expr.encode(), arr, vals, C.c_int(len(names)), C.byref(out)) != 0:
        raise HostError(lib.ph_last_error().decode())
    return out.value


def easyrec_requests(items, user=None, context_features=(), item_features=(), batch_count=100):
    """rank::EasyrecAlgoDataGenerator (service/rank/algo_data.go:173-350) over items [(item_id, properties), ...]: the
    PBRequests RankService would hand to algorithm.Run for one request, one dict per batch."""
    lib = load_host_library()
    lib.ph_easyrec_requests.restype = C.c_longlong
    args = (json.dumps(list(context_features)).encode(), json.dumps(list(item_features)).encode(),
            json.dumps([{"item_id": i, "properties": p or {}} for i, p in items]).encode(),
            json.dumps(user or {}).encode(), C.c_int(batch_count))
    need = lib.ph_easyrec_requests(*args, None, C.c_ulonglong(0))
    if need < 0:
        raise HostError(lib.ph_last_error().decode())
    buf = C.create_string_buffer(int(need))
    lib.ph_easyrec_requests(*args, buf, C.c_ulonglong(need))
    return json.loads(buf.value.decode())


def alink_fm_scores(body):
    """alinkFMResponseFunc + GetScore (algorithm/eas/fm_response.go:28-53): the ALINK_FM response body -> P(label 1) per item."""
    lib = load_host_library()
    lib.ph_alink_fm_scores.restype = C.c_longlong
    n = lib.ph_alink_fm_scores(body.encode(), None, C.c_ulonglong(0))
    if n < 0:
        raise HostError(lib.ph_last_error().decode())
    buf = (C.c_double * int(n))()
    lib.ph_alink_fm_scores(body.encode(), buf, C.c_ulonglong(n))
    return list(buf)


def tfserving_scores(outputs):
    """tfservingResponseFunc (algorithm/tfserving/response.go:51-63): Outputs [][]float64 -> one score per value."""
    import numpy as np
    lib = load_host_library()
    lib.ph_tfserving_scores.restype = C.c_longlong
    o = np.ascontiguousarray(outputs, dtype=np.float64).reshape(len(outputs), -1)
    buf = (C.c_double * max(1, o.size))()
    n = lib.ph_tfserving_scores(o.ctypes.data_as(C.c_void_p), C.c_int(o.shape[0]), C.c_int(o.shape[1]), buf, C.c_ulonglong(o.size))
    if n < 0:
        raise HostError(lib.ph_last_error().decode())
    return list(buf)[:int(n)]


def parse_float32(text):
    """strconv.ParseFloat(text, 32) with the error ignored, as vector_recall.go:78-79 reads a user-vector element."""
    lib = load_host_library()
    lib.ph_parse_float32.restype = C.c_float
    return float(lib.ph_parse_float32(text.encode()))


def parse_embedding(text, sep=","):
    """sort/dpp_sort.go:224-233 embedding text ("{v1,v2,...}") -> list of float."""
    lib = load_host_library()
    lib.ph_parse_embedding.restype = C.c_longlong
    n = lib.ph_parse_embedding(text.encode(), sep.encode(), None, C.c_ulonglong(0))
    buf = (C.c_double * int(n))()
    lib.ph_parse_embedding(text.encode(), sep.encode(), buf, C.c_ulonglong(n))
    return list(buf)


def recall_cache_roundtrip(ids, scores, model):
    """Formats the recall-result cache string (vector_recall.go:103-110) and parses it back (:35-58)."""
    lib = load_host_library()
    lib.ph_recall_cache_roundtrip.restype = C.c_longlong
    arr = (C.c_char_p * len(ids))(*[i.encode() for i in ids])
    sc = (C.c_double * len(ids))(*[float(s) for s in scores])
    out = C.create_string_buffer(1 << 20)
    cache = C.create_string_buffer(1 << 20)
    lib.ph_recall_cache_roundtrip(arr, sc, C.c_int(len(ids)), model.encode(), out, C.c_ulonglong(1 << 20), cache,
                                  C.c_ulonglong(1 << 20))
    return cache.value.decode(), json.loads(out.value.decode())


def dosort_head(scores, has_emb, size, candidate_count=0, min_score_percent=0.0, miss_threshold=0.5, always_sort=False):
    """The head of DPPSort/SSDSort.doSort (sort/dpp_sort.go:280-300, sort/ssd_sort.go:301-331) plus the embedding-miss
    guard (:246-249): returns (indices of the list the reference holds when it loads embeddings, missed flag)."""
    import numpy as np
    lib = load_host_library()
    lib.ph_dosort_head.restype = C.c_longlong
    sc = np.ascontiguousarray(scores, dtype=np.float64)
    he = np.ascontiguousarray(has_emb, dtype=np.uint8)
    n = int(sc.shape[0])
    out = np.empty(max(n, 1), dtype=np.int32)
    missed = C.c_int(0)
    cnt = lib.ph_dosort_head(sc.ctypes.data_as(C.c_void_p), he.ctypes.data_as(C.c_void_p), C.c_int(n), C.c_int(size),
                             C.c_int(candidate_count), C.c_double(min_score_percent), C.c_double(miss_threshold),
                             C.c_int(1 if always_sort else 0), out.ctypes.data_as(C.c_void_p), C.byref(missed))
    if cnt < 0:
        raise HostError(lib.ph_last_error().decode())
    return out[:cnt].copy(), bool(missed.value)


def encode_fields(spec, rows):
    """Item-feature column sets -> the u32 field matrix of Engine.set_item_fields (host mirror ingest::FieldEncoder).
    spec: [{"column": name, "vocab": [values...]} | {"column": name, "id": True}, ...];
    rows: one dict per item with the properties a FeatureDao fetched (module/feature_hologres_dao.go:644-675); a NULL
    column is absent or None.  Absent / unknown values encode as 0xFFFFFFFF."""
    import numpy as np
    lib = load_host_library()
    lib.ph_encode_fields.restype = C.c_longlong
    sj = json.dumps(spec).encode()
    rj = json.dumps(rows).encode()
    n = lib.ph_encode_fields(sj, rj, None, C.c_ulonglong(0))
    if n < 0:
        raise HostError(lib.ph_last_error().decode())
    out = np.empty(max(int(n), 1), dtype=np.uint32)
    lib.ph_encode_fields(sj, rj, out.ctypes.data_as(C.c_void_p), C.c_ulonglong(n))
    F = max(1, len(spec))
    return out[:n].reshape(-1, F).copy()
