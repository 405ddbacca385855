"""ctypes binding of libpairec_gpu.so (include/pairec_gpu.h) — what the cgo stub in INTEGRATION.md does from Go.

Arrays cross the boundary as raw pointers: numpy arrays for PRG_MEM_HOST, integer device addresses
(`tensor.data_ptr()`) for PRG_MEM_DEVICE.  torch is never imported here.
"""
import ctypes as C
import os

import numpy as np

MEM_HOST, MEM_DEVICE = 0, 1
F32, F64 = 0, 1
MODEL_FM, MODEL_MLP, MODEL_FM_MLP = 0, 1, 2

_HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None

EXPORTS = [
    "prg_init", "prg_destroy", "prg_last_error", "prg_version", "prg_sync", "prg_stream", "prg_set_item_matrix",
    "prg_stage_item_matrix", "prg_commit_item_matrix",
    "prg_set_item_fields", "prg_set_feature_table", "prg_set_fm_bias", "prg_set_mlp", "prg_set_diversity_matrix",
    "prg_recall_topk", "prg_recall_local_keys", "prg_merge_keys", "prg_shard_sample_len", "prg_shard_sample",
    "prg_shard_candidates", "prg_shard_check", "prg_rank", "prg_sort_desc_host", "prg_sort_desc",
    "prg_dpp", "prg_ssd", "prg_recommend", "prg_recommend_from_keys", "prg_lookup", "prg_launch_count", "prg_recall_stats", "prg_recall_filter", "prg_timing",
    "prg_batcher_start", "prg_batcher_recommend", "prg_batcher_stats", "prg_batcher_stop", "prg_batcher_drive",
    "prg_set_user_fields", "prg_set_rank_score", "prg_rank_ex", "prg_recommend_ex", "prg_recommend_from_keys_ex",
    "prg_batcher_recommend_ex", "prg_item_dim", "prg_dpp_ex",
    "prg_set_prerank", "prg_shard_pack_owner", "prg_shard_check_owner", "prg_group_create", "prg_group_size", "prg_group_recommend", "prg_group_destroy",
]


class PrgError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libpairec_gpu error {code}: {msg}")
        self.code = code


class DppParams(C.Structure):
    """prg_dpp_params (DPPSortConfig / abtest knobs of sort/dpp_sort.go)."""
    _fields_ = [("alpha", C.c_double), ("top_n", C.c_int32), ("window_size", C.c_int32), ("norm_mode", C.c_int32),
                ("normalize_emb", C.c_int32), ("candidate_count", C.c_int32), ("min_score_percent", C.c_double),
                ("no_positive_sim", C.c_int32), ("reserved", C.c_int32)]

    def __init__(self, alpha=1.0, top_n=50, window_size=10, norm_mode=0, normalize_emb=1, candidate_count=0,
                 min_score_percent=0.0, no_positive_sim=0):
        super().__init__(alpha, top_n, window_size, norm_mode, normalize_emb, candidate_count, min_score_percent,
                         no_positive_sim, 0)


class SsdParams(C.Structure):
    """prg_ssd_params (SSDSortConfig / abtest knobs of sort/ssd_sort.go)."""
    _fields_ = [("gamma", C.c_double), ("top_n", C.c_int32), ("window_size", C.c_int32), ("norm_mode", C.c_int32),
                ("normalize_emb", C.c_int32), ("use_ssd_star", C.c_int32), ("candidate_count", C.c_int32),
                ("min_score_percent", C.c_double)]

    def __init__(self, gamma=0.25, top_n=50, window_size=5, norm_mode=0, normalize_emb=1, use_ssd_star=0,
                 candidate_count=0, min_score_percent=0.0):
        super().__init__(gamma, top_n, window_size, norm_mode, normalize_emb, use_ssd_star, candidate_count,
                         min_score_percent)


class UserFeatures(C.Structure):
    """prg_user_features: a batch's user / context features (service/rank/algo_data.go:104-118)."""
    _fields_ = [("ids", C.c_void_p), ("dense", C.c_void_p)]


class BatcherConfig(C.Structure):
    """prg_batcher_config."""
    _fields_ = [("max_batch", C.c_int32), ("max_wait_us", C.c_int32), ("recall_k", C.c_int32), ("model", C.c_int32),
                ("dpp", DppParams)]


def lib_path():
    # PRG_LIB selects another build of the same library (kernel-variant experiments); never a different backend
    return os.environ.get("PRG_LIB") or os.path.join(_HERE, "libpairec_gpu.so")


def load_library():
    """Loads the CUDA library; raises if it has not been built (no fallback of any kind)."""
    global _lib
    if _lib is None:
        p = lib_path()
        if not os.path.exists(p):
            raise OSError(f"{p} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(make -C pairec_b200/csrc). pairec_b200 has no CPU fallback.")
        _lib = C.CDLL(p)
        _lib.prg_last_error.restype = C.c_char_p
        _lib.prg_version.restype = C.c_char_p
        _lib.prg_stream.restype = C.c_void_p
        _lib.prg_launch_count.restype = C.c_uint64
        _lib.prg_item_dim.restype = C.c_uint32
        _lib.prg_destroy.restype = None
        _lib.prg_batcher_stop.restype = None
        _lib.prg_group_destroy.restype = None
        for name in EXPORTS:
            getattr(_lib, name)  # raises AttributeError if a declared symbol is not exported
    return _lib


def _ptr(a):
    """numpy array -> void*, int (device address) -> void*, None -> NULL."""
    if a is None:
        return C.c_void_p(0)
    if isinstance(a, np.ndarray):
        assert a.flags["C_CONTIGUOUS"], "array must be C-contiguous"
        return C.c_void_p(a.ctypes.data)
    return C.c_void_p(int(a))


def _np(a, dt):
    return np.ascontiguousarray(a, dtype=dt)


class Engine:
    """One prg_handle (one GPU)."""

    def __init__(self, device=0, **cfg):
        self._lib = load_library()
        self._h = C.c_void_p(0)
        import json
        cfg = dict(cfg, device=device)
        if os.environ.get("PRG_CFG"):  # experiment switches (A/B of kernel variants), e.g. '{"mlp_one_tile":1}'
            cfg.update(json.loads(os.environ["PRG_CFG"]))
        rc = self._lib.prg_init(json.dumps(cfg).encode(), C.byref(self._h))
        if rc != 0:
            raise PrgError(rc, self._lib.prg_last_error().decode())
        self.dim = 0
        self.n_fields = 0
        self.fdim = 0

    def _ck(self, rc):
        if rc != 0:
            raise PrgError(rc, self._lib.prg_last_error().decode())

    def close(self):
        if self._h:
            self._lib.prg_destroy(self._h)
            self._h = C.c_void_p(0)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def sync(self):
        self._ck(self._lib.prg_sync(self._h))

    @property
    def stream(self):
        return self._lib.prg_stream(self._h)

    @property
    def launches(self):
        return int(self._lib.prg_launch_count(self._h))

    def recall_stats(self):
        a, b = C.c_int32(0), C.c_int32(0)
        self._ck(self._lib.prg_recall_stats(self._h, C.byref(a), C.byref(b)))
        f = C.c_int32(0)
        self._ck(self._lib.prg_recall_filter(self._h, C.byref(f)))
        return {"fallback_queries": a.value, "max_candidates": b.value,
                "filter": ("none", "ffma2", "tf32", "bf16", "int8")[f.value]}

    STAGES = ("scan", "scan_dense", "select", "gather_fm", "mlp", "sort", "dpp", "other")

    def timing(self, enable, read=False):
        """Per-stage device milliseconds / span counts measured with CUDA events inside the library."""
        if not read:
            self._ck(self._lib.prg_timing(self._h, C.c_int(int(enable)), None, None))
            return None
        ms = (C.c_double * 8)()
        n = (C.c_uint64 * 8)()
        self._ck(self._lib.prg_timing(self._h, C.c_int(int(enable)), ms, n))
        return {s: {"ms": ms[i], "spans": int(n[i])} for i, s in enumerate(self.STAGES)}

    # ---------------------------------------------------------------- tables
    def set_item_matrix(self, data, rows=None, dim=None, row_base=0, mem=MEM_HOST):
        if mem == MEM_HOST:
            data = _np(data, np.float32)
            rows, dim = data.shape
        self._ck(self._lib.prg_set_item_matrix(self._h, _ptr(data), C.c_uint64(rows), C.c_uint32(dim),
                                               C.c_uint64(row_base), C.c_int(mem)))
        self.dim = dim

    def stage_item_matrix(self, data, rows=None, dim=None, row_base=0, mem=MEM_HOST):
        """Upload + index a new item matrix beside the live one (no handle lock held); commit_item_matrix swaps."""
        if mem == MEM_HOST:
            data = _np(data, np.float32)
            rows, dim = data.shape
        self._ck(self._lib.prg_stage_item_matrix(self._h, _ptr(data), C.c_uint64(rows), C.c_uint32(dim),
                                                 C.c_uint64(row_base), C.c_int(mem)))
        self._staged_dim = dim

    def commit_item_matrix(self):
        self._ck(self._lib.prg_commit_item_matrix(self._h))
        self.dim = self._staged_dim

    def set_item_fields(self, ids, rows=None, n_fields=None, mem=MEM_HOST):
        if mem == MEM_HOST:
            ids = _np(ids, np.uint32)
            rows, n_fields = ids.shape
        self._ck(self._lib.prg_set_item_fields(self._h, _ptr(ids), C.c_uint64(rows), C.c_uint32(n_fields), C.c_int(mem)))
        self.n_fields = n_fields

    def set_feature_table(self, table, factors, linear=None, rows=None, fdim=None, mem=MEM_HOST):
        if mem == MEM_HOST:
            factors = _np(factors, np.float32)
            rows, fdim = factors.shape
            linear = None if linear is None else _np(linear, np.float32)
        self._ck(self._lib.prg_set_feature_table(self._h, C.c_int(table), _ptr(factors), _ptr(linear), C.c_uint64(rows),
                                                 C.c_uint32(fdim), C.c_int(mem)))
        self.fdim = fdim

    def set_user_fields(self, n_user_fields, n_user_dense=0):
        """User fields use feature tables n_fields .. n_fields + n_user_fields - 1; call before set_mlp."""
        self._ck(self._lib.prg_set_user_fields(self._h, C.c_uint32(n_user_fields), C.c_uint32(n_user_dense)))
        self.n_user_fields, self.n_user_dense = n_user_fields, n_user_dense

    def set_prerank(self, model, keep):
        """General (pre-)rank stage of the fused path: `model` scores the recall set, the best `keep` go on (0 = off)."""
        self._ck(self._lib.prg_set_prerank(self._h, C.c_int(model), C.c_int(keep)))

    def set_rank_score(self, coef):
        c = _np(coef, np.float64)
        self._ck(self._lib.prg_set_rank_score(self._h, _ptr(c), C.c_int(c.shape[0])))

    def _user(self, B, user_ids, user_dense):
        """-> (prg_user_features*, keep-alive arrays)"""
        if user_ids is None and user_dense is None:
            return None, ()
        ids = None if user_ids is None else _np(user_ids, np.uint32).reshape(B, -1)
        dense = None if user_dense is None else _np(user_dense, np.float32).reshape(B, -1)
        if ids is not None:
            assert ids.shape[1] == getattr(self, "n_user_fields", 0), "user_ids must be [B, n_user_fields]"
        if dense is not None:
            assert dense.shape[1] == getattr(self, "n_user_dense", 0), "user_dense must be [B, n_user_dense]"
        uf = UserFeatures(None if ids is None else ids.ctypes.data, None if dense is None else dense.ctypes.data)
        return C.byref(uf), (uf, ids, dense)

    def set_fm_bias(self, w0):
        self._ck(self._lib.prg_set_fm_bias(self._h, C.c_float(w0)))

    def set_mlp(self, dims, W, bias):
        L = len(W)
        Wc = [_np(w, np.uint16) for w in W]
        bc = [_np(b, np.float32) for b in bias]
        wp = (C.c_void_p * L)(*[w.ctypes.data for w in Wc])
        bp = (C.c_void_p * L)(*[b.ctypes.data for b in bc])
        d = np.array(dims, dtype=np.uint32)
        self._ck(self._lib.prg_set_mlp(self._h, C.c_int(L), _ptr(d), wp, bp))
        self.heads = int(dims[-1])

    def set_diversity_matrix(self, data, rows=None, dim=None, dtype=None, mem=MEM_HOST):
        if mem == MEM_HOST:
            data = np.ascontiguousarray(data)
            assert data.dtype in (np.float32, np.float64)
            dtype = F32 if data.dtype == np.float32 else F64
            rows, dim = data.shape
        self._ck(self._lib.prg_set_diversity_matrix(self._h, _ptr(data), C.c_uint64(rows), C.c_uint32(dim),
                                                    C.c_int(dtype), C.c_int(mem)))

    # ---------------------------------------------------------------- recall
    def recall_topk(self, q, k):
        """Host buffers in/out (the e2e path).  Returns rows u32 [B,k], scores f32 [B,k], n i32 [B]."""
        q = self._queries(q)
        B = q.shape[0]
        rows = np.empty((B, k), dtype=np.uint32)
        scores = np.empty((B, k), dtype=np.float32)
        n = np.empty(B, dtype=np.int32)
        self._ck(self._lib.prg_recall_topk(self._h, _ptr(q), C.c_int(B), C.c_int(k), _ptr(rows), _ptr(scores), _ptr(n),
                                           C.c_int(MEM_HOST)))
        return rows, scores, n

    def _queries(self, q):
        """The C ABI takes no vector length: a query of the wrong width would be read past its end
        (service/recall/vector_recall.go:72-82 silently skips malformed 'i:v' pairs, so short vectors do occur)."""
        q = _np(q, np.float32)
        dim = int(self._lib.prg_item_dim(self._h))
        if q.ndim != 2 or q.shape[1] != dim:
            raise PrgError(1, f"query vectors must be [B, {dim}], got {q.shape}")
        return q

    def recall_topk_dev(self, q_ptr, B, k, rows_ptr, scores_ptr, n_ptr):
        self._ck(self._lib.prg_recall_topk(self._h, _ptr(q_ptr), C.c_int(B), C.c_int(k), _ptr(rows_ptr),
                                           _ptr(scores_ptr), _ptr(n_ptr), C.c_int(MEM_DEVICE)))

    def recall_local_keys_dev(self, q_ptr, B, k, keys_ptr):
        self._ck(self._lib.prg_recall_local_keys(self._h, _ptr(q_ptr), C.c_int(B), C.c_int(k), _ptr(keys_ptr)))

    # global-threshold sharded recall (device pointers; the two all-gathers are the caller's)
    def shard_sample_len(self, k):
        return int(self._lib.prg_shard_sample_len(C.c_int(k)))

    def shard_sample_dev(self, q_ptr, Bg, k, G, out_ptr):
        self._ck(self._lib.prg_shard_sample(self._h, _ptr(q_ptr), C.c_int(Bg), C.c_int(k), C.c_int(G), _ptr(out_ptr)))

    def shard_candidates_dev(self, q_ptr, Bg, k, G, all_samples_ptr, out_ptr):
        self._ck(self._lib.prg_shard_candidates(self._h, _ptr(q_ptr), C.c_int(Bg), C.c_int(k), C.c_int(G),
                                                _ptr(all_samples_ptr), _ptr(out_ptr)))

    def shard_check_dev(self, gathered_ptr, G, Bg, k, retry_ptr):
        self._ck(self._lib.prg_shard_check(self._h, _ptr(gathered_ptr), C.c_int(G), C.c_int(Bg), C.c_int(k),
                                           _ptr(retry_ptr)))

    def shard_pack_owner_dev(self, cand_ptr, Bg, B, k, out_ptr):
        self._ck(self._lib.prg_shard_pack_owner(self._h, _ptr(cand_ptr), C.c_int(Bg), C.c_int(B), C.c_int(k), _ptr(out_ptr)))

    def shard_check_owner_dev(self, received_ptr, G, B, k, q0, retry_ptr):
        self._ck(self._lib.prg_shard_check_owner(self._h, _ptr(received_ptr), C.c_int(G), C.c_int(B), C.c_int(k), C.c_int(q0),
                                                 _ptr(retry_ptr)))

    def merge_keys(self, keys_ptr, G, B, k, rows=None, scores=None, n=None, mem=MEM_HOST):
        if mem == MEM_HOST:
            rows = np.empty((B, k), dtype=np.uint32)
            scores = np.empty((B, k), dtype=np.float32)
            n = np.empty(B, dtype=np.int32)
        self._ck(self._lib.prg_merge_keys(self._h, _ptr(keys_ptr), C.c_int(G), C.c_int(B), C.c_int(k), _ptr(rows),
                                          _ptr(scores), _ptr(n), C.c_int(mem)))
        return rows, scores, n

    # ---------------------------------------------------------------- rank
    def rank(self, model, rows, user_ids=None, user_dense=None, score_map=False):
        """rows [B, n] -> Item.Score [B, n] (and, score_map=True, every head's score [B, n, heads])."""
        rows = _np(rows, np.uint32)
        B, n = rows.shape
        out = np.empty((B, n), dtype=np.float64)
        uf, keep = self._user(B, user_ids, user_dense)
        smap = np.empty((B, n, getattr(self, "heads", 1)), dtype=np.float64) if score_map else None
        self._ck(self._lib.prg_rank_ex(self._h, C.c_int(model), _ptr(rows), C.c_int(B), C.c_int(n), uf, _ptr(out),
                                       _ptr(smap), C.c_int(MEM_HOST)))
        return (out, smap) if score_map else out

    def rank_dev(self, model, rows_ptr, B, n, out_ptr, user_ids_ptr=None, user_dense_ptr=None):
        uf = None
        if user_ids_ptr is not None or user_dense_ptr is not None:
            uf = C.byref(UserFeatures(user_ids_ptr, user_dense_ptr))
        self._ck(self._lib.prg_rank_ex(self._h, C.c_int(model), _ptr(rows_ptr), C.c_int(B), C.c_int(n), uf, _ptr(out_ptr),
                                       None, C.c_int(MEM_DEVICE)))

    # ---------------------------------------------------------------- sort
    def sort_desc(self, score):
        score = _np(score, np.float64)
        B, n = score.shape
        perm = np.empty((B, n), dtype=np.int32)
        self._ck(self._lib.prg_sort_desc(self._h, _ptr(score), C.c_int(B), C.c_int(n), _ptr(perm), C.c_int(MEM_HOST)))
        return perm

    def sort_desc_dev(self, score_ptr, B, n, perm_ptr):
        self._ck(self._lib.prg_sort_desc(self._h, _ptr(score_ptr), C.c_int(B), C.c_int(n), _ptr(perm_ptr), C.c_int(MEM_DEVICE)))

    # ---------------------------------------------------------------- DPP
    def dpp(self, rows, score, params, hook=None, use_table=True):
        """hook [B, n, hook_dim] f64: embeddings from registered hooks (sort/dpp_sort.go:362-370); use_table=False = hooks only."""
        score = _np(score, np.float64)
        B, n = score.shape
        rows = None if rows is None else _np(rows, np.uint32)
        hook = None if hook is None else _np(hook, np.float64)
        T = params.top_n
        idx = np.full((B, T), -1, dtype=np.int32)
        cnt = np.zeros(B, dtype=np.int32)
        st = np.zeros(B, dtype=np.int32)
        self._ck(self._lib.prg_dpp_ex(self._h, _ptr(rows), _ptr(score), _ptr(hook), C.c_int(0 if hook is None else hook.shape[2]),
                                      C.c_int(1 if use_table else 0), C.c_int(B), C.c_int(n), C.byref(params), _ptr(idx),
                                      _ptr(cnt), _ptr(st), C.c_int(MEM_HOST)))
        return idx, cnt, st

    def ssd(self, rows, score, params):
        rows = _np(rows, np.uint32)
        score = _np(score, np.float64)
        B, n = rows.shape
        T = params.top_n
        idx = np.full((B, T), -1, dtype=np.int32)
        cnt = np.zeros(B, dtype=np.int32)
        st = np.zeros(B, dtype=np.int32)
        self._ck(self._lib.prg_ssd(self._h, _ptr(rows), _ptr(score), C.c_int(B), C.c_int(n), C.byref(params), _ptr(idx),
                                   _ptr(cnt), _ptr(st), C.c_int(MEM_HOST)))
        return idx, cnt, st

    # ---------------------------------------------------------------- fused
    def recommend(self, q, recall_k, model, params, user_ids=None, user_dense=None):
        q = self._queries(q)
        B = q.shape[0]
        T = params.top_n
        rows = np.empty((B, T), dtype=np.uint32)
        scores = np.empty((B, T), dtype=np.float64)
        n = np.empty(B, dtype=np.int32)
        uf, keep = self._user(B, user_ids, user_dense)
        self._ck(self._lib.prg_recommend_ex(self._h, _ptr(q), C.c_int(B), C.c_int(recall_k), C.c_int(model),
                                            C.byref(params), uf, _ptr(rows), _ptr(scores), _ptr(n), C.c_int(MEM_HOST)))
        return rows, scores, n

    def recommend_dev(self, q_ptr, B, recall_k, model, params, rows_ptr, scores_ptr, n_ptr, user_ids_ptr=None,
                      user_dense_ptr=None):
        uf = None
        if user_ids_ptr is not None or user_dense_ptr is not None:
            uf = C.byref(UserFeatures(user_ids_ptr, user_dense_ptr))
        self._ck(self._lib.prg_recommend_ex(self._h, _ptr(q_ptr), C.c_int(B), C.c_int(recall_k), C.c_int(model),
                                            C.byref(params), uf, _ptr(rows_ptr), _ptr(scores_ptr), _ptr(n_ptr),
                                            C.c_int(MEM_DEVICE)))


def _engine_recommend_from_keys_dev(self, keys_ptr, G, g_stride, B, k, model, params, rows_ptr, scores_ptr, n_ptr,
                                    user_ids_ptr=None, user_dense_ptr=None):
    uf = None
    if user_ids_ptr is not None or user_dense_ptr is not None:
        uf = C.byref(UserFeatures(user_ids_ptr, user_dense_ptr))
    self._ck(self._lib.prg_recommend_from_keys_ex(self._h, _ptr(keys_ptr), C.c_int(G), C.c_uint64(g_stride), C.c_int(B),
                                                  C.c_int(k), C.c_int(model), C.byref(params), uf, _ptr(rows_ptr),
                                                  _ptr(scores_ptr), _ptr(n_ptr), C.c_int(MEM_DEVICE)))


Engine.recommend_from_keys_dev = _engine_recommend_from_keys_dev


def sort_desc_host(score):
    """prg_sort_desc_host: Go pdqsort tie order, host only."""
    lib = load_library()
    score = _np(score, np.float64)
    perm = np.empty(score.shape[0], dtype=np.int32)
    rc = lib.prg_sort_desc_host(_ptr(score), C.c_int(score.shape[0]), _ptr(perm))
    if rc != 0:
        raise PrgError(rc, lib.prg_last_error().decode())
    return perm


def lookup(value, present):
    lib = load_library()
    value = _np(value, np.float64)
    present = _np(present, np.uint8)
    out = np.empty_like(value)
    rc = lib.prg_lookup(_ptr(value), _ptr(present), C.c_int(value.shape[0]), _ptr(out))
    if rc != 0:
        raise PrgError(rc, lib.prg_last_error().decode())
    return out


class Group:
    """prg_group: G Engines of this process (one per GPU of the box, or several on one GPU in tests), each holding one row
    shard of the item matrix and replicas of the other tables; one call serves a batch over all of them, the exchanges
    run as peer stores over NVLink inside the library (no NCCL, no torch.distributed)."""

    def __init__(self, engines):
        self._lib = engines[0]._lib
        self._engines = list(engines)   # keep the members alive
        arr = (C.c_void_p * len(engines))(*[e._h for e in engines])
        self._g = C.c_void_p(0)
        rc = self._lib.prg_group_create(arr, C.c_int(len(engines)), C.byref(self._g))
        if rc != 0:
            raise PrgError(rc, self._lib.prg_last_error().decode())

    def recommend(self, q, recall_k, model, params, user_ids=None, user_dense=None):
        """q [n, dim] -> rows u32 [n, top_n], scores f64 [n, top_n], counts i32 [n], redone (bool)."""
        q = self._engines[0]._queries(q)
        n = q.shape[0]
        T = params.top_n
        rows = np.empty((n, T), dtype=np.uint32)
        scores = np.empty((n, T), dtype=np.float64)
        cnt = np.empty(n, dtype=np.int32)
        redone = C.c_int32(0)
        uf, keep = self._engines[0]._user(n, user_ids, user_dense)
        rc = self._lib.prg_group_recommend(self._g, _ptr(q), C.c_int(n), C.c_int(recall_k), C.c_int(model), C.byref(params),
                                           uf, _ptr(rows), _ptr(scores), _ptr(cnt), C.byref(redone))
        if rc != 0:
            raise PrgError(rc, self._lib.prg_last_error().decode())
        return rows, scores, cnt, bool(redone.value)

    def close(self):
        if self._g:
            self._lib.prg_group_destroy(self._g)
            self._g = C.c_void_p(0)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Batcher:
    """prg_batcher: coalesces concurrent single-request calls (one per host thread) into prg_recommend batches.
    ctypes releases the GIL during the call, so Python threads behave like the reference's request goroutines."""

    def __init__(self, engine, recall_k, model, dpp, max_batch=64, max_wait_us=0):
        self._lib = engine._lib
        self._eng = engine
        self.dim = engine.dim
        self.top_n = int(dpp.top_n)
        cfg = BatcherConfig(max_batch, max_wait_us, recall_k, model, dpp)
        self._b = C.c_void_p(0)
        rc = self._lib.prg_batcher_start(engine._h, C.byref(cfg), C.byref(self._b))
        if rc != 0:
            raise PrgError(rc, self._lib.prg_last_error().decode())

    def recommend(self, q, user_ids=None, user_dense=None):
        """One request: q [dim] f32 (+ its user features) -> (rows [n] u32, scores [n] f64).  Blocks until served."""
        q = _np(q, np.float32)
        assert q.shape == (self.dim,)
        rows = np.empty(self.top_n, dtype=np.uint32)
        scores = np.empty(self.top_n, dtype=np.float64)
        n = C.c_int32(0)
        ids = None if user_ids is None else _np(user_ids, np.uint32)
        dense = None if user_dense is None else _np(user_dense, np.float32)
        rc = self._lib.prg_batcher_recommend_ex(self._b, _ptr(q), _ptr(ids), _ptr(dense), _ptr(rows), _ptr(scores),
                                                C.byref(n))
        if rc != 0:
            raise PrgError(rc, self._lib.prg_last_error().decode())
        return rows[:n.value], scores[:n.value]

    def drive(self, q_pool, n_threads, per_thread):
        """Closed-loop load from n_threads native client threads -> (latency_us [n], wall_s, rows [n_pool, top_n], n)."""
        q_pool = _np(q_pool, np.float32)
        n_pool = q_pool.shape[0]
        lat = np.zeros(n_threads * per_thread, dtype=np.float32)
        rows = np.full((n_pool, self.top_n), 0xFFFFFFFF, dtype=np.uint32)
        n = np.zeros(n_pool, dtype=np.int32)
        wall = C.c_double(0)
        rc = self._lib.prg_batcher_drive(self._b, _ptr(q_pool), C.c_int(n_pool), C.c_int(n_threads), C.c_int(per_thread),
                                         _ptr(lat), C.byref(wall), _ptr(rows), _ptr(n))
        if rc != 0:
            raise PrgError(rc, self._lib.prg_last_error().decode())
        return lat, wall.value, rows, n

    def stats(self):
        nr, nb = C.c_uint64(0), C.c_uint64(0)
        hist = (C.c_uint64 * 9)()
        self._lib.prg_batcher_stats(self._b, C.byref(nr), C.byref(nb), hist)
        return {"requests": nr.value, "batches": nb.value, "size_hist": list(hist)}

    def close(self):
        if self._b:
            self._lib.prg_batcher_stop(self._b)
            self._b = C.c_void_p(0)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
