"""ctypes front-end of the CPU oracle (oracle/oracle.c).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this package;
nothing under pairec_b200/ does.  Every function restates the reference file:line cited in oracle.c.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "liboracle.so")
_lib = None


def build(force=False):
    """Compile oracle.c with gcc (recipe: oracle/Makefile)."""
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(
            os.path.join(_HERE, "oracle.c")):
        subprocess.run(["make", "-C", _HERE], check=True, stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.orc_key_score.restype = C.c_float
        _lib.orc_key_score.argtypes = [C.c_uint64]
        _lib.orc_key_row.restype = C.c_uint32
        _lib.orc_key_row.argtypes = [C.c_uint64]
        _lib.orc_make_key.restype = C.c_uint64
        _lib.orc_make_key.argtypes = [C.c_float, C.c_uint32]
        _lib.orc_sigmoid.restype = C.c_float
        _lib.orc_exp.restype = C.c_double
        _lib.orc_exp.argtypes = [C.c_double]
        _lib.orc_sigmoid.argtypes = [C.c_float]
        _lib.orc_rank_score_expr.restype = C.c_double
        _lib.orc_bf16_to_f32.restype = C.c_float
        _lib.orc_bf16_to_f32.argtypes = [C.c_uint16]
        _lib.orc_f32_to_bf16.restype = C.c_uint16
        _lib.orc_f32_to_bf16.argtypes = [C.c_float]
    return _lib


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


class DppParams(C.Structure):
    _fields_ = [("alpha", C.c_double), ("top_n", C.c_int32), ("window_size", C.c_int32), ("norm_mode", C.c_int32),
                ("normalize_emb", C.c_int32), ("candidate_count", C.c_int32), ("min_score_percent", C.c_double)]


def num_threads():
    return int(lib().orc_num_threads())


# ------------------------------------------------------------------ keys
def keys_split(keys):
    """u64 order keys -> (rows u32, scores f32, n valid per list)."""
    keys = np.ascontiguousarray(keys, dtype=np.uint64)
    rows = (np.uint64(0xFFFFFFFF) - (keys & np.uint64(0xFFFFFFFF))).astype(np.uint32)
    o = (keys >> np.uint64(32)).astype(np.uint32)
    u = np.where(o & np.uint32(0x80000000), o & np.uint32(0x7FFFFFFF), ~o).astype(np.uint32)
    u = np.where(o == 1, np.uint32(0x7FC00000), u).astype(np.uint32)
    scores = u.view(np.float32).copy()
    empty = keys == 0
    rows[empty] = 0xFFFFFFFF
    scores[empty] = -np.inf
    n = (~empty).sum(axis=-1).astype(np.int32)
    return rows, scores, n


# ------------------------------------------------------------------ recall
def recall_topk(E, Q, k, row_base=0, n_threads=0):
    E = np.ascontiguousarray(E, dtype=np.float32)
    Q = np.ascontiguousarray(Q, dtype=np.float32)
    B = Q.shape[0]
    out = np.zeros((B, k), dtype=np.uint64)
    rc = lib().orc_recall_topk(_p(E, C.c_float), C.c_uint64(E.shape[0]), C.c_uint32(E.shape[1]), C.c_uint64(row_base),
                               _p(Q, C.c_float), C.c_int(B), C.c_int(k), _p(out, C.c_uint64), C.c_int(n_threads))
    assert rc == 0
    return out


def recall_scores(E, Q):
    E = np.ascontiguousarray(E, dtype=np.float32)
    Q = np.ascontiguousarray(Q, dtype=np.float32)
    out = np.zeros((Q.shape[0], E.shape[0]), dtype=np.float32)
    lib().orc_recall_scores(_p(E, C.c_float), C.c_uint64(E.shape[0]), C.c_uint32(E.shape[1]), _p(Q, C.c_float),
                            C.c_int(Q.shape[0]), _p(out, C.c_float))
    return out


def merge_keys(keys):
    keys = np.ascontiguousarray(keys, dtype=np.uint64)
    G, B, k = keys.shape
    out = np.zeros((B, k), dtype=np.uint64)
    lib().orc_merge_keys(_p(keys, C.c_uint64), C.c_int(G), C.c_int(B), C.c_int(k), _p(out, C.c_uint64))
    return out


# ------------------------------------------------------------------ gather + FM / MLP
def gather_fm(fields, factors, linear, w0, rows, want_x=True, user_ids=None, user_dense=None):
    """One request's candidates.  user_ids [U] / user_dense [n_dense]: that request's user features
    (service/rank/algo_data.go:104-118); the tables of the U user fields follow the F item tables in factors / linear."""
    fields = np.ascontiguousarray(fields, dtype=np.uint32)
    rows = np.ascontiguousarray(rows, dtype=np.uint32).reshape(-1)
    F = fields.shape[1]
    U = 0 if user_ids is None else len(user_ids)
    nd = 0 if user_dense is None else len(user_dense)
    assert len(factors) == F + U
    fdim = factors[0].shape[1]
    fac = [np.ascontiguousarray(f, dtype=np.float32) for f in factors]
    lin = [None if l is None else np.ascontiguousarray(l, dtype=np.float32) for l in linear]
    fp = (C.POINTER(C.c_float) * (F + U))(*[_p(f, C.c_float) for f in fac])
    lp = (C.POINTER(C.c_float) * (F + U))(*[C.POINTER(C.c_float)() if l is None else _p(l, C.c_float) for l in lin])
    tr = np.array([f.shape[0] for f in fac], dtype=np.uint64)
    n = rows.shape[0]
    logit = np.zeros(n, dtype=np.float32)
    x = np.zeros((n, (F + U) * fdim + nd), dtype=np.float32) if want_x else None
    uid = np.ascontiguousarray(user_ids, dtype=np.uint32) if U else None
    ud = np.ascontiguousarray(user_dense, dtype=np.float32) if nd else None
    rc = lib().orc_gather_fm_user(_p(fields, C.c_uint32), C.c_uint64(fields.shape[0]), C.c_uint32(F), fp, lp,
                                  _p(tr, C.c_uint64), C.c_uint32(fdim), C.c_float(w0), _p(rows, C.c_uint32), C.c_int(n),
                                  C.c_uint32(U), _p(uid, C.c_uint32) if U else None, C.c_uint32(nd),
                                  _p(ud, C.c_float) if nd else None,
                                  _p(logit, C.c_float), _p(x, C.c_float) if want_x else None)
    assert rc == 0
    return logit, x


def exp(x):
    """libm exp exactly as oracle.c calls it (element by element; fixtures only)."""
    f = lib().orc_exp
    return np.array([f(float(v)) for v in np.asarray(x, dtype=np.float64).ravel()], dtype=np.float64)


def sigmoid(logit):
    logit = np.asarray(logit, dtype=np.float32)
    return (1.0 / (1.0 + np.exp(-logit.astype(np.float64)))).astype(np.float32)


def f32_to_bf16(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    u = a.view(np.uint32).astype(np.uint64)
    r = ((u + 0x7FFF + ((u >> 16) & 1)) >> 16).astype(np.uint16)
    return r


def bf16_to_f32(h):
    return (np.ascontiguousarray(h, dtype=np.uint16).astype(np.uint32) << 16).view(np.float32)


def mlp_forward(x, dims, W, bias):
    x = np.ascontiguousarray(x, dtype=np.float32)
    n = x.shape[0]
    L = len(W)
    Wc = [np.ascontiguousarray(w, dtype=np.uint16) for w in W]
    bc = [np.ascontiguousarray(b, dtype=np.float32) for b in bias]
    wp = (C.POINTER(C.c_uint16) * L)(*[_p(w, C.c_uint16) for w in Wc])
    bp = (C.POINTER(C.c_float) * L)(*[_p(b, C.c_float) for b in bc])
    d = np.array(dims, dtype=np.uint32)
    O = int(d[-1])
    out = np.zeros((n, O), dtype=np.float32)
    rc = lib().orc_mlp_forward(_p(x, C.c_float), C.c_int(n), C.c_int(L), _p(d, C.c_uint32), wp, bp, _p(out, C.c_float))
    assert rc == 0
    return out[:, 0].copy() if O == 1 else out


# ------------------------------------------------------------------ sort / lookup
def go_sort(score, descending=True):
    score = np.ascontiguousarray(score, dtype=np.float64)
    perm = np.zeros(score.shape[0], dtype=np.int32)
    lib().orc_go_sort(_p(score, C.c_double), C.c_int(score.shape[0]), C.c_int(1 if descending else 0), _p(perm, C.c_int32))
    return perm


def stable_sort_desc(score):
    score = np.ascontiguousarray(score, dtype=np.float64)
    perm = np.zeros(score.shape[0], dtype=np.int32)
    lib().orc_stable_sort_desc(_p(score, C.c_double), C.c_int(score.shape[0]), _p(perm, C.c_int32))
    return perm


def algo_score_sort(score, field, switch_threshold):
    score = np.ascontiguousarray(score, dtype=np.float64)
    field = np.ascontiguousarray(field, dtype=np.float64)
    perm = np.zeros(score.shape[0], dtype=np.int32)
    lib().orc_algo_score_sort(_p(score, C.c_double), _p(field, C.c_double), C.c_int(score.shape[0]),
                              C.c_double(switch_threshold), _p(perm, C.c_int32))
    return perm


def lookup(value, present):
    value = np.ascontiguousarray(value, dtype=np.float64)
    present = np.ascontiguousarray(present, dtype=np.uint8)
    out = np.zeros_like(value)
    lib().orc_lookup(_p(value, C.c_double), _p(present, C.c_uint8), C.c_int(value.shape[0]), _p(out, C.c_double))
    return out


def rank_score_expr(values, coefs):
    v = np.ascontiguousarray(values, dtype=np.float64)
    c = np.ascontiguousarray(coefs, dtype=np.float64)
    return float(lib().orc_rank_score_expr(_p(v, C.c_double), _p(c, C.c_double), C.c_int(v.shape[0])))


# ------------------------------------------------------------------ DPP
def dpp_request(emb, score, top_n, alpha=1.0, window_size=10, norm_mode=0, normalize_emb=1, candidate_count=0,
                min_score_percent=0.0):
    emb = np.ascontiguousarray(emb, dtype=np.float64)
    score = np.ascontiguousarray(score, dtype=np.float64)
    n, D = emb.shape
    p = DppParams(alpha, top_n, window_size, norm_mode, normalize_emb, candidate_count, min_score_percent)
    out = np.full(max(top_n, n), -1, dtype=np.int32)
    st = C.c_int32(0)
    ny = lib().orc_dpp_request(_p(emb, C.c_double), _p(score, C.c_double), C.c_int(n), C.c_int(D), C.byref(p),
                               _p(out, C.c_int32), C.byref(st))
    return out[:ny].copy(), int(st.value)


def dpp_substitute(sub_row, dim):
    """The fixed direction a candidate without an embedding receives (oracle.c orc_dpp_substitute)."""
    fn = lib().orc_dpp_substitute
    fn.restype = C.c_float
    fn.argtypes = [C.c_uint32, C.c_uint32]
    return np.array([fn(sub_row, d) for d in range(dim)], dtype=np.float32)


def dpp_request_ex(emb, score, top_n, present=None, hook=None, use_table=True, no_positive_sim=0, alpha=1.0,
                   window_size=10, norm_mode=0, normalize_emb=1, candidate_count=0, min_score_percent=0.0, want_L=False):
    """orc_dpp_request_ex: missing embeddings (present mask), hook embeddings [n, hook_dim], hook-only path."""
    score = np.ascontiguousarray(score, dtype=np.float64)
    n = score.shape[0]
    emb = None if emb is None else np.ascontiguousarray(emb, dtype=np.float64)
    D = 0 if emb is None else emb.shape[1]
    hook = None if hook is None else np.ascontiguousarray(hook, dtype=np.float64)
    present = None if present is None else np.ascontiguousarray(present, dtype=np.uint8)
    p = DppParams(alpha, top_n, window_size, norm_mode, normalize_emb, candidate_count, min_score_percent)
    out = np.full(max(top_n, n), -1, dtype=np.int32)
    st = C.c_int32(0)
    Ld = np.full(n, np.nan) if want_L else None
    L0 = np.full(n, np.nan) if want_L else None
    ny = lib().orc_dpp_request_ex(None if emb is None else _p(emb, C.c_double),
                                  None if present is None else _p(present, C.c_uint8),
                                  None if hook is None else _p(hook, C.c_double),
                                  C.c_int(0 if hook is None else hook.shape[1]), C.c_int(1 if use_table else 0),
                                  C.c_int(no_positive_sim), _p(score, C.c_double), C.c_int(n), C.c_int(D), C.byref(p),
                                  _p(out, C.c_int32), C.byref(st), None if Ld is None else _p(Ld, C.c_double),
                                  None if L0 is None else _p(L0, C.c_double))
    if want_L:
        return out[:ny].copy(), int(st.value), Ld, L0
    return out[:ny].copy(), int(st.value)


class SsdParams(C.Structure):
    _fields_ = [("gamma", C.c_double), ("top_n", C.c_int32), ("window_size", C.c_int32), ("norm_mode", C.c_int32),
                ("normalize_emb", C.c_int32), ("use_ssd_star", C.c_int32), ("candidate_count", C.c_int32),
                ("min_score_percent", C.c_double)]


def ssd_request(emb, score, top_n, gamma=0.25, window_size=5, norm_mode=0, normalize_emb=1, use_ssd_star=0,
                candidate_count=0, min_score_percent=0.0):
    emb = np.ascontiguousarray(emb, dtype=np.float64)
    score = np.ascontiguousarray(score, dtype=np.float64)
    n, D = emb.shape
    p = SsdParams(gamma, top_n, window_size, norm_mode, normalize_emb, use_ssd_star, candidate_count, min_score_percent)
    out = np.full(max(top_n, n), -1, dtype=np.int32)
    st = C.c_int32(0)
    ny = lib().orc_ssd_request(_p(emb, C.c_double), _p(score, C.c_double), C.c_int(n), C.c_int(D), C.byref(p),
                               _p(out, C.c_int32), C.byref(st))
    return out[:ny].copy(), int(st.value)


def dpp_kernel_matrix(emb, rel, alpha=1.0, normalize=1):
    emb = np.ascontiguousarray(emb, dtype=np.float64)
    rel = np.ascontiguousarray(rel, dtype=np.float64)
    n, D = emb.shape
    L = np.zeros((n, n), dtype=np.float64)
    lib().orc_dpp_kernel_matrix(_p(emb, C.c_double), _p(rel, C.c_double), C.c_int(n), C.c_int(D), C.c_double(alpha),
                                C.c_int(normalize), _p(L, C.c_double))
    return L
