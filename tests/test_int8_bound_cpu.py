"""CPU check of the int8 filter's survivor test (pairec_b200/csrc/recall_i8.cu), restated in numpy float32 with the
kernel's own formulas: every row whose EXACT score (the oracle's fmaf chain) reaches the threshold must satisfy
I >= T_r.  The GPU tests prove the kernel; this proves the bound on inputs built to stress it, without a GPU."""
import numpy as np
import pytest

F = np.float32
SLACK = F(1.01)


def _quantise_rows(E):
    am = np.abs(E).max(axis=1).astype(F)
    s = (am / F(127.0)).astype(F)
    with np.errstate(divide="ignore", invalid="ignore"):
        x8 = np.clip(np.rint((E / s[:, None]).astype(F)), -127, 127)
    x8 = np.where(am[:, None] == 0, 0, x8).astype(np.int64)
    with np.errstate(divide="ignore"):
        a_r = (F(1.0 - 1e-6) / s).astype(F)                               # +inf for all-zero rows
    hl = (F(0.5) * np.abs(x8).sum(axis=1).astype(F) * SLACK).astype(F)
    return x8, a_r, hl


def _uniform_pass(Q, tau, d):
    """queries of a uniform pass: q' = q * fl(1 / tau_q), ONE scale for the pass, C = max_q 1.01 (L1(Q_q) / 2 + d / 4 + 1)"""
    sc = (F(1.0) / tau.astype(F)).astype(F)
    Qs = (Q * sc[:, None]).astype(F)
    t = F(np.abs(Qs).max()) / F(127.0)
    Q8 = np.clip(np.rint((Qs / t).astype(F)), -127, 127).astype(np.int64)
    C = ((F(0.5) * np.abs(Q8).sum(axis=1).astype(F) + F(d // 4 + 1)) * SLACK).astype(F).max()
    return Q8, F(t), F(C)


def _thresholds(a_r, hl, t, C):
    inv_t = F(1.0) / t
    with np.errstate(over="ignore", invalid="ignore"):
        T = (a_r * inv_t).astype(F)
        T = np.where(np.isfinite(a_r) & ~(T < F(1e30)), F(np.nan), (T - hl - C).astype(F))
    T = np.where(np.isnan(T), F(-2.0e9), np.clip(T, F(-2.0e9), F(2.0e9)))      # fmaxf drops a NaN
    return np.ceil(T.astype(np.float64)).astype(np.int64)


def _check(oracle, E, Q, k):
    d = E.shape[1]
    E = np.ascontiguousarray(E, dtype=F)
    Q = np.ascontiguousarray(Q, dtype=F)
    keys = oracle.recall_topk(E, Q, k)
    rows, scores, n = oracle.keys_split(keys)
    assert (n == k).all()
    tau = scores[:, k - 1]                     # the k-th best exact score: the top-k rows are the rows that reach it
    assert (tau > 0).all(), "uniform pass: thresholds must be positive"
    x8, a_r, hl = _quantise_rows(E)
    Q8, t, C = _uniform_pass(Q, tau, d)
    Ti = _thresholds(a_r, hl, t, C)
    acc = x8 @ Q8.T                            # exact in int64 (the tensor core: s32)
    keep = acc >= Ti[:, None]
    for q in range(Q.shape[0]):
        missed = [int(r) for r in rows[q] if not keep[r, q]]
        assert not missed, f"query {q}: rows {missed[:5]} reach the threshold exactly but fail the int8 test"
    return keep.sum(axis=0).mean() / k


@pytest.mark.parametrize("d", [64, 128])
def test_random_rows(oracle_lib, d):
    rng = np.random.default_rng(d)
    E = (rng.standard_normal((60_000, d)) / np.sqrt(d)).astype(F)
    Q = (rng.standard_normal((8, d)) / np.sqrt(d)).astype(F)
    ratio = _check(oracle_lib, E, Q, 300)
    assert ratio < 4.0                          # the bound is loose, not useless (the kernels' lists hold 4 x the expected rows)


@pytest.mark.parametrize("d", [64, 128])
def test_all_positive_rows_and_queries(oracle_lib, d):
    rng = np.random.default_rng(1000 + d)
    E = (rng.random((40_000, d)) + 0.1).astype(F)
    Q = (rng.random((6, d)) + 0.1).astype(F)
    _check(oracle_lib, E, Q, 200)


@pytest.mark.parametrize("d", [64, 128])
def test_half_way_quantisation_levels(oracle_lib, d):
    # every element half-way between two int8 levels of its row; rows and queries positive: all errors add up
    rng = np.random.default_rng(2000 + d)
    n = 40_000
    lvl = rng.integers(0, 126, size=(n, d)).astype(F) + F(0.5)
    lvl[:, 0] = 127.0
    E = (lvl * (2.0 ** rng.integers(-12, -6, size=(n, 1)))).astype(F) / F(127.0)
    ql = rng.integers(0, 126, size=(5, d)).astype(F) + F(0.5)
    ql[:, 1] = 127.0
    Q = (ql / F(127.0 * 8.0)).astype(F)
    _check(oracle_lib, E, Q, 500)


def test_rows_over_six_decades_one_hot_and_sparse_queries(oracle_lib):
    rng = np.random.default_rng(7)
    d = 64
    E = (rng.standard_normal((50_000, d)) / 8).astype(F)
    E *= (10.0 ** rng.uniform(-3, 3, size=(E.shape[0], 1))).astype(F)
    E[::11] = 0                                 # all-zero rows: a_r = +inf, never survive a positive threshold
    Q = (rng.standard_normal((6, d)) / 8).astype(F)
    Q[0] = 0
    Q[0, 5] = 1.0
    Q[1, 8:] = 0
    Q[2] *= F(1e-20)
    Q[3] *= F(1e20)
    _check(oracle_lib, E, Q, 100)


def test_one_dominant_element_per_row(oracle_lib):
    rng = np.random.default_rng(9)
    n, d = 50_000, 64
    E = (rng.standard_normal((n, d)) * 0.01).astype(F)
    E[np.arange(n), rng.integers(0, d, size=n)] = (10.0 + rng.random(n)).astype(F)
    Q = (rng.standard_normal((5, d)) / 8).astype(F)
    _check(oracle_lib, E, Q, 200)
