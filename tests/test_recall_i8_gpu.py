"""GPU parity of the int8 filter index (pairec_b200/csrc/recall_i8.cu: dim 64, at most 64 queries per pass) against the CPU
oracle and against the bf16 filter, through the C ABI.  The filter only prunes: rows and score bits must not depend on it."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _data(n, d, b, seed):
    rng = np.random.default_rng(seed)
    E = (rng.standard_normal((n, d)) / np.sqrt(d)).astype(np.float32)
    Q = (rng.standard_normal((b, d)) / np.sqrt(d)).astype(np.float32)
    return E, Q


def _check(engine, oracle, E, Q, k, row_base=0, want_filter="int8"):
    engine.set_item_matrix(E, row_base=row_base)
    rows, scores, n = engine.recall_topk(Q, k)
    st = engine.recall_stats()
    assert st["filter"] == want_filter, st
    keys = oracle.recall_topk(E, Q, k, row_base=row_base)
    orows, oscores, on = oracle.keys_split(keys)
    assert (n == on).all()
    assert (rows == orows).all(), "recalled row ids differ from the oracle"
    assert (scores.view(np.uint32) == oscores.view(np.uint32)).all(), "scores are not bit-identical"
    return st


@pytest.mark.parametrize("n,b,k", [(1_000_000, 64, 1000), (300_001, 5, 200), (262_144 + 511, 33, 1000), (400_000, 1, 50)])
def test_int8_filter_matches_oracle(engine, oracle_lib, n, b, k):
    # (row counts that are not multiples of the 512-row tile, a row base, full and partial query blocks)
    E, Q = _data(n, 64, b, seed=n % 1000 + b)
    st = _check(engine, oracle_lib, E, Q, k, row_base=4321)
    assert st["fallback_queries"] == 0
    assert st["max_candidates"] >= k


def test_int8_and_bf16_filters_give_the_same_bits(oracle_lib):
    from pairec_b200 import Engine
    E, Q = _data(700_000, 64, 64, seed=5)
    out = {}
    for name, cfg in (("int8", dict()), ("bf16", dict(scan_int8=0))):
        eng = Engine(0, **cfg)
        try:
            eng.set_item_matrix(E)
            out[name] = eng.recall_topk(Q, 1000)
            st = eng.recall_stats()
            assert st["filter"] == name and st["fallback_queries"] == 0, st
            out[name + "_cand"] = st["max_candidates"]
        finally:
            eng.close()
    for a, b in zip(out["int8"], out["bf16"]):
        assert (np.asarray(a).view(np.uint32) == np.asarray(b).view(np.uint32)).all()
    # the int8 bound is wider: more rows survive the filter, but within the candidate capacity (4 x the expected count)
    assert out["bf16_cand"] < out["int8_cand"] < 3 * out["bf16_cand"]


def test_more_than_64_queries_keep_the_bf16_filter(engine, oracle_lib):
    E, Q = _data(300_000, 64, 65, seed=9)
    _check(engine, oracle_lib, E, Q, 100, want_filter="bf16")


def test_dim128_keeps_the_bf16_filter(engine, oracle_lib):
    E, Q = _data(300_000, 128, 8, seed=13)
    _check(engine, oracle_lib, E, Q, 100, want_filter="bf16")


def test_worst_case_quantisation_error(engine, oracle_lib):
    # every element sits half-way between two int8 levels of its row (largest rounding error), rows and queries positive so
    # that nothing cancels: the bound must still keep every true winner
    n, d = 400_000, 64
    rng = np.random.default_rng(17)
    lvl = rng.integers(0, 126, size=(n, d)).astype(np.float32) + np.float32(0.5)
    lvl[:, 0] = 127.0                                           # the row maximum fixes the scale: s_r = 2^e
    E = (lvl * (2.0 ** rng.integers(-12, -6, size=(n, 1)))).astype(np.float32) / np.float32(127.0)
    ql = rng.integers(0, 126, size=(7, d)).astype(np.float32) + np.float32(0.5)
    ql[:, 1] = 127.0
    Q = (ql / np.float32(127.0 * 8.0)).astype(np.float32)
    _check(engine, oracle_lib, E, Q, 1000)


def test_one_large_element_per_row(engine, oracle_lib):
    # one element 1000 x the others: all other elements quantise to 0; the queries look only at the small ones
    n, d = 400_000, 64
    rng = np.random.default_rng(19)
    E = (rng.standard_normal((n, d)) * 0.01).astype(np.float32)
    big = rng.integers(0, d, size=n)
    E[np.arange(n), big] = (10.0 + rng.random(n)).astype(np.float32) * np.where(rng.random(n) < 0.5, -1, 1).astype(np.float32)
    Q = (rng.standard_normal((9, d)) / 8).astype(np.float32)
    Q[0] = 0
    Q[0, 5] = 1.0          # one-hot query
    Q[1, :] = 0
    Q[1, 7] = -3.0
    _check(engine, oracle_lib, E, Q, 500)


def test_rows_over_six_decades_and_sparse_queries(engine, oracle_lib):
    E, Q = _data(400_000, 64, 12, seed=23)
    rng = np.random.default_rng(29)
    E *= (10.0 ** rng.uniform(-3, 3, size=(E.shape[0], 1))).astype(np.float32)
    Q[3, 8:] = 0
    Q[4] *= np.float32(1e-20)
    Q[5] *= np.float32(1e20)
    _check(engine, oracle_lib, E, Q, 300)


def test_nan_inf_zero_rows_and_negative_thresholds(engine, oracle_lib):
    # negative thresholds force the per-query form of the test; zero rows score exactly 0 and must be kept when tau <= 0;
    # NaN / inf rows always survive the filter and are settled by the exact re-score
    n, d = 400_000, 64
    rng = np.random.default_rng(31)
    E = rng.random((n, d), dtype=np.float32) + np.float32(0.1)
    Q = -(rng.random((6, d), dtype=np.float32) + np.float32(0.1))
    Q[3] = -Q[3]
    E[1234, 3] = np.inf
    E[99_999, 7] = np.nan
    E[200_000:200_600] = 0       # 600 zero rows: the best rows of every negative-score query
    E[300_000, 1] = -np.inf
    st = _check(engine, oracle_lib, E, Q, 700)
    # (the zero rows sit in tiles the strided sample does not visit: the thresholds of the negative-score queries come out
    # wrong for either filter and those queries are redone densely.)  After a batch with redone queries the library
    # takes the bf16 index for the next batches of the matrix: the int8 bound is the wider one
    assert st["fallback_queries"] > 0
    rows, scores, n = engine.recall_topk(Q, 700)
    assert engine.recall_stats()["filter"] == "bf16"
    orows, oscores, _ = oracle_lib.keys_split(oracle_lib.recall_topk(E, Q, 700))
    assert (rows == orows).all() and (scores.view(np.uint32) == oscores.view(np.uint32)).all()


def test_all_positive_rows_stay_within_the_candidate_capacity(engine, oracle_lib):
    # no cancellation: scores of a query differ by a few percent of their size while the int8 bound is ~ 1/127 of
    # |x|_1 |q|_inf — the widest it gets relative to the spread of the scores; still no list overflows
    n, d = 400_000, 64
    rng = np.random.default_rng(59)
    E = rng.random((n, d), dtype=np.float32) + np.float32(0.1)
    Q = rng.random((16, d), dtype=np.float32) + np.float32(0.1)
    Q[8:] = -Q[8:]
    st = _check(engine, oracle_lib, E, Q, 700)
    assert st["fallback_queries"] == 0


def test_zero_rows_with_positive_thresholds_and_a_nan_query(engine, oracle_lib):
    E, Q = _data(400_000, 64, 5, seed=37)
    E[::7] = 0
    Q[2, 11] = np.nan            # every score of this query is NaN: the query is redone densely, the others are not
    Q[4] = 0
    _check(engine, oracle_lib, E, Q, 400)


def test_subnormal_rows_and_tiny_queries(engine, oracle_lib):
    E, Q = _data(300_000, 64, 4, seed=41)
    E[::3] *= np.float32(1e-38)
    E[1::3] *= np.float32(1e-44)
    Q[1] *= np.float32(1e-30)
    Q[2] *= np.float32(1e-40)
    _check(engine, oracle_lib, E, Q, 500)


def test_adversarial_order_falls_back_exactly(engine, oracle_lib):
    n, d = 400_000, 64
    rng = np.random.default_rng(43)
    E = (rng.standard_normal((n, d)) * 0.01).astype(np.float32)
    n_tiles = (n + 255) // 256
    stride = n_tiles // max(64, n_tiles // 128)        # sampling plan of recall.cu
    tile = np.arange(n) // 256
    E[:, 0] = np.where(tile % stride == 0, 0.0, 1.0 + rng.random(n) * 0.5).astype(np.float32)
    Q = np.zeros((2, d), dtype=np.float32)
    Q[0, 0] = 1.0
    Q[1, 1] = -1.0
    st = _check(engine, oracle_lib, E, Q, 1000)
    assert st["fallback_queries"] >= 1


def test_snapshot_swap_rebuilds_the_int8_index(engine, oracle_lib):
    E1, Q = _data(300_000, 64, 4, seed=47)
    _check(engine, oracle_lib, E1, Q, 100)
    E2, _ = _data(350_000, 64, 4, seed=53)
    engine.stage_item_matrix(E2)
    engine.commit_item_matrix()
    rows, scores, n = engine.recall_topk(Q, 100)
    assert engine.recall_stats()["filter"] == "int8"
    orows, oscores, _ = oracle_lib.keys_split(oracle_lib.recall_topk(E2, Q, 100))
    assert (rows == orows).all() and (scores.view(np.uint32) == oscores.view(np.uint32)).all()


# ---- dim 128: GROUP-mode passes of up to 256 queries over the int8 index (recall_scan_i8g_kernel; the C5 shard pass)

def _data128(n, b, seed):
    return _data(n, 128, b, seed)


@pytest.mark.parametrize("n,b,k", [(350_000, 100, 200), (300_001, 256, 1000), (400_000, 300, 100), (262_144 + 255, 65, 50)])
def test_dim128_group_passes_use_the_int8_index(engine, oracle_lib, n, b, k):
    E, Q = _data128(n, b, seed=b)
    st = _check(engine, oracle_lib, E, Q, k, row_base=99)
    assert st["fallback_queries"] == 0


def test_dim128_int8_and_bf16_give_the_same_bits(oracle_lib):
    from pairec_b200 import Engine
    E, Q = _data128(500_000, 150, seed=3)
    out = {}
    for name, cfg in (("int8", dict()), ("bf16", dict(scan_int8=0)), ("bf16g", dict(scan_int8=0, scan_groups=1))):
        eng = Engine(0, **cfg)
        try:
            eng.set_item_matrix(E)
            out[name] = eng.recall_topk(Q, 1000)
            st = eng.recall_stats()
            assert st["filter"] == name[:4] and st["fallback_queries"] == 0, st
        finally:
            eng.close()
    for other in ("bf16", "bf16g"):
        for a, b in zip(out["int8"], out[other]):
            assert (np.asarray(a).view(np.uint32) == np.asarray(b).view(np.uint32)).all()


def test_dim128_negative_thresholds_nan_rows_and_padding_groups(engine, oracle_lib):
    # 70 queries: one full block of 64 + 6 in a block whose other groups are padding; negative thresholds force the
    # per-query form; NaN / inf rows survive for every group that has a list and for none that has not
    n, d = 400_000, 128
    rng = np.random.default_rng(7)
    E = rng.random((n, d), dtype=np.float32) + np.float32(0.1)
    Q = -(rng.random((70, d), dtype=np.float32) + np.float32(0.1))
    Q[5] = -Q[5]
    Q[66] = -Q[66]
    E[1234, 3] = np.inf
    E[99_999, 7] = np.nan
    E[200_000] = 0
    _check(engine, oracle_lib, E, Q, 300)


def test_dim128_positive_batch_with_nan_rows_and_zero_rows(engine, oracle_lib):
    E, Q = _data128(400_000, 130, seed=11)
    E[10, 0] = np.nan
    E[20, 1] = np.inf
    E[::9] = 0
    Q[17] = 0
    _check(engine, oracle_lib, E, Q, 500)


def test_dim128_worst_case_quantisation_error(engine, oracle_lib):
    n, d = 300_000, 128
    rng = np.random.default_rng(13)
    lvl = rng.integers(0, 126, size=(n, d)).astype(np.float32) + np.float32(0.5)
    lvl[:, 0] = 127.0
    E = (lvl * (2.0 ** rng.integers(-12, -6, size=(n, 1)))).astype(np.float32) / np.float32(127.0)
    ql = rng.integers(0, 126, size=(80, d)).astype(np.float32) + np.float32(0.5)
    ql[:, 1] = 127.0
    Q = (ql / np.float32(127.0 * 8.0)).astype(np.float32)
    _check(engine, oracle_lib, E, Q, 1000)


def test_dim128_adversarial_order_falls_back_exactly(engine, oracle_lib):
    n, d = 400_000, 128
    rng = np.random.default_rng(17)
    E = (rng.standard_normal((n, d)) * 0.01).astype(np.float32)
    n_tiles = (n + 255) // 256
    stride = n_tiles // max(64, n_tiles // 128)
    tile = np.arange(n) // 256
    E[:, 0] = np.where(tile % stride == 0, 0.0, 1.0 + rng.random(n) * 0.5).astype(np.float32)
    Q = (rng.standard_normal((96, d)) * 0.01).astype(np.float32)
    Q[:, 0] = 0
    Q[0] = 0
    Q[0, 0] = 1.0
    Q[70] = 0
    Q[70, 0] = 2.0
    st = _check(engine, oracle_lib, E, Q, 1000)
    assert st["fallback_queries"] >= 2


# ---- dim 64, more than 64 queries: the GROUP-mode pass over the bf16 index through the 16-epilogue-warp kernel
# (recall_scan_grp_kernel<false>, config scan_grp16) against recall_tc.cu's form of the same pass

@pytest.mark.parametrize("b", [70, 128, 129, 256, 300])
def test_dim64_group_pass_kernels_agree(oracle_lib, b):
    from pairec_b200 import Engine
    E, Q = _data(500_000, 64, b, seed=200 + b)
    E[77, 5] = np.nan
    E[78] = 0
    out = []
    for grp16 in (1, 0):
        eng = Engine(0, scan_grp16=grp16)
        try:
            eng.set_item_matrix(E, row_base=31)
            out.append(eng.recall_topk(Q, 500))
            st = eng.recall_stats()
            assert st["filter"] == "bf16" and st["fallback_queries"] == 0, st
        finally:
            eng.close()
    for a, c in zip(out[0], out[1]):
        assert (np.asarray(a).view(np.uint32) == np.asarray(c).view(np.uint32)).all()
    orows, oscores, _ = oracle_lib.keys_split(oracle_lib.recall_topk(E, Q, 500, row_base=31))
    assert (out[0][0] == orows).all() and (out[0][1].view(np.uint32) == oscores.view(np.uint32)).all()


def test_dim64_group_pass_negative_thresholds(engine, oracle_lib):
    n, d = 400_000, 64
    rng = np.random.default_rng(301)
    E = rng.random((n, d), dtype=np.float32) + np.float32(0.1)
    Q = -(rng.random((140, d), dtype=np.float32) + np.float32(0.1))
    Q[5] = -Q[5]
    Q[130] = -Q[130]
    E[1234, 3] = np.inf
    _check(engine, oracle_lib, E, Q, 300, want_filter="bf16")
