"""GPU parity of the recall path against the CPU oracle, through the C ABI (prg_recall_topk)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _data(n, d, b, seed):
    rng = np.random.default_rng(seed)
    E = (rng.standard_normal((n, d)) / np.sqrt(d)).astype(np.float32)
    Q = (rng.standard_normal((b, d)) / np.sqrt(d)).astype(np.float32)
    return E, Q


def _check(engine, oracle, E, Q, k, row_base=0):
    engine.set_item_matrix(E, row_base=row_base)
    rows, scores, n = engine.recall_topk(Q, k)
    keys = oracle.recall_topk(E, Q, k, row_base=row_base)
    orows, oscores, on = oracle.keys_split(keys)
    assert (n == on).all()
    assert (rows == orows).all(), "recalled row ids differ from the oracle"
    assert (scores.view(np.uint32) == oscores.view(np.uint32)).all(), "scores are not bit-identical"


@pytest.mark.parametrize("n,d,b,k", [(1000, 64, 1, 10), (5000, 64, 7, 100), (4097, 128, 64, 50), (300, 64, 3, 1000),
                                      (70000, 64, 65, 100)])
def test_dense_path_matches_oracle(engine, oracle_lib, n, d, b, k):
    E, Q = _data(n, d, b, seed=n + b)
    _check(engine, oracle_lib, E, Q, k)


@pytest.mark.parametrize("n,d,b,k", [(400_000, 64, 64, 1000), (300_001, 128, 5, 200), (1_000_000, 64, 17, 1000)])
def test_sampled_path_matches_oracle(engine, oracle_lib, n, d, b, k):
    E, Q = _data(n, d, b, seed=7)
    _check(engine, oracle_lib, E, Q, k, row_base=12345)
    st = engine.recall_stats()
    assert st["fallback_queries"] == 0
    assert st["max_candidates"] >= k


def test_ties_zero_query_and_duplicates(engine, oracle_lib):
    # zero query: every score ties at +0 -> rows 0..k-1; duplicated rows tie pairwise
    E, Q = _data(300_000, 64, 4, seed=3)
    E[150_000:] = E[:150_000]
    Q[0] = 0
    _check(engine, oracle_lib, E, Q, 100)


def test_adversarial_order_falls_back_exactly(engine, oracle_lib):
    # every large score sits in a tile the strided sample never visits -> the candidate list overflows and the
    # query is redone through the dense path; the result must still be exact
    n, d = 400_000, 64
    rng = np.random.default_rng(11)
    E = (rng.standard_normal((n, d)) * 0.01).astype(np.float32)
    n_tiles = (n + 255) // 256
    stride = n_tiles // max(64, n_tiles // 128)        # sampling plan of recall.cu
    tile = np.arange(n) // 256
    E[:, 0] = np.where(tile % stride == 0, 0.0, 1.0 + rng.random(n) * 0.5).astype(np.float32)
    Q = np.zeros((2, d), dtype=np.float32)
    Q[0, 0] = 1.0
    Q[1, 1] = -1.0
    _check(engine, oracle_lib, E, Q, 1000)
    assert engine.recall_stats()["fallback_queries"] >= 1


def test_nan_and_inf_scores(engine, oracle_lib):
    E, Q = _data(3000, 64, 2, seed=5)
    E[10, 0] = np.nan
    E[20, 1] = np.inf
    E[30, 2] = -np.inf
    _check(engine, oracle_lib, E, Q, 3000)


def test_k_larger_than_rows(engine, oracle_lib):
    E, Q = _data(100, 64, 2, seed=9)
    engine.set_item_matrix(E)
    rows, scores, n = engine.recall_topk(Q, 256)
    assert (n == 100).all()
    assert (rows[:, 100:] == 0xFFFFFFFF).all() and np.isneginf(scores[:, 100:]).all()
    keys = oracle_lib.recall_topk(E, Q, 256)
    orows, _, _ = oracle_lib.keys_split(keys)
    assert (rows == orows).all()


def test_errors_do_not_abort(engine):
    from pairec_b200 import PrgError
    with pytest.raises(PrgError):
        engine.set_item_matrix(np.zeros((10, 48), dtype=np.float32))  # unsupported dim


def test_exact_ffma2_scan_path_matches_oracle(oracle_lib):
    # the full pass through the exact FFMA2 scan (config scan_ffma2) instead of the tensor-core filter + re-score
    from pairec_b200 import Engine
    eng = Engine(0, scan_ffma2=1)
    try:
        E, Q = _data(500_000, 64, 33, seed=21)
        _check(eng, oracle_lib, E, Q, 500)
        assert eng.recall_stats()["fallback_queries"] == 0
    finally:
        eng.close()


def test_filter_margin_holds_for_scaled_rows(engine, oracle_lib):
    # rows with very different norms (1e-3 .. 1e3): the TF32 filter margin is per row, the result stays exact
    E, Q = _data(400_000, 64, 8, seed=23)
    rng = np.random.default_rng(5)
    E *= (10.0 ** rng.uniform(-3, 3, size=(E.shape[0], 1))).astype(np.float32)
    _check(engine, oracle_lib, E, Q, 300)


def test_dim128_sampled_path(engine, oracle_lib):
    E, Q = _data(300_000, 128, 64, seed=29)
    _check(engine, oracle_lib, E, Q, 1000)


@pytest.mark.parametrize("b", [100, 200, 300])
def test_many_queries_per_pass(engine, oracle_lib, b):
    # 2 / 4 query blocks per pass of the tensor-core filter (one read of the matrix), and more than one pass
    E, Q = _data(350_000, 64, b, seed=41 + b)
    _check(engine, oracle_lib, E, Q, 200)
    assert engine.recall_stats()["fallback_queries"] == 0


@pytest.mark.parametrize("b", [100, 200, 300])
def test_many_queries_per_pass_dim128(oracle_lib, b):
    # config scan128_nqb: 2 / 4 query blocks per pass at dim 128 (2-stage ring), and more than one pass
    from pairec_b200 import Engine
    eng = Engine(0, scan128_nqb=1 if b != 100 else 0)   # b == 100 also covers the 64-queries-per-pass form
    try:
        E, Q = _data(350_000, 128, b, seed=43 + b)
        _check(eng, oracle_lib, E, Q, 200)
        assert eng.recall_stats()["fallback_queries"] == 0
    finally:
        eng.close()


@pytest.mark.parametrize("d,b", [(64, 70), (128, 9)])
def test_tf32_filter_variant_matches_oracle(oracle_lib, d, b):
    # config scan_tf32: the tensor-core filter over the fp32 rows themselves (no bf16 shadow index)
    from pairec_b200 import Engine
    eng = Engine(0, scan_tf32=1)
    try:
        E, Q = _data(400_000, d, b, seed=51)
        _check(eng, oracle_lib, E, Q, 400)
        assert eng.recall_stats()["fallback_queries"] == 0
    finally:
        eng.close()


def test_bf16_filter_worst_case_rounding(engine, oracle_lib):
    # elements that sit exactly between two bf16 values (largest rounding error of the shadow index), positive rows
    # and a positive query so that no cancellation hides the error: the margin must still keep every true winner
    n, d = 400_000, 64
    rng = np.random.default_rng(61)
    base = (rng.integers(128, 256, size=(n, d)).astype(np.float32)) / 256.0       # 8 significant bits
    E = (base * (1.0 + 2.0 ** -8) * (2.0 ** rng.integers(-3, 3, size=(n, 1)))).astype(np.float32)  # half-way cases
    Q = (rng.integers(128, 256, size=(5, d)).astype(np.float32) / 256.0 * (1.0 + 2.0 ** -8)).astype(np.float32)
    _check(engine, oracle_lib, E, Q, 1000)


def test_subnormal_rows_and_queries(engine, oracle_lib):
    E, Q = _data(300_000, 64, 3, seed=67)
    E[::3] *= np.float32(1e-38)          # subnormal / near-subnormal rows
    Q[1] *= np.float32(1e-30)
    Q[2] *= np.float32(1e-40)
    _check(engine, oracle_lib, E, Q, 500)


def test_matrix_replaced_rebuilds_filter_index(engine, oracle_lib):
    E1, Q = _data(300_000, 64, 4, seed=71)
    _check(engine, oracle_lib, E1, Q, 100)
    E2, _ = _data(350_000, 64, 4, seed=73)
    _check(engine, oracle_lib, E2, Q, 100)


def test_negative_thresholds_use_the_per_query_filter(engine, oracle_lib):
    # every score is negative (positive rows, negative queries): the sampled threshold is < 0, so the uniform
    # (query / tau) form of the filter does not apply and the per-query form must give the same exact result
    n, d = 400_000, 64
    rng = np.random.default_rng(83)
    E = rng.random((n, d), dtype=np.float32) + np.float32(0.1)
    Q = -(rng.random((6, d), dtype=np.float32) + np.float32(0.1))
    Q[3] = -Q[3]    # one query with positive scores in the same pass
    _check(engine, oracle_lib, E, Q, 700)
    assert engine.recall_stats()["fallback_queries"] == 0


def test_large_k_uses_the_split_refine_path(engine, oracle_lib):
    # k = 3000: the survivors' exact keys do not fit in shared memory, so re-score and select run as separate kernels
    E, Q = _data(400_000, 64, 5, seed=91)
    _check(engine, oracle_lib, E, Q, 3000)
    assert engine.recall_stats()["fallback_queries"] == 0


def test_clustered_scores_refine_select(engine, oracle_lib):
    # many near-identical scores: the radix select of the refine kernel has to walk several digit passes
    n, d = 400_000, 64
    rng = np.random.default_rng(97)
    base = (rng.standard_normal((1, d)) / np.sqrt(d)).astype(np.float32)
    E = (base + rng.standard_normal((n, d)).astype(np.float32) * np.float32(1e-4)).astype(np.float32)
    Q = (base * np.float32(1.0) + rng.standard_normal((3, d)).astype(np.float32) * np.float32(1e-3)).astype(np.float32)
    _check(engine, oracle_lib, E, Q, 1000)


# ---- GROUP mode of the tensor-core filter (passes of more than 64 queries: survivors recorded per (row, group of 16
# queries), exact re-score against the group; recall_tc.cu / recall.cu rescore_group_kernel)

def test_group_mode_equals_per_query_mode(oracle_lib):
    from pairec_b200 import Engine
    E, Q = _data(500_000, 64, 150, seed=101)
    out = []
    for groups in (1, 0):
        eng = Engine(0, scan_groups=groups)
        try:
            eng.set_item_matrix(E, row_base=777)
            out.append(eng.recall_topk(Q, 1000))
            assert eng.recall_stats()["fallback_queries"] == 0
        finally:
            eng.close()
    for a, b in zip(out[0], out[1]):
        assert (np.asarray(a).view(np.uint32) == np.asarray(b).view(np.uint32)).all()
    keys = oracle_lib.recall_topk(E, Q, 1000, row_base=777)
    orows, oscores, on = oracle_lib.keys_split(keys)
    assert (out[0][0] == orows).all() and (out[0][1].view(np.uint32) == oscores.view(np.uint32)).all()


def test_group_mode_negative_thresholds_nan_and_padding(engine, oracle_lib):
    # 70 queries = one full block + 6 queries in a block whose other groups are padding; negative thresholds force the
    # per-query form of the filter test; one query and some rows carry NaN / inf
    n, d = 400_000, 64
    rng = np.random.default_rng(103)
    E = rng.random((n, d), dtype=np.float32) + np.float32(0.1)
    Q = -(rng.random((70, d), dtype=np.float32) + np.float32(0.1))
    Q[5] = -Q[5]
    Q[66] = -Q[66]
    E[1234, 3] = np.inf
    E[99_999, 7] = np.nan
    E[200_000] = 0
    _check(engine, oracle_lib, E, Q, 300)


def test_group_mode_adversarial_order_falls_back_exactly(engine, oracle_lib):
    # as test_adversarial_order_falls_back_exactly, with 96 queries: the group lists overflow, the affected queries are
    # flagged through their segment counts and redone densely
    n, d = 400_000, 64
    rng = np.random.default_rng(107)
    E = (rng.standard_normal((n, d)) * 0.01).astype(np.float32)
    n_tiles = (n + 255) // 256
    stride = n_tiles // max(64, n_tiles // 128)
    tile = np.arange(n) // 256
    E[:, 0] = np.where(tile % stride == 0, 0.0, 1.0 + rng.random(n) * 0.5).astype(np.float32)
    Q = (rng.standard_normal((96, d)) * 0.01).astype(np.float32)
    Q[:, 0] = 0          # the other queries do not see the adversarial column
    Q[0] = 0
    Q[0, 0] = 1.0
    Q[70] = 0
    Q[70, 0] = 2.0
    _check(engine, oracle_lib, E, Q, 1000)
    assert engine.recall_stats()["fallback_queries"] >= 2


def test_group_mode_large_k_keeps_the_split_path(engine, oracle_lib):
    E, Q = _data(400_000, 64, 80, seed=109)
    _check(engine, oracle_lib, E, Q, 3000)
    assert engine.recall_stats()["fallback_queries"] == 0
