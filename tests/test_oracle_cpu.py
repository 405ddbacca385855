"""CPU tests of the oracle (no GPU): the reference's one known-answer test for this path, the committed golden
fixtures, an independent pure-Python restatement of DPP, and structural properties."""
import os

import numpy as np
import pytest

from tests import ref_py, synth

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_rank_score_expression_kat(oracle_lib):
    # utils/ast/ast_test.go:12-27 — "${ctr} + ${click} + ${price}" with ctr=0.1, click=0.3, price=0.1 evaluates to 0.5
    assert oracle_lib.rank_score_expr([0.1, 0.3, 0.1], [1.0, 1.0, 1.0]) == 0.5


def test_lookup_default(oracle_lib):
    # algorithm/lookup.go:44-49
    out = oracle_lib.lookup(np.array([0.2, 9.0, 0.7]), np.array([1, 0, 1]))
    assert out.tolist() == [0.2, 0.5, 0.7]


def test_golden_recall(oracle_lib):
    g = np.load(os.path.join(G, "recall_small.npz"))
    keys = oracle_lib.recall_topk(g["E"], g["Q"], int(g["k"]), row_base=int(g["row_base"]))
    assert (keys == g["keys"]).all()
    rows, scores, n = oracle_lib.keys_split(keys)
    # independent check: numpy stable argsort of fp32 fmaf-chain scores
    s = oracle_lib.recall_scores(g["E"], g["Q"])
    ref = np.argsort(-s, axis=1, kind="stable")[:, :int(g["k"])]
    assert (rows == ref + int(g["row_base"])).all()
    assert (np.diff(keys.astype(np.uint64).view(np.int64), axis=1) < 0).all()  # strictly descending keys


def test_recall_thread_count_does_not_change_result(oracle_lib):
    rng = np.random.default_rng(1)
    E = rng.standard_normal((20000, 64)).astype(np.float32)
    Q = rng.standard_normal((3, 64)).astype(np.float32)
    a = oracle_lib.recall_topk(E, Q, 50, n_threads=1)
    b = oracle_lib.recall_topk(E, Q, 50, n_threads=5)
    assert (a == b).all()


def test_shard_merge_equals_unsharded(oracle_lib):
    rng = np.random.default_rng(2)
    E = rng.standard_normal((9000, 64)).astype(np.float32)
    E[4000:4100] = E[:100]          # ties across shards
    Q = rng.standard_normal((4, 64)).astype(np.float32)
    full = oracle_lib.recall_topk(E, Q, 64)
    parts = [oracle_lib.recall_topk(E[a:b], Q, 64, row_base=a) for a, b in ((0, 3000), (3000, 6000), (6000, 9000))]
    merged = oracle_lib.merge_keys(np.stack(parts))
    assert (merged == full).all()


def test_golden_rank(oracle_lib):
    g = np.load(os.path.join(G, "rank_small.npz"))
    factors = [g[f"factors{t}"] for t in range(8)]
    linear = [g[f"linear{t}"] for t in range(8)]
    logit, x = oracle_lib.gather_fm(g["fields"], factors, linear, float(g["w0"]), g["rows"])
    assert (logit.view(np.uint32) == g["fm_logit"].view(np.uint32)).all()
    assert (x == g["x"]).all()
    W = [g[f"W{l}"] for l in range(3)]
    b = [g[f"b{l}"] for l in range(3)]
    mlp = oracle_lib.mlp_forward(x, g["dims"].tolist(), W, b)
    assert (mlp.view(np.uint32) == g["mlp_logit"].view(np.uint32)).all()
    assert (oracle_lib.sigmoid(logit).view(np.uint32) == g["fm_score"].view(np.uint32)).all()


def test_fm_against_numpy_float64(oracle_lib):
    fields, factors, linear = synth.rank_tables(n_items=200, n_fields=6, max_rows=100)
    rows = np.arange(200, dtype=np.uint32)
    logit, x = oracle_lib.gather_fm(fields, factors, linear, 0.1, rows)
    v = np.stack([factors[f][fields[:, f]] for f in range(6)], axis=1).astype(np.float64)   # [n, F, 16]
    lin = 0.1 + sum(linear[f][fields[:, f]].astype(np.float64) for f in range(6))
    inter = ((v.sum(1) ** 2) - (v ** 2).sum(1)).sum(1)
    want = lin + 0.5 * inter
    assert np.allclose(logit, want, rtol=0, atol=1e-6)
    assert (x.reshape(200, 6, 16) == v.astype(np.float32)).all()


def test_mlp_against_numpy_float64(oracle_lib):
    dims = [64, 64, 1]
    W, b = synth.mlp_weights(dims)
    rng = np.random.default_rng(3)
    x = (rng.standard_normal((40, 64)) * 0.5).astype(np.float32)
    got = oracle_lib.mlp_forward(x, dims, W, b)
    a = oracle_lib.bf16_to_f32(oracle_lib.f32_to_bf16(x)).astype(np.float64)   # the tower input is bf16(x)
    a = np.maximum(a @ oracle_lib.bf16_to_f32(W[0]).astype(np.float64).T + b[0], 0)
    want = a @ oracle_lib.bf16_to_f32(W[1]).astype(np.float64).T + b[1]
    # bf16x2 activations carry >= 16 significant bits
    assert np.allclose(got, want[:, 0], rtol=0, atol=2e-5)


def test_user_features_against_numpy_float64(oracle_lib):
    """features = userFeatures, then itemFeatures (service/rank/algo_data.go:104-118): FM over user + item fields,
    tower input [item | user | dense]; multi-head output."""
    F, U, nd = 5, 2, 3
    fields, factors, linear = synth.rank_tables(n_items=200, n_fields=F + U, max_rows=100)
    fields = np.ascontiguousarray(fields[:, :F])
    rows = np.arange(200, dtype=np.uint32)
    uid = np.array([7, 0xFFFFFFFF], dtype=np.uint32)
    dense = np.array([0.25, -1.5, 3.0], dtype=np.float32)
    logit, x = oracle_lib.gather_fm(fields, factors, linear, 0.1, rows, user_ids=uid, user_dense=dense)
    v_item = np.stack([factors[f][fields[:, f]] for f in range(F)], axis=1).astype(np.float64)
    v_user = np.zeros((200, U, 16))
    v_user[:, 0] = factors[F][7]
    v = np.concatenate([v_user, v_item], axis=1)
    lin = 0.1 + float(linear[F][7]) + sum(linear[f][fields[:, f]].astype(np.float64) for f in range(F))
    want = lin + 0.5 * ((v.sum(1) ** 2) - (v ** 2).sum(1)).sum(1)
    assert np.allclose(logit, want, rtol=0, atol=1e-6)
    assert x.shape == (200, (F + U) * 16 + nd)
    assert (x[:, :F * 16].reshape(200, F, 16) == v_item.astype(np.float32)).all()
    assert (x[:, F * 16:(F + 1) * 16] == factors[F][7]).all() and (x[:, (F + 1) * 16:(F + 2) * 16] == 0).all()
    assert (x[:, (F + U) * 16:] == dense).all()
    # without user features the two entry points agree bit for bit
    a, _ = oracle_lib.gather_fm(fields, factors[:F], linear[:F], 0.1, rows)
    b, _ = oracle_lib.gather_fm(fields, factors, linear, 0.1, rows, user_ids=np.full(U, 0xFFFFFFFF, dtype=np.uint32))
    assert (a.view(np.uint32) == b.view(np.uint32)).all()
    # three heads
    dims = [x.shape[1] + 13, 64, 3]      # any input width
    dims[0] = x.shape[1]
    W, bs = synth.mlp_weights(dims)
    got = oracle_lib.mlp_forward(x, dims, W, bs)
    a0 = oracle_lib.bf16_to_f32(oracle_lib.f32_to_bf16(x)).astype(np.float64)
    a1 = np.maximum(a0 @ oracle_lib.bf16_to_f32(W[0]).astype(np.float64).T + bs[0], 0)
    want = a1 @ oracle_lib.bf16_to_f32(W[1]).astype(np.float64).T + bs[1]
    assert got.shape == (200, 3) and np.allclose(got, want, rtol=0, atol=2e-5)


def test_bf16_round_to_nearest_even(oracle_lib):
    lib = oracle_lib.lib()
    assert lib.orc_f32_to_bf16(1.0) == 0x3F80
    assert lib.orc_f32_to_bf16(np.float32(1.0 + 2 ** -8)) == 0x3F80      # tie -> even
    assert lib.orc_f32_to_bf16(np.float32(1.0 + 3 * 2 ** -8)) == 0x3F82  # tie -> even (up)
    a = np.random.default_rng(0).standard_normal(1000).astype(np.float32)
    assert (oracle_lib.f32_to_bf16(a) == np.array([lib.orc_f32_to_bf16(float(v)) for v in a], dtype=np.uint16)).all()


def test_golden_dpp_and_python_restatement(oracle_lib):
    g = np.load(os.path.join(G, "dpp_small.npz"))
    emb = g["emb"].astype(np.float64)
    idx, st = oracle_lib.dpp_request(emb, g["score"], 25, alpha=1.0, window_size=10)
    assert st == int(g["status"]) and (idx == g["idx"]).all()
    idx2, st2 = oracle_lib.dpp_request(emb, g["score2"], 12, alpha=2.0, window_size=5, norm_mode=2)
    assert st2 == int(g["status2"]) and (idx2 == g["idx2"]).all()
    # independent pure-Python float64 restatement (tests/ref_py.py)
    L = ref_py.kernel_matrix(emb.tolist(), g["score"].tolist(), 1.0)
    assert ref_py.dpp_with_window(L, 25, 10) == idx.tolist()
    Lc = oracle_lib.dpp_kernel_matrix(emb, g["score"], alpha=1.0)
    assert (np.array(L) == Lc).all(), "kernel matrix differs from the Python restatement"


def test_dpp_kernel_matrix_small_products_take_the_serial_gemm(oracle_lib):
    """gonum dgemmParallel hands a product with fewer than four 64 x 64 blocks of C (at most 64 items) to dgemmSerial: one
    DotUnitary over all of k instead of k blocks of 64 — a different rounding once D + 1 > 64.  The C oracle and the
    independent Python restatement agree bit for bit on both sides of the threshold, and the two forms really differ."""
    rng = np.random.default_rng(64)
    D = 96
    for n in (40, 64, 65):
        emb = rng.standard_normal((n, D))
        rel = rng.random(n)
        Lc = oracle_lib.dpp_kernel_matrix(emb, rel, alpha=1.0)
        Lp = np.array(ref_py.kernel_matrix(emb.tolist(), rel.tolist(), 1.0))
        assert (Lc.view(np.uint64) == Lp.view(np.uint64)).all(), n
    b = rng.standard_normal(97).tolist()
    n_diff = 0
    for _ in range(50):
        a = rng.standard_normal(97).tolist()
        n_diff += ref_py.gemm_nt(a, b, True) != ref_py.gemm_nt(a, b, False)
    assert n_diff > 0, "the serial and the blocked summation orders should round differently somewhere"
    # the selection follows the same rule: 40 candidates, top 12
    emb = rng.standard_normal((40, D))
    score = rng.random(40)
    idx, st = oracle_lib.dpp_request(emb, score, 12, alpha=1.0, window_size=10)
    L = ref_py.kernel_matrix(emb.tolist(), score.tolist(), 1.0)
    assert st == 0 and ref_py.dpp_with_window(L, 12, 10) == idx.tolist()


def test_dpp_kernel_matrix_matches_closed_form(oracle_lib):
    rng = np.random.default_rng(5)
    emb = rng.standard_normal((30, 12))
    rel = rng.random(30)
    L = oracle_lib.dpp_kernel_matrix(emb, rel, alpha=1.5)
    e = emb / np.linalg.norm(emb, axis=1, keepdims=True)
    S = (1 + e @ e.T) / 2
    q = np.exp(1.5 * rel)
    assert np.allclose(L, q[:, None] * S * q[None, :], rtol=1e-12)


def test_dpp_edge_cases(oracle_lib):
    rng = np.random.default_rng(6)
    emb = rng.standard_normal((25, 8))
    score = rng.random(25)
    # ctx.Size > n in windowed mode: index 0 is repeated once the candidates run out (dpp_sort.go:477-499)
    idx, st = oracle_lib.dpp_request(emb, score, 50, window_size=10)
    assert st == 0 and len(idx) == 50 and sorted(set(idx[:25].tolist())) == list(range(25)) and (idx[25:] == 0).all()
    # ctx.Size <= window: one call, clamped to n
    idx, st = oracle_lib.dpp_request(emb[:5], score[:5], 8, window_size=10)
    assert len(idx) == 5 and sorted(idx.tolist()) == list(range(5))
    # duplicate embeddings + alpha 0: second copy has d2 ~ 0 -> early stop + lowest-index fill (:539-548)
    emb2 = np.repeat(rng.standard_normal((1, 8)), 6, axis=0)
    idx, st = oracle_lib.dpp_request(emb2, np.zeros(6), 4, alpha=0.0, window_size=10)
    assert idx.tolist() == [0, 1, 2, 3]
    # all-zero scores with z-score normalisation -> error path, items unchanged
    idx, st = oracle_lib.dpp_request(emb, np.zeros(25), 10, norm_mode=1)
    assert st == 1 and idx.tolist() == list(range(25))
    # CandidateCount / MinScorePercent truncation (:280-300)
    idx, st = oracle_lib.dpp_request(emb, score, 5, candidate_count=10)
    top10 = set(np.argsort(-score)[:10].tolist())
    assert set(idx.tolist()) <= top10


def test_go_sort_properties(oracle_lib):
    g = np.load(os.path.join(G, "sort_small.npz"))
    assert (oracle_lib.go_sort(g["score"]) == g["go_perm"]).all()
    assert (oracle_lib.stable_sort_desc(g["score"]) == g["stable_perm"]).all()
    rng = np.random.default_rng(7)
    for n in (0, 1, 2, 12, 13, 49, 50, 51, 333, 1000, 4097):
        s = rng.random(n)
        p = oracle_lib.go_sort(s)
        assert sorted(p.tolist()) == list(range(n))
        assert (np.diff(s[p]) <= 0).all()
        assert (p == oracle_lib.stable_sort_desc(s)).all()     # distinct scores: every correct sort agrees
    # already-sorted input (recall output feeding ItemRankScore) is left untouched, ties included
    s = -np.sort(-np.round(rng.random(1000), 2))
    assert (oracle_lib.go_sort(s) == np.arange(1000)).all()
    # heavy ties: still a sorted permutation; tie groups are the same sets as the stable order
    s = np.round(rng.random(500), 1)
    p, q = oracle_lib.go_sort(s), oracle_lib.stable_sort_desc(s)
    assert (s[p] == s[q]).all() and sorted(p.tolist()) == list(range(500))


def test_algo_score_sort_switch(oracle_lib):
    # sort/algo_score_sort.go:45-49
    score = np.array([0.2, 0.9, 0.5])
    field = np.array([3.0, 1.0, 2.0])
    assert oracle_lib.algo_score_sort(score, field, switch_threshold=0.95).tolist() == [0, 2, 1]   # by field
    assert oracle_lib.algo_score_sort(score, field, switch_threshold=0.5).tolist() == [1, 2, 0]    # by current score


def test_key_order_properties(oracle_lib):
    lib = oracle_lib.lib()
    vals = [float("-inf"), -1.5, -0.0, 0.0, 1e-30, 2.5, float("inf")]
    ks = [lib.orc_make_key(v, 10) for v in vals]
    assert ks == sorted(ks) and len(set(ks)) == len(ks)
    assert lib.orc_make_key(float("nan"), 10) < ks[0]           # NaN ranks below -inf
    assert lib.orc_make_key(1.0, 3) > lib.orc_make_key(1.0, 4)  # ties: lower row first
    assert lib.orc_key_row(lib.orc_make_key(1.0, 12345)) == 12345 and lib.orc_key_score(lib.orc_make_key(-2.5, 1)) == -2.5


def test_ssd_edge_cases(oracle_lib):
    rng = np.random.default_rng(8)
    emb = rng.standard_normal((40, 12))
    score = rng.random(40)
    idx, st = oracle_lib.ssd_request(emb, score, 15, gamma=0.25, window_size=5)
    assert st == 0 and len(idx) == 15 and len(set(idx.tolist())) == 15
    assert idx[0] == int(np.argmax(score))                       # first pick = best score (ssd_sort.go:394)
    # gamma == 0: sorted list returned unchanged (:304-307)
    idx0, st0 = oracle_lib.ssd_request(emb, score, 15, gamma=0.0)
    assert st0 == 1 and idx0.tolist() == oracle_lib.go_sort(score).tolist()
    # ctx.Size > n: T = min(N, ctx.Size) (:395)
    idx1, st1 = oracle_lib.ssd_request(emb[:6], score[:6], 30, gamma=0.25)
    assert st1 == 0 and sorted(idx1.tolist()) == list(range(6))
    # a large gamma makes the pick sequence diversity driven: orthogonal-ish items first
    e2 = np.eye(4)[[0, 0, 1, 2]] + 1e-3 * rng.standard_normal((4, 4))
    idx2, _ = oracle_lib.ssd_request(e2, np.array([0.9, 0.8, 0.1, 0.1]), 3, gamma=10.0, window_size=5)
    assert idx2[0] == 0 and 1 not in idx2.tolist()[:3]          # the near-duplicate of item 0 is not picked


def test_reciprocal_division_is_correctly_rounded_for_f32_operands():
    """dpp_cluster.cu phase 2a replaces av/scale (both f32 values widened to f64) by q0 = av*y, rem = fma(-q0, scale, av),
    q = fma(rem, y, q0) with y = RN(1/scale).  Check against exact rational arithmetic that q == RN(av/scale)."""
    import random
    import struct
    from fractions import Fraction

    def fma(x, y, z):
        return float(Fraction(x) * Fraction(y) + Fraction(z))

    def f32(x):
        return struct.unpack("f", struct.pack("f", x))[0]

    rng = random.Random(7)
    cases = []
    for i in range(20000):
        if i % 3 == 0:
            a = float(rng.getrandbits(23) | (1 << 23))
            b = float(rng.getrandbits(23) | (1 << 23)) * 2.0 ** rng.randint(0, 3)
        else:
            a = f32(abs(rng.gauss(0, 1)) * 10 ** rng.uniform(-6, 6))
            b = f32(abs(rng.gauss(0, 1)) * 10 ** rng.uniform(-6, 6))
        if a == 0 or b == 0:
            continue
        cases.append((min(a, b), max(a, b)))
    cases += [(1.0, 1.0), (f32(1e-45), f32(3e38)), (f32(1.17549435e-38), f32(3.4028235e38)), (3.0, 3.0), (1.0, 3.0)]
    for a, b in cases:
        y = float(Fraction(1) / Fraction(b))
        q0 = a * y
        q = fma(fma(-q0, b, a), y, q0)
        assert q == float(Fraction(a) / Fraction(b)), (a, b)


def test_go_sort_order_three_restatements_agree(oracle_lib):
    """Go's pdqsort (sort.Sort / sort.Slice, go1.19+) restated three times — oracle.c, the library's host entry
    prg_sort_desc_host (csrc/sort.cu) and, written separately, tests/ref_py.go_sort_perm — must give the same permutation
    on random, tied, sorted, reversed, organ-pipe, constant and periodic inputs (the tie order is part of "bit-exact final
    ordering"; no Go toolchain here to ask the real thing: baseline/go/sort/b200_sort_test.go does that elsewhere)."""
    from pairec_b200.binding import sort_desc_host
    rng = np.random.default_rng(0)
    n_cases = 0
    for n in (2, 5, 12, 13, 49, 50, 51, 100, 333, 1000, 2000):
        a = rng.random(n)
        for s in (a, np.round(a, 1), np.sort(a), np.sort(a)[::-1].copy(), np.zeros(n), (np.arange(n) % 7).astype(float),
                  np.concatenate([np.arange(n // 2), np.arange(n - n // 2)[::-1]]).astype(float)):
            s = np.ascontiguousarray(s, dtype=np.float64)
            want = oracle_lib.go_sort(s).tolist()
            assert ref_py.go_sort_perm(s.tolist(), True) == want, n
            assert sort_desc_host(s).tolist() == want, n
            assert ref_py.go_sort_perm(s.tolist(), False) == oracle_lib.go_sort(s, descending=False).tolist(), n
            n_cases += 1
    assert n_cases == 77
