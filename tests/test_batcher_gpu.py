"""Cross-call request batcher (prg_batcher_*): many host threads, one request each, coalesced into prg_recommend
batches.  Every request must get exactly what prg_recommend returns for its query alone, whatever batch it lands in
(the staged oracle checks the batch call itself in test_pipeline_gpu.py)."""
import threading

import numpy as np
import pytest

from tests import synth

pytestmark = pytest.mark.gpu


def _setup(eng, n_items, d, seed):
    rng = np.random.default_rng(seed)
    E = (rng.standard_normal((n_items, d)) / 8).astype(np.float32)
    fields, factors, linear = synth.rank_tables(n_items=n_items, n_fields=32)
    D = synth.diversity(n_items=n_items, dim=32)
    eng.set_item_matrix(E)
    eng.set_item_fields(fields)
    for t, (f, l) in enumerate(zip(factors, linear)):
        eng.set_feature_table(t, f, l)
    eng.set_fm_bias(0.05)
    eng.set_diversity_matrix(D)
    return rng


@pytest.mark.parametrize("max_batch,max_wait_us,n_threads", [(64, 0, 96), (16, 200, 40), (1, 0, 5)])
def test_batched_requests_equal_direct_calls(max_batch, max_wait_us, n_threads):
    from pairec_b200 import DppParams, Engine
    from pairec_b200.binding import MODEL_FM, Batcher
    n_items, d, k, T, per_thread = 300_000, 64, 400, 15, 4
    eng = Engine(0)
    try:
        rng = _setup(eng, n_items, d, 5)
        Q = (rng.standard_normal((n_threads * per_thread, d)) / 8).astype(np.float32)
        p = DppParams(top_n=T, alpha=1.0, window_size=10)
        want_rows, want_scores, want_n = [], [], []
        for i in range(0, Q.shape[0], 64):   # direct batch calls (checked against the oracle elsewhere)
            r, s, n = eng.recommend(Q[i:i + 64], k, MODEL_FM, p)
            want_rows.append(r), want_scores.append(s), want_n.append(n)
        want_rows, want_scores, want_n = np.concatenate(want_rows), np.concatenate(want_scores), np.concatenate(want_n)
        bat = Batcher(eng, k, MODEL_FM, p, max_batch=max_batch, max_wait_us=max_wait_us)
        errors = []

        def client(t):
            try:
                for j in range(per_thread):
                    i = t * per_thread + j
                    rows, scores = bat.recommend(Q[i])
                    assert len(rows) == want_n[i]
                    assert (rows == want_rows[i, :want_n[i]]).all(), f"request {i}: rows differ"
                    assert (scores.view(np.uint64) == want_scores[i, :want_n[i]].view(np.uint64)).all()
            except Exception as ex:  # noqa: BLE001
                errors.append(ex)

        th = [threading.Thread(target=client, args=(t,)) for t in range(n_threads)]
        [t.start() for t in th]
        [t.join(timeout=120) for t in th]
        assert not errors, errors[0]
        st = bat.stats()
        assert st["requests"] == n_threads * per_thread
        assert sum(st["size_hist"]) == st["batches"] <= st["requests"]
        if max_batch == 1:
            assert st["batches"] == st["requests"]
        elif n_threads > 2 * max_batch // 3:
            assert st["batches"] < st["requests"], "concurrent callers were never coalesced"
        bat.close()
    finally:
        eng.close()


def test_batcher_reports_errors_to_every_caller_and_stops_cleanly():
    from pairec_b200 import DppParams, Engine, PrgError
    from pairec_b200.binding import MODEL_MLP, Batcher
    eng = Engine(0)
    try:
        _setup(eng, 100_000, 64, 6)
        p = DppParams(top_n=8)
        bat = Batcher(eng, 100, MODEL_MLP, p, max_batch=8)   # no tower was set: every batch fails with PRG_ESTATE
        got = []

        def client():
            try:
                bat.recommend(np.zeros(64, dtype=np.float32))
                got.append(None)
            except PrgError as ex:
                got.append(ex.code)

        th = [threading.Thread(target=client) for _ in range(12)]
        [t.start() for t in th]
        [t.join(timeout=60) for t in th]
        assert len(got) == 12 and all(c not in (None, 0) for c in got)
        bat.close()
        bat.close()   # idempotent
    finally:
        eng.close()


def test_batched_requests_survive_a_repaired_recall():
    """Adversarial row order (the sampled threshold fails for one query of the batch): the batcher's asynchronous call
    has already copied the unrepaired results to the host when the deferred check trips; the repair must copy them
    again before any caller is woken.  Every request still gets what the direct call returns."""
    from pairec_b200 import DppParams, Engine
    from pairec_b200.binding import MODEL_FM, Batcher
    n, d, k, T = 400_000, 64, 500, 12
    rng = np.random.default_rng(11)
    E = (rng.standard_normal((n, d)) * 0.01).astype(np.float32)
    n_tiles = (n + 255) // 256
    stride = n_tiles // max(64, n_tiles // 128)
    tile = np.arange(n) // 256
    E[:, 0] = np.where(tile % stride == 0, 0.0, 1.0 + rng.random(n) * 0.5).astype(np.float32)
    Q = (rng.standard_normal((24, d)) * 0.01).astype(np.float32)
    Q[::3] = 0
    Q[::3, 0] = 1.0                      # every third request defeats the strided sample
    fields, factors, linear = synth.rank_tables(n_items=n, n_fields=32)
    D = synth.diversity(n_items=n, dim=32)
    eng = Engine(0)
    try:
        eng.set_item_matrix(E)
        eng.set_item_fields(fields)
        for t, (f, l) in enumerate(zip(factors, linear)):
            eng.set_feature_table(t, f, l)
        eng.set_fm_bias(0.05)
        eng.set_diversity_matrix(D)
        p = DppParams(top_n=T, alpha=1.0, window_size=10)
        want_rows, want_scores, want_n = eng.recommend(Q, k, MODEL_FM, p)
        assert eng.recall_stats()["fallback_queries"] >= 1
        bat = Batcher(eng, k, MODEL_FM, p, max_batch=8, max_wait_us=0)
        errors = []

        def client(i):
            try:
                for _ in range(3):
                    rows, scores = bat.recommend(Q[i])
                    assert len(rows) == want_n[i]
                    assert (rows == want_rows[i, :want_n[i]]).all(), f"request {i}: rows differ"
                    assert (scores.view(np.uint64) == want_scores[i, :want_n[i]].view(np.uint64)).all()
            except Exception as ex:  # noqa: BLE001
                errors.append(ex)

        th = [threading.Thread(target=client, args=(i,)) for i in range(Q.shape[0])]
        [t.start() for t in th]
        [t.join(timeout=120) for t in th]
        assert not errors, errors[0]
        bat.close()
    finally:
        eng.close()
