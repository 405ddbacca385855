"""Item-matrix snapshot swap (the reference's vector DAO switches partition tables in the background,
module/vector_hologres_dao.go:40-61): prg_stage_item_matrix builds the new snapshot beside the live one,
prg_commit_item_matrix swaps them between two batches.  Parity of every answer against the CPU oracle."""
import threading

import numpy as np
import pytest

from pairec_b200 import Engine
from pairec_b200.binding import PrgError

pytestmark = pytest.mark.gpu


def _data(n, d, b, seed):
    rng = np.random.default_rng(seed)
    E = (rng.standard_normal((n, d)) / np.sqrt(d)).astype(np.float32)
    Q = (rng.standard_normal((b, d)) / np.sqrt(d)).astype(np.float32)
    return E, Q


def _same(engine, oracle, E, Q, k, row_base=0):
    rows, scores, n = engine.recall_topk(Q, k)
    orows, oscores, on = oracle.keys_split(oracle.recall_topk(E, Q, k, row_base=row_base))
    return bool((n == on).all() and (rows == orows).all()
                and (scores.view(np.uint32) == oscores.view(np.uint32)).all())


def test_stage_then_commit_swaps_snapshots(oracle_lib):
    eng = Engine(device=0)
    try:
        E1, Q = _data(300_000, 64, 9, seed=5)
        E2, _ = _data(420_000, 64, 9, seed=6)      # a different row count: every derived table changes size
        with pytest.raises(PrgError):
            eng.commit_item_matrix()                # nothing staged
        eng.set_item_matrix(E1)
        assert _same(eng, oracle_lib, E1, Q, 200)
        eng.stage_item_matrix(E2, row_base=1000)
        assert _same(eng, oracle_lib, E1, Q, 200), "staging must not disturb the live snapshot"
        eng.commit_item_matrix()
        assert _same(eng, oracle_lib, E2, Q, 200, row_base=1000), "after the commit the new snapshot answers"
        with pytest.raises(PrgError):
            eng.commit_item_matrix()                # the staged slot is empty again
        # staging twice keeps the later one; an uncommitted snapshot is released with the handle
        E3, _ = _data(280_000, 64, 9, seed=7)
        eng.stage_item_matrix(E1)
        eng.stage_item_matrix(E3)
        eng.commit_item_matrix()
        assert _same(eng, oracle_lib, E3, Q, 200)
        eng.stage_item_matrix(E1)
    finally:
        eng.close()


def test_requests_are_served_while_a_snapshot_is_staged(oracle_lib):
    eng = Engine(device=0)
    try:
        E1, Q = _data(300_000, 64, 5, seed=11)
        E2, _ = _data(300_000, 64, 5, seed=12)
        eng.set_item_matrix(E1)
        want1 = oracle_lib.keys_split(oracle_lib.recall_topk(E1, Q, 100))
        want2 = oracle_lib.keys_split(oracle_lib.recall_topk(E2, Q, 100))
        stop = threading.Event()
        bad, served = [], [0]

        def serve():          # every answer must come from ONE snapshot, old or new, never a mixture
            while not stop.is_set():
                rows, scores, n = eng.recall_topk(Q, 100)
                ok1 = (rows == want1[0]).all() and (scores.view(np.uint32) == want1[1].view(np.uint32)).all()
                ok2 = (rows == want2[0]).all() and (scores.view(np.uint32) == want2[1].view(np.uint32)).all()
                if not (ok1 or ok2):
                    bad.append(served[0])
                served[0] += 1

        t = threading.Thread(target=serve)
        t.start()
        for _ in range(3):
            eng.stage_item_matrix(E2)
            eng.commit_item_matrix()
            eng.stage_item_matrix(E1)
            eng.commit_item_matrix()
        eng.stage_item_matrix(E2)
        eng.commit_item_matrix()
        stop.set()
        t.join()
        assert not bad, f"answers from a torn snapshot at requests {bad[:5]}"
        assert served[0] > 0
        rows, scores, n = eng.recall_topk(Q, 100)
        assert (rows == want2[0]).all()
    finally:
        eng.close()
