"""GPU parity of feature gather + FM through prg_rank: FM scores are bit-exact."""
import numpy as np
import pytest

from tests import synth

pytestmark = pytest.mark.gpu


def _load(engine, fields, factors, linear, w0=0.1):
    engine.set_item_fields(fields)
    for t, (f, l) in enumerate(zip(factors, linear)):
        engine.set_feature_table(t, f, l)
    engine.set_fm_bias(w0)


@pytest.mark.parametrize("n_fields", [32, 5, 1])
def test_fm_scores_bit_exact(engine, oracle_lib, n_fields):
    from pairec_b200.binding import MODEL_FM
    fields, factors, linear = synth.rank_tables(n_items=4000, n_fields=n_fields)
    _load(engine, fields, factors, linear)
    rng = np.random.default_rng(0)
    rows = rng.integers(0, 4000, size=(7, 333)).astype(np.uint32)
    rows[0, -5:] = 0xFFFFFFFF           # padding
    rows[1, 0] = 4000 + 17              # out-of-range item row -> padding semantics
    got = engine.rank(MODEL_FM, rows)
    logit, _ = oracle_lib.gather_fm(fields, factors, linear, 0.1, rows.reshape(-1), want_x=False)
    want = oracle_lib.sigmoid(logit).astype(np.float64).reshape(rows.shape)
    want[rows == 0xFFFFFFFF] = 0.0
    want[1, 0] = oracle_lib.sigmoid(np.float32(0.0))  # live == false -> logit 0
    assert (got.view(np.uint64) == want.view(np.uint64)).all()


def test_fm_out_of_range_ids_and_missing_linear(engine, oracle_lib):
    from pairec_b200.binding import MODEL_FM
    fields, factors, linear = synth.rank_tables(n_items=1000, n_fields=8)
    fields[::7, 3] = 0xFFFFFFF0          # id beyond the table -> the field contributes zeros
    linear[2] = None
    _load(engine, fields, factors, linear, w0=-0.3)
    rows = np.arange(1000, dtype=np.uint32).reshape(1, -1)
    got = engine.rank(MODEL_FM, rows)
    logit, _ = oracle_lib.gather_fm(fields, factors, linear, -0.3, rows.reshape(-1), want_x=False)
    want = oracle_lib.sigmoid(logit).astype(np.float64).reshape(rows.shape)
    assert (got.view(np.uint64) == want.view(np.uint64)).all()
