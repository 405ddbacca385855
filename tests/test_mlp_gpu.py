"""GPU parity of the tcgen05 dense tower through prg_rank: 1e-5 relative on scores."""
import numpy as np
import pytest

from tests import synth

pytestmark = pytest.mark.gpu

RTOL = 1e-5  # north-star tolerance for float rank scores


def _setup(engine, n_items, n_fields, dims, seed=6):
    fields, factors, linear = synth.rank_tables(n_items=n_items, n_fields=n_fields)
    # larger factors than the FM tests so that the tower is not in its linear regime
    factors = [f * 8 for f in factors]
    engine.set_item_fields(fields)
    for t, (f, l) in enumerate(zip(factors, linear)):
        engine.set_feature_table(t, f, l)
    engine.set_fm_bias(0.02)
    W, b = synth.mlp_weights(dims, seed=seed)
    engine.set_mlp(dims, W, b)
    return fields, factors, linear, W, b


# shape (3, 333): 999 rows = 8 row blocks -> the CTA-pair kernel (tcgen05 cta_group::2, TMA-stored activations);
# shape (2, 320): 640 rows = 5 row blocks (odd) -> the single-CTA persistent kernel
@pytest.mark.parametrize("shape", [(3, 333), (2, 320)])
@pytest.mark.parametrize("dims,n_fields", [([512, 512, 256, 128, 1], 32), ([64, 64, 1], 4), ([256, 192, 64, 1], 16)])
def test_mlp_scores_within_tolerance(engine, oracle_lib, dims, n_fields, shape):
    from pairec_b200.binding import MODEL_MLP, MODEL_FM_MLP
    n_items = 3000
    fields, factors, linear, W, b = _setup(engine, n_items, n_fields, dims)
    rng = np.random.default_rng(1)
    rows = rng.integers(0, n_items, size=shape).astype(np.uint32)   # not a multiple of the 128-row tile
    rows[-1, -3:] = 0xFFFFFFFF
    got = engine.rank(MODEL_MLP, rows)
    fm_logit, x = oracle_lib.gather_fm(fields, factors, linear, 0.02, rows.reshape(-1), want_x=True)
    mlp_logit = oracle_lib.mlp_forward(x, dims, W, b)
    want = oracle_lib.sigmoid(mlp_logit).astype(np.float64).reshape(rows.shape)
    live = rows != 0xFFFFFFFF
    assert (got[~live] == 0).all()
    rel = np.abs(got[live] - want[live]) / np.abs(want[live])
    assert rel.max() <= RTOL, f"max relative error {rel.max():.3e}"
    # logits spread over a useful range (the test is not vacuous)
    assert np.ptp(mlp_logit) > 0.05

    got2 = engine.rank(MODEL_FM_MLP, rows)
    want2 = oracle_lib.sigmoid((fm_logit + mlp_logit).astype(np.float32)).astype(np.float64).reshape(rows.shape)
    rel2 = np.abs(got2[live] - want2[live]) / np.abs(want2[live])
    assert rel2.max() <= RTOL


def test_mlp_rejects_unsupported_shapes(engine):
    from pairec_b200 import PrgError
    W, b = synth.mlp_weights([64, 100, 1])
    with pytest.raises(PrgError):
        engine.set_mlp([64, 100, 1], W, b)
