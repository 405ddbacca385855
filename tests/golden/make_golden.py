"""Generates the committed golden fixtures (tests/golden/*.npz) from seeded inputs through the CPU oracle.

The reference holds no golden vectors for this path and cannot be built or imported here (no Go
toolchain), so the fixtures freeze the oracle's own outputs: a later change to the oracle or to the CUDA path that
moves any of these bits fails tests/test_golden_*.py.  Run from the repo root:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
from tests import synth  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def main():
    rng = np.random.default_rng(20260925)
    # recall: 3000 x 64, 5 queries, k = 40 (with a duplicated block -> ties)
    E = (rng.standard_normal((3000, 64)) / 8).astype(np.float32)
    E[1500:1600] = E[100:200]
    Q = (rng.standard_normal((5, 64)) / 8).astype(np.float32)
    keys = oracle.recall_topk(E, Q, 40, row_base=7)
    np.savez_compressed(os.path.join(OUT, "recall_small.npz"), E=E, Q=Q, k=40, row_base=7, keys=keys)

    # gather + FM + MLP: 500 items, 8 fields
    fields, factors, linear = synth.rank_tables(n_items=500, n_fields=8, max_rows=300, seed=11)
    factors = [f * 8 for f in factors]
    rows = rng.integers(0, 500, size=64).astype(np.uint32)
    logit, x = oracle.gather_fm(fields, factors, linear, 0.05, rows)
    dims = [128, 64, 64, 1]
    W, b = synth.mlp_weights(dims, seed=12)
    mlp = oracle.mlp_forward(x, dims, W, b)
    d = dict(fields=fields, rows=rows, w0=np.float32(0.05), fm_logit=logit, x=x, dims=np.array(dims), mlp_logit=mlp,
             fm_score=oracle.sigmoid(logit), score=oracle.sigmoid((logit + mlp).astype(np.float32)))
    for t in range(8):
        d[f"factors{t}"] = factors[t]
        d[f"linear{t}"] = linear[t]
    for l in range(3):
        d[f"W{l}"] = W[l]
        d[f"b{l}"] = b[l]
    np.savez_compressed(os.path.join(OUT, "rank_small.npz"), **d)

    # DPP: 80 candidates, 24-d, top 25, window 10 (+ a second case with norm_mode 2 on sorted scores)
    emb = rng.standard_normal((80, 24)).astype(np.float32)
    score = rng.random(80)
    idx, st = oracle.dpp_request(emb.astype(np.float64), score, 25, alpha=1.0, window_size=10)
    score2 = -np.sort(-rng.random(80))
    idx2, st2 = oracle.dpp_request(emb.astype(np.float64), score2, 12, alpha=2.0, window_size=5, norm_mode=2)
    np.savez_compressed(os.path.join(OUT, "dpp_small.npz"), emb=emb, score=score, idx=idx, status=st, score2=score2,
                        idx2=idx2, status2=st2)

    # sorts: ties included
    s = np.round(rng.random(300), 2)
    np.savez_compressed(os.path.join(OUT, "sort_small.npz"), score=s, go_perm=oracle.go_sort(s),
                        stable_perm=oracle.stable_sort_desc(s))
    print("golden fixtures written to", OUT)


if __name__ == "__main__":
    main()
