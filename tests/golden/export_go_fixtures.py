"""Exports the JSON fixtures that baseline/go/*_test.go replays through the REAL reference code (sort.DPP,
DPPWithWindow, DPPSort.KernelMatrix, SSDSort.SSDWithSlidingWindow, LookupPolicy.Run, sort.Sort order) on a machine with
a Go toolchain.  Inputs are seeded; expected outputs are the CPU oracle's (oracle/oracle.c) — the same oracle the CUDA
path is tested against.  float64 values are written with repr() (shortest string that round-trips), which Go's
encoding/json parses back to the same bits.  Run from the repo root:  python tests/golden/export_go_fixtures.py
tests/test_go_fixtures_cpu.py checks that the committed files still reproduce from the oracle."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402

OUT = os.path.join(ROOT, "baseline", "go", "testdata")


def _f(a):
    return [float(x) for x in np.asarray(a, dtype=np.float64).ravel()]


def _m(a):
    return [[float(x) for x in row] for row in np.asarray(a, dtype=np.float64)]


def dpp_fixture(name, n, dim, top_n, window, seed, alpha=1.0, norm_mode=0, normalize_emb=True, hook_dim=0, use_table=True,
                ensure_pos=True, sort_scores=False, zero_scores=False):
    rng = np.random.default_rng(seed)
    emb = rng.standard_normal((n, dim)).astype(np.float32).astype(np.float64) if use_table else None
    hook = (rng.standard_normal((n, hook_dim)) * 0.7) if hook_dim else None
    score = rng.random(n)
    if sort_scores:
        score = -np.sort(-score)
    if zero_scores:
        score[:] = 0.0
    idx, st, Ld, L0 = oracle.dpp_request_ex(emb, score, top_n, hook=hook, use_table=use_table,
                                            no_positive_sim=0 if ensure_pos else 1, alpha=alpha, window_size=window,
                                            norm_mode=norm_mode, normalize_emb=1 if normalize_emb else 0, want_L=True)
    return {"name": name, "emb": _m(emb) if use_table else [], "hook": _m(hook) if hook_dim else [], "score": _f(score),
            "alpha": alpha, "top_n": top_n, "window": window, "norm_mode": norm_mode, "normalize_emb": normalize_emb,
            "ensure_positive_sim": ensure_pos, "expect_idx": [int(i) for i in idx] if st == 0 else [],
            "expect_l_diag": _f(Ld) if st == 0 else [], "expect_l_row0": _f(L0) if st == 0 else [], "expect_status": st,
            # quality terms exp(alpha * score) as the oracle's libm gives them (norm_mode 0 only): Go's math.Exp can differ in
            # the last bit, and the test that pins gonum's Dgemm order takes them from here
            "expect_r": _f(oracle.exp(alpha * score)) if (st == 0 and norm_mode == 0) else []}


def ssd_fixture(name, n, dim, top_n, window, seed, gamma=0.25, norm_mode=0, use_ssd_star=False):
    rng = np.random.default_rng(seed)
    emb = rng.standard_normal((n, dim)).astype(np.float32).astype(np.float64)
    emb *= (1.0 + 0.3 * np.sin(np.arange(n)))[:, None]     # SSD's quality term uses the residual norms
    score = rng.random(n)
    idx, st = oracle.ssd_request(emb, score, top_n, gamma=gamma, window_size=window, norm_mode=norm_mode, normalize_emb=0,
                                 use_ssd_star=1 if use_ssd_star else 0)
    return {"name": name, "emb": _m(emb), "score": _f(score), "gamma": gamma, "top_n": top_n, "window": window,
            "norm_mode": norm_mode, "use_ssd_star": use_ssd_star, "expect_idx": [int(i) for i in idx], "expect_status": st}


def sort_fixture(name, n, seed, decimals=None):
    rng = np.random.default_rng(seed)
    s = rng.random(n)
    if decimals is not None:
        s = np.round(s, decimals)          # ties: Go's pdqsort is not stable, the oracle restates its order
    return {"name": name, "score": _f(s), "expect_perm": [int(i) for i in oracle.go_sort(s)]}


def lookup_fixture(seed):
    rng = np.random.default_rng(seed)
    v = rng.random(64)
    present = rng.random(64) < 0.7
    return {"field_name": "score", "value": _f(v), "present": [bool(p) for p in present],
            "expect": _f(oracle.lookup(v, present.astype(np.uint8)))}


def fixtures():
    return {
        # BASELINE.json configs[3] shape: n = 1000 candidates, D = 128, top 50, window 10
        "dpp_c4.json": dpp_fixture("c4", 1000, 128, 50, 10, seed=101),
        "dpp_small.json": dpp_fixture("small", 80, 24, 25, 10, seed=102),
        "dpp_single_call.json": dpp_fixture("single_call", 120, 16, 8, 10, seed=103, alpha=2.0),
        "dpp_raw_emb.json": dpp_fixture("raw_emb", 90, 16, 20, 7, seed=104, alpha=0.5, normalize_emb=False),
        "dpp_zscore.json": dpp_fixture("zscore", 100, 16, 20, 10, seed=105, norm_mode=1),
        "dpp_minmax.json": dpp_fixture("minmax", 100, 16, 20, 10, seed=106, norm_mode=2, sort_scores=True),
        "dpp_zero_scores.json": dpp_fixture("zero_scores", 50, 8, 10, 10, seed=107, norm_mode=1, zero_scores=True),
        "dpp_topn_gt_n.json": dpp_fixture("topn_gt_n", 25, 16, 50, 10, seed=108),
        # both sides of gonum's serial / blocked Dgemm switch (at most 64 items: one DotUnitary over all of k; more: 64-wide
        # k blocks) with D + 1 > 64, where the two orders round L differently (oracle.c g_gemm_serial)
        "dpp_gemm_serial_48.json": dpp_fixture("gemm_serial_48", 48, 96, 12, 10, seed=113),
        "dpp_gemm_serial_64.json": dpp_fixture("gemm_serial_64", 64, 80, 20, 10, seed=114),
        "dpp_gemm_blocked_65.json": dpp_fixture("gemm_blocked_65", 65, 80, 20, 10, seed=115),
        "dpp_hook_table.json": dpp_fixture("hook_table", 200, 32, 30, 10, seed=109, hook_dim=6),
        "dpp_hook_only.json": dpp_fixture("hook_only", 200, 0, 30, 10, seed=110, hook_dim=12, use_table=False),
        "dpp_hook_raw.json": dpp_fixture("hook_raw", 200, 0, 30, 10, seed=111, hook_dim=12, use_table=False, normalize_emb=False),
        "dpp_hook_nopos.json": dpp_fixture("hook_nopos", 200, 0, 30, 10, seed=112, hook_dim=12, use_table=False, ensure_pos=False),
        "ssd_small.json": ssd_fixture("small", 300, 32, 30, 5, seed=201),
        "ssd_star.json": ssd_fixture("star", 200, 16, 20, 4, seed=202, gamma=0.5, use_ssd_star=True),
        "ssd_zscore.json": ssd_fixture("zscore", 200, 16, 20, 5, seed=203, norm_mode=1),
        "sort_distinct.json": sort_fixture("distinct", 1000, seed=301),
        "sort_ties.json": sort_fixture("ties", 1000, seed=302, decimals=2),
        "sort_short.json": sort_fixture("short", 11, seed=303, decimals=1),
        "lookup.json": lookup_fixture(seed=401),
    }


def main():
    os.makedirs(OUT, exist_ok=True)
    for name, fx in fixtures().items():
        with open(os.path.join(OUT, name), "w") as f:
            json.dump(fx, f, separators=(",", ":"))
    print("fixtures written to", OUT)


if __name__ == "__main__":
    main()
