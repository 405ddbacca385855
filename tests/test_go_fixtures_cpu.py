"""The JSON fixtures baseline/go/*_test.go replays through the real reference code (on a machine with Go) must keep
reproducing from the oracle: a change to oracle.c that moves any expected index or L element fails here first."""
import json
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DATA = os.path.join(ROOT, "baseline", "go", "testdata")


def test_committed_go_fixtures_reproduce_from_the_oracle(oracle_lib):
    from tests.golden import export_go_fixtures as ex
    fresh = ex.fixtures()
    assert sorted(fresh) == sorted(f for f in os.listdir(DATA) if f.endswith(".json"))
    for name, fx in fresh.items():
        with open(os.path.join(DATA, name)) as f:
            committed = json.load(f)
        assert committed == json.loads(json.dumps(fx)), f"{name} differs from what the oracle produces now"


def test_go_harness_files_name_the_reference_symbols():
    """Cheap guard against the harness drifting from the reference API it binds (no Go toolchain here to compile it)."""
    src = {n: open(os.path.join(ROOT, "baseline", "go", d, n)).read()
           for d, n in (("sort", "b200_dpp_test.go"), ("sort", "b200_ssd_test.go"), ("sort", "b200_sort_test.go"),
                        ("sort", "b200_ctx_test.go"), ("algorithm", "b200_lookup_test.go"))}
    assert "DPPWithWindow(" in src["b200_dpp_test.go"] and "s.KernelMatrix(ctx, items, lenEmb, hasTable)" in src["b200_dpp_test.go"]
    assert "RegisterEmbeddingHook(" in src["b200_dpp_test.go"]
    assert "s.SSDWithSlidingWindow(items, ctx)" in src["b200_ssd_test.go"]
    assert "gosort.Sort(gosort.Reverse(ItemScoreSlice(items)))" in src["b200_sort_test.go"]
    assert "NewLookupPolicy()" in src["b200_lookup_test.go"] and "p.Run(batch)" in src["b200_lookup_test.go"]
    ref = "/root/reference"
    if os.path.isdir(ref):   # the signatures the harness relies on, read from the reference where it is available
        dpp = open(os.path.join(ref, "sort", "dpp_sort.go")).read()
        assert "func DPPWithWindow(L *mat.Dense, topN int, windowSize int) []int" in dpp
        assert "func (s *DPPSort) KernelMatrix(context *context.RecommendContext, items []*module.Item, lenEmb int, hasTable bool) (*mat.Dense, error)" in dpp
        assert "func RegisterEmbeddingHook(name string, fn EmbeddingHookFunc)" in dpp
        ssd = open(os.path.join(ref, "sort", "ssd_sort.go")).read()
        assert "func (s *SSDSort) SSDWithSlidingWindow(items []*module.Item, ctx *context.RecommendContext) []*module.Item" in ssd
        for field in ("alpha", "windowSize", "normalizeEmb", "ensurePosSimilarity", "embeddingHookNames"):
            assert field in dpp
        for field in ("gamma", "windowSize", "useSSDStar", "tableName"):
            assert field in ssd
