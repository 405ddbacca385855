package sort

import (
	gosort "sort"
	"testing"
)

// ssd_*.json: inputs in the order they reach doSort; doSort sorts by score first (ssd_sort.go:301), then
// SSDWithSlidingWindow (:346-486) picks min(N, ctx.Size) items.
type b200SSDFixture struct {
	Name       string      `json:"name"`
	Emb        [][]float64 `json:"emb"` // as held by the cache: AFTER the load-time normalisation when normalize_emb
	Score      []float64   `json:"score"`
	Gamma      float64     `json:"gamma"`
	TopN       int         `json:"top_n"`
	Window     int         `json:"window"`
	NormMode   int         `json:"norm_mode"`
	UseSSDStar bool        `json:"use_ssd_star"`
	ExpectIdx  []int       `json:"expect_idx"` // indices into the INPUT list
}

func TestB200SSDWithSlidingWindow(t *testing.T) {
	for _, name := range b200Glob(t, "ssd_*.json") {
		var f b200SSDFixture
		b200Load(t, name, &f)
		items := b200Items(f.Score, f.Emb)
		gosort.Sort(gosort.Reverse(ItemScoreSlice(items))) // doSort :301
		s := &SSDSort{gamma: f.Gamma, windowSize: f.Window, useSSDStar: f.UseSSDStar, normalizeEmb: true,
			ensurePosSimilarity: true, tableName: "b200"}
		params := map[string]interface{}{}
		if f.NormMode != 0 {
			params["ssd_norm_quality_score"] = f.NormMode
		}
		ctx := newB200Context(f.TopN, params)
		out := s.SSDWithSlidingWindow(items, ctx)
		if len(out) != len(f.ExpectIdx) {
			t.Fatalf("%s: %d items, oracle has %d", name, len(out), len(f.ExpectIdx))
		}
		for i, it := range out {
			if got := b200Index(t, it); got != f.ExpectIdx[i] {
				t.Fatalf("%s: position %d is item %d, oracle has %d", name, i, got, f.ExpectIdx[i])
			}
		}
	}
}
