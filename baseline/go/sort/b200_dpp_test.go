package sort

import (
	"math"
	"testing"

	"github.com/alibaba/pairec/v2/context"
	"github.com/alibaba/pairec/v2/module"
	"gonum.org/v1/gonum/floats"
	"gonum.org/v1/gonum/mat"
)

// dpp_*.json / dpp_hook_*.json (tests/golden/export_go_fixtures.py)
type b200DPPFixture struct {
	Name          string      `json:"name"`
	Emb           [][]float64 `json:"emb"`  // table embeddings as loaded (BEFORE the load-time normalisation, dpp_sort.go:234-237); may be empty
	Hook          [][]float64 `json:"hook"` // hook embeddings per item (GenerateEmbedding output); may be empty
	Score         []float64   `json:"score"`
	Alpha         float64     `json:"alpha"`
	TopN          int         `json:"top_n"`
	Window        int         `json:"window"`
	NormMode      int         `json:"norm_mode"`
	NormalizeEmb  bool        `json:"normalize_emb"`
	EnsurePosSim  bool        `json:"ensure_positive_sim"`
	ExpectIdx     []int       `json:"expect_idx"`
	ExpectLDiag   []float64   `json:"expect_l_diag"`   // diag(L), bit-exact
	ExpectLRow0   []float64   `json:"expect_l_row0"`   // row 0 of L, bit-exact
	ExpectStatus  int         `json:"expect_status"`   // 1: KernelMatrix returns an error ("all item score is zero")
	ExpectR       []float64   `json:"expect_r"`        // exp(alpha*score) from the oracle's libm (norm_mode 0 only; may be empty)
}

// b200SameBits / b200LastBits: L elements are compared bit for bit; a difference of a few units in the last place is
// reported separately, because two inputs of L are platform dependent in the last bit on the Go side and cannot be pinned
// by any fixture: math.Exp (assembly kernels on amd64 / arm64 / s390x, the oracle calls libm) and, in the z-score mode,
// floats.Sum (its amd64 kernel peels an element when the slice is not 16-byte aligned).
func b200SameBits(a, b float64) bool { return math.Float64bits(a) == math.Float64bits(b) }
func b200LastBits(got, want float64) bool {
	ulp := math.Abs(math.Nextafter(want, math.Inf(1)) - want)
	return math.Abs(got-want) <= 8*ulp
}

func b200LoadNormalised(f *b200DPPFixture) [][]float64 {
	// loadEmbeddingCache: floats.Norm / floats.Scale(1/normV) at load (dpp_sort.go:234-237)
	out := make([][]float64, len(f.Emb))
	for i, e := range f.Emb {
		v := append([]float64(nil), e...)
		if f.NormalizeEmb {
			floats.Scale(1/floats.Norm(v, 2), v)
		}
		out[i] = v
	}
	return out
}

func b200CheckIdx(t *testing.T, name string, got []int, want []int) {
	if len(got) != len(want) {
		t.Fatalf("%s: %d indices, oracle has %d", name, len(got), len(want))
	}
	for i := range got {
		if got[i] != want[i] {
			t.Fatalf("%s: pick %d is item %d, oracle picked %d", name, i, got[i], want[i])
		}
	}
}

// DPPWithWindow / DPP (dpp_sort.go:477-551) on an L assembled with gonum exactly as KernelMatrix does
// (:428-431 features, :463-472 the three dense products) — needs nothing but gonum and the two exported functions.
func TestB200DPPFromGonumKernel(t *testing.T) {
	for _, name := range b200Glob(t, "dpp_[^h]*.json") {
		var f b200DPPFixture
		b200Load(t, name, &f)
		if f.NormMode != 0 || f.ExpectStatus != 0 {
			continue // relevance normalisation lives inside KernelMatrix: TestB200KernelMatrix covers it
		}
		emb := b200LoadNormalised(&f)
		n, d := len(emb), len(emb[0])
		feat := mat.NewDense(n, d+1, nil)
		raw := make([]float64, n)
		expDiff := 0
		for i := range emb {
			row := append(append([]float64(nil), emb[i]...), 1)
			floats.Scale(1/math.Sqrt2, row)
			feat.SetRow(i, row)
			raw[i] = math.Exp(f.Alpha * f.Score[i])
			if len(f.ExpectR) == n { // pin gonum's summation order independently of math.Exp's last bit
				if !b200SameBits(raw[i], f.ExpectR[i]) {
					expDiff++
				}
				raw[i] = f.ExpectR[i]
			}
		}
		if expDiff > 0 {
			t.Logf("%s: math.Exp differs from libm exp in the last bit for %d of %d items (platform dependent; the oracle's values are used)", name, expDiff, n)
		}
		var sim mat.Dense
		sim.Mul(feat, feat.T())
		ru := mat.NewDense(n, n, nil)
		for i, v := range raw {
			ru.Set(i, i, v)
		}
		var L mat.Dense
		L.Mul(ru, &sim)
		L.Mul(&L, ru)
		for i := 0; i < n; i++ {
			if math.Float64bits(L.At(i, i)) != math.Float64bits(f.ExpectLDiag[i]) {
				t.Fatalf("%s: L[%d][%d] = %v, oracle %v", name, i, i, L.At(i, i), f.ExpectLDiag[i])
			}
			if math.Float64bits(L.At(0, i)) != math.Float64bits(f.ExpectLRow0[i]) {
				t.Fatalf("%s: L[0][%d] = %v, oracle %v", name, i, L.At(0, i), f.ExpectLRow0[i])
			}
		}
		b200CheckIdx(t, name, DPPWithWindow(&L, f.TopN, f.Window), f.ExpectIdx)
	}
}

// The real (*DPPSort).KernelMatrix (dpp_sort.go:372-475): table path, hook-only path (normalizeEmb /
// ensurePosSimilarity variants, :432-447) and hook + table (concat re-normalised, :418-421).
func TestB200KernelMatrix(t *testing.T) {
	for _, name := range b200Glob(t, "dpp_*.json") {
		var f b200DPPFixture
		b200Load(t, name, &f)
		hasTable := len(f.Emb) > 0
		var emb [][]float64
		if hasTable {
			emb = b200LoadNormalised(&f)
		}
		items := b200Items(f.Score, emb)
		s := &DPPSort{alpha: f.Alpha, windowSize: f.Window, normalizeEmb: f.NormalizeEmb, ensurePosSimilarity: f.EnsurePosSim}
		lenEmb := 0
		if hasTable {
			lenEmb = len(f.Emb[0])
		}
		if len(f.Hook) > 0 {
			hookName := "b200_hook_" + name
			hooks := f.Hook
			RegisterEmbeddingHook(hookName, func(_ *context.RecommendContext, item *module.Item) []float64 {
				return append([]float64(nil), hooks[b200Index(t, item)]...)
			})
			s.embeddingHookNames = []string{hookName}
			lenEmb += len(f.Hook[0])
		}
		params := map[string]interface{}{}
		if f.NormMode != 0 {
			params["dpp_norm_relevance_score"] = f.NormMode
		}
		ctx := newB200Context(f.TopN, params)
		L, err := s.KernelMatrix(ctx, items, lenEmb, hasTable)
		if f.ExpectStatus != 0 {
			if err == nil {
				t.Fatalf("%s: oracle expects KernelMatrix to fail", name)
			}
			continue
		}
		if err != nil {
			t.Fatalf("%s: %v", name, err)
		}
		lastBit := 0
		for i := range items {
			for _, c := range [2][3]float64{{L.At(i, i), f.ExpectLDiag[i], float64(i)}, {L.At(0, i), f.ExpectLRow0[i], 0}} {
				if b200SameBits(c[0], c[1]) {
					continue
				}
				if !b200LastBits(c[0], c[1]) {
					t.Fatalf("%s: L[%d][%d] = %v, oracle %v", name, int(c[2]), i, c[0], c[1])
				}
				lastBit++
			}
		}
		if lastBit > 0 {
			t.Logf("%s: %d elements of L differ from the oracle in the last bits (math.Exp / floats.Sum are platform dependent; "+
				"TestB200DPPFromGonumKernel pins the summation order with the oracle's quality terms)", name, lastBit)
		}
		b200CheckIdx(t, name, DPPWithWindow(L, f.TopN, f.Window), f.ExpectIdx)
	}
}
