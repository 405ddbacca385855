package sort

// Shared helpers of the B200 parity replay (baseline/go/README.md of pairec_b200).  In-package test file: copy into
// the reference's sort/ directory.

import (
	"encoding/json"
	"os"
	"path/filepath"
	"strconv"
	"testing"

	"github.com/alibaba/pairec/v2/context"
	"github.com/alibaba/pairec/v2/module"
	"github.com/aliyun/aliyun-pairec-config-go-sdk/v2/model"
)

func b200FixtureDir(t *testing.T) string {
	dir := os.Getenv("PAIREC_B200_FIXTURES")
	if dir == "" {
		t.Skip("PAIREC_B200_FIXTURES is not set")
	}
	return dir
}

func b200Load(t *testing.T, name string, out interface{}) {
	raw, err := os.ReadFile(filepath.Join(b200FixtureDir(t), name))
	if err != nil {
		t.Fatalf("fixture %s: %v", name, err)
	}
	if err := json.Unmarshal(raw, out); err != nil {
		t.Fatalf("fixture %s: %v", name, err)
	}
}

func b200Glob(t *testing.T, pattern string) []string {
	m, err := filepath.Glob(filepath.Join(b200FixtureDir(t), pattern))
	if err != nil || len(m) == 0 {
		t.Fatalf("no fixtures match %s (%v)", pattern, err)
	}
	for i := range m {
		m[i] = filepath.Base(m[i])
	}
	return m
}

// newB200Context: Size = ctx.Size (items to return); ExperimentResult must answer GetExperimentParams() with an
// object whose GetFloat / GetInt return the supplied defaults (dpp_sort.go:373-382, ssd_sort.go:352-368 read their
// knobs through it and dereference it without a nil check).  abtest overrides ride in `params`.
func newB200Context(size int, params map[string]interface{}) *context.RecommendContext {
	ctx := context.NewRecommendContext()
	ctx.Size = size
	ctx.RecommendId = "b200-parity"
	// [UNVERIFIED-UPSTREAM] constructor of the SDK's ExperimentResult; the SDK source is not in the reference tree.
	res := model.NewExperimentResult("b200_parity", &model.ExperimentContext{RequestId: "b200-parity", Uid: "u"})
	if len(params) > 0 {
		layer := model.NewLayerParams()
		for k, v := range params {
			layer.AddParam(k, v)
		}
		res.LayerParamsMap["b200"] = layer
		res.Init()
	}
	ctx.ExperimentResult = res
	return ctx
}

func b200Items(score []float64, emb [][]float64) []*module.Item {
	items := make([]*module.Item, len(score))
	for i := range score {
		it := module.NewItem(strconv.Itoa(i))
		it.Score = score[i]
		if emb != nil {
			it.Embedding = append([]float64(nil), emb[i]...) // SSD mutates embeddings in place (ssd_sort.go:423-431)
		}
		items[i] = it
	}
	return items
}

func b200Index(t *testing.T, it *module.Item) int {
	i, err := strconv.Atoi(string(it.Id))
	if err != nil {
		t.Fatalf("item id %v", it.Id)
	}
	return i
}
