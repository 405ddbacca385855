package algorithm

import (
	"encoding/json"
	"math"
	"os"
	"path/filepath"
	"testing"

	"github.com/alibaba/pairec/v2/algorithm/response"
	"github.com/alibaba/pairec/v2/recconf"
)

// lookup.json: LookupPolicy.Run (lookup.go:37-51): score = features[FieldName].(float64), 0.5 when the key is absent,
// nil for an empty batch.
type b200LookupFixture struct {
	FieldName string    `json:"field_name"`
	Value     []float64 `json:"value"`
	Present   []bool    `json:"present"`
	Expect    []float64 `json:"expect"`
}

func TestB200LookupPolicy(t *testing.T) {
	dir := os.Getenv("PAIREC_B200_FIXTURES")
	if dir == "" {
		t.Skip("PAIREC_B200_FIXTURES is not set")
	}
	raw, err := os.ReadFile(filepath.Join(dir, "lookup.json"))
	if err != nil {
		t.Fatal(err)
	}
	var f b200LookupFixture
	if err := json.Unmarshal(raw, &f); err != nil {
		t.Fatal(err)
	}
	p := NewLookupPolicy()
	if err := p.Init(&recconf.AlgoConfig{Name: "b200", Type: "LOOKUP", LookupConf: recconf.LookupConfig{FieldName: f.FieldName}}); err != nil {
		t.Fatal(err)
	}
	if out, err := p.Run([]map[string]interface{}{}); out != nil || err != nil {
		t.Fatalf("empty batch: (%v, %v), want (nil, nil)", out, err)
	}
	batch := make([]map[string]interface{}, len(f.Value))
	for i := range batch {
		batch[i] = map[string]interface{}{"other": "x"}
		if f.Present[i] {
			batch[i][f.FieldName] = f.Value[i]
		}
	}
	out, err := p.Run(batch)
	if err != nil {
		t.Fatal(err)
	}
	res := out.([]response.AlgoResponse)
	if len(res) != len(f.Expect) {
		t.Fatalf("%d responses, want %d", len(res), len(f.Expect))
	}
	for i, r := range res {
		if math.Float64bits(r.GetScore()) != math.Float64bits(f.Expect[i]) {
			t.Fatalf("response %d: %v, oracle %v", i, r.GetScore(), f.Expect[i])
		}
	}
}
