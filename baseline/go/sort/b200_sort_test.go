package sort

import (
	gosort "sort"
	"testing"
)

// sort_*.json: ItemRankScoreSort.doSort (item_rank_score.go:26-32) = sort.Sort(sort.Reverse(ItemScoreSlice)); the
// oracle restates Go's pdqsort, so the expected permutation includes the (unstable) order of tied scores.
type b200SortFixture struct {
	Name       string    `json:"name"`
	Score      []float64 `json:"score"`
	ExpectPerm []int     `json:"expect_perm"` // expect_perm[i] = input index of the item at output position i
}

func TestB200ItemRankScoreOrder(t *testing.T) {
	for _, name := range b200Glob(t, "sort_*.json") {
		var f b200SortFixture
		b200Load(t, name, &f)
		items := b200Items(f.Score, nil)
		gosort.Sort(gosort.Reverse(ItemScoreSlice(items)))
		for i, it := range items {
			if got := b200Index(t, it); got != f.ExpectPerm[i] {
				t.Fatalf("%s: position %d holds item %d (score %v), oracle has item %d", name, i, got, it.Score, f.ExpectPerm[i])
			}
		}
	}
}
