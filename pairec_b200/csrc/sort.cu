// sort.cu — score sorts.
//
//  * prg_sort_desc_host: host restatement of Go's sort.Sort(sort.Reverse(ItemScoreSlice(items)))
//    (sort/item_rank_score.go:29 with Less of sort/item_score.go:15-18).  Go >= 1.19 sorts with pattern-defeating
//    quicksort, which is deterministic but not stable; to keep the reference's order among equal scores the same
//    algorithm is run on a permutation here.  Sequential, n is a few thousand at most: host logic, no GPU.
//  * prg_sort_desc: batched device sort used between rank and DPP inside the fused path.  Total order
//    (score descending, input position ascending) — identical to the reference whenever scores are distinct, and the
//    documented deviation (stable instead of pdqsort tie order) otherwise.
#include "handle.h"
#include <cstdlib>
#include <vector>

namespace prg {

// ------------------------------------------------------------------ host: Go pdqsort on a permutation
namespace gosort {
struct Ctx {
  const double* score;
  int32_t* perm;
  // sort.Reverse(ItemScoreSlice): Less(i, j) = items[j].Score < items[i].Score
  bool less(int i, int j) const { return score[perm[j]] < score[perm[i]]; }
  void swap(int i, int j) { int32_t t = perm[i]; perm[i] = perm[j]; perm[j] = t; }
};
static int bit_len(unsigned x) { int n = 0; for (; x; x >>= 1) ++n; return n; }

static void insertion_sort(Ctx& c, int a, int b) {
  for (int i = a + 1; i < b; ++i)
    for (int j = i; j > a && c.less(j, j - 1); --j) c.swap(j, j - 1);
}
static void sift_down(Ctx& c, int lo, int hi, int first) {
  for (int root = lo;;) {
    int child = 2 * root + 1;
    if (child >= hi) return;
    if (child + 1 < hi && c.less(first + child, first + child + 1)) ++child;
    if (!c.less(first + root, first + child)) return;
    c.swap(first + root, first + child);
    root = child;
  }
}
static void heap_sort(Ctx& c, int a, int b) {
  const int first = a, hi = b - a;
  for (int i = (hi - 1) / 2; i >= 0; --i) sift_down(c, i, hi, first);
  for (int i = hi - 1; i >= 0; --i) { c.swap(first, first + i); sift_down(c, 0, i, first); }
}
static void break_patterns(Ctx& c, int a, int b) {
  const int length = b - a;
  if (length < 8) return;
  uint64_t rnd = (uint64_t)length;  // xorshift seeded with the length
  const unsigned modulus = 1u << bit_len((unsigned)length);
  const int idx = a + (length / 4) * 2 - 1;
  for (int i = 0; i < 3; ++i) {
    rnd ^= rnd << 13; rnd ^= rnd >> 7; rnd ^= rnd << 17;
    int other = (int)((unsigned)rnd & (modulus - 1));
    if (other >= length) other -= length;
    c.swap(idx - 1 + i, a + other);
  }
}
static int median3(Ctx& c, int a, int b, int d, int& swaps) {
  auto order2 = [&](int& x, int& y) { if (c.less(y, x)) { int t = x; x = y; y = t; ++swaps; } };
  order2(a, b); order2(b, d); order2(a, b);
  return b;
}
enum Hint { kUnknown, kIncreasing, kDecreasing };
static int choose_pivot(Ctx& c, int a, int b, Hint& hint) {
  const int l = b - a;
  int swaps = 0, i = a + l / 4, j = a + l / 4 * 2, k = a + l / 4 * 3;
  if (l >= 8) {
    if (l >= 50) {  // Tukey ninther
      i = median3(c, i - 1, i, i + 1, swaps);
      j = median3(c, j - 1, j, j + 1, swaps);
      k = median3(c, k - 1, k, k + 1, swaps);
    }
    j = median3(c, i, j, k, swaps);
  }
  hint = swaps == 0 ? kIncreasing : (swaps == 12 ? kDecreasing : kUnknown);
  return j;
}
static bool partial_insertion_sort(Ctx& c, int a, int b) {
  int i = a + 1;
  for (int step = 0; step < 5; ++step) {
    while (i < b && !c.less(i, i - 1)) ++i;
    if (i == b) return true;
    if (b - a < 50) return false;
    c.swap(i, i - 1);
    if (i - a >= 2)
      for (int j = i - 1; j >= 1 && c.less(j, j - 1); --j) c.swap(j, j - 1);
    if (b - i >= 2)
      for (int j = i + 1; j < b && c.less(j, j - 1); ++j) c.swap(j, j - 1);
  }
  return false;
}
static int partition_equal(Ctx& c, int a, int b, int pivot) {
  c.swap(a, pivot);
  int i = a + 1, j = b - 1;
  for (;;) {
    while (i <= j && !c.less(a, i)) ++i;
    while (i <= j && c.less(a, j)) --j;
    if (i > j) break;
    c.swap(i, j); ++i; --j;
  }
  return i;
}
static int partition(Ctx& c, int a, int b, int pivot, bool& already) {
  c.swap(a, pivot);
  int i = a + 1, j = b - 1;
  while (i <= j && c.less(i, a)) ++i;
  while (i <= j && !c.less(j, a)) --j;
  if (i > j) { c.swap(j, a); already = true; return j; }
  c.swap(i, j); ++i; --j;
  for (;;) {
    while (i <= j && c.less(i, a)) ++i;
    while (i <= j && !c.less(j, a)) --j;
    if (i > j) break;
    c.swap(i, j); ++i; --j;
  }
  c.swap(j, a);
  already = false;
  return j;
}
static void pdqsort(Ctx& c, int a, int b, int limit) {
  bool balanced = true, partitioned = true;
  for (;;) {
    const int length = b - a;
    if (length <= 12) { insertion_sort(c, a, b); return; }
    if (limit == 0) { heap_sort(c, a, b); return; }
    if (!balanced) { break_patterns(c, a, b); --limit; }
    Hint hint;
    int pivot = choose_pivot(c, a, b, hint);
    if (hint == kDecreasing) {
      for (int i = a, j = b - 1; i < j; ++i, --j) c.swap(i, j);
      pivot = (b - 1) - (pivot - a);
      hint = kIncreasing;
    }
    if (balanced && partitioned && hint == kIncreasing && partial_insertion_sort(c, a, b)) return;
    if (a > 0 && !c.less(a - 1, pivot)) { a = partition_equal(c, a, b, pivot); continue; }
    bool already;
    const int mid = partition(c, a, b, pivot, already);
    partitioned = already;
    const int left = mid - a, right = b - mid, thr = length / 8;
    if (left < right) { balanced = left >= thr; pdqsort(c, a, mid, limit); a = mid + 1; }
    else { balanced = right >= thr; pdqsort(c, mid + 1, b, limit); b = mid; }
  }
}
}  // namespace gosort

// ------------------------------------------------------------------ device: batched bitonic sort of (score, idx)
__device__ __forceinline__ uint64_t f64_ord(double d) {
  uint64_t u = (uint64_t)__double_as_longlong(d);
  if ((u & 0x7FFFFFFFFFFFFFFFull) > 0x7FF0000000000000ull) return 0ull;  // NaN sorts last
  return (u >> 63) ? ~u : (u | 0x8000000000000000ull);
}

// One CTA per list.  smem: P2 u64 keys + P2 i32 indices.  perm[i] = input index at output position i.
// Optionally also writes the list in sorted order (rows_o / scores_o: what the fused path feeds to DPP; saves the
// separate apply-permutation launch).
__global__ void __launch_bounds__(1024) sort_desc_kernel(const double* __restrict__ score, int n, int32_t* __restrict__ perm,
                                                         const uint32_t* __restrict__ rows, uint32_t* __restrict__ rows_o,
                                                         double* __restrict__ scores_o) {
  pdl_wait();                 // chained launch: the predecessor's writes are visible from here on
  pdl_launch_dependents();
  extern __shared__ __align__(16) uint8_t sort_smem[];
  uint32_t P2 = 32;
  while (P2 < (uint32_t)n) P2 <<= 1;
  uint64_t* key = reinterpret_cast<uint64_t*>(sort_smem);
  int32_t* idx = reinterpret_cast<int32_t*>(key + P2);
  const double* s = score + (size_t)blockIdx.x * n;
  for (uint32_t i = threadIdx.x; i < P2; i += blockDim.x) {
    key[i] = (i < (uint32_t)n) ? f64_ord(s[i]) : 0ull;
    idx[i] = (i < (uint32_t)n) ? (int32_t)i : 0x7FFFFFFF;  // padding sorts after every real element
  }
  __syncthreads();
  for (uint32_t size = 2; size <= P2; size <<= 1) {
    for (uint32_t stride = size >> 1; stride > 0; stride >>= 1) {
      for (uint32_t i = threadIdx.x; i < (P2 >> 1); i += blockDim.x) {
        const uint32_t pos = 2 * i - (i & (stride - 1));
        const uint64_t ka = key[pos], kb = key[pos + stride];
        const int32_t ia = idx[pos], ib = idx[pos + stride];
        const bool a_after_b = (ka < kb) || (ka == kb && ia > ib);  // a belongs after b in the final order
        const bool fwd = (pos & size) == 0;
        if (a_after_b == fwd) { key[pos] = kb; key[pos + stride] = ka; idx[pos] = ib; idx[pos + stride] = ia; }
      }
      __syncthreads();
    }
  }
  int32_t* o = perm + (size_t)blockIdx.x * n;
  for (uint32_t i = threadIdx.x; i < (uint32_t)n; i += blockDim.x) {
    const int32_t src = idx[i];
    o[i] = src;
    if (rows_o) {
      rows_o[(size_t)blockIdx.x * n + i] = rows[(size_t)blockIdx.x * n + src];
      scores_o[(size_t)blockIdx.x * n + i] = s[src];
    }
  }
}

int sort_desc_device(prg_handle* h, const double* score_dev, int B, int n, int32_t* perm_dev, const uint32_t* rows_dev,
                     uint32_t* rows_sorted, double* scores_sorted) {
  if (n > 8192) return fail(PRG_EUNSUPPORTED, "prg_sort_desc: n > 8192");
  uint32_t P2 = 32;
  while (P2 < (uint32_t)n) P2 <<= 1;
  const size_t smem = (size_t)P2 * 12;
  StageScope span(h, ST_SORT);
  if (smem > 48 * 1024)
    PRG_CUDA(cudaFuncSetAttribute(sort_desc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const unsigned threads = P2 / 2 < 1024 ? (P2 / 2 < 32 ? 32 : P2 / 2) : 1024;
  PRG_CUDA(launch_chained(h, sort_desc_kernel, dim3(B), dim3(threads), smem, 1, score_dev, n, perm_dev, rows_dev, rows_sorted,
                          scores_sorted));
  count_launch(h);
  return PRG_OK;
}

}  // namespace prg

using namespace prg;

extern "C" {

int prg_sort_desc_host(const double* score, int n, int32_t* out_perm) {
  if (n < 0 || (n > 0 && (!score || !out_perm))) return fail(PRG_EINVAL, "null buffer");
  for (int i = 0; i < n; ++i) out_perm[i] = i;
  if (n <= 1) return PRG_OK;
  gosort::Ctx c{score, out_perm};
  gosort::pdqsort(c, 0, n, gosort::bit_len((unsigned)n));
  return PRG_OK;
}

int prg_sort_desc(prg_handle* h, const double* score, int B, int n, int32_t* out_perm, int mem) {
  if (!h) return fail(PRG_EINVAL, "null handle");
  if (!score || !out_perm || B <= 0 || n <= 0) return fail(PRG_EINVAL, "bad arguments");
  std::lock_guard<std::mutex> lk(h->mu);
  PRG_CUDA(cudaSetDevice(h->device));
  prg::resolve_pending(h);
  if (mem == PRG_MEM_DEVICE) return sort_desc_device(h, score, B, n, out_perm, nullptr, nullptr, nullptr);
  const size_t cnt = (size_t)B * n;
  PRG_TRY(h->sort_in.ensure(cnt * 8));
  PRG_TRY(h->sort_perm.ensure(cnt * 4));
  PRG_CUDA(cudaMemcpyAsync(h->sort_in.p, score, cnt * 8, cudaMemcpyHostToDevice, h->stream));
  PRG_TRY(sort_desc_device(h, (const double*)h->sort_in.p, B, n, (int32_t*)h->sort_perm.p, nullptr, nullptr, nullptr));
  PRG_CUDA(cudaMemcpyAsync(out_perm, h->sort_perm.p, cnt * 4, cudaMemcpyDeviceToHost, h->stream));
  PRG_CUDA(cudaStreamSynchronize(h->stream));
  return PRG_OK;
}

}  // extern "C"
