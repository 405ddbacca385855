// recall.cu — exact inner-product top-k over the HBM-resident item matrix.
//
// Replaces the remote faiss VectorRetrieval.Search behind algorithm/faiss/vector_client.go:32-42 that
// service/recall/vector_recall.go:88 reaches through algorithm.Run.  Arithmetic contract (defined here, mirrored by
// oracle/oracle.c orc_recall_topk): score(row,q) = fmaf chain over dims 0..d-1 from +0, one accumulator; total
// order = 64-bit key (score desc, row asc).
//
// Pipeline per block of <=64 queries:
//   1. scan<DENSE> over a strided sample of tiles      -> sample keys
//   2. select (r-th largest of the sample)             -> per-query threshold key tau
//   3. scan<THRESH> over the whole matrix              -> candidates with key >= tau (expected ~4k per query)
//   4. select (exact top-k of the candidates, sorted)  -> keys / rows / scores
// A query whose candidate list under- or overflows (adversarial row order) is redone through the dense path
// (all keys materialised, same select), so the result is exact for every input.
//
// The scan kernel is persistent (one CTA per SM): one producer warp streams 256-row tiles through a 3-stage
// TMA/mbarrier ring (SWIZZLE_128B so the per-row LDS.128 reads are bank-conflict free), eight consumer warps hold a
// 4-row x 16-query register tile per thread and issue packed FFMA2; the threshold test is fused into the tile
// epilogue, so the only HBM traffic is the matrix itself (algorithmic bytes = rows*dim*4 per <=64-query block).
#include "handle.h"
#include "recall.h"
#include "bitonic.cuh"
#include <cmath>
#include <cstdlib>
#include <vector>

namespace prg {

template <int DIM>
constexpr size_t scan_smem_bytes() {
  return (size_t)kStages * kStageBytes + (size_t)DIM * kQB * 4 + kQB * 4 + kQB * 8 +
         2 * kStages * 8 + kQB * 4;
}

template <int DIM, int MODE>
__global__ void __launch_bounds__(kScanThreads, 1)
recall_scan_kernel(const __grid_constant__ CUtensorMap emap, const ScanParams pin) {
  extern __shared__ __align__(1024) uint8_t smem[];  // SWIZZLE_128B boxes need 1024-B aligned destinations
  ScanParams p = pin;
  if (MODE == SCAN_DENSE && gridDim.y > 1) {  // one grid row per block of 64 queries (the sample of a large batch)
    const int y = (int)blockIdx.y;
    p.Q = pin.Q + (size_t)y * kQB * DIM;
    p.nq = pin.nq - y * kQB < kQB ? pin.nq - y * kQB : kQB;
    p.dense = pin.dense + (size_t)y * kQB * pin.dense_stride;
  }
  float* stage_base = reinterpret_cast<float*>(smem);
  float* Qs = reinterpret_cast<float*>(smem + (size_t)kStages * kStageBytes);  // [DIM][64]
  float* tauf = Qs + DIM * kQB;                                                // [64]
  uint64_t* tau64 = reinterpret_cast<uint64_t*>(tauf + kQB);                   // [64]
  uint64_t* full = tau64 + kQB;
  uint64_t* empty = full + kStages;
  uint32_t* s_cnt = reinterpret_cast<uint32_t*>(empty + kStages);                // [64] candidates appended by this CTA

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  constexpr int KH = DIM / 64;

  for (int i = tid; i < DIM * kQB; i += kScanThreads) {
    const int q = i / DIM, dd = i - q * DIM;
    Qs[dd * kQB + q] = (q < p.nq) ? p.Q[(size_t)q * DIM + dd] : 0.f;
  }
  if (tid < kQB) {
    uint64_t t = 0;
    float tf = __int_as_float(0x7F800000);  // +inf: padded queries never pass the pre-test
    if (tid < p.nq) {
      if (MODE == SCAN_THRESH) t = p.tau[tid];
      tf = (t == 0) ? __int_as_float(0xFF800000) : key_score(t);
    }
    tau64[tid] = t;
    tauf[tid] = tf;
    s_cnt[tid] = 0;
  }
  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], kConsumerWarps);
    }
    mbar_fence_init();
  }
  __syncthreads();

  const uint32_t my_tiles = (p.n_tiles > blockIdx.x) ? (p.n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

  if (warp == kConsumerWarps) {
    // ------------------------------------------------------------ producer: one lane drives TMA
    if (lane == 0) {
      tma_prefetch_desc(&emap);
      uint32_t it = 0;
      for (uint32_t i = 0; i < my_tiles; ++i) {
        const uint32_t t = blockIdx.x + i * gridDim.x;
        const int row0 = (int)(t * p.tile_stride * (uint32_t)kTileRows);
        for (int h = 0; h < KH; ++h, ++it) {
          const uint32_t s = it % kStages, ph = (it / kStages) & 1u;
          mbar_wait(&empty[s], ph ^ 1u);
          mbar_arrive_expect_tx(&full[s], kStageBytes);
          float* dst = stage_base + (size_t)s * kStageFloats;
          tma_load_2d(dst, &emap, h * 64, row0, &full[s], kEvictFirst);
          tma_load_2d(dst + kSubTileFloats, &emap, h * 64 + 32, row0, &full[s], kEvictFirst);
        }
      }
    }
    return;
  }

  // -------------------------------------------------------------- consumers
  const int wr = warp / WQ, wq = warp - wr * WQ;
  const int rloc0 = wr * (32 * TM) + lane;  // this thread's rows: rloc0 + 32*m
  const int qbase = wq * TQ;
  const int sw = lane & 7;                  // == row & 7 for every row of this thread (swizzle phase)
  uint32_t it = 0;

  for (uint32_t i = 0; i < my_tiles; ++i) {
    float2 acc[TM][TQ / 2];
#pragma unroll
    for (int m = 0; m < TM; ++m)
#pragma unroll
      for (int c = 0; c < TQ / 2; ++c) acc[m][c] = make_float2(0.f, 0.f);

    for (int h = 0; h < KH; ++h, ++it) {
      const uint32_t s = it % kStages, ph = (it / kStages) & 1u;
      mbar_wait(&full[s], ph);
      const float* st = stage_base + (size_t)s * kStageFloats;
      const float* qs = Qs + (size_t)h * 64 * kQB + qbase;
#pragma unroll kScanUnroll
      for (int c = 0; c < 16; ++c) {
        const int sub = c >> 3, ch = c & 7;
        float4 xv[TM];
#pragma unroll
        for (int m = 0; m < TM; ++m) {
          const int r = rloc0 + 32 * m;
          xv[m] = *reinterpret_cast<const float4*>(st + sub * kSubTileFloats + r * 32 + ((ch ^ sw) << 2));
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float4* qrow = reinterpret_cast<const float4*>(qs + (c * 4 + j) * kQB);
          float2 qv[TQ / 2];
#pragma unroll
          for (int v = 0; v < TQ / 4; ++v) {
            const float4 t4 = qrow[v];
            qv[2 * v] = make_float2(t4.x, t4.y);
            qv[2 * v + 1] = make_float2(t4.z, t4.w);
          }
#pragma unroll
          for (int m = 0; m < TM; ++m) {
            const float x = (j == 0) ? xv[m].x : (j == 1) ? xv[m].y : (j == 2) ? xv[m].z : xv[m].w;
            const float2 xd = make_float2(x, x);
#pragma unroll
            for (int c2 = 0; c2 < TQ / 2; ++c2) ffma2(acc[m][c2], xd, qv[c2]);
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[s]);
    }

    // ------------------------------------------------------------ tile epilogue
    const uint32_t t = blockIdx.x + i * gridDim.x;
    const uint64_t tile_row0 = (uint64_t)t * p.tile_stride * kTileRows;
    if (MODE == SCAN_THRESH) {
      // pre-test: one predicate per row of the register tile (64 FSETP per thread-tile in total); a pass is rare
      // (~1 score per warp-tile), and the append below costs one shared-memory atomic + one fire-and-forget store:
      // no global round trip stalls the warp.
      bool any_m[TM];
#pragma unroll
      for (int m = 0; m < TM; ++m) any_m[m] = false;
#pragma unroll
      for (int c2 = 0; c2 < TQ / 2; ++c2) {
        const float t0 = tauf[qbase + 2 * c2], t1 = tauf[qbase + 2 * c2 + 1];
#pragma unroll
        for (int m = 0; m < TM; ++m) {
          any_m[m] |= !(acc[m][c2].x < t0);
          any_m[m] |= !(acc[m][c2].y < t1);
        }
      }
#pragma unroll
      for (int m = 0; m < TM; ++m) {
        if (any_m[m]) {
          const uint64_t lrow = tile_row0 + (uint64_t)(rloc0 + 32 * m);
          if (lrow < p.n_rows) {
            const uint32_t grow = (uint32_t)(p.row_base + lrow);
#pragma unroll
            for (int c2 = 0; c2 < TQ / 2; ++c2) {
#pragma unroll
              for (int e = 0; e < 2; ++e) {
                const int q = qbase + 2 * c2 + e;
                const float sc = e ? acc[m][c2].y : acc[m][c2].x;
                if (!(sc < tauf[q])) {
                  const uint64_t key = make_key(sc, grow);
                  if (q < p.nq && key >= tau64[q]) {
                    const uint32_t pos = atomicAdd(&s_cnt[q], 1u);
                    if (pos < p.seg_cap) p.cand[((size_t)q * gridDim.x + blockIdx.x) * p.seg_cap + pos] = key;
                  }
                }
              }
            }
          }
        }
      }
    } else {
#pragma unroll
      for (int m = 0; m < TM; ++m) {
        const uint64_t lrow = tile_row0 + (uint64_t)(rloc0 + 32 * m);
        const bool valid = lrow < p.n_rows;
        const uint32_t grow = (uint32_t)(p.row_base + lrow);
        const uint64_t slot = (uint64_t)t * kTileRows + (uint64_t)(rloc0 + 32 * m);
#pragma unroll
        for (int c2 = 0; c2 < TQ / 2; ++c2) {
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int q = qbase + 2 * c2 + e;
            if (q < p.nq) {
              const float sc = e ? acc[m][c2].y : acc[m][c2].x;
              p.dense[(size_t)q * p.dense_stride + slot] = valid ? make_key(sc, grow) : 0ull;
            }
          }
        }
      }
    }
  }
  if (MODE == SCAN_THRESH) {
    asm volatile("bar.sync 1, %0;" ::"n"(kConsumerWarps * 32) : "memory");  // consumer warps only
    if (tid < p.nq) p.seg_cnt[(size_t)tid * gridDim.x + blockIdx.x] = s_cnt[tid];
  }
}

// ------------------------------------------------------------------ select: exact k-th / sorted top-k of a key list
template <int MODE>
__global__ void __launch_bounds__(1024, 1) select_kernel(const SelectParams p) {
  __shared__ uint32_t hist[256];
  __shared__ uint64_t s_prefix;
  __shared__ uint32_t s_want, s_fill, s_zero;
  extern __shared__ uint64_t sk[];  // MODE_TOPK: K2 keys

  const int q = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (p.skip && p.skip[q]) return;
  const uint64_t* keys = p.keys + (size_t)q * p.stride;
  uint32_t m = p.counts ? p.counts[q] : p.fixed_m;
  bool overflow = m > p.cap;
  if (p.seg_counts) {
    // pack the per-CTA segments: exclusive prefix of the (clamped) lengths, then a strided copy
    __shared__ uint32_t seg_off[513];
    if (tid == 0) {
      uint32_t run = 0;
      bool ov = false;
      for (uint32_t sgi = 0; sgi < p.n_seg; ++sgi) {
        uint32_t c = p.seg_counts[(size_t)q * p.n_seg + sgi];
        if (c > p.seg_cap) { c = p.seg_cap; ov = true; }
        seg_off[sgi] = run;
        run += c;
      }
      seg_off[p.n_seg] = run;
      s_want = ov ? 1u : 0u;
    }
    __syncthreads();
    m = seg_off[p.n_seg];
    overflow = (s_want != 0u) || m > p.cap;
    if (m > p.cap) m = p.cap;
    uint64_t* dst = p.compact + (size_t)q * p.stride;
    if (p.seg_rows) {
      // tensor-core survivors: re-score with the exact fmaf chain of the arithmetic contract (dims ascending, one
      // accumulator) — the approximate score never reaches the output
      __shared__ float qv[128];
      for (uint32_t d = tid; d < p.dim; d += 1024) qv[d] = p.Q[(size_t)q * p.dim + d];
      __syncthreads();
      for (uint32_t sgi = warp; sgi < p.n_seg; sgi += 32) {
        const uint32_t o = seg_off[sgi], c = seg_off[sgi + 1] - o;
        const uint32_t* src = p.seg_rows + ((size_t)q * p.n_seg + sgi) * p.seg_cap;
        for (uint32_t i = lane; i < c; i += 32) {
          if (o + i >= p.cap) break;
          const uint32_t grow = src[i];
          const float4* x = reinterpret_cast<const float4*>(p.E + ((size_t)grow - p.row_base) * p.dim);
          float acc = 0.f;
          for (uint32_t d4 = 0; d4 < p.dim / 4; ++d4) {
            const float4 xv = x[d4];
            acc = __fmaf_rn(xv.x, qv[4 * d4], acc);
            acc = __fmaf_rn(xv.y, qv[4 * d4 + 1], acc);
            acc = __fmaf_rn(xv.z, qv[4 * d4 + 2], acc);
            acc = __fmaf_rn(xv.w, qv[4 * d4 + 3], acc);
          }
          dst[o + i] = make_key(acc, grow);
        }
      }
    } else {
      for (uint32_t sgi = warp; sgi < p.n_seg; sgi += 32) {
        const uint32_t o = seg_off[sgi], c = seg_off[sgi + 1] - o;
        const uint64_t* src = p.keys + ((size_t)q * p.n_seg + sgi) * p.seg_cap;
        for (uint32_t i = lane; i < c; i += 32)
          if (o + i < p.cap) dst[o + i] = src[i];
      }
    }
    __syncthreads();
    keys = dst;
  }
  if (p.max_count && tid == 0) atomicMax(p.max_count, m);
  if (overflow && m > p.cap) m = p.cap;

  if (tid == 0) { s_zero = 0; s_fill = 0; }
  __syncthreads();
  {
    uint32_t z = 0;
    for (uint32_t i = tid; i < m; i += 1024) z += (keys[i] == 0ull);
    if (z) atomicAdd(&s_zero, z);
  }
  __syncthreads();
  const uint32_t valid = m - s_zero;
  const uint32_t k = (uint32_t)p.k;

  uint64_t kth = 1;  // every valid key is >= 1
  if (valid > k || (MODE == SEL_KTH && valid == k)) {
    uint64_t prefix = 0, mask = 0;
    uint32_t want = k;
    for (int pass = 0; pass < 8; ++pass) {
      const int shift = 56 - 8 * pass;
      if (tid < 256) hist[tid] = 0;
      __syncthreads();
      for (uint32_t i = tid; i < m; i += 1024) {
        const uint64_t key = keys[i];
        if (key != 0ull && (key & mask) == prefix) atomicAdd(&hist[(uint32_t)(key >> shift) & 255u], 1u);
      }
      __syncthreads();
      if (warp == 0) {
        // lane l owns digits [8l, 8l+8); suffix sums from the top digit down
        uint32_t c[8], tot = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) { c[j] = hist[lane * 8 + j]; tot += c[j]; }
        uint32_t above = 0;  // keys in digits strictly above this lane's block
        {
          uint32_t run = tot;
#pragma unroll
          for (int off = 1; off < 32; off <<= 1) {
            const uint32_t v = __shfl_down_sync(0xffffffffu, run, off);
            if (lane + off < 32) run += v;
          }
          above = run - tot;
        }
        if (above < want && above + tot >= want) {
          uint32_t a = above;
          int d = 7;
          for (; d >= 0; --d) {
            if (a + c[d] >= want) break;
            a += c[d];
          }
          s_prefix = prefix | ((uint64_t)(lane * 8 + d) << shift);
          s_want = want - a;
        }
      }
      __syncthreads();
      prefix = s_prefix;
      want = s_want;
      mask |= (0xFFull << shift);
      __syncthreads();
    }
    kth = prefix;
  } else if (MODE == SEL_KTH) {
    kth = 0;  // fewer than k valid keys: no threshold
  }

  if (MODE == SEL_KTH) {
    if (tid == 0) p.tau[q] = kth;
    return;
  }

  // compaction of keys >= kth (exactly min(k, valid) of them: keys are distinct) and bitonic sort, descending
  uint32_t K2 = 32;
  while (K2 < k) K2 <<= 1;
  for (uint32_t i = tid; i < K2; i += 1024) sk[i] = 0ull;
  __syncthreads();
  for (uint32_t i = tid; i < m; i += 1024) {
    const uint64_t key = keys[i];
    if (key != 0ull && key >= kth) {
      const uint32_t pos = atomicAdd(&s_fill, 1u);
      if (pos < K2) sk[pos] = key;
    }
  }
  __syncthreads();
  for (uint32_t size = 2; size <= K2; size <<= 1) {
    for (uint32_t stride = size >> 1; stride > 0; stride >>= 1) {
      for (uint32_t i = tid; i < (K2 >> 1); i += 1024) {
        const uint32_t pos = 2 * i - (i & (stride - 1));
        const uint64_t a = sk[pos], b = sk[pos + stride];
        const bool desc = (pos & size) == 0;
        if ((a < b) == desc) { sk[pos] = b; sk[pos + stride] = a; }
      }
      __syncthreads();
    }
  }
  const uint32_t n_out = valid < k ? valid : k;
  for (uint32_t i = tid; i < (uint32_t)p.k_out; i += 1024) {
    const uint64_t key = (i < n_out) ? sk[i] : 0ull;
    const size_t o = (size_t)q * p.k_out + i;
    if (p.out_keys) p.out_keys[o] = key;
    if (p.out_row) p.out_row[o] = key ? key_row(key) : 0xFFFFFFFFu;
    if (p.out_score) p.out_score[o] = key ? key_score(key) : __int_as_float(0xFF800000);
  }
  if (tid == 0) {
    if (p.out_n) p.out_n[q] = (int32_t)n_out;
    int fl = overflow ? 1 : (n_out < p.expect ? 2 : 0);
    if (!fl && p.tau_check && n_out > 0 && sk[n_out - 1] < p.tau_check[q]) fl = 2;  // a row outside the list could win
    if (p.flags) p.flags[q] = fl;
  }
}

// r-th largest key of a long list when r is small (the sampled threshold: r ~ 32 of ~78k keys).  One histogram pass
// over 11-bit digits, refined until the bucket holding the answer fits in shared memory, then that bucket is sorted:
// typically 2 passes over the keys instead of the 9 of the generic radix select.  Exact (keys are distinct).
constexpr uint32_t kKthCap = 4096;
__global__ void __launch_bounds__(1024, 1) select_kth_kernel(const SelectParams p) {
  pdl_wait();                 // chained launch: the predecessor's writes are visible from here on
  pdl_launch_dependents();
  __shared__ uint32_t hist[2048];
  __shared__ uint64_t s_prefix, s_mask;
  __shared__ uint32_t s_want, s_bucket, s_fill;
  extern __shared__ uint64_t sk[];  // kKthCap keys
  const int q = blockIdx.x, tid = threadIdx.x;
  if (p.skip && p.skip[q]) return;
  const uint64_t* keys = p.keys + (size_t)q * p.stride;
  const uint32_t m = p.fixed_m;
  uint64_t prefix = 0, mask = 0;
  uint32_t want = (uint32_t)p.k;
  uint64_t answer = 0;
  bool done = false;
  for (int level = 0; level < 6 && !done; ++level) {
    const int bits = (level < 5) ? 11 : 9;
    const int shift = (level < 5) ? (53 - 11 * level) : 0;
    for (int i = tid; i < 2048; i += 1024) hist[i] = 0;
    __syncthreads();
    for (uint32_t i = tid; i < m; i += 1024) {
      const uint64_t key = keys[i];
      if (key != 0ull && (key & mask) == prefix) atomicAdd(&hist[(uint32_t)(key >> shift) & ((1u << bits) - 1u)], 1u);
    }
    __syncthreads();
    if (tid < 32) {  // digits from the top: lane l owns [64l, 64l+64)
      uint32_t tot = 0;
      for (int j = 0; j < 64; ++j) tot += hist[tid * 64 + j];
      uint32_t run = tot;
#pragma unroll
      for (int off = 1; off < 32; off <<= 1) {
        const uint32_t v = __shfl_down_sync(0xffffffffu, run, off);
        if (tid + off < 32) run += v;
      }
      const uint32_t above = run - tot;
      const uint32_t grand = __shfl_sync(0xffffffffu, run, 0);
      if (grand < want) {
        if (tid == 0) s_bucket = 0xFFFFFFFFu;  // fewer than r valid keys: no threshold
      } else if (tot > 0 && above < want && above + tot >= want) {
        uint32_t a = above;
        int d = 63;
        for (; d >= 0; --d) {
          const uint32_t c = hist[tid * 64 + d];
          if (a + c >= want) break;
          a += c;
        }
        s_prefix = prefix | ((uint64_t)(tid * 64 + d) << shift);
        s_mask = mask | ((uint64_t)((1u << bits) - 1u) << shift);
        s_want = want - a;
        s_bucket = hist[tid * 64 + d];
      }
      if (tid == 0) s_fill = 0;
    }
    __syncthreads();
    if (s_bucket == 0xFFFFFFFFu) { answer = 0; done = true; break; }
    prefix = s_prefix; mask = s_mask; want = s_want;
    const uint32_t bucket = s_bucket;
    __syncthreads();
    if (bucket <= kKthCap || level == 5) {
      for (uint32_t i = tid; i < kKthCap; i += 1024) sk[i] = 0ull;
      __syncthreads();
      for (uint32_t i = tid; i < m; i += 1024) {
        const uint64_t key = keys[i];
        if (key != 0ull && (key & mask) == prefix) {
          const uint32_t pos = atomicAdd(&s_fill, 1u);
          if (pos < kKthCap) sk[pos] = key;
        }
      }
      __syncthreads();
      uint32_t K2 = 32;
      while (K2 < bucket && K2 < kKthCap) K2 <<= 1;
      for (uint32_t size = 2; size <= K2; size <<= 1) {
        for (uint32_t stride = size >> 1; stride > 0; stride >>= 1) {
          for (uint32_t i = tid; i < (K2 >> 1); i += 1024) {
            const uint32_t pos = 2 * i - (i & (stride - 1));
            const uint64_t a = sk[pos], b = sk[pos + stride];
            if ((a < b) == ((pos & size) == 0)) { sk[pos] = b; sk[pos + stride] = a; }
          }
          __syncthreads();
        }
      }
      answer = sk[want - 1];
      done = true;
    }
  }
  if (tid == 0) p.tau[q] = answer;
}

// Exact re-score of the tensor-core survivors: one WARP per (segment, query); every candidate row gets the fmaf chain
// of the arithmetic contract (dims ascending, one accumulator) — the approximate score never reaches an output.
// Rows are fetched COALESCED (a half-warp reads one 256-B row per instruction, 16 x 16 B) into a padded shared-memory
// tile and each lane then walks its own row from there: with one thread per row every 16-B load touched 32 different
// lines and the LSU, not DRAM, bounded the kernel (ncu r5: 55 us for ~290 k rows = 9 GB/s per SM, mio/lg throttle).
//
// One warp walks a GROUP of `seg_group` consecutive segments of one query as one flattened list: a row shard of a
// G-GPU run sees G times the queries with 1/G of the survivors each, and a warp per (segment, query) then holds ~4
// rows — 28 idle lanes and 75 k tiny CTAs (71 us at G = 8 against 25 us for the same number of rows at G = 1).
template <int DIM>
__global__ void __launch_bounds__(32) rescore_kernel(const uint32_t* __restrict__ seg_rows, const uint32_t* __restrict__ seg_cnt,
                                                     uint32_t n_seg, uint32_t seg_cap, uint32_t seg_group,
                                                     const float* __restrict__ E, const float* __restrict__ Q,
                                                     uint64_t row_base, uint64_t* __restrict__ seg_keys) {
  pdl_wait();                 // chained launch: the predecessor's writes are visible from here on
  pdl_launch_dependents();
  constexpr int F4 = DIM / 4;            // 16-B pieces per row
  constexpr int RPI = 32 / F4;           // rows fetched per load instruction (2 at dim 64, 1 at dim 128)
  constexpr int PITCH = DIM + 4;         // floats; keeps both the row-wise stores and the lane-per-row loads conflict free
  __shared__ __align__(16) float tile[32 * PITCH];
  __shared__ float qv[DIM];
  const uint32_t q = blockIdx.y, lane = threadIdx.x;
  const uint32_t s0 = blockIdx.x * seg_group;             // first segment of this warp's group (seg_group <= 32)
  // lane j < seg_group: length of segment s0 + j; inclusive scan -> the flattened list's offsets
  uint32_t c = 0;
  if (lane < seg_group && s0 + lane < n_seg) {
    c = seg_cnt[(size_t)q * n_seg + s0 + lane];
    if (c > seg_cap) c = seg_cap;
  }
  uint32_t incl = c;
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    const uint32_t v = __shfl_up_sync(0xffffffffu, incl, off);
    if (lane >= (uint32_t)off) incl += v;
  }
  const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
  if (total == 0) return;
  const size_t seg0 = ((size_t)q * n_seg + s0) * seg_cap;
  const uint32_t* src = seg_rows + seg0;
  uint64_t* dst = seg_keys + seg0;
  for (uint32_t d = lane; d < DIM; d += 32) qv[d] = Q[(size_t)q * DIM + d];
  const uint32_t sub = lane / F4, piece = lane % F4;
  for (uint32_t base = 0; base < total; base += 32) {
    const uint32_t n = total - base < 32 ? total - base : 32;
    // item base + lane lives in segment sg = #{j : incl_j <= item}, at position item - excl_sg
    const uint32_t item = base + lane;
    uint32_t sg = 0, excl = 0;
    for (uint32_t j = 0; j < seg_group; ++j) {
      const uint32_t e = __shfl_sync(0xffffffffu, incl, j);
      if (e <= item) { sg = j + 1; excl = e; }
    }
    const uint32_t slot = sg * seg_cap + (item - excl);   // offset from the group's first segment
    const uint32_t my_row = lane < n ? src[slot] : 0u;
    __syncwarp();
    float4 xv[32 / RPI];
#pragma unroll
    for (int r = 0; r < 32 / RPI; ++r) {
      const uint32_t row_in_chunk = r * RPI + sub;
      const uint32_t grow = __shfl_sync(0xffffffffu, my_row, row_in_chunk);
      xv[r] = row_in_chunk < n ? __ldg(reinterpret_cast<const float4*>(E + ((size_t)grow - row_base) * DIM) + piece)
                               : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int r = 0; r < 32 / RPI; ++r)
      *reinterpret_cast<float4*>(&tile[(r * RPI + sub) * PITCH + piece * 4]) = xv[r];
    __syncwarp();
    if (lane < n) {
      const float4* x = reinterpret_cast<const float4*>(&tile[lane * PITCH]);
      float acc = 0.f;
#pragma unroll
      for (int d4 = 0; d4 < F4; ++d4) {
        const float4 v = x[d4];
        acc = __fmaf_rn(v.x, qv[4 * d4], acc);
        acc = __fmaf_rn(v.y, qv[4 * d4 + 1], acc);
        acc = __fmaf_rn(v.z, qv[4 * d4 + 2], acc);
        acc = __fmaf_rn(v.w, qv[4 * d4 + 3], acc);
      }
      dst[slot] = make_key(acc, my_row);
    }
  }
}

// GROUP mode of the tensor-core filter (ScanParams::grp_rows, passes of more than 64 queries): the filter recorded a
// surviving row once per group of 16 queries.  One warp takes the list of one (group, filter CTA): 32 rows at a time are
// copied coalesced into a shared-memory tile (cp.async, no staging registers), the group's 16 queries sit beside it, and
// the 32 x 16 exact scores of the chunk are computed as a 4 x 4 REGISTER TILE per lane — lane (r8, q4) owns rows
// {r8 + 8 i} x queries {q4 + 4 j}: per four dims it reads four row pieces and four query pieces (eight conflict-free
// 16-byte shared-memory loads) for 64 FMAs.  Each (row, query) score is still one fmaf chain of the arithmetic contract
// (dims ascending, one accumulator).  The keys that reach the query's threshold are appended to the (query, CTA) segment
// the refine step reads (ballot + prefix count among the lanes that share the query: the warp owns those 16 segments).
// Keys below tau are dropped here: every row whose exact key reaches tau is in the list (filter guarantee), and the
// refine step's check needs nothing else.  A list that overflowed its capacity poisons the 16 segment counts
// (> seg_cap), which the refine step reports as an overflow of those queries (dense redo).
// (First form: one row per lane against 16 broadcast query pieces — 17 shared-memory loads per 64 FMAs and 128 staging
// registers at dim 128: 0.46 ms for the ~1 M records of a c5 shard step, ncu i8h, against ~0.08 ms of FMA issue.)
__device__ __forceinline__ void cp_async16(void* dst, const void* src, uint32_t src_bytes) {   // src_bytes 0: zero fill
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(dst)), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}

template <int DIM>
__global__ void __launch_bounds__(32) rescore_group_kernel(const uint32_t* __restrict__ grp_rows, const uint32_t* __restrict__ grp_cnt,
                                                           uint32_t n_seg, uint32_t grp_cap, uint32_t seg_cap, int nq,
                                                           const float* __restrict__ E, const float* __restrict__ Q,
                                                           uint64_t row_base, const uint64_t* __restrict__ tau,
                                                           uint64_t* __restrict__ seg_keys, uint32_t* __restrict__ seg_cnt) {
  pdl_wait();                 // chained launch: the predecessor's writes are visible from here on
  pdl_launch_dependents();
  constexpr int F4 = DIM / 4;            // 16-B pieces per row
  constexpr int RPI = 32 / F4;           // rows copied per instruction (2 at dim 64, 1 at dim 128)
  constexpr int NLD = 32 / RPI;          // copy instructions per 32 rows
  constexpr int PITCH = DIM + 4;         // floats; rows r, r + 1 are 4 banks apart: 8 rows x 16 B per load are conflict free
  __shared__ __align__(16) float tile[32 * PITCH];
  __shared__ __align__(16) float qs[kGrpQ * PITCH];
  const uint32_t seg = blockIdx.x, gi = blockIdx.y, lane = threadIdx.x;
  const int q0 = (int)gi * kGrpQ;
  const uint32_t total = grp_cnt[(size_t)gi * n_seg + seg];
  if (total > grp_cap || total == 0) {   // lane j < 16: query q0 + j
    if (lane < (uint32_t)kGrpQ && q0 + (int)lane < nq) seg_cnt[(size_t)(q0 + (int)lane) * n_seg + seg] = total ? 0xFFFFFFFFu : 0u;
    return;
  }
  const uint32_t* src = grp_rows + ((size_t)gi * n_seg + seg) * grp_cap;
  uint32_t my_row = lane < total ? src[lane] : 0u;
  const uint32_t r8 = lane >> 2, q4 = lane & 3u;
  // thresholds of this lane's four queries (a query slot past nq gets the impossible threshold)
  uint64_t tq[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) tq[j] = (q0 + (int)q4 + 4 * j < nq) ? tau[q0 + (int)q4 + 4 * j] : ~0ull;
  const uint32_t sub = lane / F4, piece = lane % F4;
#pragma unroll
  for (int u = 0; u < kGrpQ / RPI; ++u) {   // the group's queries (zero rows past nq)
    const uint32_t j = u * RPI + sub;
    const bool in = q0 + (int)j < nq;
    cp_async16(&qs[j * PITCH + piece * 4], Q + (in ? (size_t)(q0 + (int)j) * DIM + piece * 4 : 0), in ? 16u : 0u);
  }
  uint32_t cnt[4] = {0u, 0u, 0u, 0u};      // keys of query q0 + q4 + 4 j so far (the same in the eight lanes that share q4)
  const uint32_t same_q = 0x11111111u << q4, below = (1u << lane) - 1u;
  for (uint32_t base = 0; base < total; base += 32) {
    const uint32_t n = total - base < 32u ? total - base : 32u;
    __syncwarp();                                               // the previous chunk's tile has been consumed
#pragma unroll
    for (int l = 0; l < NLD; ++l) {
      const uint32_t row_in_chunk = l * RPI + sub;
      const uint32_t grow = __shfl_sync(0xffffffffu, my_row, row_in_chunk);
      const bool in = row_in_chunk < n;
      cp_async16(&tile[row_in_chunk * PITCH + piece * 4], E + (in ? ((size_t)grow - row_base) * DIM + piece * 4 : 0), in ? 16u : 0u);
    }
    const uint32_t this_row = my_row;
    if (base + 32 < total) my_row = base + 32 + lane < total ? src[base + 32 + lane] : 0u;   // the next chunk's ids meanwhile
    cp_async_wait_all();
    __syncwarp();
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
#pragma unroll 2
    for (int d4 = 0; d4 < F4; ++d4) {
      float4 x[4], w[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) x[i] = *reinterpret_cast<const float4*>(&tile[(r8 + 8 * i) * PITCH + 4 * d4]);
#pragma unroll
      for (int j = 0; j < 4; ++j) w[j] = *reinterpret_cast<const float4*>(&qs[(q4 + 4 * j) * PITCH + 4 * d4]);
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          acc[i][j] = __fmaf_rn(x[i].x, w[j].x, acc[i][j]);
          acc[i][j] = __fmaf_rn(x[i].y, w[j].y, acc[i][j]);
          acc[i][j] = __fmaf_rn(x[i].z, w[j].z, acc[i][j]);
          acc[i][j] = __fmaf_rn(x[i].w, w[j].w, acc[i][j]);
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const uint32_t row_in_chunk = r8 + 8 * i;
      const uint32_t grow = __shfl_sync(0xffffffffu, this_row, row_in_chunk);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint64_t key = make_key(acc[i][j], grow);
        const bool pass = row_in_chunk < n && key >= tq[j];
        const uint32_t b = __ballot_sync(0xffffffffu, pass) & same_q;   // the lanes that share this lane's query
        if (pass) {
          const uint32_t pos = cnt[j] + __popc(b & below);
          if (pos < seg_cap) seg_keys[((size_t)(q0 + (int)q4 + 4 * j) * n_seg + seg) * seg_cap + pos] = key;
        }
        cnt[j] += __popc(b);
      }
    }
  }
  if (r8 == 0) {
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (q0 + (int)q4 + 4 * j < nq) seg_cnt[(size_t)(q0 + (int)q4 + 4 * j) * n_seg + seg] = cnt[j];
  }
}

// Select step of the tensor-core filter: one CTA per query packs the query's exact (re-scored) survivor keys into
// SHARED memory, selects and sorts the top k there.  Replaces select_kernel<TOPK> on that path (ncu r5: 35 us per 64
// queries: a tid-0 serial walk over the 148 segment counts, ten passes over the keys through L2, and an 8-pass radix
// select whose first passes all hit one histogram bin).
//   1. exclusive scan of the (clamped) per-CTA segment lengths -> m survivors, m <= cap (else overflow, flag 1)
//   2. survivor i: segment by binary search, key from the re-scored segment
//   3. radix select that starts at the highest bit in which the keys differ and stops as soon as the keys above
//      the current bucket plus the bucket fit the sort buffer (typically one or two 8-bit passes for ~5 k keys)
//   4. bitonic sort of that buffer, first k out (keys are distinct: no tie handling left)
struct RefineParams {
  const uint64_t* seg_keys;    // [nq][n_seg][seg_cap] exact keys (rescore_kernel)
  const uint32_t* seg_counts;  // [nq][n_seg]
  uint32_t n_seg, seg_cap, cap;  // cap = keys the shared buffer holds
  const uint64_t* tau_check;   // the k-th exact key must reach the sampled threshold (else flag 2)
  int k, k_out;
  uint32_t expect;
  uint32_t sort_cap;           // power of two >= 2 * pow2ceil(k)
  uint64_t* out_keys;          // [nq][k_out]
  int32_t* flags;
  uint32_t* max_count;
};

// NT threads per CTA: 1024 for the single-GPU shape (64 queries x ~4.6 k survivors, one CTA per SM), 256 for row shards
// (G times the queries, 1/G of the survivors each: several CTAs per SM instead of 3.5 waves of one).
template <int NT, bool FAST = false>   // FAST: the final sort by bitonic.cuh (NT == 1024)
__global__ void __launch_bounds__(NT, NT == 1024 ? 1 : 2) refine_select_kernel(const RefineParams p) {
  pdl_wait();                 // chained launch: the predecessor's writes are visible from here on
  pdl_launch_dependents();
  extern __shared__ uint64_t rs_keys[];          // [cap] exact keys, then [sort_cap] sort buffer
  __shared__ uint32_t seg_off[520];
  __shared__ uint32_t hist[256];
  __shared__ unsigned long long s_or;
  __shared__ uint64_t s_prefix;
  __shared__ uint32_t s_want, s_above, s_bucket, s_fill, s_ov;
  const int q = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  uint64_t* keys = rs_keys;
  uint64_t* sk = rs_keys + p.cap;

  // ---- 1. segment offsets
  if (tid == 0) { s_ov = 0; s_or = 0ull; s_fill = 0; }
  __syncthreads();
  for (uint32_t sgi = tid; sgi < p.n_seg; sgi += NT) {
    uint32_t c = p.seg_counts[(size_t)q * p.n_seg + sgi];
    if (c > p.seg_cap) { c = p.seg_cap; s_ov = 1; }
    seg_off[sgi + 1] = c;
  }
  __syncthreads();
  if (warp == 0) {  // inclusive scan of seg_off[1..n_seg] (n_seg <= 512): 16 entries per lane
    const uint32_t per = (p.n_seg + 31) / 32;
    uint32_t sum = 0;
    for (uint32_t j = 0; j < per; ++j) { const uint32_t i = lane * per + j; if (i < p.n_seg) sum += seg_off[i + 1]; }
    uint32_t run = sum;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const uint32_t v = __shfl_up_sync(0xffffffffu, run, off);
      if (lane >= off) run += v;
    }
    uint32_t base = run - sum;
    for (uint32_t j = 0; j < per; ++j) {
      const uint32_t i = lane * per + j;
      if (i < p.n_seg) { const uint32_t c = seg_off[i + 1]; seg_off[i + 1] = base + c; base += c; }
    }
    if (lane == 0) seg_off[0] = 0;
  }
  __syncthreads();
  uint32_t m = seg_off[p.n_seg];
  bool overflow = s_ov != 0u;
  if (m > p.cap) { m = p.cap; overflow = true; }
  if (p.max_count && tid == 0) atomicMax(p.max_count, m);

  // ---- 2. pack the exact keys into shared memory
  unsigned long long my_or = 0ull;
  uint64_t key0 = 0ull;
  for (uint32_t i = tid; i < m; i += NT) {
    uint32_t lo = 0, hi = p.n_seg;       // last segment with seg_off[s] <= i
    while (hi - lo > 1) { const uint32_t mid = (lo + hi) >> 1; if (seg_off[mid] <= i) lo = mid; else hi = mid; }
    const uint64_t key = p.seg_keys[((size_t)q * p.n_seg + lo) * p.seg_cap + (i - seg_off[lo])];
    keys[i] = key;
    if (i == tid) key0 = key;
    my_or |= key ^ key0;
  }
  __syncthreads();
  // bits in which any two keys differ: OR over (key ^ keys[0])
  if (tid < m) my_or |= key0 ^ keys[0];
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) my_or |= __shfl_xor_sync(0xffffffffu, my_or, off);
  if (lane == 0 && my_or) atomicOr(&s_or, my_or);
  __syncthreads();

  // ---- 3. radix select from the highest differing bit; stop when (keys above the bucket) + bucket fit sort_cap
  const uint32_t k = (uint32_t)p.k;
  uint64_t lo_key = 1ull;  // keys >= lo_key go to the sort buffer
  if (m > p.sort_cap) {
    const int top0 = 64 - __clzll((long long)s_or);   // bits [top0, 64) are common to all keys (s_or != 0: m > 1 distinct)
    int top = top0;
    uint64_t prefix = 0ull, mask = 0ull;
    uint32_t want = k, above_total = 0;
    while (true) {
      const int shift = top > 8 ? top - 8 : 0;
      const uint32_t dmask = (1u << (top - shift)) - 1u;
      if (tid < 256) hist[tid] = 0;
      __syncthreads();
      for (uint32_t i = tid; i < m; i += NT) {
        const uint64_t key = keys[i];
        const bool in = (key & mask) == prefix;
        const uint32_t dg = (uint32_t)(key >> shift) & dmask;
        // warp-aggregated histogram: one atomic per distinct digit and warp
        const unsigned act = __ballot_sync(__activemask(), in);
        if (in) {
          const unsigned same = __match_any_sync(act, dg);
          if (lane == __ffs(same) - 1) atomicAdd(&hist[dg], __popc(same));
        }
      }
      __syncthreads();
      if (warp == 0) {  // lane l owns digits [8l, 8l+8); suffix sums from the top digit down
        uint32_t c[8], tot = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) { c[j] = hist[lane * 8 + j]; tot += c[j]; }
        uint32_t run = tot;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
          const uint32_t v = __shfl_down_sync(0xffffffffu, run, off);
          if (lane + off < 32) run += v;
        }
        const uint32_t above = run - tot;
        if (above < want && above + tot >= want) {
          uint32_t a = above;
          int d = 7;
          for (; d >= 0; --d) {
            if (a + c[d] >= want) break;
            a += c[d];
          }
          s_prefix = prefix | ((uint64_t)(lane * 8 + d) << shift);
          s_want = want - a;
          s_above = a;
          s_bucket = c[d];
        }
      }
      __syncthreads();
      prefix = s_prefix;
      want = s_want;
      above_total += s_above;
      mask |= ((uint64_t)dmask << shift);
      const uint32_t bucket = s_bucket;
      top = shift;
      __syncthreads();
      if (above_total + bucket <= p.sort_cap || shift == 0) break;
    }
    // every key shares the bits at and above top0, so "in the bucket or above it" is a plain comparison
    lo_key = (top0 >= 64 ? 0ull : ((keys[0] >> top0) << top0)) | prefix;
  }
  for (uint32_t i = tid; i < p.sort_cap; i += NT) sk[i] = 0ull;
  __syncthreads();
  for (uint32_t i = tid; i < m; i += NT) {
    const uint64_t key = keys[i];
    if (key >= lo_key) {
      const uint32_t pos = atomicAdd(&s_fill, 1u);
      if (pos < p.sort_cap) sk[pos] = key;
    }
  }
  __syncthreads();
  const uint32_t filled = s_fill < p.sort_cap ? s_fill : p.sort_cap;
  uint32_t K2 = 32;
  while (K2 < filled) K2 <<= 1;
  // ---- 4. bitonic sort, descending
  bool sorted = false;
  if constexpr (FAST && NT == 1024) {
    if (K2 <= 1024u && p.cap >= 1024u) {   // one key per thread; the packed-key region is free and serves as exchange buffer
      uint64_t e = (uint32_t)tid < p.sort_cap ? sk[tid] : 0ull;
      e = bitonic_sort_1024(e, keys, [](uint64_t a, uint64_t b) { return a > b; });
      if ((uint32_t)tid < p.sort_cap) sk[tid] = e;
      __syncthreads();
      sorted = true;
    }
  }
  if (!sorted) {
  for (uint32_t size = 2; size <= K2; size <<= 1) {
    for (uint32_t stride = size >> 1; stride > 0; stride >>= 1) {
      for (uint32_t i = tid; i < (K2 >> 1); i += NT) {
        const uint32_t pos = 2 * i - (i & (stride - 1));
        const uint64_t a = sk[pos], b = sk[pos + stride];
        const bool desc = (pos & size) == 0;
        if ((a < b) == desc) { sk[pos] = b; sk[pos + stride] = a; }
      }
      __syncthreads();
    }
  }
  }
  const uint32_t n_out = m < k ? m : k;
  for (uint32_t i = tid; i < (uint32_t)p.k_out; i += NT)
    p.out_keys[(size_t)q * p.k_out + i] = (i < n_out) ? sk[i] : 0ull;
  if (tid == 0) {
    int fl = overflow ? 1 : (n_out < p.expect ? 2 : 0);
    if (!fl && p.tau_check && n_out > 0 && sk[n_out - 1] < p.tau_check[q]) fl = 2;  // a row outside the list could win
    if (p.flags) p.flags[q] = fl;
  }
}

// keys -> rows / scores / counts (used by the shard-merge path where keys already are sorted): one warp per 32 keys,
// the per-request count is a ballot + one atomic per warp (out_n is zeroed by the launcher)
__global__ void keys_unpack_kernel(const uint64_t* keys, int total, int k, uint32_t* out_row, float* out_score,
                                   int32_t* out_n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const bool in = i < total;
  const uint64_t key = in ? keys[i] : 0ull;
  if (in) {
    out_row[i] = key ? key_row(key) : 0xFFFFFFFFu;
    out_score[i] = key ? key_score(key) : __int_as_float(0xFF800000);
  }
  // lanes of a warp may straddle requests (k is not a multiple of 32 in general): group the lanes by request
  const int b = in ? i / k : -1;
  const unsigned same = __match_any_sync(0xffffffffu, b);
  const unsigned nz = __ballot_sync(0xffffffffu, key != 0ull);
  const int lane = threadIdx.x & 31;
  if (in && lane == __ffs(same) - 1 && (same & nz)) atomicAdd(&out_n[b], __popc(same & nz));
}

// ------------------------------------------------------------------ host side
int recall_build_map(prg_handle* h) {
  h->E_map_ok = false;
  PFN_encodeTiled enc = get_encode_tiled();
  if (!enc) return fail(PRG_ECUDA, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t gdim[2] = {(cuuint64_t)h->E_dim, (cuuint64_t)h->E_rows};
  cuuint64_t gstride[1] = {(cuuint64_t)h->E_dim * 4};
  cuuint32_t box[2] = {32, (cuuint32_t)kTileRows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(&h->E_map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(h->E), gdim, gstride, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(PRG_ECUDA, "cuTensorMapEncodeTiled failed: " + std::to_string((int)r));
  h->E_map_ok = true;
  return PRG_OK;
}

template <int DIM, int MODE>
static int launch_scan(prg_handle* h, const ScanParams& p) {
  const size_t smem = scan_smem_bytes<DIM>();
  PRG_CUDA(cudaFuncSetAttribute(recall_scan_kernel<DIM, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  if (p.n_tiles == 0) return PRG_OK;
  StageScope span(h, MODE == SCAN_THRESH ? ST_SCAN : ST_SCAN_DENSE);
  const unsigned gx = p.n_tiles < (uint32_t)h->sm_count ? p.n_tiles : (unsigned)h->sm_count;
  const dim3 grid(gx, (MODE == SCAN_DENSE && p.q_blocks > 1) ? (unsigned)p.q_blocks : 1u);
  recall_scan_kernel<DIM, MODE><<<grid, kScanThreads, smem, h->stream>>>(h->E_map, p);
  PRG_CUDA(cudaGetLastError());
  count_launch(h);
  return PRG_OK;
}

static int scan(prg_handle* h, int mode, const ScanParams& p) {
  if (h->E_dim == 64) return mode == SCAN_THRESH ? launch_scan<64, SCAN_THRESH>(h, p) : launch_scan<64, SCAN_DENSE>(h, p);
  if (h->E_dim == 128)
    return mode == SCAN_THRESH ? launch_scan<128, SCAN_THRESH>(h, p) : launch_scan<128, SCAN_DENSE>(h, p);
  return fail(PRG_EUNSUPPORTED, "item matrix dim must be 64 or 128");
}

// The r best keys of a sample (r ~ 32 of 10 k - 80 k keys), sorted — or just the r-th — in two passes over the keys:
//   1. every thread takes the maximum of its strided share; T0 = the r-th largest of the NT maxima (rank counting in
//      shared memory).  The maxima are distinct keys, so at least r keys reach T0 and the r-th largest key is >= T0;
//   2. the keys >= T0 (r plus a handful when the large keys are spread over the threads) are collected in shared
//      memory, sorted, and the first r written out.
// If more than kTopRCap keys reach T0 (the large keys sit in a few threads' shares) the query is left to the generic
// select (done[q] = 0).  The generic kernels cost 29 us (KTH, 64 x 78 k keys) and 73 us (TOPK, 512 x 9.7 k keys).
constexpr uint32_t kTopRCap = 2048;
struct TopRParams {
  const uint64_t* keys;   // [q][stride]
  uint64_t stride;
  uint32_t m, r;
  uint64_t* out_keys;     // nullable: [q][r] sorted descending, 0-padded
  uint64_t* tau;          // nullable: [q] the r-th largest (0 if fewer than r valid keys)
  int32_t* done;          // [q]
};
template <int NT>
__global__ void __launch_bounds__(NT) sample_topr_kernel(const TopRParams p) {
  pdl_wait();                 // chained launch: the predecessor's writes are visible from here on
  pdl_launch_dependents();
  __shared__ uint64_t mx[256];
  __shared__ uint64_t cand[kTopRCap];
  __shared__ uint64_t s_t0;
  __shared__ uint32_t s_cnt;
  const int q = blockIdx.x, tid = threadIdx.x;
  const uint64_t* keys = p.keys + (size_t)q * p.stride;
  const ulonglong2* keys2 = reinterpret_cast<const ulonglong2*>(keys);   // rows of the sample are 2 KiB multiples
  const uint32_t m2 = p.m >> 1;
  if (tid == 0) { s_cnt = 0; s_t0 = 0ull; }
  uint64_t best = 0ull;
  constexpr int U = 8;   // independent 16-B loads in flight per thread (64 CTAs x 78 k keys: latency-, not bandwidth-bound)
  {
    uint32_t i = tid;
    for (; i + (U - 1) * NT < m2; i += U * NT) {
      ulonglong2 v[U];
#pragma unroll
      for (int u = 0; u < U; ++u) v[u] = keys2[i + u * NT];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const uint64_t a = v[u].x > v[u].y ? v[u].x : v[u].y;
        best = a > best ? a : best;
      }
    }
    for (; i < m2; i += NT) {
      const ulonglong2 v = keys2[i];
      const uint64_t a = v.x > v.y ? v.x : v.y;
      best = a > best ? a : best;
    }
  }
  if ((p.m & 1u) && tid == 0) { const uint64_t a = keys[p.m - 1]; best = a > best ? a : best; }
  // 256 group maxima (groups of NT/256 adjacent threads): ranking NT values against each other is O(NT^2) — with 1024
  // threads almost every warp holds one of the larger maxima and walks all 1024 (33 us)
  constexpr int GS = NT / 256;
  uint64_t gbest = best;
#pragma unroll
  for (int off = 1; off < GS; off <<= 1) {
    const uint64_t o = __shfl_xor_sync(0xffffffffu, gbest, off);
    gbest = o > gbest ? o : gbest;
  }
  if ((tid & (GS - 1)) == 0) mx[tid / GS] = gbest;
  __syncthreads();
  if (tid < 256) {
    const uint64_t mine = mx[tid];
    uint32_t rank = 0;   // only a maximum that can be among the r largest needs its exact rank: stop counting at r
    for (int u0 = 0; u0 < 256 && rank < p.r; u0 += 8) {
#pragma unroll
      for (int u = u0; u < u0 + 8; ++u) {
        const uint64_t v = mx[u];
        rank += (v > mine || (v == mine && u < tid)) ? 1u : 0u;
      }
    }
    if (rank == p.r - 1) s_t0 = mine;
  }
  __syncthreads();
  const uint64_t lim = s_t0 ? s_t0 : 1ull;
  auto take = [&](uint64_t key) {
    if (key >= lim) { const uint32_t pos = atomicAdd(&s_cnt, 1u); if (pos < kTopRCap) cand[pos] = key; }
  };
  {
    uint32_t i = tid;
    for (; i + (U - 1) * NT < m2; i += U * NT) {
      ulonglong2 v[U];
#pragma unroll
      for (int u = 0; u < U; ++u) v[u] = keys2[i + u * NT];
#pragma unroll
      for (int u = 0; u < U; ++u) { take(v[u].x); take(v[u].y); }
    }
    for (; i < m2; i += NT) { const ulonglong2 v = keys2[i]; take(v.x); take(v.y); }
  }
  if ((p.m & 1u) && tid == 0) {
    const uint64_t a = keys[p.m - 1];
    if (a >= lim) { const uint32_t pos = atomicAdd(&s_cnt, 1u); if (pos < kTopRCap) cand[pos] = a; }
  }
  __syncthreads();
  const uint32_t c = s_cnt;
  if (c > kTopRCap) { if (tid == 0) p.done[q] = 0; return; }
  uint32_t K2 = 32;
  while (K2 < c) K2 <<= 1;
  for (uint32_t i = c + tid; i < K2; i += NT) cand[i] = 0ull;
  __syncthreads();
  for (uint32_t size = 2; size <= K2; size <<= 1) {
    for (uint32_t stride = size >> 1; stride > 0; stride >>= 1) {
      for (uint32_t i = tid; i < (K2 >> 1); i += NT) {
        const uint32_t pos = 2 * i - (i & (stride - 1));
        const uint64_t a = cand[pos], b = cand[pos + stride];
        if ((a < b) == ((pos & size) == 0)) { cand[pos] = b; cand[pos + stride] = a; }
      }
      __syncthreads();
    }
  }
  if (p.out_keys)
    for (uint32_t i = tid; i < p.r; i += NT) p.out_keys[(size_t)q * p.r + i] = i < c ? cand[i] : 0ull;
  if (tid == 0) {
    if (p.tau) p.tau[q] = c >= p.r ? cand[p.r - 1] : 0ull;
    p.done[q] = 1;
  }
}

// Exact re-score of the filter survivors of nq queries.  `per_seg` = expected survivors per (segment, query): a warp takes
// as many consecutive segments as give it about two chunks of 32 rows.
static int launch_rescore(prg_handle* h, const float* q_dev, int nq, uint32_t n_seg, uint32_t seg_cap, double per_seg) {
  StageScope span(h, ST_SELECT);
  uint32_t group = 1;
  while (group < 32 && per_seg * group < 24.0) group <<= 1;
  const dim3 grid((n_seg + group - 1) / group, (unsigned)nq);
  if (h->E_dim == 64)
    PRG_CUDA(launch_chained(h, rescore_kernel<64>, grid, dim3(32), 0, 1, (const uint32_t*)h->seg_rows.p,
                            (const uint32_t*)h->cand_cnt.p, n_seg, seg_cap, group, h->E, q_dev, h->E_row_base,
                            (uint64_t*)h->seg_keys.p));
  else
    PRG_CUDA(launch_chained(h, rescore_kernel<128>, grid, dim3(32), 0, 1, (const uint32_t*)h->seg_rows.p,
                            (const uint32_t*)h->cand_cnt.p, n_seg, seg_cap, group, h->E, q_dev, h->E_row_base,
                            (uint64_t*)h->seg_keys.p));
  PRG_CUDA(cudaGetLastError());
  count_launch(h);
  return PRG_OK;
}

// config "scan_groups": 1 / 0 force GROUP mode on / off; default (-1) = on at dim 64, where the epilogue of a 256-query
// pass is the limit (c4 shard pass 0.320 -> 0.298 ms), off at dim 128, where the MMAs and the loads are and the two forms
// measure within the run-to-run spread of each other (c5 shard pass 2.94 / 3.11 ms in gpurun r3g, 3.18 / 3.02 ms in r3i;
// the re-score of a group costs 16 dot products per recorded row instead of one).
static bool scan_groups_on(const prg_handle* h) {
  if (h->scan_filter != SCAN_FILTER_BF16) return false;   // the group form of the filter exists over the bf16 index only
  // (dim 128 with the int8 index: its pass of more than 64 queries exists in GROUP form only, recall_i8.cu)
  return h->scan_groups < 0 ? (h->E_dim == 64 || scan_i8g_available(h)) : h->scan_groups != 0;
}

// GROUP mode: exact scores of the recorded (row, group of 16 queries) pairs -> the (query, segment) key lists + counts
static int launch_rescore_groups(prg_handle* h, const float* q_dev, int nq, uint32_t n_seg, uint32_t seg_cap) {
  StageScope span(h, ST_SELECT);
  const dim3 grid(n_seg, (unsigned)((nq + kGrpQ - 1) / kGrpQ));
  const uint32_t grp_cap = kGrpQ * seg_cap;
  if (h->E_dim == 64)
    PRG_CUDA(launch_chained(h, rescore_group_kernel<64>, grid, dim3(32), 0, 1, (const uint32_t*)h->seg_rows.p,
                            (const uint32_t*)h->grp_cnt.p, n_seg, grp_cap, seg_cap, nq, h->E, q_dev, h->E_row_base,
                            (const uint64_t*)h->tau.p, (uint64_t*)h->seg_keys.p, (uint32_t*)h->cand_cnt.p));
  else
    PRG_CUDA(launch_chained(h, rescore_group_kernel<128>, grid, dim3(32), 0, 1, (const uint32_t*)h->seg_rows.p,
                            (const uint32_t*)h->grp_cnt.p, n_seg, grp_cap, seg_cap, nq, h->E, q_dev, h->E_row_base,
                            (const uint64_t*)h->tau.p, (uint64_t*)h->seg_keys.p, (uint32_t*)h->cand_cnt.p));
  PRG_CUDA(cudaGetLastError());
  count_launch(h);
  return PRG_OK;
}
// the filter pass [q0, q0 + nq) of a GROUP-mode call: where its groups' lists and lengths live (h->seg_rows is shared with
// the per-query mode: 16 queries x seg_cap rows per (group, segment) = the same bytes)
static void scan_group_outputs(prg_handle* h, ScanParams& sc, int q0, uint32_t n_seg, uint32_t seg_cap) {
  const size_t g0 = (size_t)q0 / kGrpQ;
  sc.grp_cap = kGrpQ * seg_cap;
  sc.grp_rows = (uint32_t*)h->seg_rows.p + g0 * n_seg * sc.grp_cap;
  sc.grp_cnt = (uint32_t*)h->grp_cnt.p + g0 * n_seg;
}

// Top-k of the exact keys, one CTA per query: 1024 threads when every query gets an SM of its own, 256 (several CTAs
// per SM) when there are more queries than SMs.
static int launch_refine(prg_handle* h, const RefineParams& rp, int nq, size_t smem) {
  StageScope span(h, ST_SELECT);
  if (nq > h->sm_count && smem <= 100 * 1024) {
    PRG_CUDA(cudaFuncSetAttribute(refine_select_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    PRG_CUDA(launch_chained(h, refine_select_kernel<256>, dim3(nq), dim3(256), smem, 1, rp));
  } else {
    // FAST: the final 1024-element sort with one key per thread, intra-warp stages by shuffle (bitonic.cuh); measured
    // on the C4 batch: select stage 0.0892 -> 0.0861 ms (profiles/r02_experiments_ab.txt)
    PRG_CUDA(cudaFuncSetAttribute(refine_select_kernel<1024, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    PRG_CUDA(launch_chained(h, refine_select_kernel<1024, true>, dim3(nq), dim3(1024), smem, 1, rp));
  }
  PRG_CUDA(cudaGetLastError());
  count_launch(h);
  return PRG_OK;
}

static int launch_select(prg_handle* h, int mode, const SelectParams& p, int nq) {
  StageScope span(h, ST_SELECT);
  if (mode == SEL_KTH) {
    PRG_CUDA(launch_chained(h, select_kth_kernel, dim3(nq), dim3(1024), kKthCap * 8, 1, p));
  } else {
    uint32_t K2 = 32;
    while (K2 < (uint32_t)p.k) K2 <<= 1;
    select_kernel<SEL_TOPK><<<nq, 1024, K2 * 8, h->stream>>>(p);
  }
  PRG_CUDA(cudaGetLastError());
  count_launch(h);
  return PRG_OK;
}

// Threshold / best-r selection over the sample: the two-pass kernel first, the generic select only for the queries it
// declined (none on well-mixed data).  mode SEL_KTH writes st.tau, SEL_TOPK st.out_keys (k_out == k == r).
static int launch_sample_select(prg_handle* h, int mode, SelectParams st, int nq) {
  const uint32_t r = (uint32_t)st.k;
  if (r == 0 || r > 256 || (st.stride & 1ull) || (reinterpret_cast<uintptr_t>(st.keys) & 15))
    return launch_select(h, mode, st, nq);
  PRG_TRY(h->topr_done.ensure((size_t)nq * 4));
  {
    StageScope span(h, ST_SELECT);
    TopRParams tp{};
    tp.keys = st.keys; tp.stride = st.stride; tp.m = st.fixed_m; tp.r = r;
    tp.out_keys = mode == SEL_TOPK ? st.out_keys : nullptr;
    tp.tau = mode == SEL_KTH ? st.tau : nullptr;
    tp.done = (int32_t*)h->topr_done.p;
    static const int force_nt = getenv("PRG_TOPR_NT") ? atoi(getenv("PRG_TOPR_NT")) : 0;   // experiments only
    const int nt = force_nt ? force_nt : (((uint64_t)st.fixed_m >= 32768 && nq <= 2 * h->sm_count) ? 1024 : 256);
    if (nt == 1024) PRG_CUDA(launch_chained(h, sample_topr_kernel<1024>, dim3(nq), dim3(1024), 0, 1, tp));
    else if (nt == 512) PRG_CUDA(launch_chained(h, sample_topr_kernel<512>, dim3(nq), dim3(512), 0, 1, tp));
    else PRG_CUDA(launch_chained(h, sample_topr_kernel<256>, dim3(nq), dim3(256), 0, 1, tp));
    PRG_CUDA(cudaGetLastError());
    count_launch(h);
  }
  st.skip = (const int32_t*)h->topr_done.p;
  return launch_select(h, mode, st, nq);
}

// Threshold from tile maxima (config "recall_tilemax", default on).  The sample scan leaves ONE value per (sample tile,
// query): the largest approximate score of the tile's 256 rows (recall_tc.cu, SCAN_TILEMAX).  tau = the r-th largest of a
// query's T tile maxima: at least r sampled rows reach it, and the share of tiles whose maximum reaches it, r / T,
// estimates the share of ROWS that do: -ln(1 - r/T) / 256 — the same statistic the r-th largest of all sample keys
// gives, from T values per query instead of 256 T keys (C4: 40 MB of keys written and read twice per batch).  tau is a
// pruning hint only; the refine step's check keeps the result exact for any tau.
// One CTA per query: rank counting in shared memory (T <= 2048).
// out_keys (nullable): the r largest as order keys [q][r], sorted descending, 0-padded — what a shard contributes to the
// global threshold (recall_shard_sample_device).
__global__ void __launch_bounds__(256) tilemax_tau_kernel(const uint32_t* __restrict__ tile_max, uint64_t stride, uint32_t T,
                                                          uint32_t r, uint64_t* __restrict__ tau,
                                                          uint64_t* __restrict__ out_keys) {
  pdl_wait();
  pdl_launch_dependents();
  __shared__ uint32_t v[2048];
  __shared__ uint32_t s_pick;
  const int q = blockIdx.x, tid = threadIdx.x;
  const uint32_t* src = tile_max + (size_t)q * stride;
  for (uint32_t i = tid; i < T; i += 256) v[i] = src[i];
  if (tid == 0) s_pick = 0u;
  if (out_keys)
    for (uint32_t j = tid; j < r; j += 256) out_keys[(size_t)q * r + j] = 0ull;
  __syncthreads();
  for (uint32_t i = tid; i < T; i += 256) {
    const uint32_t mine = v[i];
    uint32_t rank = 0;   // values that come before `mine` in (value desc, index asc) order
    // no early exit: the loop is a stream of independent broadcast reads (8 in flight), not a chain of dependent ones
    // (ncu launch list, r2d: 29 us for T = 305 with the data-dependent exit)
    uint32_t u = 0;
    for (; u + 8 <= T; u += 8) {
#pragma unroll
      for (uint32_t t = 0; t < 8; ++t) {
        const uint32_t o = v[u + t];
        rank += (o > mine || (o == mine && u + t < i)) ? 1u : 0u;
      }
    }
    for (; u < T; ++u) {
      const uint32_t o = v[u];
      rank += (o > mine || (o == mine && u < i)) ? 1u : 0u;
    }
    // ordered score in the key's high word, row bits zero: a row whose exact score equals the threshold still reaches it
    if (rank < r && out_keys) out_keys[(size_t)q * r + rank] = mine > 1u ? ((uint64_t)mine << 32) : 0ull;
    if (rank == r - 1) s_pick = mine;   // exactly one element has rank r - 1 (if T >= r)
  }
  __syncthreads();
  if (tid == 0 && tau) tau[q] = (T >= r && s_pick > 1u) ? ((uint64_t)s_pick << 32) : 0ull;
}

// Scores of the strided tile sample for B queries -> h->sample_keys [B][slots].  With the bf16 filter index the
// tensor-core kernel scores it (approximate keys: tau is only a pruning hint); otherwise the exact FFMA2 kernel, one
// launch with one grid row per block of 64 queries.
static int score_sample(prg_handle* h, const float* q_dev, int B, uint32_t sample_tiles, uint32_t tile_stride,
                        uint64_t slots, bool use_tc) {
  ScanParams sp{};
  sp.n_rows = h->E_rows; sp.row_base = h->E_row_base;
  sp.n_tiles = sample_tiles; sp.tile_stride = tile_stride;
  sp.dense_stride = slots;
  if (use_tc && scan_tc_dense_available(h)) {
    sp.row_norm = (const float*)h->row_norm.p;   // the producer copies the tile's norms although DENSE does not use them
    sp.Q = q_dev; sp.nq = B;
    sp.dense = (uint64_t*)h->sample_keys.p;
    return launch_scan_tc_dense(h, sp);          // one launch: grid = sample tiles x blocks of 64 queries
  }
  sp.Q = q_dev; sp.nq = B; sp.q_blocks = (B + kQB - 1) / kQB;
  sp.dense = (uint64_t*)h->sample_keys.p;
  return scan(h, SCAN_DENSE, sp);
}

// Dense path for query block [q0, q0+nq): all keys materialised, then the same select.
static int recall_dense(prg_handle* h, const float* q_dev, int nq, int k, int k_out, uint64_t* keys_out) {
  const uint32_t n_tiles = (uint32_t)((h->E_rows + kTileRows - 1) / kTileRows);
  const uint64_t slots = (uint64_t)n_tiles * kTileRows;
  PRG_TRY(h->dense_keys.ensure((size_t)nq * slots * 8));
  ScanParams sp{};
  sp.Q = q_dev; sp.nq = nq; sp.n_rows = h->E_rows; sp.row_base = h->E_row_base;
  sp.n_tiles = n_tiles; sp.tile_stride = 1;
  sp.dense = (uint64_t*)h->dense_keys.p; sp.dense_stride = slots;
  PRG_TRY(scan(h, SCAN_DENSE, sp));
  SelectParams se{};
  se.keys = (const uint64_t*)h->dense_keys.p; se.stride = slots; se.counts = nullptr;
  se.fixed_m = (uint32_t)slots; se.cap = (uint32_t)slots; se.k = k; se.k_out = k_out;
  se.expect = 0; se.out_keys = keys_out;
  PRG_TRY(launch_select(h, SEL_TOPK, se, nq));
  return PRG_OK;
}

static size_t refine_smem_bytes(uint32_t cand_cap, int k) {
  uint32_t k_pow2 = 32;
  while (k_pow2 < (uint32_t)k) k_pow2 <<= 1;
  const uint32_t sort_cap = (k_pow2 - (uint32_t)k >= 16) ? k_pow2 : 2 * k_pow2;
  return ((size_t)cand_cap + sort_cap) * 8;
}

constexpr uint64_t kSampledMinRows = 1u << 18;  // below this the dense path is cheaper than sampling

int recall_topk_device(prg_handle* h, const float* q_dev, int B, int k, uint64_t* keys_out, bool defer) {
  if (!h->E || !h->E_map_ok) return fail(PRG_ESTATE, "item matrix not set (prg_set_item_matrix)");
  if (B <= 0 || k <= 0) return fail(PRG_EINVAL, "B and k must be positive");
  if (k > 4096) return fail(PRG_EUNSUPPORTED, "k > 4096");
  if (h->E_rows + h->E_row_base > 0xFFFFFFFFull) return fail(PRG_EUNSUPPORTED, "global row ids must fit u32");
  if (h->E_rows >= (1ull << 31)) return fail(PRG_EUNSUPPORTED, "more than 2^31 rows per shard");
  const uint32_t dim = h->E_dim;
  h->last_fallback = 0;
  h->last_max_cand = 0;
  h->last_filter = PRG_FILTER_NONE;

  const uint32_t n_tiles = (uint32_t)((h->E_rows + kTileRows - 1) / kTileRows);
  const bool sampled = h->E_rows >= kSampledMinRows && (uint64_t)k * 64 <= h->E_rows;

  // sampling plan: ~1/128 of the tiles, strided across the whole matrix
  const int nblk = (B + kQB - 1) / kQB;
  const size_t QT = (size_t)nblk * kQB;  // query slots (blocks of 64)
  h->pending.active = false;
  if (!sampled) {
    for (int q0 = 0; q0 < B; q0 += kQB) {
      const int nq = (B - q0 < kQB) ? (B - q0) : kQB;
      PRG_TRY(recall_dense(h, q_dev + (size_t)q0 * dim, nq, k, k, keys_out + (size_t)q0 * k));
    }
    return PRG_OK;
  }
  uint32_t sample_tiles = n_tiles / 128;
  if (sample_tiles < 64) sample_tiles = 64;
  if (sample_tiles > 2048) sample_tiles = 2048;
  const uint32_t tile_stride = n_tiles / sample_tiles;
  const uint64_t slots = (uint64_t)sample_tiles * kTileRows;
  const double f = (double)slots / (double)h->E_rows;
  const double target = 4.0 * (k < 1024 ? 1024 : k);
  uint32_t r_rank = (uint32_t)(target * f + 0.999);
  if (r_rank < 24) r_rank = 24;
  uint32_t cand_cap = (uint32_t)(4.0 * (double)r_rank / f);
  cand_cap = (cand_cap + 1023) & ~1023u;
  const uint32_t n_seg = n_tiles < (uint32_t)h->sm_count ? n_tiles : (uint32_t)h->sm_count;
  uint32_t seg_cap = (uint32_t)(8.0 * ((double)r_rank / f) / n_seg) + 32;
  seg_cap = (seg_cap + 15) & ~15u;
  const bool use_tc = !h->scan_ffma2;
  PRG_TRY(h->sample_keys.ensure(QT * slots * 8));
  PRG_TRY(h->cand_keys.ensure(QT * cand_cap * 8));
  PRG_TRY(h->seg_keys.ensure(QT * n_seg * seg_cap * 8));
  if (use_tc) PRG_TRY(h->seg_rows.ensure(QT * n_seg * seg_cap * 4));
  if (use_tc && !h->row_norm.p) PRG_TRY(build_row_norms(h));
  PRG_TRY(h->cand_cnt.ensure(QT * n_seg * 4 + 4));
  PRG_TRY(h->tau.ensure(QT * 8));
  PRG_TRY(h->flags.ensure(QT * 4));
  const size_t seg_q = (size_t)n_seg * seg_cap;  // segment slots per query

  // All query blocks go through each phase together: the scans run once per block of <= 64 queries, the selects
  // and the re-score once for the whole batch (one CTA per query), and there is ONE status read-back.
  // (the survivor-count maximum is cleared up front: no memset between the kernels of the chain, see launch_chained)
  uint32_t* max_cnt = (uint32_t*)h->cand_cnt.p + QT * n_seg;
  PRG_CUDA(cudaMemsetAsync(max_cnt, 0, 4, h->stream));
  // share of a tile's rows expected above the threshold, x = 256 * target / rows: the tile-maximum estimate needs x << 1
  const double tile_x = (double)kTileRows * target / (double)h->E_rows;
  // tiles of the sample expected to hold a row above the threshold; below 24 the estimate is too noisy (small catalogs)
  const double tile_r = (double)sample_tiles * (1.0 - exp(-tile_x));
  if (h->recall_tilemax && use_tc && scan_tc_dense_available(h) && tile_x <= 0.35 && tile_r >= 24.0 && sample_tiles <= 2048) {
    // 1'. one maximum per (sample tile, query);  2'. threshold = r_t-th largest tile maximum (experimental, see above)
    const uint32_t r_t = (uint32_t)(tile_r + 0.999);
    ScanParams sp{};
    sp.n_rows = h->E_rows; sp.row_base = h->E_row_base;
    sp.n_tiles = sample_tiles; sp.tile_stride = tile_stride;
    sp.row_norm = (const float*)h->row_norm.p;
    sp.Q = q_dev; sp.nq = B;
    sp.dense = (uint64_t*)h->sample_keys.p;   // used as u32 [B][sample_tiles]
    sp.dense_stride = sample_tiles;
    PRG_TRY(launch_scan_tc_tilemax(h, sp));
    {
      StageScope span(h, ST_SELECT);
      PRG_CUDA(launch_chained(h, tilemax_tau_kernel, dim3(B), dim3(256), 0, 1, (const uint32_t*)h->sample_keys.p,
                              (uint64_t)sample_tiles, sample_tiles, r_t, (uint64_t*)h->tau.p, (uint64_t*)nullptr));
      count_launch(h);
    }
  } else {
  // 1. sample
  PRG_TRY(score_sample(h, q_dev, B, sample_tiles, tile_stride, slots, use_tc));
  // 2. threshold = r-th largest sample key
  SelectParams st{};
  st.keys = (const uint64_t*)h->sample_keys.p; st.stride = slots; st.fixed_m = (uint32_t)slots;
  st.cap = st.fixed_m; st.k = (int)r_rank; st.tau = (uint64_t*)h->tau.p;
  PRG_TRY(launch_sample_select(h, SEL_KTH, st, B));
  }
  // 3. full pass with the threshold test fused into the tile epilogue
  const int pass_q = use_tc ? scan_tc_max_queries(h) : kQB;  // queries per pass over the matrix
  // GROUP mode (config "scan_groups", default on): more than 64 queries and the on-chip refine below
  const bool grouped = use_tc && scan_groups_on(h) && B > kQB && n_seg <= 512 && refine_smem_bytes(cand_cap, k) <= 200 * 1024 &&
                       (uint64_t)QT * n_seg * seg_cap < (1ull << 32);
  if (grouped) {
    // a GROUP-mode pass writes the lists / lengths of every group of its kernel's query capacity (recall_tc.cu: NQB * 4
    // groups — 16 for a last pass of 129..192 queries; a row whose norm bound is +inf survives for all of them): size
    // both for whole passes, not for whole blocks of 64 queries
    const size_t QTp = (QT + (size_t)pass_q - 1) / (size_t)pass_q * (size_t)pass_q;
    PRG_TRY(h->grp_cnt.ensure(QTp / kGrpQ * n_seg * 4));
    PRG_TRY(h->seg_rows.ensure(QTp * n_seg * seg_cap * 4));
  }
  bool i8_ok = use_tc && ((!grouped && scan_i8_available(h)) || (grouped && scan_i8g_available(h)));
  if (i8_ok && h->i8_backoff > 0) { --h->i8_backoff; i8_ok = false; }
  for (int q0 = 0; q0 < B; q0 += pass_q) {
    ScanParams sc{};
    sc.Q = q_dev + (size_t)q0 * dim; sc.nq = (B - q0 < pass_q) ? (B - q0) : pass_q;
    sc.n_rows = h->E_rows; sc.row_base = h->E_row_base;
    sc.n_tiles = n_tiles; sc.tile_stride = 1;
    sc.tau = (const uint64_t*)h->tau.p + q0; sc.cand = (uint64_t*)h->seg_keys.p + (size_t)q0 * seg_q; sc.seg_cap = seg_cap;
    sc.seg_cnt = (uint32_t*)h->cand_cnt.p + (size_t)q0 * n_seg;
    sc.row_norm = (const float*)h->row_norm.p;
    sc.cand_rows = use_tc ? (uint32_t*)h->seg_rows.p + (size_t)q0 * seg_q : nullptr;
    if (grouped) scan_group_outputs(h, sc, q0, n_seg, seg_cap);
    if (i8_ok && grouped) {
      PRG_TRY(launch_scan_i8g(h, sc, n_seg));
      h->last_filter = PRG_FILTER_INT8;
    } else if (i8_ok && sc.nq <= kQB) {
      PRG_TRY(launch_scan_i8(h, sc, n_seg));
      h->last_filter = PRG_FILTER_INT8;
    } else if (use_tc) {
      // dim 64, 129..256 queries: the 16-epilogue-warp form of the GROUP pass (recall_i8.cu); up to 128 queries half of its
      // warps would idle and recall_tc.cu's form is the faster one (c4 shard of a 2-GPU run: scan 0.110 against 0.134 ms)
      if (grouped && h->scan_grp16 && dim == 64 && sc.nq > 128) PRG_TRY(launch_scan_g16(h, sc, n_seg));
      else PRG_TRY(launch_scan_tc(h, sc));
      h->last_filter = h->scan_filter == SCAN_FILTER_BF16 ? PRG_FILTER_BF16 : PRG_FILTER_TF32;
    } else {
      PRG_TRY(scan(h, SCAN_THRESH, sc));
      h->last_filter = PRG_FILTER_FFMA2;
    }
  }
  uint32_t k_pow2 = 32;
  while (k_pow2 < (uint32_t)k) k_pow2 <<= 1;
  const uint32_t sort_cap = (k_pow2 - (uint32_t)k >= 16) ? k_pow2 : 2 * k_pow2;
  const size_t refine_smem = ((size_t)cand_cap + sort_cap) * 8;
  const double per_seg = ((double)r_rank / f) / n_seg;   // expected survivors per (segment, query)
  if (use_tc && n_seg <= 512 && refine_smem <= 200 * 1024) {
    // 4. exact re-score, then the top-k of the exact keys in shared memory
    if (grouped) PRG_TRY(launch_rescore_groups(h, q_dev, B, n_seg, seg_cap));
    else PRG_TRY(launch_rescore(h, q_dev, B, n_seg, seg_cap, per_seg));
    RefineParams rp{};
    rp.seg_keys = (const uint64_t*)h->seg_keys.p; rp.seg_counts = (const uint32_t*)h->cand_cnt.p;
    rp.n_seg = n_seg; rp.seg_cap = seg_cap; rp.cap = cand_cap;
    rp.tau_check = (const uint64_t*)h->tau.p;
    rp.k = k; rp.k_out = k; rp.expect = (uint32_t)((uint64_t)k < h->E_rows ? (uint64_t)k : h->E_rows);
    rp.sort_cap = sort_cap; rp.out_keys = keys_out; rp.flags = (int32_t*)h->flags.p; rp.max_count = max_cnt;
    PRG_TRY(launch_refine(h, rp, B, refine_smem));
  } else {
  if (use_tc) PRG_TRY(launch_rescore(h, q_dev, B, n_seg, seg_cap, per_seg));   // large k: the keys do not fit on chip
  // 4. exact top-k of the candidates
  SelectParams se{};
  se.keys = (const uint64_t*)h->seg_keys.p; se.stride = cand_cap; se.counts = nullptr;
  se.seg_counts = (const uint32_t*)h->cand_cnt.p; se.n_seg = n_seg; se.seg_cap = seg_cap;
  se.compact = (uint64_t*)h->cand_keys.p;
  if (use_tc) se.tau_check = (const uint64_t*)h->tau.p;
  se.cap = cand_cap; se.k = k; se.k_out = k;
  se.expect = (uint32_t)((uint64_t)k < h->E_rows ? (uint64_t)k : h->E_rows);
  se.out_keys = keys_out; se.flags = (int32_t*)h->flags.p;
  se.max_count = max_cnt;
  PRG_TRY(launch_select(h, SEL_TOPK, se, B));
  }
  // 5. per-query status -> pinned host memory; checked now, or later when deferred (fused path)
  if (h->host_flags_cap < (size_t)B) {
    if (h->host_flags) cudaFreeHost(h->host_flags);
    h->host_flags = nullptr;
    h->host_flags_cap = 0;
    const size_t cap = (size_t)B < 1024 ? 1024 : (size_t)B;
    PRG_CUDA(cudaHostAlloc((void**)&h->host_flags, (cap + 1) * 4, cudaHostAllocDefault));
    h->host_flags_cap = cap;
  }
  if (!h->flags_ev) PRG_CUDA(cudaEventCreateWithFlags(&h->flags_ev, cudaEventDisableTiming));
  PRG_CUDA(cudaMemcpyAsync(h->host_flags, h->flags.p, (size_t)B * 4, cudaMemcpyDeviceToHost, h->stream));
  PRG_CUDA(cudaMemcpyAsync(h->host_flags + h->host_flags_cap, max_cnt, 4, cudaMemcpyDeviceToHost, h->stream));
  PRG_CUDA(cudaEventRecord(h->flags_ev, h->stream));
  h->pending.active = true;
  h->pending.fused = false;
  h->pending.host_row = nullptr; h->pending.host_score = nullptr; h->pending.host_n = nullptr;
  h->pending.done = nullptr; h->pending.seq = 0;
  h->pending.B = B; h->pending.k = k; h->pending.q_dev = q_dev; h->pending.keys_out = keys_out;
  if (defer) return PRG_OK;
  bool repaired = false;
  return recall_resolve(h, &repaired);
}

int recall_resolve(prg_handle* h, bool* repaired) {
  *repaired = false;
  if (!h->pending.active) return PRG_OK;
  h->pending.active = false;
  PRG_CUDA(cudaEventSynchronize(h->flags_ev));
  const int B = h->pending.B, k = h->pending.k;
  const int32_t mx = h->host_flags[h->host_flags_cap];
  if (mx > h->last_max_cand) h->last_max_cand = mx;
  for (int q = 0; q < B; ++q) {
    if (h->host_flags[q] != 0) {  // candidate list under/overflowed: redo this query through the dense path
      ++h->last_fallback;
      *repaired = true;
      PRG_TRY(recall_dense(h, h->pending.q_dev + (size_t)q * h->E_dim, 1, k, k, h->pending.keys_out + (size_t)q * k));
    }
  }
  // the int8 bound lets more rows through than the bf16 one: when a batch overflowed under it, the next batches use bf16
  if (h->last_fallback > 0 && h->last_filter == PRG_FILTER_INT8) h->i8_backoff = 64;
  return PRG_OK;
}

int keys_to_outputs(prg_handle* h, const uint64_t* keys_dev, int B, int k, uint32_t* out_row, float* out_score,
                    int32_t* out_n) {
  const int total = B * k;
  PRG_CUDA(cudaMemsetAsync(out_n, 0, (size_t)B * 4, h->stream));
  keys_unpack_kernel<<<(total + 255) / 256, 256, 0, h->stream>>>(keys_dev, total, k, out_row, out_score, out_n);
  PRG_CUDA(cudaGetLastError());
  count_launch(h);
  return PRG_OK;
}

// Shard merge: keys_dev is the all-gather output [G][B][k]; per query the G*k keys are distinct
// (global rows), so the same exact select yields the replica-identical global top-k.
__global__ void regroup_kernel(const uint64_t* in, uint64_t g_stride, uint64_t* out, int G, int B, int k) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t total = (size_t)G * B * k;
  if (i >= total) return;
  const int j = (int)(i % k);
  const int b = (int)((i / k) % B);
  const int g = (int)(i / ((size_t)k * B));
  out[((size_t)b * G + g) * k + j] = in[(size_t)g * g_stride + (size_t)b * k + j];
}

int merge_keys_device(prg_handle* h, const uint64_t* keys_dev, int G, uint64_t g_stride, int B, int k,
                      uint64_t* keys_out) {
  if (G <= 0 || B <= 0 || k <= 0) return fail(PRG_EINVAL, "G, B, k must be positive");
  if (k > 4096) return fail(PRG_EUNSUPPORTED, "k > 4096");
  const size_t total = (size_t)G * B * k;
  PRG_TRY(h->dense_keys.ensure(total * 8));
  regroup_kernel<<<(unsigned)((total + 255) / 256), 256, 0, h->stream>>>(keys_dev, g_stride, (uint64_t*)h->dense_keys.p, G,
                                                                          B, k);
  PRG_CUDA(cudaGetLastError());
  count_launch(h);
  SelectParams se{};
  se.keys = (const uint64_t*)h->dense_keys.p; se.stride = (uint64_t)G * k; se.fixed_m = (uint32_t)(G * k);
  se.cap = se.fixed_m; se.k = k; se.k_out = k; se.out_keys = keys_out;
  PRG_TRY(launch_select(h, SEL_TOPK, se, B));
  return PRG_OK;
}

// ------------------------------------------------------------------ sharded recall with ONE global threshold per query
// (SURVEY §8e).  With the protocol above (prg_recall_local_keys) every shard refines its own ~4k survivors per query
// to a local top-k, so the refine work of a rank grows with the number of shards G (it sees all B*G queries).  Here
// the threshold is global: each shard samples 1/128 of its rows and contributes its r best sample keys per query; after
// a first, tiny all-gather every rank takes the r-th largest of the G*r keys as tau (r/f ~ 4k rows of the WHOLE
// catalog reach it), filters its shard against that tau (~4k/G survivors per query), re-scores them exactly and emits
// them sorted (at most k).  The second all-gather carries those lists plus one status word per query; the union of the
// lists holds every row whose exact key reaches tau, so if at least k of them exist the merged top-k is the global
// top-k.  The two failure modes — a list that overflowed (tau too low) and fewer than k rows reaching tau (tau too
// high), both < 1e-10 per query for row orders a strided sample represents — are detected identically on every rank
// from the gathered data (shard_check), and the host then redoes the batch with the exact local protocol.
constexpr int kShardSampleStride = 128;   // one tile in 128 is sampled
constexpr uint32_t kShardMaxSampleKeys = 2048;

struct ShardPlan {
  bool sampled;
  uint32_t n_tiles, sample_tiles, tile_stride, r;
  uint64_t slots;
  uint32_t n_seg, seg_cap, cand_cap, sort_cap;
  size_t refine_smem;
};

static uint32_t shard_r(int k) {   // sample keys per query and shard; depends on k only, so every rank agrees on it
  const uint32_t target = 4u * (uint32_t)(k < 1024 ? 1024 : k);
  return (target + kShardSampleStride - 1) / kShardSampleStride;
}

static ShardPlan shard_plan(const prg_handle* h, int k, int G) {
  ShardPlan pl{};
  pl.n_tiles = (uint32_t)((h->E_rows + kTileRows - 1) / kTileRows);
  pl.r = shard_r(k);
  pl.sample_tiles = pl.n_tiles / kShardSampleStride;
  // a shard joins the sampled protocol when its sample can hold r keys per query; G equal shards decide alike
  pl.sampled = !h->scan_ffma2 && pl.sample_tiles >= 8 && (uint64_t)pl.sample_tiles * kTileRows >= 4ull * pl.r &&
               (uint64_t)k * 64 <= h->E_rows * (uint64_t)G;
  if (!pl.sampled) return pl;
  pl.tile_stride = pl.n_tiles / pl.sample_tiles;
  pl.slots = (uint64_t)pl.sample_tiles * kTileRows;
  const double expect = (double)pl.r * kShardSampleStride / (double)G;   // survivors per query on this shard
  pl.cand_cap = ((uint32_t)(4.0 * expect) + 1024 + 1023) & ~1023u;
  pl.n_seg = pl.n_tiles < (uint32_t)h->sm_count ? pl.n_tiles : (uint32_t)h->sm_count;
  pl.seg_cap = ((uint32_t)(8.0 * expect / pl.n_seg) + 32 + 15) & ~15u;
  uint32_t k_pow2 = 32;
  while (k_pow2 < (uint32_t)k) k_pow2 <<= 1;
  pl.sort_cap = (k_pow2 - (uint32_t)k >= 16) ? k_pow2 : 2 * k_pow2;
  pl.refine_smem = ((size_t)pl.cand_cap + pl.sort_cap) * 8;
  return pl;
}

int shard_sample_len(int k) { return (int)shard_r(k); }

// phase 1: the r best keys of this shard's sample, per query (sorted, 0-padded); zeros if the shard is not sampled
int recall_shard_sample_device(prg_handle* h, const float* q_dev, int Bg, int k, int G, uint64_t* out) {
  if (!h->E || !h->E_map_ok) return fail(PRG_ESTATE, "item matrix not set (prg_set_item_matrix)");
  if (Bg <= 0 || k <= 0 || G <= 0) return fail(PRG_EINVAL, "Bg, k, G must be positive");
  if (k > 4096) return fail(PRG_EUNSUPPORTED, "k > 4096");
  const ShardPlan pl = shard_plan(h, k, G);
  if ((uint64_t)pl.r * (uint64_t)G > kShardMaxSampleKeys) return fail(PRG_EUNSUPPORTED, "G * sample keys per query > 2048");
  if (!pl.sampled) {
    PRG_CUDA(cudaMemsetAsync(out, 0, (size_t)Bg * pl.r * 8, h->stream));
    return PRG_OK;
  }
  const uint32_t dim = h->E_dim;
  const int nblk = (Bg + kQB - 1) / kQB;
  PRG_TRY(h->sample_keys.ensure((size_t)nblk * kQB * pl.slots * 8));
  if (!h->row_norm.p) PRG_TRY(build_row_norms(h));
  {
    // config "recall_tilemax" (same on every rank): the shard contributes its r largest TILE MAXIMA instead
    // of its r largest sample keys; the r-th largest of the gathered maxima estimates the same row share (the share of
    // the G * T sample tiles that hold a row above tau is r / (G T) ~ 256 * target / rows for small shares).
    const double rows_all = (double)h->E_rows * G;
    const double tile_x = (double)kTileRows * 4.0 * (k < 1024 ? 1024 : k) / rows_all;
    const double tile_r = (double)pl.sample_tiles * G * (1.0 - exp(-tile_x));
    if (h->recall_tilemax && scan_tc_dense_available(h) && tile_x <= 0.35 && tile_r >= 24.0 && pl.sample_tiles <= 2048) {
      ScanParams sp{};
      sp.n_rows = h->E_rows; sp.row_base = h->E_row_base;
      sp.n_tiles = pl.sample_tiles; sp.tile_stride = pl.tile_stride;
      sp.row_norm = (const float*)h->row_norm.p;
      sp.Q = q_dev; sp.nq = Bg;
      sp.dense = (uint64_t*)h->sample_keys.p;   // used as u32 [Bg][sample_tiles]
      sp.dense_stride = pl.sample_tiles;
      PRG_TRY(launch_scan_tc_tilemax(h, sp));
      StageScope span(h, ST_SELECT);
      PRG_CUDA(launch_chained(h, tilemax_tau_kernel, dim3(Bg), dim3(256), 0, 1, (const uint32_t*)h->sample_keys.p,
                              (uint64_t)pl.sample_tiles, pl.sample_tiles, pl.r, (uint64_t*)nullptr, out));
      count_launch(h);
      return PRG_OK;
    }
  }
  PRG_TRY(score_sample(h, q_dev, Bg, pl.sample_tiles, pl.tile_stride, pl.slots, true));
  SelectParams st{};
  st.keys = (const uint64_t*)h->sample_keys.p; st.stride = pl.slots; st.fixed_m = (uint32_t)pl.slots;
  st.cap = st.fixed_m; st.k = (int)pl.r; st.k_out = (int)pl.r; st.out_keys = out;
  return launch_sample_select(h, SEL_TOPK, st, Bg);
}

// tau[q] = r-th largest of the G*r gathered sample keys (0 = no threshold when fewer than r are valid)
__global__ void __launch_bounds__(256) shard_tau_kernel(const uint64_t* __restrict__ all, int G, int Bg, uint32_t r,
                                                        uint64_t* __restrict__ tau) {
  __shared__ uint64_t sk[kShardMaxSampleKeys];
  const int q = blockIdx.x, tid = threadIdx.x;
  const uint32_t m = (uint32_t)G * r;
  uint32_t K2 = 32;
  while (K2 < m) K2 <<= 1;
  for (uint32_t i = tid; i < K2; i += 256) {
    const uint32_t g = i / r, j = i - g * r;
    sk[i] = i < m ? all[((size_t)g * Bg + q) * r + j] : 0ull;
  }
  __syncthreads();
  for (uint32_t size = 2; size <= K2; size <<= 1) {
    for (uint32_t stride = size >> 1; stride > 0; stride >>= 1) {
      for (uint32_t i = tid; i < (K2 >> 1); i += 256) {
        const uint32_t pos = 2 * i - (i & (stride - 1));
        const uint64_t a = sk[pos], b = sk[pos + stride];
        if ((a < b) == ((pos & size) == 0)) { sk[pos] = b; sk[pos + stride] = a; }
      }
      __syncthreads();
    }
  }
  if (tid == 0) tau[q] = sk[r - 1];
}

__global__ void shard_status_kernel(const int32_t* __restrict__ flags, int Bg, uint64_t* __restrict__ status) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q < Bg) status[q] = flags ? (uint64_t)(uint32_t)flags[q] : 0ull;
}

// phase 2: out = [Bg][k] exact keys of this shard's rows that reach the global tau (sorted, at most k, 0-padded)
// followed by Bg status words (0 ok, 1 = the candidate list overflowed)
int recall_shard_candidates_device(prg_handle* h, const float* q_dev, int Bg, int k, int G, const uint64_t* all_samples,
                                   uint64_t* out) {
  if (!h->E || !h->E_map_ok) return fail(PRG_ESTATE, "item matrix not set (prg_set_item_matrix)");
  if (Bg <= 0 || k <= 0 || G <= 0) return fail(PRG_EINVAL, "Bg, k, G must be positive");
  if (k > 4096) return fail(PRG_EUNSUPPORTED, "k > 4096");
  if (h->E_rows + h->E_row_base > 0xFFFFFFFFull) return fail(PRG_EUNSUPPORTED, "global row ids must fit u32");
  const ShardPlan pl = shard_plan(h, k, G);
  if ((uint64_t)pl.r * (uint64_t)G > kShardMaxSampleKeys) return fail(PRG_EUNSUPPORTED, "G * sample keys per query > 2048");
  uint64_t* status = out + (size_t)Bg * k;
  const uint32_t dim = h->E_dim;
  const size_t QT = (size_t)((Bg + kQB - 1) / kQB) * kQB;
  PRG_TRY(h->tau.ensure(QT * 8));
  PRG_TRY(h->flags.ensure(QT * 4));
  h->pending.active = false;
  {
    StageScope span(h, ST_SELECT);
    shard_tau_kernel<<<Bg, 256, 0, h->stream>>>(all_samples, G, Bg, pl.r, (uint64_t*)h->tau.p);
    PRG_CUDA(cudaGetLastError());
    count_launch(h);
  }
  if (!pl.sampled) {   // small shard: its exact local top-k is a superset of what the protocol needs
    for (int q0 = 0; q0 < Bg; q0 += kQB) {
      const int nq = (Bg - q0 < kQB) ? (Bg - q0) : kQB;
      PRG_TRY(recall_dense(h, q_dev + (size_t)q0 * dim, nq, k, k, out + (size_t)q0 * k));
    }
    shard_status_kernel<<<(Bg + 255) / 256, 256, 0, h->stream>>>(nullptr, Bg, status);
    PRG_CUDA(cudaGetLastError());
    count_launch(h);
    return PRG_OK;
  }
  if (pl.n_seg > 512 || pl.refine_smem > 200 * 1024) return fail(PRG_EUNSUPPORTED, "k too large for the global-threshold protocol");
  const size_t seg_q = (size_t)pl.n_seg * pl.seg_cap;
  PRG_TRY(h->seg_keys.ensure(QT * seg_q * 8));
  PRG_TRY(h->seg_rows.ensure(QT * seg_q * 4));
  if (!h->row_norm.p) PRG_TRY(build_row_norms(h));
  PRG_TRY(h->cand_cnt.ensure(QT * pl.n_seg * 4 + 4));
  uint32_t* max_cnt = (uint32_t*)h->cand_cnt.p + QT * pl.n_seg;
  PRG_CUDA(cudaMemsetAsync(max_cnt, 0, 4, h->stream));
  const int pass_q = scan_tc_max_queries(h);
  const bool grouped = scan_groups_on(h) && Bg > kQB && (uint64_t)QT * seg_q < (1ull << 32);   // GROUP mode, as in recall_topk_device
  if (grouped) {   // whole passes (see recall_topk_device)
    const size_t QTp = (QT + (size_t)pass_q - 1) / (size_t)pass_q * (size_t)pass_q;
    PRG_TRY(h->grp_cnt.ensure(QTp / kGrpQ * pl.n_seg * 4));
    PRG_TRY(h->seg_rows.ensure(QTp * seg_q * 4));
  }
  for (int q0 = 0; q0 < Bg; q0 += pass_q) {
    ScanParams sc{};
    sc.Q = q_dev + (size_t)q0 * dim; sc.nq = (Bg - q0 < pass_q) ? (Bg - q0) : pass_q;
    sc.n_rows = h->E_rows; sc.row_base = h->E_row_base;
    sc.n_tiles = pl.n_tiles; sc.tile_stride = 1;
    sc.tau = (const uint64_t*)h->tau.p + q0; sc.seg_cap = pl.seg_cap;
    sc.seg_cnt = (uint32_t*)h->cand_cnt.p + (size_t)q0 * pl.n_seg;
    sc.row_norm = (const float*)h->row_norm.p;
    sc.cand_rows = (uint32_t*)h->seg_rows.p + (size_t)q0 * seg_q;
    if (grouped) scan_group_outputs(h, sc, q0, pl.n_seg, pl.seg_cap);
    h->last_filter = (grouped && scan_i8g_available(h)) ? PRG_FILTER_INT8 : (h->scan_filter == SCAN_FILTER_BF16 ? PRG_FILTER_BF16 : PRG_FILTER_TF32);
    if (grouped && scan_i8g_available(h)) PRG_TRY(launch_scan_i8g(h, sc, pl.n_seg));   // dim 128: int8 index (recall_i8.cu)
    else if (grouped && h->scan_grp16 && dim == 64 && sc.nq > 128) PRG_TRY(launch_scan_g16(h, sc, pl.n_seg));
    else PRG_TRY(launch_scan_tc(h, sc));
  }
  {
    if (grouped) PRG_TRY(launch_rescore_groups(h, q_dev, Bg, pl.n_seg, pl.seg_cap));
    else PRG_TRY(launch_rescore(h, q_dev, Bg, pl.n_seg, pl.seg_cap, (double)pl.r * kShardSampleStride / (double)G / pl.n_seg));
    RefineParams rp{};
    rp.seg_keys = (const uint64_t*)h->seg_keys.p; rp.seg_counts = (const uint32_t*)h->cand_cnt.p;
    rp.n_seg = pl.n_seg; rp.seg_cap = pl.seg_cap; rp.cap = pl.cand_cap;
    rp.k = k; rp.k_out = k; rp.expect = 0; rp.sort_cap = pl.sort_cap;
    rp.out_keys = out; rp.flags = (int32_t*)h->flags.p; rp.max_count = max_cnt;
    PRG_TRY(launch_refine(h, rp, Bg, pl.refine_smem));
    StageScope span(h, ST_SELECT);
    shard_status_kernel<<<(Bg + 255) / 256, 256, 0, h->stream>>>((const int32_t*)h->flags.p, Bg, status);
    PRG_CUDA(cudaGetLastError());
    count_launch(h);
  }
  return PRG_OK;
}

// phase 3 check, identical on every rank: gathered = G blocks of [Bg*k keys | Bg status words]; a query needs the exact
// protocol if a shard reported an overflow or fewer than k gathered keys reach its tau.  retry[0] |= 1, retry[1] += count.
__global__ void __launch_bounds__(128) shard_check_kernel(const uint64_t* __restrict__ keys, uint64_t key_stride,
                                                          const uint64_t* __restrict__ status, uint64_t status_stride, int G,
                                                          int Bq, int k, const uint64_t* __restrict__ tau,
                                                          int32_t* __restrict__ retry) {
  // one WARP per query, lane g walks shard g's list: the G binary searches (dependent loads through L2) run side by side
  // instead of one after the other in a single thread (46 us -> a few us at G = 8, Bg = 512).
  // Shard g's lists: keys + g * key_stride ([Bq][k], sorted descending, 0-padded); its status words: status + g * status_stride.
  const int q = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  if (q >= Bq) return;
  const uint64_t t = tau[q];
  const uint64_t lim = t ? t : 1ull;
  bool bad = false;
  uint32_t reach = 0;
  for (int g = lane; g < G; g += 32) {
    if (status[(size_t)g * status_stride + q] != 0ull) bad = true;
    const uint64_t* list = keys + (size_t)g * key_stride + (size_t)q * k;   // count keys >= max(t, 1)
    int lo = 0, hi = k;
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (list[mid] >= lim) lo = mid + 1; else hi = mid; }
    reach += (uint32_t)lo;
  }
  reach = __reduce_add_sync(0xffffffffu, reach);
  bad = __any_sync(0xffffffffu, bad);
  if (t != 0ull && reach < (uint32_t)k) bad = true;
  if (bad && lane == 0) { atomicOr(&retry[0], 1); atomicAdd(&retry[1], 1); }
}

int shard_check_device(prg_handle* h, const uint64_t* gathered, int G, int Bg, int k, int32_t* retry_dev, const uint64_t* tau) {
  if (!h->tau.p) return fail(PRG_ESTATE, "prg_shard_candidates has not run on this handle");
  const uint64_t blk = (uint64_t)Bg * k + Bg;
  shard_check_kernel<<<(Bg + 3) / 4, 128, 0, h->stream>>>(gathered, blk, gathered + (size_t)Bg * k, blk, G, Bg, k,
                                                         tau ? tau : (const uint64_t*)h->tau.p, retry_dev);
  PRG_CUDA(cudaGetLastError());
  count_launch(h);
  return PRG_OK;
}

// [Bg][k] lists + [Bg] status words (prg_shard_candidates' output) -> G owner-major chunks of [B][k] lists + [B] status words:
// chunk o is what rank o needs of this shard (its own B queries), so ONE all-to-all moves B*k + B words per pair of ranks
// instead of an all-gather of everything to everybody
__global__ void shard_pack_owner_kernel(const uint64_t* __restrict__ in, int Bg, int B, int k, uint64_t* __restrict__ out) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t chunk = (size_t)B * k + B, total = (size_t)(Bg / B) * chunk;
  if (i >= total) return;
  const size_t o = i / chunk, r = i - o * chunk;
  out[i] = r < (size_t)B * k ? in[o * (size_t)B * k + r] : in[(size_t)Bg * k + o * B + (r - (size_t)B * k)];
}

int shard_pack_owner_device(prg_handle* h, const uint64_t* in, int Bg, int B, int k, uint64_t* out) {
  if (B <= 0 || Bg % B != 0) return fail(PRG_EINVAL, "Bg must be a multiple of B");
  const size_t total = (size_t)Bg * k + Bg;
  shard_pack_owner_kernel<<<(unsigned)((total + 255) / 256), 256, 0, h->stream>>>(in, Bg, B, k, out);
  PRG_CUDA(cudaGetLastError());
  count_launch(h);
  return PRG_OK;
}

// the check for the B queries [q0, q0 + B) this rank owns, on what the all-to-all delivered: G chunks of [B*k | B]
int shard_check_owner_device(prg_handle* h, const uint64_t* received, int G, int B, int k, int q0, int32_t* retry_dev) {
  if (!h->tau.p) return fail(PRG_ESTATE, "prg_shard_candidates has not run on this handle");
  const uint64_t blk = (uint64_t)B * k + B;
  shard_check_kernel<<<(B + 3) / 4, 128, 0, h->stream>>>(received, blk, received + (size_t)B * k, blk, G, B, k,
                                                        (const uint64_t*)h->tau.p + q0, retry_dev);
  PRG_CUDA(cudaGetLastError());
  count_launch(h);
  return PRG_OK;
}

}  // namespace prg
