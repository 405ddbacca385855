// recall_tc.cu — tensor-core candidate filter for the recall scan: moves the full-matrix pass
// from the FP32-issue limit (recall.cu, 64 queries/pass: 81.9 GFLOP of FFMA2 per 2.56 GB) to the HBM roofline while
// the RESULT stays bit-exact.
//
//   filter (this kernel):  approx(row,q) = TF32 tcgen05.mma of the item tile against the query block, fp32 accum in
//                          TMEM; a row is kept for query q unless  approx < tau_f[q] - c*||x_row||*||q||,  where
//                          tau_f is the exact sampled threshold (recall.cu steps 1-2), ||x_row|| a precomputed upper
//                          bound of the row norm and c = 1.05 * 2^-9 bounds |approx - exact| (two TF32 operand
//                          truncations of <= 2^-10 each, Cauchy-Schwarz, fp32 accumulation slack).  Every row whose
//                          EXACT key reaches tau therefore survives; a few percent extra rows survive too.
//   refine (select_kernel): survivors (~4.5k rows per query) are re-scored with the exact fmaf chain of the arithmetic
//                          contract while they are packed, then the exact radix select / sort runs as before, and
//                          the k-th exact key is checked against tau (else the query falls back to the dense path).
//
// Structure: persistent, one CTA per SM, 10 warps: warp 0 = TMA producer (same 256-row x 64-dim SWIZZLE_128B stages
// as the FFMA2 scan — the fp32 rows are consumed by kind::tf32 as they lie in HBM, no conversion pass), warp 1 = TMEM
// allocator + the single MMA-issuing thread (M=128, N=64, K=8; two 128-row halves per stage, accumulators double
// buffered in 256 TMEM columns), warps 2-9 = epilogue (tcgen05.ld 32x32b.x32 -> 64 FFMA + 64 FSETP per row).
// Survivor rows go to the per-CTA, per-query segments with one shared-memory atomic + one store.
//
// BF16 filter index (default): the same filter over a bf16 SHADOW of the item matrix (round-to-nearest copy built once
// when the matrix is set, next to the row norms): the full pass then streams rows*dim*2 bytes instead of rows*dim*4
// (kind::f16 MMA, one 128-B swizzle row per 64 dims).  The bound grows to c = 1.05 * 2^-8: |bf16(x)-x| <= 2^-9 |x|
// for both operands, so |approx - exact| <= (2^-8 + 2^-18) * sum|x_d||q_d| <= ... * ||x|| * ||q||; products of two
// bf16 are exact in fp32, the 5 % slack covers the fp32 accumulation of either side, and both norm bounds carry a
// +1e-30 absolute term for subnormal elements.  The fp32 matrix stays the authority: survivors are re-scored from
// it, so every emitted row and score is still bit-identical to the oracle.  Config "scan_tf32":1 selects the
// fp32-operand filter above.
#include "recall.h"
#include <cuda_bf16.h>
#include <type_traits>

namespace prg {

constexpr int kTcThreads = 320;
constexpr int kTcEpiWarps = 8;
constexpr float kTcMarginTf32 = 1.05f / 512.f;  // c = 1.05 * 2^-9
constexpr float kTcMarginBf16 = 1.05f / 256.f;  // c = 1.05 * 2^-8

// NQB = query blocks (of 64) handled per pass over the matrix: the tile is read from HBM once and multiplied against
// every block (the shard of a G-GPU run sees 64*G queries per step: one pass instead of G).
// A stage is 256 rows x 64 dims of fp32 (64 KiB, two 32-float TMA boxes) for the tf32 filter and 256 rows x DIM of
// bf16 (DIM/64 boxes of 32 KiB) for the bf16 filter.
// 256 queries per pass over the bf16 index (NQB == 4): a stage is HALF a tile (128 rows, the unit of the per-half
// accumulator barriers below) — the ring then turns over at half-tile granularity: a half's buffer is requested again as
// soon as its own MMAs are through, not when the whole tile's are.  With whole-tile stages the dim-128 pass had two
// 64-KiB stages and ran at (MMA + load latency) / 2 per tile: tensor pipe 51 % busy, DRAM 49 % (ncu r3d, C5 shard shape).
template <int NQB, bool BF>
constexpr bool tc_half_stages() { return BF && NQB == 4; }
template <int DIM, int NQB, bool BF>
constexpr int tc_stage_bytes() { return BF ? (tc_half_stages<NQB, BF>() ? kTileRows / 2 : kTileRows) * DIM * 2 : kStageBytes; }
template <int DIM, int NQB, bool BF>
constexpr int tc_stages() {
  return BF ? (DIM == 64 ? (NQB <= 2 ? 5 : 8) : (NQB == 1 ? 3 : (NQB == 2 ? 2 : 4))) : ((NQB == 1 && DIM == 64) ? 3 : 2);
}
// The row norms of a tile travel with it (one 1 KiB bulk copy completing on the tile's `full` barrier) into a ring of
// S + 3 slots: the producer may run S tiles ahead of the MMA warp and the MMA warp 2 tiles (accumulator buffers) ahead
// of the epilogue, which reads the norms.
template <int DIM, int NQB, bool BF>
constexpr int tc_norm_ring() { return tc_stages<DIM, NQB, BF>() + 3; }
template <int DIM, int NQB, bool BF>
constexpr size_t scan_tc_smem_bytes() {
  return (size_t)tc_stages<DIM, NQB, BF>() * tc_stage_bytes<DIM, NQB, BF>() + (size_t)NQB * DIM * kQB * (BF ? 2 : 4) /*Q operands*/ +
         (size_t)NQB * 3 * kQB * 4 /*tauf, qn, s_cnt*/ + (2 * tc_stages<DIM, NQB, BF>() + 4) * 8 + 16 +
         (size_t)tc_norm_ring<DIM, NQB, BF>() * kTileRows * 4 /*row-norm ring*/;
}

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}

// Epilogue survivor test, shared by both operand types: keep row r for query q unless approx < tau_f[q] - nr * c||q||.
// The 64 queries of a block are tested in four groups of 16 and only a group with a survivor is walked again to append
// rows: ~5 k survivors per query and 10 M rows make a survivor a 3 % event per row but a 60 % event per WARP and tile,
// and walking all 64 queries for each of those was half of the epilogue's instructions.
__device__ __forceinline__ float fmax3(float a, float b, float c) {  // SASS FMNMX3; NaN operands are ignored
  float r;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
  return r;
}

// MODE = SCAN_DENSE: no threshold — every accumulator is written out as an order key (approximate score, row) into
// p.dense[q][launch tile * 256 + r].  Used for the strided SAMPLE from which the threshold is estimated: tau is a pruning
// hint only (the filter keeps a superset of {exact >= tau} for any tau and the refine step checks that set), so the
// sample does not need exact scores, and the tensor cores score it in a fraction of the FFMA2 kernel's time.
template <int DIM, int NQB, bool BF, int MODE = SCAN_THRESH, bool GROUP = false>   // GROUP: ScanParams::grp_rows form
__global__ void __launch_bounds__(kTcThreads, 1)
recall_scan_tc_kernel(const __grid_constant__ CUtensorMap emap, const ScanParams p_in) {
  extern __shared__ __align__(1024) uint8_t smem[];
  constexpr bool kDenseMode = MODE != SCAN_THRESH;   // DENSE and TILEMAX: no thresholds, one grid row per query block
  // DENSE launches carry one grid row per block of 64 queries (the sample is a few dozen tiles: tiles x query blocks
  // CTAs fill the machine, tiles alone do not); grid row y sees its own slice of the queries and of the output
  ScanParams p_adj = p_in;
  if constexpr (kDenseMode) {
    const int y = (int)blockIdx.y;
    p_adj.Q = p_in.Q + (size_t)y * kQB * DIM;
    p_adj.nq = p_in.nq - y * kQB < kQB ? p_in.nq - y * kQB : kQB;
    if constexpr (MODE == SCAN_DENSE) p_adj.dense = p_in.dense + (size_t)y * kQB * p_in.dense_stride;
  }
  const ScanParams& p = kDenseMode ? p_adj : p_in;
  constexpr int KH = BF ? 1 : DIM / 64;                    // stages per tile
  constexpr int kTcStages = tc_stages<DIM, NQB, BF>();
  constexpr int kStageB = tc_stage_bytes<DIM, NQB, BF>();
  constexpr int kQBytes = DIM * kQB * (BF ? 2 : 4);        // one query block as a B operand
  constexpr int NBUF = NQB <= 2 ? 2 : 1;                   // accumulator buffers (128 TMEM columns per block and buffer)
  // NQB = 4 fills all 512 TMEM columns with ONE tile's accumulators (256 rows x 256 queries), so whole tiles cannot be
  // double buffered.  The two 128-row HALVES of a tile are independent though (own MMAs, own TMEM columns, own four
  // epilogue warps): with one full / empty barrier pair per half the MMAs of a half run while the other half is being
  // read, and the next tile's first half starts as soon as its four warps are through.  Measured before (C5 shard shape,
  // 12.5 M x 128, 256 queries per pass): 1.10 ms per pass = MMA (2 x 1024 clk) + TMEM read (4096 clk) per tile back to back.
  constexpr bool kHalfBars = NBUF == 1 && BF;              // (bf16 index: one stage per tile)
  constexpr uint32_t kTmemCols = NQB == 1 ? 256 : 512;
  constexpr int QTOT = NQB * kQB;
  constexpr float kMargin = BF ? kTcMarginBf16 : kTcMarginTf32;
  constexpr bool kDense = MODE != SCAN_THRESH;
  constexpr bool kTileMax = MODE == SCAN_TILEMAX;
  uint8_t* stage_base = smem;
  uint8_t* Qb = smem + (size_t)kTcStages * kStageB;  // B operand, K-major SWIZZLE_128B: tf32 [DIM/32][QTOT q][32 f32],
                                                     // bf16 [DIM/64][QTOT q][64 bf16] — ALL queries of the pass are the N
                                                     // dimension of one MMA (N = 64 / 128 / 256)
  float2* tq = reinterpret_cast<float2*>(Qb + (size_t)NQB * kQBytes);           // [NQB*64] {tau_f, c*||q||}
  uint32_t* s_cnt = reinterpret_cast<uint32_t*>(tq + QTOT);
  uint64_t* full = reinterpret_cast<uint64_t*>(s_cnt + QTOT);
  uint64_t* empty = full + kTcStages;
  uint64_t* tfull = empty + kTcStages;   // [2] accumulator buffer ready
  uint64_t* tempty = tfull + 2;          // [2] accumulator buffer drained
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
  constexpr int kNormRing = tc_norm_ring<DIM, NQB, BF>();
  float* nrm = reinterpret_cast<float*>(tmem_slot + 4);   // [kNormRing][256], 16-B aligned

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  // Uniform threshold: when every query of the pass has a positive, well-scaled threshold the B operand is staged as
  // q / tau_q, so "exact >= tau_q" becomes "exact' >= 1" for every query and the survivor test needs ONE per-row
  // scalar, t_r = (1 - 1e-6) - ||x_r|| * max_q(c ||q'_q||): the epilogue is then a running maximum (FMNMX3) over the
  // accumulators plus one compare per 16 queries instead of an FFMA + FSETP per accumulator (the bf16 pass was
  // epilogue-bound at 0.40 ms with the per-query form).  The 1e-6 covers fl(1/tau)*tau != 1; taking the largest
  // c||q'|| for all queries only lets a few more rows through.  Otherwise (a threshold <= 0, none, or extreme) the
  // per-query form below is used for the whole pass.
  __shared__ float s_scale[NQB * kQB];
  __shared__ float s_ss[NQB * kQB];     // squared norms of the staged (scaled) queries
  __shared__ uint32_t s_cmax;
  if (tid == 0) {
    s_cmax = 0u;
    tma_prefetch_desc(&emap);
    for (int s = 0; s < kTcStages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&tfull[b], 1); mbar_init(&tempty[b], kHalfBars ? kTcEpiWarps / 2 : kTcEpiWarps); }
    mbar_fence_init();
  }
  const uint32_t my_tiles = (p.n_tiles > blockIdx.x) ? (p.n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  // one tile = SPT stages of data + its row norms
  constexpr bool kHalfStage = tc_half_stages<NQB, BF>();
  constexpr int SPT = kHalfStage ? 2 : KH;               // stages per tile
  constexpr int kStageRows = kHalfStage ? kTileRows / 2 : kTileRows;
  auto issue_tile = [&](uint32_t i, uint32_t it) {
    const uint32_t t = blockIdx.x + i * gridDim.x;
    const int row0 = (int)(t * p.tile_stride * (uint32_t)kTileRows);
    for (int h = 0; h < SPT; ++h, ++it) {
      const uint32_t s = it % kTcStages, ph = (it / kTcStages) & 1u;
      mbar_wait(&empty[s], ph ^ 1u);
      mbar_arrive_expect_tx(&full[s], kStageB + (h == 0 ? kTileRows * 4 : 0));
      uint8_t* dst = stage_base + (size_t)s * kStageB;
      if (h == 0)   // the norm array is padded to whole tiles (build_row_norms), so the copy never runs past it
        bulk_load_1d(nrm + (size_t)(i % kNormRing) * kTileRows, p.row_norm + (size_t)row0, kTileRows * 4, &full[s]);
      if constexpr (BF) {
#pragma unroll
        for (int sub = 0; sub < DIM / 64; ++sub)   // one box = kStageRows rows x 64 bf16 (128 B); half stages: rows [128 h, +128)
          tma_load_2d(dst + sub * (kStageRows * 128), &emap, sub * 64, row0 + (kHalfStage ? h * kStageRows : 0), &full[s], kEvictFirst);
      } else {
        tma_load_2d(dst, &emap, h * 64, row0, &full[s], kEvictFirst);
        tma_load_2d(dst + kSubTileFloats * 4, &emap, h * 64 + 32, row0, &full[s], kEvictFirst);
      }
    }
  };
  // The first tiles are requested at once, by the thread that initialised the barriers: the matrix and its norms are
  // tables no kernel of the per-batch chain writes, so on a chained launch the ring fills while the previous kernel
  // drains (everything below pdl_wait() reads what that kernel wrote: thresholds, queries) and while the query
  // operands are staged.
  constexpr uint32_t kPrefetchTiles = kTcStages / SPT;
  const uint32_t n_pre = my_tiles < kPrefetchTiles ? my_tiles : kPrefetchTiles;
  if (tid == 0)
    for (uint32_t i = 0; i < n_pre; ++i) issue_tile(i, i * SPT);
  pdl_wait();
  int bad = 0;
  for (int q = tid; q < QTOT; q += kTcThreads) {
    float sc = 0.f;
    if (kDense) bad = 1;
    else if (q < p.nq) {
      const uint64_t t = p.tau[q];
      const float tf = (t == 0) ? __int_as_float(0xFF800000) : key_score(t);
      if (tf > 0x1p-60f && tf < 0x1p60f) sc = 1.0f / tf; else bad = 1;
    }
    s_scale[q] = sc;
  }
  const bool scaled = __syncthreads_or(bad) == 0 && !kDense;   // (also publishes the mbarrier initialisation)

  // query blocks -> K-major SWIZZLE_128B B operands (16-B chunk index XOR row-in-group); padded queries are zero rows.
  // 16 bytes per thread and iteration, four iterations' loads in flight: the element-wise form of this loop and a
  // per-query serial norm loop cost 15-20 us per pass at 256 queries (ncu r3d: a quarter of a c4-shard pass).  The squared
  // norm of the (scaled) query is summed on the way: the DIM/4 pieces of a query are held by consecutive lanes.
  {
    constexpr int C4 = DIM / 4;                       // 16-B pieces per query: 16 or 32 (<= one warp)
    constexpr int kIters = (C4 * QTOT + kTcThreads - 1) / kTcThreads;
#pragma unroll 4
    for (int itq = 0; itq < kIters; ++itq) {
      const int i4 = tid + itq * kTcThreads;
      const bool in = i4 < C4 * QTOT;
      const int q = in ? i4 / C4 : 0, c4 = i4 - q * C4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (in && q < p.nq) v = __ldg(reinterpret_cast<const float4*>(p.Q + (size_t)q * DIM) + c4);
      if (scaled) { const float sc = s_scale[q]; v.x *= sc; v.y *= sc; v.z *= sc; v.w *= sc; }
      float ss = fmaf(v.x, v.x, fmaf(v.y, v.y, fmaf(v.z, v.z, v.w * v.w)));
#pragma unroll
      for (int off = 1; off < C4; off <<= 1) ss += __shfl_xor_sync(0xffffffffu, ss, off);   // lanes of one query only
      if (in) {
        const int dd = c4 * 4;
        if constexpr (BF) {
          const int sub = dd >> 6, e = dd & 63, ch = e >> 3;     // 4 bf16 = half of a 16-B chunk
          __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
          uint2 pk;
          pk.x = *reinterpret_cast<uint32_t*>(&lo);
          pk.y = *reinterpret_cast<uint32_t*>(&hi);
          *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(Qb) + (size_t)sub * (QTOT * 64) + q * 64 +
                                    ((ch ^ (q & 7)) << 3) + (e & 7)) = pk;
        } else {
          const int sub = dd >> 5, ch = (dd & 31) >> 2;          // 4 f32 = one 16-B chunk
          *reinterpret_cast<float4*>(reinterpret_cast<float*>(Qb) + (size_t)sub * (QTOT * 32) + q * 32 +
                                     ((ch ^ (q & 7)) << 2)) = v;
        }
        if (c4 == 0) s_ss[q] = ss;
      }
    }
  }
  __syncthreads();
  for (int q = tid; q < QTOT; q += kTcThreads) {
    float tf = __int_as_float(0x7F800000), nq2 = 0.f;
    if (q < p.nq && !kDense) {
      const uint64_t t = p.tau[q];
      tf = (t == 0) ? __int_as_float(0xFF800000) : key_score(t);
      // (the pairwise order of the sum differs from a serial one by a few ulp; the bound is inflated by 1e-4)
      nq2 = (sqrtf(s_ss[q]) * 1.0001f + 1e-30f) * kMargin;
      if (scaled) atomicMax(&s_cmax, __float_as_uint(nq2));  // non-negative floats order like their bit patterns
    }
    tq[q] = make_float2(tf, nq2);
    s_cnt[q] = 0;
  }
  // TILEMAX: per-query maxima of the current tile, two buffers of 64 ordered scores, over tq (2 x 64 words, unused here)
  uint32_t* tmax = reinterpret_cast<uint32_t*>(tq);
  if constexpr (kTileMax) {
    __syncthreads();   // the loop above wrote tq
    for (int i = tid; i < 2 * kQB; i += kTcThreads) tmax[i] = 0u;
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(kTmemCols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  // the generic-proxy writes of Qb must be visible to the tensor core (async proxy) before the first MMA
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ---------------------------------------------------------------- TMA producer
    if (lane == 0) {
      for (uint32_t i = n_pre; i < my_tiles; ++i) issue_tile(i, i * SPT);
      // every tile of this CTA is requested: the next kernel of the chain may start its prologue (it orders itself
      // behind this grid with pdl_wait)
      pdl_launch_dependents();
    }
  } else if (warp == 1) {
    // ---------------------------------------------------------------- MMA issuer
    if (lane == 0) {
      // D=f32, A=B=tf32 (format 2) or bf16 (format 1), both K-major, N=64, M=128
      constexpr uint32_t fmt = BF ? 1u : 2u;
      // One MMA covers every query block of the pass: N = QTOT.  With one N = 64 MMA per block (round 2's form) a tile at
      // 256 queries per pass took 32 (dim 64) / 64 (dim 128) instructions of 32 tensor-pipe cycles each, and the pass ran
      // at the rate the issuing thread could feed them: tensor pipe 24 % / 51 % busy, DRAM 24 % / 49 % (ncu r3d), the
      // loads themselves good for 7 TB/s at this ring depth (tools/ubench_tma_stream.cu).
      constexpr uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(QTOT >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      constexpr int kSubs = BF ? DIM / 64 : 2;           // 128-B-wide sub-tiles per stage
      constexpr uint32_t kQSubBytes = (uint32_t)QTOT * 128u;   // one 128-B-wide K slice of all queries
      const uint32_t qb_addr = smem_u32(Qb);
      uint32_t it = 0;
      if constexpr (kHalfBars) {
        static_assert(!kHalfBars || kHalfStage, "per-half accumulator barriers come with half-tile stages");
        for (uint32_t i = 0; i < my_tiles; ++i) {
#pragma unroll
          for (int half = 0; half < 2; ++half, ++it) {
            const uint32_t s = it % kTcStages, ph = (it / kTcStages) & 1u;
            mbar_wait(&full[s], ph);                        // this half's 128 rows (and, with half 0, the tile's norms)
            tc_fence_after();
            const uint32_t st_addr = smem_u32(stage_base + (size_t)s * kStageB);
            mbar_wait(&tempty[half], (i & 1u) ^ 1u);        // this half's four epilogue warps have drained the previous tile
            tc_fence_after();
            const uint32_t d_addr = tmem_base + (uint32_t)half * (uint32_t)QTOT;
#pragma unroll
            for (int sub = 0; sub < kSubs; ++sub) {
              const uint64_t adesc = umma_desc_k_sw128(st_addr + (uint32_t)sub * (kStageRows * 128));
              const uint64_t bdesc = umma_desc_k_sw128(qb_addr + (uint32_t)sub * kQSubBytes);
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_bf16(d_addr, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (sub | k) != 0 ? 1u : 0u);
            }
            umma_commit(&tfull[half]);
            umma_commit(&empty[s]);
          }
        }
      } else {
      for (uint32_t i = 0; i < my_tiles; ++i) {
        const uint32_t buf = (NBUF == 2) ? (i & 1u) : 0u;
        const uint32_t use = (NBUF == 2) ? (i >> 1) : i;  // how often this buffer has been used before
        mbar_wait(&tempty[buf], (use & 1u) ^ 1u);         // epilogue has drained this accumulator buffer
        tc_fence_after();
        for (int h = 0; h < KH; ++h, ++it) {
          const uint32_t s = it % kTcStages, ph = (it / kTcStages) & 1u;
          mbar_wait(&full[s], ph);
          tc_fence_after();
          const uint32_t st_addr = smem_u32(stage_base + (size_t)s * kStageB);
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            const uint32_t d_addr = tmem_base + (buf * 2u + (uint32_t)half) * (uint32_t)QTOT;
#pragma unroll
            for (int sub = 0; sub < kSubs; ++sub) {
              // a sub-tile is 256 rows of 128 B; rows [128*half, +128) start 16 KiB in; the matching K slice of the
              // queries is QTOT rows of 128 B
              const uint64_t adesc = umma_desc_k_sw128(st_addr + (uint32_t)sub * (kTileRows * 128) + (uint32_t)half * (128 * 128));
              const uint64_t bdesc = umma_desc_k_sw128(qb_addr + (uint32_t)(h * kSubs + sub) * kQSubBytes);
#pragma unroll
              for (int k = 0; k < 4; ++k) {  // UMMA_K = 32 B (8 tf32 / 16 bf16) -> +2 in 16-B units
                if constexpr (BF)
                  umma_bf16(d_addr, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (sub | k) != 0 ? 1u : 0u);
                else
                  umma_tf32(d_addr, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (h | sub | k) != 0 ? 1u : 0u);
              }
            }
          }
          umma_commit(&empty[s]);
        }
        umma_commit(&tfull[buf]);
      }
      }
    }
  } else {
    // ---------------------------------------------------------------- epilogue: 8 warps, 256 rows
    const int ew = warp - 2;
    const int half = ew >> 2, quarter = warp & 3;  // TMEM lanes [32*(warp%4), +32) of half-tile `half`
    const int row_local = half * 128 + quarter * 32 + lane;
    auto tile_row = [&](uint32_t i) -> uint64_t {
      return (uint64_t)(blockIdx.x + i * gridDim.x) * p.tile_stride * kTileRows + (uint64_t)row_local;
    };
    const uint32_t seg_stride = gridDim.x * p.seg_cap, seg_base = blockIdx.x * p.seg_cap;
    const float cmax = __uint_as_float(s_cmax);
    // the tile loop is instantiated once per threshold form (the choice is uniform for the launch), so that neither
    // form pays registers for the other
    auto run = [&](auto scaled_tag, auto group_tag) {
    constexpr bool kScaled = decltype(scaled_tag)::value;
    // GROUP mode (ScanParams::grp_rows): a survivor is appended once per group of 16 queries — a shared-memory atomic and a
    // store per (row, group) — instead of walking the group's 16 accumulators again to append it per query.  With 128 / 256
    // queries per pass the per-query walk (entered by 60 % of the warps for every block of 64 queries: one lane with a
    // survivor is enough) made the epilogue, not HBM or the MMAs, the limit of the pass (~200 instructions per warp and
    // block; ~60 here).  The exact re-score (recall.cu rescore_group_kernel) sorts out which of the 16 queries it was.
    constexpr bool kGroup = decltype(group_tag)::value && !kDense;
    const uint32_t grp_stride = gridDim.x * p.grp_cap;                       // slots between two groups' lists
    uint32_t* const grp_base = p.grp_rows + (size_t)blockIdx.x * p.grp_cap;  // this CTA's list of group 0
    auto append_groups = [&](int blk, uint32_t* scb, uint32_t grow, bool a0, bool a1, bool a2, bool a3) {
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        if (g == 0 ? a0 : g == 1 ? a1 : g == 2 ? a2 : a3) {
          const uint32_t pos = atomicAdd(&scb[g], 1u);
          if (pos < p.grp_cap) grp_base[(uint32_t)(blk * 4 + g) * grp_stride + pos] = grow;   // 32-bit offsets (< 2^32 slots)
        }
      }
    };
    for (uint32_t i = 0; i < my_tiles; ++i) {
      const uint32_t buf = (NBUF == 2) ? (i & 1u) : 0u;          // TMEM buffer (column offset)
      const uint32_t use = (NBUF == 2) ? (i >> 1) : i;
      const uint32_t bar = kHalfBars ? (uint32_t)half : buf;     // barrier pair: per half-tile, or per buffer
      const uint64_t lrow = tile_row(i);
      const bool valid = lrow < p.n_rows;
      const uint32_t grow = (uint32_t)(p.row_base + lrow);
      mbar_wait(&tfull[bar], use & 1u);
      tc_fence_after();
      // The norms landed before the tile's `full` barrier completed, which the MMA thread observed before it issued the
      // MMAs whose commit completes `tfull`: reading them after the tfull wait is ordered behind the copy.  (Waiting on
      // `full` itself from here would be wrong: that barrier may already be a whole phase further, the epilogue being
      // up to two tiles behind the MMA warp.)
      const float nr = nrm[(size_t)(i % kNormRing) * kTileRows + row_local];
      if constexpr (kGroup && kScaled) {
        // GROUP mode, uniform threshold — the form every well-scaled batch takes: the blocks' accumulators go through two
        // register buffers (block b + 1 travels from tensor memory while block b is reduced to its four group maxima), the
        // survivor test leaves one bit per (block, group), and the appends of the whole tile are made at the end: their
        // shared-memory atomics are independent and go out back to back instead of one latency after the other per block.
        const float t_r = fmaf(-nr, cmax, 1.0f - 1e-6f);
        const uint32_t tcol = tmem_base + ((uint32_t)(quarter * 32) << 16) + (buf * 2u + (uint32_t)half) * (uint32_t)QTOT;
        uint32_t a0[32], a1[32], b0[32], b1[32];
        uint32_t hits = 0u;
        tmem_ld32_nowait(tcol, a0);
        tmem_ld32_nowait(tcol + 32u, a1);
        tmem_ld_wait32(a0);
        tmem_ld_wait32(a1);
#pragma unroll
        for (int blk = 0; blk < NQB; ++blk) {
          uint32_t (&c0)[32] = (blk & 1) ? b0 : a0;
          uint32_t (&c1)[32] = (blk & 1) ? b1 : a1;
          uint32_t (&n0)[32] = (blk & 1) ? a0 : b0;
          uint32_t (&n1)[32] = (blk & 1) ? a1 : b1;
          if (blk + 1 < NQB) {
            tmem_ld32_nowait(tcol + (uint32_t)(blk + 1) * 64u, n0);
            tmem_ld32_nowait(tcol + (uint32_t)(blk + 1) * 64u + 32u, n1);
          } else {             // every accumulator of this tile is in registers: the MMA warp may overwrite the buffer
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[bar]);
          }
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            const uint32_t* v = g < 2 ? c0 : c1;
            const int o = (g & 1) * 16;
            float a = fmax3(__uint_as_float(v[o]), __uint_as_float(v[o + 1]), __uint_as_float(v[o + 2]));
#pragma unroll
            for (int j = 3; j < 15; j += 2) a = fmax3(a, __uint_as_float(v[o + j]), __uint_as_float(v[o + j + 1]));
            a = fmaxf(a, __uint_as_float(v[o + 15]));
            hits |= !(a < t_r) ? (1u << (blk * 4 + g)) : 0u;   // (NaN maxima and t_r = -inf / NaN survive, as above)
          }
          if (blk + 1 < NQB) {
            tmem_ld_wait32(n0);
            tmem_ld_wait32(n1);
          }
        }
        if (valid && hits) {
          uint32_t pos[NQB * 4];
#pragma unroll
          for (int b = 0; b < NQB * 4; ++b)
            pos[b] = (hits >> b & 1u) ? atomicAdd(&s_cnt[(b >> 2) * kQB + (b & 3)], 1u) : 0xFFFFFFFFu;
#pragma unroll
          for (int b = 0; b < NQB * 4; ++b)
            if (pos[b] < p.grp_cap) grp_base[(uint32_t)b * grp_stride + pos[b]] = grow;
        }
        continue;
      }
#pragma unroll 1
      for (int blk = 0; blk < NQB; ++blk) {
        const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (buf * 2u + (uint32_t)half) * (uint32_t)QTOT +
                               (uint32_t)blk * 64u;
        uint32_t v0[32], v1[32];
        tmem_ld32_nowait(taddr, v0);
        tmem_ld32_nowait(taddr + 32u, v1);
        tmem_ld_wait();
        if (blk == NQB - 1) {  // last accumulator read of this tile: the MMA warp may overwrite the buffer
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tempty[bar]);
        }
        const int nqb = p.nq - blk * kQB;  // queries of this block (may exceed 64; <= 0 for an unused block)
        uint32_t* scb = s_cnt + blk * kQB;
        const uint32_t qoff = (uint32_t)(blk * kQB);
        if constexpr (kTileMax) {
          // Sample for the threshold estimate, reduced on the spot: only the MAXIMUM score of each (tile, query) leaves
          // the kernel (recall.cu: the r-th largest tile maximum is the threshold) — 4 bytes per tile and query instead
          // of 8 per row and query.  Warp maximum per query (REDUX on the ordered score), lane q % 32 keeps query q's.
          static_assert(!kTileMax || NQB == 1, "TILEMAX is launched with one query block per grid row");
          uint32_t keep0 = 0u, keep1 = 0u;
#pragma unroll
          for (int q = 0; q < 32; ++q) {
            const uint32_t o0 = __reduce_max_sync(0xffffffffu, valid ? f32_ord(__uint_as_float(v0[q])) : 0u);
            const uint32_t o1 = __reduce_max_sync(0xffffffffu, valid ? f32_ord(__uint_as_float(v1[q])) : 0u);
            if (lane == q) { keep0 = o0; keep1 = o1; }
          }
          uint32_t* tm = tmax + (i & 1u) * kQB;
          atomicMax(&tm[lane], keep0);
          atomicMax(&tm[32 + lane], keep1);
          asm volatile("bar.sync 1, %0;" ::"n"(kTcEpiWarps * 32) : "memory");   // the 8 epilogue warps: tile complete
          if (ew == 0) {
            const uint32_t tile = blockIdx.x + i * gridDim.x;      // launch tile index (< p.n_tiles)
            uint32_t* out = reinterpret_cast<uint32_t*>(p_in.dense);
            const uint32_t qg0 = (uint32_t)blockIdx.y * kQB;
#pragma unroll
            for (int hq = 0; hq < 2; ++hq) {
              const int q = hq * 32 + lane;
              if (q < nqb) out[(size_t)(qg0 + (uint32_t)q) * p.dense_stride + tile] = tm[q];
              tm[q] = 0u;   // free for tile i + 2: the other warps get there only through the barrier of tile i + 1
            }
          }
        } else if constexpr (kDense) {
          // lanes of a warp hold consecutive rows: for a fixed query the 32 keys are one contiguous 256-B store
          const uint64_t slot = (uint64_t)(blockIdx.x + i * gridDim.x) * kTileRows + (uint64_t)row_local;
#pragma unroll
          for (int q = 0; q < 64; ++q) {
            if (q < nqb) {
              const float sc = __uint_as_float(q < 32 ? v0[q & 31] : v1[q & 31]);
              p.dense[(size_t)(qoff + (uint32_t)q) * p.dense_stride + slot] = valid ? make_key(sc, grow) : 0ull;
            }
          }
        } else if constexpr (kScaled) {
          // one threshold for the whole row: running maxima over four groups of 16 queries (NaN accumulators are
          // ignored by max; a row whose norm bound is +inf has t_r = -inf or NaN and always survives)
          const float t_r = fmaf(-nr, cmax, 1.0f - 1e-6f);
          float m[4];
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            const uint32_t* v = g < 2 ? v0 : v1;
            const int o = (g & 1) * 16;
            float a = fmax3(__uint_as_float(v[o]), __uint_as_float(v[o + 1]), __uint_as_float(v[o + 2]));
#pragma unroll
            for (int j = 3; j < 15; j += 2) a = fmax3(a, __uint_as_float(v[o + j]), __uint_as_float(v[o + j + 1]));
            m[g] = fmaxf(a, __uint_as_float(v[o + 15]));
          }
          const bool any0 = !(m[0] < t_r), any1 = !(m[1] < t_r), any2 = !(m[2] < t_r), any3 = !(m[3] < t_r);
          if constexpr (kGroup) {
            if (valid && (any0 || any1 || any2 || any3)) append_groups(blk, scb, grow, any0, any1, any2, any3);
          } else
          if (valid && (any0 || any1 || any2 || any3)) {
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              if (g == 0 ? any0 : g == 1 ? any1 : g == 2 ? any2 : any3) {
                // only the lanes that own a survivor get here (typically one per warp): collect the group's survivors
                // in a bit mask and append them one by one — a per-query loop over all 16 would be executed by the
                // whole warp with a (compiler-aggregated) shared-memory atomic per query
                uint32_t mask = 0;
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                  const float sc = __uint_as_float(g < 2 ? v0[(g & 1) * 16 + j] : v1[(g & 1) * 16 + j]);
                  if (!(sc < t_r)) mask |= 1u << j;
                }
                const int left = nqb - g * 16;   // real queries in this group
                if (left < 16) mask &= left > 0 ? ((1u << left) - 1u) : 0u;
                while (mask) {
                  const int q = g * 16 + __ffs(mask) - 1;
                  mask &= mask - 1;
                  const uint32_t pos = atomicAdd(&scb[q], 1u);
                  if (pos < p.seg_cap) p.cand_rows[(qoff + (uint32_t)q) * seg_stride + seg_base + pos] = grow;
                }
              }
            }
          }
        } else {
          // thr_q = tau_f[q] - ||x_row|| * c*||q||; {tau_f, c||q||} pairs are re-read from shared memory per tile (the
          // mbarrier wait above is a compiler memory barrier, so nothing is hoisted into 128 live registers)
          const float2* tqb = tq + blk * kQB;
          const float4* tq4 = reinterpret_cast<const float4*>(tqb);
          bool any[4] = {false, false, false, false};  // queries [0,16) [16,32) [32,48) [48,64)
#pragma unroll
          for (int q = 0; q < 32; q += 2) {
            const float4 a4 = tq4[q >> 1];
            const float4 b4 = tq4[16 + (q >> 1)];
            any[q >> 4] |= !(__uint_as_float(v0[q]) < fmaf(-nr, a4.y, a4.x));
            any[q >> 4] |= !(__uint_as_float(v0[q + 1]) < fmaf(-nr, a4.w, a4.z));
            any[2 + (q >> 4)] |= !(__uint_as_float(v1[q]) < fmaf(-nr, b4.y, b4.x));
            any[2 + (q >> 4)] |= !(__uint_as_float(v1[q + 1]) < fmaf(-nr, b4.w, b4.z));
          }
          if constexpr (kGroup) {
            if (valid && (any[0] || any[1] || any[2] || any[3])) append_groups(blk, scb, grow, any[0], any[1], any[2], any[3]);
          } else
          if (valid && (any[0] || any[1] || any[2] || any[3])) {
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              if (any[g]) {
                uint32_t mask = 0;
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                  const int q = g * 16 + j;
                  const float sc = __uint_as_float(g < 2 ? v0[q & 31] : v1[q & 31]);
                  if (!(sc < fmaf(-nr, tqb[q].y, tqb[q].x))) mask |= 1u << j;
                }
                const int left = nqb - g * 16;
                if (left < 16) mask &= left > 0 ? ((1u << left) - 1u) : 0u;
                while (mask) {
                  const int q = g * 16 + __ffs(mask) - 1;
                  mask &= mask - 1;
                  const uint32_t pos = atomicAdd(&scb[q], 1u);
                  if (pos < p.seg_cap) p.cand_rows[(qoff + (uint32_t)q) * seg_stride + seg_base + pos] = grow;
                }
              }
            }
          }
        }
      }
    }
    };
    constexpr bool grouped = !kDense && GROUP;   // (an instantiation of its own: the per-query form keeps its registers)
    if (scaled) run(std::true_type{}, std::integral_constant<bool, grouped>{});
    else run(std::false_type{}, std::integral_constant<bool, grouped>{});
    asm volatile("bar.sync 1, %0;" ::"n"(kTcEpiWarps * 32) : "memory");
    if (grouped) {   // every group of the pass gets its length (0 for the padding groups of the last block)
      for (int i = tid - 64; i < NQB * 4; i += kTcEpiWarps * 32)
        p.grp_cnt[(size_t)i * gridDim.x + blockIdx.x] = s_cnt[(i >> 2) * kQB + (i & 3)];
    } else if (!kDense)
      for (int q = tid - 64; q < p.nq; q += kTcEpiWarps * 32) p.seg_cnt[(size_t)q * gridDim.x + blockIdx.x] = s_cnt[q];
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kTmemCols));
  }
}

// upper bounds of the row norms (sqrt in fp64, inflated, rounded up to f32, plus an absolute term that covers subnormal
// elements) and, for the bf16 filter, the round-to-nearest bf16 shadow of the row
__global__ void row_norm_kernel(const float* __restrict__ E, uint64_t rows, int dim, float* __restrict__ out,
                                __nv_bfloat16* __restrict__ shadow) {
  const uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  const float4* x = reinterpret_cast<const float4*>(E + r * dim);
  double ss = 0.0;
  for (int i = 0; i < dim / 4; ++i) {
    const float4 v = x[i];
    ss += (double)v.x * v.x + (double)v.y * v.y + (double)v.z * v.z + (double)v.w * v.w;
    if (shadow) {
      __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
      uint2 pk;
      pk.x = *reinterpret_cast<uint32_t*>(&lo);
      pk.y = *reinterpret_cast<uint32_t*>(&hi);
      reinterpret_cast<uint2*>(shadow + r * dim)[i] = pk;
    }
  }
  float f = (float)(sqrt(ss) * (1.0 + 1e-6) + 1e-30);
  f = (f == f && f < __int_as_float(0x7F800000)) ? __uint_as_float(__float_as_uint(f) + 1u) : __int_as_float(0x7F800000);
  out[r] = f;  // NaN / inf rows get +inf: they always survive the filter and are settled by the exact re-score
}

// Built lazily by the first recall after prg_set_item_matrix (which releases row_norm): norms, and the bf16 shadow +
// its tensor map unless the tf32 filter is configured.
int build_row_norms(prg_handle* h) {
  const bool bf = h->scan_filter == SCAN_FILTER_BF16;
  h->E16_map_ok = false;
  if (bf) PRG_TRY(h->E16.ensure((size_t)h->E_rows * h->E_dim * 2));
  const size_t padded = (size_t)((h->E_rows + kTileRows - 1) / kTileRows) * kTileRows;
  PRG_TRY(h->row_norm.ensure(padded * 4));
  PRG_CUDA(cudaMemsetAsync(h->row_norm.p, 0, padded * 4, h->stream));   // whole tiles: the scan copies 256 norms per tile
  const unsigned grid = (unsigned)((h->E_rows + 255) / 256);
  row_norm_kernel<<<grid, 256, 0, h->stream>>>(h->E, h->E_rows, (int)h->E_dim, (float*)h->row_norm.p,
                                               bf ? (__nv_bfloat16*)h->E16.p : nullptr);
  PRG_CUDA(cudaGetLastError());
  PRG_CUDA(cudaStreamSynchronize(h->stream));
  count_launch(h);
  if (bf) {
    PFN_encodeTiled enc = get_encode_tiled();
    if (!enc) return fail(PRG_ECUDA, "cuTensorMapEncodeTiled entry point not available");
    cuuint64_t gdim[2] = {(cuuint64_t)h->E_dim, (cuuint64_t)h->E_rows};
    cuuint64_t gstride[1] = {(cuuint64_t)h->E_dim * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)kTileRows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(&h->E16_map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, h->E16.p, gdim, gstride, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(PRG_ECUDA, "cuTensorMapEncodeTiled (bf16 shadow) failed: " + std::to_string((int)r));
    box[1] = (cuuint32_t)(kTileRows / 2);   // half-tile boxes: the 256-queries-per-pass kernels' stages (tc_half_stages)
    r = enc(&h->E16_map_h, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, h->E16.p, gdim, gstride, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(PRG_ECUDA, "cuTensorMapEncodeTiled (bf16 shadow, half tiles) failed: " + std::to_string((int)r));
    h->E16_map_ok = true;
    PRG_TRY(build_i8_index(h));   // dim 64: the int8 index of the <= 64-query pass (recall_i8.cu)
  }
  return PRG_OK;
}

template <int DIM, int NQB, bool BF, bool GROUP>
static int launch_tc_g(prg_handle* h, const ScanParams& p) {
  const size_t smem = scan_tc_smem_bytes<DIM, NQB, BF>();
  PRG_CUDA(cudaFuncSetAttribute(recall_scan_tc_kernel<DIM, NQB, BF, SCAN_THRESH, GROUP>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)smem));
  if (p.n_tiles == 0) return PRG_OK;
  if (BF && !h->E16_map_ok) return fail(PRG_ESTATE, "bf16 filter index not built");
  StageScope span(h, ST_SCAN);
  const unsigned grid = p.n_tiles < (uint32_t)h->sm_count ? p.n_tiles : (unsigned)h->sm_count;
  PRG_CUDA(launch_chained(h, recall_scan_tc_kernel<DIM, NQB, BF, SCAN_THRESH, GROUP>, dim3(grid), dim3(kTcThreads), smem, 1,
                          BF ? (tc_half_stages<NQB, BF>() ? h->E16_map_h : h->E16_map) : h->E_map, p));
  count_launch(h);
  return PRG_OK;
}
template <int DIM, int NQB, bool BF>
static int launch_tc(prg_handle* h, const ScanParams& p) {
  if constexpr (BF) {   // GROUP mode exists over the bf16 index only (recall.cu asks for it there)
    if (p.grp_rows) return launch_tc_g<DIM, NQB, BF, true>(h, p);
  } else {
    if (p.grp_rows) return fail(PRG_EINVAL, "launch_scan_tc: group mode needs the bf16 filter index");
  }
  return launch_tc_g<DIM, NQB, BF, false>(h, p);
}

// sample scoring on the tensor cores (bf16 index): any number of queries per launch, one grid row per block of 64
template <int DIM>
static int launch_tc_dense(prg_handle* h, const ScanParams& p) {
  const size_t smem = scan_tc_smem_bytes<DIM, 1, true>();
  PRG_CUDA(cudaFuncSetAttribute(recall_scan_tc_kernel<DIM, 1, true, SCAN_DENSE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)smem));
  if (p.n_tiles == 0 || p.nq <= 0) return PRG_OK;
  StageScope span(h, ST_SCAN_DENSE);
  const unsigned gx = p.n_tiles < (uint32_t)h->sm_count ? p.n_tiles : (unsigned)h->sm_count;
  const unsigned gy = (unsigned)((p.nq + kQB - 1) / kQB);
  if (gy > 65535u) return fail(PRG_EINVAL, "launch_scan_tc_dense: too many queries");
  PRG_CUDA(launch_chained(h, recall_scan_tc_kernel<DIM, 1, true, SCAN_DENSE>, dim3(gx, gy), dim3(kTcThreads), smem, 1,
                          h->E16_map, p));
  count_launch(h);
  return PRG_OK;
}

// config "recall_tilemax": sample scoring that emits one maximum per (tile, query):
// p.dense = u32 [nq][dense_stride], slot = launch tile
template <int DIM>
static int launch_tc_tilemax(prg_handle* h, const ScanParams& p) {
  const size_t smem = scan_tc_smem_bytes<DIM, 1, true>();
  PRG_CUDA(cudaFuncSetAttribute(recall_scan_tc_kernel<DIM, 1, true, SCAN_TILEMAX>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)smem));
  if (p.n_tiles == 0 || p.nq <= 0) return PRG_OK;
  StageScope span(h, ST_SCAN_DENSE);
  const unsigned gx = p.n_tiles < (uint32_t)h->sm_count ? p.n_tiles : (unsigned)h->sm_count;
  const unsigned gy = (unsigned)((p.nq + kQB - 1) / kQB);
  if (gy > 65535u) return fail(PRG_EINVAL, "launch_scan_tc_tilemax: too many queries");
  PRG_CUDA(launch_chained(h, recall_scan_tc_kernel<DIM, 1, true, SCAN_TILEMAX>, dim3(gx, gy), dim3(kTcThreads), smem, 1,
                          h->E16_map, p));
  count_launch(h);
  return PRG_OK;
}
int launch_scan_tc_tilemax(prg_handle* h, const ScanParams& p) {
  if (!scan_tc_dense_available(h)) return fail(PRG_ESTATE, "bf16 filter index not built");
  if (h->E_dim == 64) return launch_tc_tilemax<64>(h, p);
  if (h->E_dim == 128) return launch_tc_tilemax<128>(h, p);
  return fail(PRG_EINVAL, "launch_scan_tc_tilemax: unsupported shape");
}

bool scan_tc_dense_available(const prg_handle* h) { return h->scan_filter == SCAN_FILTER_BF16 && h->E16_map_ok; }

int launch_scan_tc_dense(prg_handle* h, const ScanParams& p) {
  if (!scan_tc_dense_available(h)) return fail(PRG_ESTATE, "bf16 filter index not built");
  if (h->E_dim == 64) return launch_tc_dense<64>(h, p);
  if (h->E_dim == 128) return launch_tc_dense<128>(h, p);
  return fail(PRG_EINVAL, "launch_scan_tc_dense: unsupported shape");
}

// queries per pass the kernel is built for: 64/128/256 at dim 64, 64 at dim 128 (shared-memory budget)
// config "scan128_nqb" (default on): 128/256 queries per pass at dim 128 over the bf16 index with a
// 2-stage ring (C5: 1024 queries per shard are 4 passes over the index instead of 16)
int scan_tc_max_queries(const prg_handle* h) {
  if (h->E_dim == 64) return 256;
  return (h->scan128_nqb && h->scan_filter == SCAN_FILTER_BF16) ? 256 : 64;
}

template <bool BF>
static int launch_scan_tc_t(prg_handle* h, const ScanParams& p) {
  if (h->E_dim == 64) {
    if (p.nq <= 64) return launch_tc<64, 1, BF>(h, p);
    if (p.nq <= 128) return launch_tc<64, 2, BF>(h, p);
    if (p.nq <= 256) return launch_tc<64, 4, BF>(h, p);
    return fail(PRG_EINVAL, "launch_scan_tc: more than 256 queries per pass");
  }
  if (h->E_dim == 128) {
    if (p.nq <= 64) return launch_tc<128, 1, BF>(h, p);
    if constexpr (BF) {
      if (h->scan128_nqb && p.nq <= 128) return launch_tc<128, 2, true>(h, p);
      if (h->scan128_nqb && p.nq <= 256) return launch_tc<128, 4, true>(h, p);
    }
    return fail(PRG_EINVAL, "launch_scan_tc: more than 64 queries per pass at dim 128");
  }
  return fail(PRG_EUNSUPPORTED, "item matrix dim must be 64 or 128");
}

int launch_scan_tc(prg_handle* h, const ScanParams& p) {
  return h->scan_filter == SCAN_FILTER_BF16 ? launch_scan_tc_t<true>(h, p) : launch_scan_tc_t<false>(h, p);
}

}  // namespace prg
