// mlp.cu — dense tower of the rank model on the 5th-gen tensor cores.
//
// Replaces the remote EasyRec / TF-Serving DNN behind algorithm/eas/easyrec_request.go:20-73 and
// algorithm/tfserving/client.go:77-115 (wire contract: one score per item, easyrec_response.go:35-70,
// tfserving/response.go:51-63).  This is the one GEMM-shaped stage of the path, so it is the one stage on
// tcgen05: per layer  out[M x N] = relu(A[M x K] * W^T + b)  with
//   A   = the activations as bf16 hi/lo pairs ("bf16x2": a = hi + lo, hi = bf16(a), lo = bf16(a - hi)), stored
//         [M][2K] (hi | lo), so one layer is 2*K/64 k-blocks against the same W tile,
//   W   = bf16 [N][K] (K-major B operand), fp32 accumulation in TMEM.
// The hi/lo split is what makes the result reproducible to 1e-5 against oracle/oracle.c orc_mlp_forward: with plain
// bf16 activations a 1e-6 accumulation-order difference flips whole bf16 ulps of hidden units.
//
// Default kernel (mlp_layer_persistent_kernel<BN, FINAL, PAIR = true>): persistent 2-CTA clusters, one per SM pair, each
// computing 256 x BN tiles with tcgen05.mma.cta_group::2 (M = 256, N = BN, K = 16): warp 0 = TMA producer (SWIZZLE_128B
// boxes, 3-stage mbarrier ring; a stage = the hi and the lo block of a 64-wide K slice + this CTA's half of the W block),
// warp 1 = TMEM allocator + (leader CTA only) the single MMA-issuing thread, warps 2-9 = epilogue (tcgen05.ld
// 32x32b.x32 -> bias + ReLU -> hi/lo split -> shared-memory staging -> TMA tensor stores; for the last hidden layer the
// fused N = 1 output layer: logit = b + sum_j w_j * a_j in fp32).  Accumulators are double buffered in TMEM.
// Fallback: the same kernel with PAIR = false (single CTA, M = 128, direct stores) for an odd number of row blocks.
// Round 2:
//  * HILO = false for the FIRST layer: the tower input is bf16(x) alone — gathered table values carry no
//    accumulation-order noise, so the single rounding is reproducible and layer 1 issues half the MMAs and reads half
//    the activation bytes; computed activations (layers 2+) keep the hi/lo pair.
//  * per-REQUEST first-layer bias (`ubias`): the user / context features of a request (service/rank/algo_data.go:104-118)
//    are the same for all of its candidates, so their share of layer 1, b + W1[:, user columns] * x_user, is computed
//    once per request (gather_fm.cu user_prefix_kernel) and enters here as the bias row of the candidate's request.
//  * NOUT output heads (easyrec_response.go:35-70 score maps / tfserving/response.go:51-63 rows of Outputs): the fused
//    last layer evaluates up to 4 heads; Item.Score = sum_o coef_o * score_o (the RankScore expression,
//    service/rank/rank_service.go:339-363, for the sums of products the device accepts).
#include "handle.h"
#include <cuda_bf16.h>
#include <cstring>
#include <vector>

namespace prg {

constexpr int kMlpBM = 128;
constexpr int kMlpBK = 64;       // bf16 elements per k-block = one 128-B swizzle row
constexpr int kMlpMaxOut = 4;    // output heads of the fused last layer

struct MlpLayerParams {
  int M;                 // valid rows
  int K;                 // input width (the A tensor has 2K columns)
  int N;                 // output width
  const float* bias;     // [N]
  uint16_t* out;         // [Mp][2N] bf16 (hi | lo)                 (hidden layers)
  const float* w_last;   // [n_out][N] f32 (bf16 values widened)    (FINAL)
  float b_last[kMlpMaxOut];  //                                     (FINAL)
  int n_out;             // output heads, 1..kMlpMaxOut             (FINAL)
  float* logit_out;      // [M][n_out]                              (FINAL)
  // per-request first-layer bias (nullable): row i takes ubias[(i / rows_per_req)][N] instead of bias[N]
  const float* ubias;
  int rows_per_req;
  // fused score epilogue (FINAL, optional): score = rows[i] == pad ? 0 : (double)(float)sigmoid((fm_logit[i] +) logit)
  const float* fm_logit;  // nullable: logit of the FM part, added first (DeepFM-shaped model)
  const uint32_t* rows;   // candidate rows (0xFFFFFFFF = padding)
  double* score_out;      // nullable: when set, scores are written instead of logits: sum_o coef[o] * score_o
  double* score_map;      // nullable: [M][n_out] every head's score (AlgoResponse.GetScoreMap())
  double coef[kMlpMaxOut];
};

// ------------------------------------------------------------------ persistent variant (used by mlp_forward_device)
// One CTA per SM loops over output tiles; the accumulator is double buffered in TMEM (2 x BN columns) so the epilogue
// of tile i (8 warps: two per TMEM lane quarter, each taking half of the columns) overlaps the MMAs of tile i+1, and
// the per-CTA prologue (TMEM alloc, barrier init, first TMA round trip) is paid once per SM instead of once per tile.
// ncu r1 of the one-tile-per-CTA kernel above: tensor pipe 30 % active — MMA and epilogue were serialised.
// A stage holds the hi AND the lo block of the activations for one 64-wide slice of K together with the ONE W block
// both are multiplied with (a = hi + lo, so hi*W and lo*W use the same weights): per slice the CTA pulls
// 16 + 16 + BN/8 KiB instead of 2 x (16 + BN/8) KiB.  Layer 1 at 128 x 256 tiles ran at the L2 bandwidth limit
// (768 MB of operand reads in 92 us = 8.3 TB/s, tensor pipe 39 % active); this takes a third of that traffic away.
// last step of the tower for one candidate: the logit, or (fused path) the rank score of gather_fm.cu's
// logit_to_score_kernel — same operations in the same order, so the score is bit-identical to the unfused path
template <int NOUT>
__device__ __forceinline__ void mlp_write_result(const MlpLayerParams& p, int row, const float (&mlp_logit)[NOUT]) {
  if (p.score_out || p.score_map) {
    const bool pad = p.rows[row] == 0xFFFFFFFFu;
    double acc = 0.0;
#pragma unroll
    for (int o = 0; o < NOUT; ++o) {
      if (o < p.n_out) {
        // a DeepFM-shaped model adds the FM logit to the first head
        const float l = (o == 0 && p.fm_logit) ? __fadd_rn(p.fm_logit[row], mlp_logit[o]) : mlp_logit[o];
        const float sc = (float)(1.0 / (1.0 + exp(-(double)l)));
        const double v = pad ? 0.0 : (double)sc;
        if (p.score_map) p.score_map[(size_t)row * p.n_out + o] = v;
        const double t = __dmul_rn(p.coef[o], v);          // left-to-right fp64, no contraction (oracle orc_rank_score_expr)
        acc = o == 0 ? t : __dadd_rn(acc, t);
      }
    }
    if (p.score_out) p.score_out[row] = acc;
  } else {
#pragma unroll
    for (int o = 0; o < NOUT; ++o)
      if (o < p.n_out) p.logit_out[(size_t)row * p.n_out + o] = mlp_logit[o];
  }
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo_half, float hi_half) {
  const __nv_bfloat162 v = __floats2bfloat162_rn(lo_half, hi_half);   // .x (low 16 bits) = lo_half, .y = hi_half
  return *reinterpret_cast<const uint32_t*>(&v);
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa_cluster(uint32_t cta_smem_addr, uint32_t cta_rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(cta_smem_addr), "r"(cta_rank));
  return r;
}
// TMA 2-D load into OUR shared memory whose completion is signalled on an mbarrier of the CTA pair's leader
// (`bar_cluster_addr` is a shared::cluster address); .cta_group::2 is what allows the barrier to live in the peer
__device__ __forceinline__ void tma_load_2d_pair(void* dst, const CUtensorMap* map, int c0, int c1, uint32_t bar_cluster_addr,
                                                 uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint"
      " [%0], [%1, {%2, %3}], [%4], %5;" ::"r"(smem_u32(dst)),
      "l"(map), "r"(c0), "r"(c1), "r"(bar_cluster_addr), "l"(policy)
      : "memory");
}
// D[tmem, both CTAs] (+)= A[smem, 128 rows per CTA] * B[smem, half of the N rows per CTA]^T: one instruction, two SMs
__device__ __forceinline__ void umma_bf16_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
// arrives (once the pair MMAs issued so far have completed) on the barrier at this offset in both CTAs of the pair
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "h"((uint16_t)3)
               : "memory");
}
// TMA tensor store shared -> global (bulk async-group completion)
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, int c0, int c1, const void* src) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%1, %2}], [%3];" ::"l"(map), "r"(c0), "r"(c1),
               "r"(smem_u32(src))
               : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar_cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar_cluster_addr) : "memory");
}

constexpr int kMlpPThreads = 320;
constexpr int kMlpEpiWarps = 8;
template <bool HILO>
constexpr int mlp_p_stages() { return HILO ? 3 : 4; }
constexpr int kMlpOutStage = 8192;   // per epilogue warp: two buffers of (32 x 32 bf16 hi block | lo block) staged for the TMA stores

template <int BN, bool PAIR, bool HILO>
constexpr size_t mlp_p_smem_bytes() {
  return (size_t)mlp_p_stages<HILO>() * ((HILO ? 2 : 1) * kMlpBM * 128 + (PAIR ? BN / 2 : BN) * 128) +
         (PAIR ? (size_t)kMlpEpiWarps * kMlpOutStage : 0) + 2 * 1024 * 4 /*bias, w_last*/ +
         2 * 2 * kMlpMaxOut * kMlpBM * 4 /*partials*/ + (2 * mlp_p_stages<HILO>() + 4) * 8 + 16;
}

// PAIR: a 2-CTA cluster computes a 256 x BN tile with tcgen05.mma.cta_group::2 (M = 256): each CTA stages its own 128
// rows of activations and HALF of the W block (BN/2 rows); the leader's MMA thread issues one instruction for both SMs
// and each SM's accumulator (its 128 rows x BN columns) lands in its own TMEM.  A single-CTA M = 128 MMA reads
// (128 + BN) x 32 B of operands from shared memory per K = 16 step — at BN = 256 that is 12 KiB per 136 tensor-pipe
// cycles, 90 B/clk of the SM's 128 B/clk with the TMA writes of the next stage on top: the tensor pipe idled 60 % of
// the time.  The pair halves the W bytes each SM reads and stages.
// Protocol: every TMA load of both CTAs signals the LEADER's full[s] (.cta_group::2 barrier in the peer); the leader's
// commits are multicast to empty[s] / tfull[b] of both CTAs; the peer's epilogue warps arrive on the leader's tempty[b].
template <int BN, bool FINAL, bool PAIR, bool HILO, int NOUT>
__global__ void __launch_bounds__(kMlpPThreads, 1)
mlp_layer_persistent_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapW,
                            const __grid_constant__ CUtensorMap mapOut, const MlpLayerParams p, const int n_mblk) {
  extern __shared__ __align__(1024) uint8_t msm[];
  constexpr int kABytes = kMlpBM * 128, kBBytes = (PAIR ? BN / 2 : BN) * 128;   // W rows staged by this CTA
  constexpr int kAParts = HILO ? 2 : 1;
  constexpr int kStageBytes = kAParts * kABytes + kBBytes;                         // A_hi | (A_lo) | W
  constexpr int kMlpStages = mlp_p_stages<HILO>();
  constexpr uint32_t kTmemCols = 2 * BN <= 64 ? 64 : 2 * BN <= 128 ? 128 : 2 * BN <= 256 ? 256 : 512;
  uint8_t* out_stage = msm + (size_t)kMlpStages * kStageBytes;                       // PAIR: [8 warps][2 buffers][hi 2 KiB | lo 2 KiB]
  float* bias_s = reinterpret_cast<float*>(out_stage + (PAIR ? kMlpEpiWarps * kMlpOutStage : 0));  // [1024]
  float* wl_s = bias_s + 1024;                                                       // [1024]
  float* part_s = wl_s + 1024;                                                       // [2 buf][2 halves][kMlpMaxOut][128]
  uint64_t* full = reinterpret_cast<uint64_t*>(part_s + 4 * kMlpMaxOut * kMlpBM);
  uint64_t* empty = full + kMlpStages;
  uint64_t* tfull = empty + kMlpStages;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n_nblk = p.N / BN;
  uint32_t cta_rank = 0;
  if constexpr (PAIR) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(cta_rank));
  // work units: output tiles, or (PAIR) pairs of vertically adjacent tiles; unit u -> row block(s), column block
  const int unit0 = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int n_workers = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const int n_tiles = (PAIR ? n_mblk / 2 : n_mblk) * n_nblk;
  const int num_kb = p.K / kMlpBK;   // 64-wide slices of K; each brings its hi and its lo activations
  const int my_tiles = (n_tiles > unit0) ? (n_tiles - unit0 + n_workers - 1) / n_workers : 0;
  auto tile_of = [&](int i, int& m_blk, int& n_blk) {
    const int t = unit0 + i * n_workers;
    const int mu = t / n_nblk;
    n_blk = t - mu * n_nblk;
    m_blk = PAIR ? 2 * mu + (int)cta_rank : mu;
  };

  for (int i = tid; i < p.N; i += kMlpPThreads) bias_s[i] = p.bias ? p.bias[i] : 0.f;
  if constexpr (FINAL) {   // last hidden width <= 256 (prg_set_mlp): head o's weights at wl_s[o * 256 + column]
    for (int i = tid; i < NOUT * 256; i += kMlpPThreads) {
      const int o = i >> 8, c = i & 255;
      wl_s[i] = (o < p.n_out && c < p.N) ? p.w_last[(size_t)o * p.N + c] : 0.f;
    }
  }
  if (tid == 0) {
    tma_prefetch_desc(&mapA);
    tma_prefetch_desc(&mapW);
    for (int s = 0; s < kMlpStages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&tfull[b], 1); mbar_init(&tempty[b], PAIR ? 2 * kMlpEpiWarps : kMlpEpiWarps); }
    mbar_fence_init();
  }
  if (warp == 1) {
    if constexpr (PAIR) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                   "n"(kTmemCols));
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                   "n"(kTmemCols));
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (PAIR) cluster_sync_all();   // the peer's barriers are initialised before anything of ours can signal them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // chained launch: weights and biases above are static; the activations (and everything this layer writes) are
  // touched only from here on, when the previous kernel of the chain has completed
  pdl_wait();
  pdl_launch_dependents();

  if (warp == 0) {
    if (lane == 0) {
      uint32_t it = 0;
      for (int i = 0; i < my_tiles; ++i) {
        int m_blk, n_blk;
        tile_of(i, m_blk, n_blk);
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const uint32_t s = it % kMlpStages, ph = (it / kMlpStages) & 1u;
          mbar_wait(&empty[s], ph ^ 1u);
          uint8_t* a_dst = msm + (size_t)s * kStageBytes;
          if constexpr (PAIR) {
            // both CTAs' loads complete on the leader's full[s]; the leader expects the bytes of both
            if (cta_rank == 0) mbar_arrive_expect_tx(&full[s], 2 * kStageBytes);
            const uint32_t lbar = mapa_cluster(smem_u32(&full[s]), 0);
            tma_load_2d_pair(a_dst, &mapA, kb * kMlpBK, m_blk * kMlpBM, lbar, kEvictNormal);
            if constexpr (HILO) tma_load_2d_pair(a_dst + kABytes, &mapA, p.K + kb * kMlpBK, m_blk * kMlpBM, lbar, kEvictNormal);
            tma_load_2d_pair(a_dst + kAParts * kABytes, &mapW, kb * kMlpBK, n_blk * BN + (int)cta_rank * (BN / 2), lbar,
                             kEvictLast);   // our half of the W block (mapW boxes are BN/2 rows here)
          } else {
            mbar_arrive_expect_tx(&full[s], kStageBytes);
            tma_load_2d(a_dst, &mapA, kb * kMlpBK, m_blk * kMlpBM, &full[s], kEvictNormal);
            if constexpr (HILO) tma_load_2d(a_dst + kABytes, &mapA, p.K + kb * kMlpBK, m_blk * kMlpBM, &full[s], kEvictNormal);
            tma_load_2d(a_dst + kAParts * kABytes, &mapW, kb * kMlpBK, n_blk * BN, &full[s], kEvictLast);
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && cta_rank == 0) {   // PAIR: only the leader issues (for both SMs)
      constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) |
                                 ((uint32_t)((PAIR ? 2 * kMlpBM : kMlpBM) >> 4) << 24);
      uint32_t it = 0;
      for (int i = 0; i < my_tiles; ++i) {
        const uint32_t buf = (uint32_t)i & 1u;
        mbar_wait(&tempty[buf], (((uint32_t)i >> 1) & 1u) ^ 1u);
        tc_fence_after();
        const uint32_t d_addr = tmem_base + buf * (uint32_t)BN;
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const uint32_t s = it % kMlpStages, ph = (it / kMlpStages) & 1u;
          mbar_wait(&full[s], ph);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(msm + (size_t)s * kStageBytes);
          const uint64_t bdesc = umma_desc_k_sw128(a_addr + kAParts * kABytes);
#pragma unroll
          for (int part = 0; part < kAParts; ++part) {   // hi, then lo, against the same W block
            const uint64_t adesc = umma_desc_k_sw128(a_addr + (uint32_t)part * kABytes);
#pragma unroll
            for (int k = 0; k < kMlpBK / 16; ++k) {
              if constexpr (PAIR)
                umma_bf16_pair(d_addr, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (kb | part | k) != 0 ? 1u : 0u);
              else
                umma_bf16(d_addr, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (kb | part | k) != 0 ? 1u : 0u);
            }
          }
          if constexpr (PAIR) umma_commit_pair(&empty[s]);
          else umma_commit(&empty[s]);
        }
        if constexpr (PAIR) umma_commit_pair(&tfull[buf]);
        else umma_commit(&tfull[buf]);
      }
    }
  } else {
    const int ew = warp - 2;
    const int quarter = warp & 3, chalf = ew >> 2;
    constexpr int kHalf = BN / 2;
    for (int i = 0; i < my_tiles; ++i) {
      const uint32_t buf = (uint32_t)i & 1u;
      int m_blk, n_blk;
      tile_of(i, m_blk, n_blk);
      const int row = m_blk * kMlpBM + quarter * 32 + lane;
      // per-request bias row (user / context share of the first layer), or the layer's bias from shared memory
      const float* ub = nullptr;
      if (p.ubias) ub = p.ubias + (size_t)((row < p.M ? row : p.M - 1) / p.rows_per_req) * p.N;
      mbar_wait(&tfull[buf], ((uint32_t)i >> 1) & 1u);
      tc_fence_after();
      float logit[NOUT];
#pragma unroll
      for (int o = 0; o < NOUT; ++o) logit[o] = 0.f;
      // The accumulator chunks of this warp are read through two register buffers: the tcgen05.ld of chunk i + 1 is in
      // flight while chunk i goes through bias / ReLU / split / staging (with one buffer the two epilogue warps of a
      // scheduler spent most of their time waiting for the load: ncu r2d, layer 1 at 35 % tensor-pipe activity after
      // its MMA count had been halved)
      constexpr int kChunks = kHalf / 32;
      uint32_t va[32], vb[32];
      const uint32_t t_tile = tmem_base + ((uint32_t)(quarter * 32) << 16) + buf * (uint32_t)BN + (uint32_t)(chalf * kHalf);
      tmem_ld32_nowait(t_tile, va);
#pragma unroll
      for (int ci = 0; ci < kChunks; ++ci) {
        const int c0 = chalf * kHalf + ci * 32;
        uint32_t (&v)[32] = (ci & 1) ? vb : va;
        uint32_t (&vn)[32] = (ci & 1) ? va : vb;
        // the request's bias row for this chunk is requested BEFORE the wait for the accumulators, so the two latencies
        // overlap (ncu r2d: the add that consumed these loads was the top stall of layer 1, 14 % of all samples)
        float4 bq[8];
        if (ub) {
#pragma unroll
          for (int w = 0; w < 8; ++w) bq[w] = __ldg(reinterpret_cast<const float4*>(ub + n_blk * BN + c0) + w);
        }
        tmem_ld_wait32(v);
        if (ci + 1 < kChunks) {
          tmem_ld32_nowait(t_tile + (uint32_t)((ci + 1) * 32), vn);
        } else {  // last read of this tile by this warp: hand the buffer back
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if constexpr (PAIR) mbar_arrive_cluster(mapa_cluster(smem_u32(&tempty[buf]), 0));   // the leader's barrier
            else mbar_arrive(&tempty[buf]);
          }
        }
        const int nb = n_blk * BN + c0;  // column in the layer's output
        uint32_t hi_w[16], lo_w[16];
        if (ub) {   // uniform branch: add the request's bias row to the accumulators in place (L1-resident, mostly one
                    // address per warp); the shared-memory bias is then skipped below by adding 0
#pragma unroll
          for (int w = 0; w < 8; ++w) {
            const float4 t = bq[w];
            v[4 * w] = __float_as_uint(__fadd_rn(__uint_as_float(v[4 * w]), t.x));
            v[4 * w + 1] = __float_as_uint(__fadd_rn(__uint_as_float(v[4 * w + 1]), t.y));
            v[4 * w + 2] = __float_as_uint(__fadd_rn(__uint_as_float(v[4 * w + 2]), t.z));
            v[4 * w + 3] = __float_as_uint(__fadd_rn(__uint_as_float(v[4 * w + 3]), t.w));
          }
        }
#pragma unroll
        for (int c = 0; c < 32; c += 2) {
          float r0 = fmaxf(ub ? __uint_as_float(v[c]) : __fadd_rn(__uint_as_float(v[c]), bias_s[nb + c]), 0.f);
          float r1 = fmaxf(ub ? __uint_as_float(v[c + 1]) : __fadd_rn(__uint_as_float(v[c + 1]), bias_s[nb + c + 1]), 0.f);
          // packed conversions (cvt.rn.bf16x2.f32: one XU instruction per PAIR of values; the XU pipe was 32 % busy)
          const uint32_t hp = pack_bf16x2(r0, r1);                       // low half = bf16(r0), high half = bf16(r1)
          const float h0f = __uint_as_float(hp << 16), h1f = __uint_as_float(hp & 0xFFFF0000u);
          const uint32_t lp = pack_bf16x2(__fsub_rn(r0, h0f), __fsub_rn(r1, h1f));
          if (FINAL) {
            const float a0 = __fadd_rn(h0f, __uint_as_float(lp << 16));
            const float a1 = __fadd_rn(h1f, __uint_as_float(lp & 0xFFFF0000u));
#pragma unroll
            for (int o = 0; o < NOUT; ++o) {
              logit[o] = __fmaf_rn(wl_s[o * 256 + nb + c], a0, logit[o]);
              logit[o] = __fmaf_rn(wl_s[o * 256 + nb + c + 1], a1, logit[o]);
            }
          } else {
            hi_w[c >> 1] = hp;
            lo_w[c >> 1] = lp;
          }
        }
        if (!FINAL) {
          uint16_t* o = p.out + (size_t)row * (2 * p.N) + (size_t)nb;
          uint4* oh = reinterpret_cast<uint4*>(o);
          uint4* ol = reinterpret_cast<uint4*>(o + p.N);
          if constexpr (PAIR) {
            // The accumulator layout gives every lane one ROW: direct stores are 32 scattered 16-B pieces per
            // instruction, half a sector each (measured: the layer kernels spent 60 us of 160 in these stores).  The
            // 32 x 32 block goes through shared memory (SWIZZLE_64B, conflict free) and out as two TMA tensor stores.
            uint8_t* stg = out_stage + (size_t)ew * kMlpOutStage + ((c0 >> 5) & 1) * 4096;
            if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");   // the stores of two chunks ago have read it
            __syncwarp();
            const int sw = (lane >> 1) & 3;
#pragma unroll
            for (int w = 0; w < 4; ++w) {
              *reinterpret_cast<uint4*>(stg + lane * 64 + ((w ^ sw) << 4)) =
                  make_uint4(hi_w[4 * w], hi_w[4 * w + 1], hi_w[4 * w + 2], hi_w[4 * w + 3]);
              *reinterpret_cast<uint4*>(stg + 2048 + lane * 64 + ((w ^ sw) << 4)) =
                  make_uint4(lo_w[4 * w], lo_w[4 * w + 1], lo_w[4 * w + 2], lo_w[4 * w + 3]);
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) {
              const int r0 = m_blk * kMlpBM + quarter * 32;
              tma_store_2d(&mapOut, nb, r0, stg);
              tma_store_2d(&mapOut, p.N + nb, r0, stg + 2048);
              asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
          } else {
#pragma unroll
            for (int w = 0; w < 4; ++w) {
              oh[w] = make_uint4(hi_w[4 * w], hi_w[4 * w + 1], hi_w[4 * w + 2], hi_w[4 * w + 3]);
              ol[w] = make_uint4(lo_w[4 * w], lo_w[4 * w + 1], lo_w[4 * w + 2], lo_w[4 * w + 3]);
            }
          }
        }
      }
      if (FINAL) {  // logit = b + (columns of half 0) + (columns of half 1): the two warps of a lane quarter combine
#pragma unroll
        for (int o = 0; o < NOUT; ++o)
          part_s[((buf * 2 + chalf) * kMlpMaxOut + o) * kMlpBM + quarter * 32 + lane] = logit[o];
        asm volatile("bar.sync 1, %0;" ::"n"(kMlpEpiWarps * 32) : "memory");
        if (chalf == 0 && row < p.M) {
          float res[NOUT];
#pragma unroll
          for (int o = 0; o < NOUT; ++o)
            res[o] = __fadd_rn(__fadd_rn(part_s[((buf * 2) * kMlpMaxOut + o) * kMlpBM + quarter * 32 + lane],
                                         part_s[((buf * 2 + 1) * kMlpMaxOut + o) * kMlpBM + quarter * 32 + lane]), p.b_last[o]);
          mlp_write_result<NOUT>(p, row, res);
        }
      }
    }
  }

  if constexpr (PAIR && !FINAL) {   // the staged output blocks must have left shared memory before the CTA retires
    if (warp >= 2 && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (PAIR) cluster_sync_all();   // no CTA leaves (or frees TMEM) while its peer may still signal / use it
  if (warp == 1) {
    tc_fence_after();
    if constexpr (PAIR) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kTmemCols));
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kTmemCols));
  }
}

// ------------------------------------------------------------------ host side
static int encode_bf16_map(CUtensorMap* map, const void* base, uint64_t cols, uint64_t rows, uint32_t box_rows,
                           uint32_t box_cols = kMlpBK, CUtensorMapSwizzle swz = CU_TENSOR_MAP_SWIZZLE_128B) {
  PFN_encodeTiled enc = get_encode_tiled();
  if (!enc) return fail(PRG_ECUDA, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstride[1] = {cols * 2};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(PRG_ECUDA, "cuTensorMapEncodeTiled(bf16) failed: " + std::to_string((int)r));
  return PRG_OK;
}

static int tile_n(uint32_t N) {
  if (N <= 256) return (N % 16 == 0 && N >= 16) ? (int)N : 0;
  return (N % 256 == 0) ? 256 : 0;
}

template <int BN, bool FINAL, bool HILO, int NOUT>
static int launch_layer(prg_handle* h, const CUtensorMap& mapA, const CUtensorMap& mapW, const CUtensorMap& mapWhalf,
                        const CUtensorMap& mapOut, const MlpLayerParams& p, int Mp) {
  const int n_mblk = Mp / kMlpBM;
  if (!h->mlp_no_pair && n_mblk % 2 == 0 && BN % 32 == 0 && h->sm_count >= 2) {
    // CTA pairs (tcgen05 cta_group::2, 256 x BN tiles): units = pairs of row blocks
    const size_t smem = mlp_p_smem_bytes<BN, true, HILO>();
    PRG_CUDA(cudaFuncSetAttribute(mlp_layer_persistent_kernel<BN, FINAL, true, HILO, NOUT>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int units = (n_mblk / 2) * (p.N / BN);
    const int pairs = units < h->sm_count / 2 ? units : h->sm_count / 2;
    PRG_CUDA(launch_chained(h, mlp_layer_persistent_kernel<BN, FINAL, true, HILO, NOUT>, dim3((unsigned)(2 * pairs)),
                            dim3(kMlpPThreads), smem, 2, mapA, mapWhalf, mapOut, p, n_mblk));
    count_launch(h);
    return PRG_OK;
  }
  const size_t smem = mlp_p_smem_bytes<BN, false, HILO>();
  PRG_CUDA(cudaFuncSetAttribute(mlp_layer_persistent_kernel<BN, FINAL, false, HILO, NOUT>,
                                cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int tiles = n_mblk * (p.N / BN);
  const unsigned grid = (unsigned)(tiles < h->sm_count ? tiles : h->sm_count);
  PRG_CUDA(launch_chained(h, mlp_layer_persistent_kernel<BN, FINAL, false, HILO, NOUT>, dim3(grid), dim3(kMlpPThreads), smem, 1,
                          mapA, mapW, mapOut, p, n_mblk));
  count_launch(h);
  return PRG_OK;
}

template <bool FINAL, bool HILO, int NOUT>
static int launch_layer_bn(prg_handle* h, int BN, const CUtensorMap& a, const CUtensorMap& w, const CUtensorMap& wh,
                           const CUtensorMap& o, const MlpLayerParams& p, int Mp) {
  switch (BN) {
    case 64: return launch_layer<64, FINAL, HILO, NOUT>(h, a, w, wh, o, p, Mp);
    case 128: return launch_layer<128, FINAL, HILO, NOUT>(h, a, w, wh, o, p, Mp);
    case 192: return launch_layer<192, FINAL, HILO, NOUT>(h, a, w, wh, o, p, Mp);
    case 256: return launch_layer<256, FINAL, HILO, NOUT>(h, a, w, wh, o, p, Mp);
    default: return fail(PRG_EUNSUPPORTED, "MLP hidden width must be 64, 128, 192, 256 or a multiple of 256");
  }
}
template <bool HILO>
static int launch_layer_any(prg_handle* h, bool final_layer, int n_out, int BN, const CUtensorMap& a, const CUtensorMap& w,
                            const CUtensorMap& wh, const CUtensorMap& o, const MlpLayerParams& p, int Mp) {
  if (!final_layer) return launch_layer_bn<false, HILO, 1>(h, BN, a, w, wh, o, p, Mp);
  if (n_out == 1) return launch_layer_bn<true, HILO, 1>(h, BN, a, w, wh, o, p, Mp);
  return launch_layer_bn<true, HILO, kMlpMaxOut>(h, BN, a, w, wh, o, p, Mp);
}

// x_dev: [Mp][K0] bf16 (the tower input: ITEM columns only, K0 = h->mlp_k_item) in h->act[0];  logit_dev: [M][n_out]
// ubias: nullable [B][dims[1]] per-request first-layer bias rows (user / context share + b1), rows_per_req = n
// score_dev / score_map != nullptr: the last layer writes rank scores (sigmoid of fm_logit + tower logit per head, 0 for
// padding rows; Item.Score = sum_o coef_o * score_o) instead of logits
int mlp_forward_device(prg_handle* h, const uint16_t* x_dev, int M, float* logit_dev, const float* fm_logit_dev,
                       const uint32_t* rows_dev, double* score_dev, double* score_map, const float* ubias,
                       int rows_per_req) {
  const int L = h->mlp_layers;
  if (L < 2) return fail(PRG_ESTATE, "MLP weights not set (prg_set_mlp)");
  const int Mp = (M + kMlpBM - 1) / kMlpBM * kMlpBM;
  StageScope span(h, ST_MLP);
  const uint16_t* in = x_dev;
  for (int l = 0; l < L - 1; ++l) {
    const uint32_t K = l == 0 ? h->mlp_k_item : h->mlp_dims[l], N = h->mlp_dims[l + 1];
    const bool final_layer = (l == L - 2);
    const int BN = tile_n(N);
    CUtensorMap mapA;
    PRG_TRY(encode_bf16_map(&mapA, in, (l == 0 ? 1ull : 2ull) * K, (uint64_t)Mp, kMlpBM));
    MlpLayerParams p{};
    p.M = M; p.K = (int)K; p.N = (int)N; p.bias = (const float*)h->mlp_b[l].p;
    if (l == 0 && ubias) { p.ubias = ubias; p.rows_per_req = rows_per_req > 0 ? rows_per_req : 1; }
    CUtensorMap mapOut = mapA;
    if (final_layer) {
      p.w_last = (const float*)h->mlp_W[L - 1].p;
      p.n_out = (int)h->mlp_dims[L];
      for (int o = 0; o < kMlpMaxOut; ++o) { p.b_last[o] = h->mlp_b_last[o]; p.coef[o] = h->rank_coef[o]; }
      p.logit_out = logit_dev;
      if (score_dev || score_map) { p.fm_logit = fm_logit_dev; p.rows = rows_dev; p.score_out = score_dev; p.score_map = score_map; }
    } else {
      uint16_t* out = (uint16_t*)h->act[(l + 1) & 1].p;
      p.out = out;
      // 32 x 32 blocks of the [Mp][2N] output for the epilogue's TMA stores
      PRG_TRY(encode_bf16_map(&mapOut, out, 2ull * N, (uint64_t)Mp, 32, 32, CU_TENSOR_MAP_SWIZZLE_64B));
      in = out;
    }
    if (l == 0) PRG_TRY(launch_layer_any<false>(h, final_layer, p.n_out, BN, mapA, h->mlp_Wmap[l], h->mlp_Wmap_half[l], mapOut, p, Mp));
    else PRG_TRY(launch_layer_any<true>(h, final_layer, p.n_out, BN, mapA, h->mlp_Wmap[l], h->mlp_Wmap_half[l], mapOut, p, Mp));
  }
  return PRG_OK;
}

size_t mlp_act_bytes(const prg_handle* h, int M) {
  uint32_t wmax = 0;
  for (int l = 0; l < h->mlp_layers; ++l) wmax = h->mlp_dims[l] > wmax ? h->mlp_dims[l] : wmax;
  const size_t Mp = ((size_t)M + kMlpBM - 1) / kMlpBM * kMlpBM;
  return Mp * 2 * wmax * 2;
}

}  // namespace prg

using namespace prg;

extern "C" int prg_set_mlp(prg_handle* h, int n_layers, const uint32_t* dims, const uint16_t* const* W,
                           const float* const* bias) {
  if (!h) return fail(PRG_EINVAL, "null handle");
  if (n_layers < 2 || n_layers > kMaxLayers || !dims || !W || !bias) return fail(PRG_EINVAL, "bad MLP description");
  if (dims[n_layers] < 1 || dims[n_layers] > (uint32_t)kMlpMaxOut)
    return fail(PRG_EUNSUPPORTED, "MLP output width (heads) must be 1..4");
  for (int l = 1; l < n_layers; ++l) {
    if (tile_n(dims[l]) == 0 || dims[l] % kMlpBK != 0)
      return fail(PRG_EUNSUPPORTED, "MLP hidden widths must be 64, 128, 192, 256 or a multiple of 256");
  }
  if (dims[n_layers - 1] > 256) return fail(PRG_EUNSUPPORTED, "last hidden width must be <= 256 (fused output layer)");
  std::lock_guard<std::mutex> lk(h->mu);
  // the input row is [item factors | user factors | user dense] (prg_set_user_fields comes first): the item columns go
  // through the tensor cores per candidate, the user / context columns once per request (user_prefix_kernel)
  const uint32_t k_user = h->n_user_fields * 16 + h->n_user_dense;
  if (dims[0] <= k_user || (dims[0] - k_user) % kMlpBK != 0)
    return fail(PRG_EUNSUPPORTED, "MLP input width minus the user/context columns (prg_set_user_fields) must be a "
                                  "positive multiple of 64");
  const uint32_t k_item = dims[0] - k_user;
  PRG_CUDA(cudaSetDevice(h->device));
  prg::resolve_pending(h);
  PRG_CUDA(cudaStreamSynchronize(h->stream));
  h->mlp_layers = 0;
  for (int l = 0; l < n_layers; ++l) {
    const size_t K = dims[l], N = dims[l + 1];
    if (!W[l]) return fail(PRG_EINVAL, "null weight matrix");
    if (l < n_layers - 1) {
      const size_t Kt = l == 0 ? k_item : K;   // columns the TMA map covers
      std::vector<uint16_t> packed;
      const uint16_t* src = W[l];
      if (l == 0 && k_user) {                  // split: [N][k_item] bf16 for the MMA, [k_user][N] f32 for the prefix kernel
        packed.resize(N * Kt);
        std::vector<float> wu((size_t)k_user * N);
        for (size_t j = 0; j < N; ++j) {
          memcpy(&packed[j * Kt], W[l] + j * K, Kt * 2);
          for (size_t c = 0; c < k_user; ++c) {
            const uint32_t u = (uint32_t)W[l][j * K + Kt + c] << 16;
            memcpy(&wu[c * N + j], &u, 4);
          }
        }
        src = packed.data();
        PRG_TRY(h->mlp_Wu.ensure(wu.size() * 4));
        PRG_CUDA(cudaMemcpy(h->mlp_Wu.p, wu.data(), wu.size() * 4, cudaMemcpyHostToDevice));
      }
      PRG_TRY(h->mlp_W[l].ensure(N * Kt * 2));
      PRG_CUDA(cudaMemcpy(h->mlp_W[l].p, src, N * Kt * 2, cudaMemcpyHostToDevice));
      PRG_TRY(h->mlp_b[l].ensure(N * 4));
      if (bias[l]) PRG_CUDA(cudaMemcpy(h->mlp_b[l].p, bias[l], N * 4, cudaMemcpyHostToDevice));
      else PRG_CUDA(cudaMemset(h->mlp_b[l].p, 0, N * 4));
      PRG_TRY(encode_bf16_map(&h->mlp_Wmap[l], h->mlp_W[l].p, Kt, N, (uint32_t)tile_n((uint32_t)N)));
      PRG_TRY(encode_bf16_map(&h->mlp_Wmap_half[l], h->mlp_W[l].p, Kt, N, (uint32_t)tile_n((uint32_t)N) / 2));
    } else {
      std::vector<float> wl(K * N);            // [heads][K] f32
      for (size_t i = 0; i < K * N; ++i) {
        uint32_t u = (uint32_t)W[l][i] << 16;
        memcpy(&wl[i], &u, 4);
      }
      PRG_TRY(h->mlp_W[l].ensure(K * N * 4));
      PRG_CUDA(cudaMemcpy(h->mlp_W[l].p, wl.data(), K * N * 4, cudaMemcpyHostToDevice));
      for (size_t o = 0; o < (size_t)kMlpMaxOut; ++o) h->mlp_b_last[o] = (bias[l] && o < N) ? bias[l][o] : 0.f;
    }
  }
  for (int l = 0; l <= n_layers; ++l) h->mlp_dims[l] = dims[l];
  h->mlp_k_item = k_item;
  h->mlp_k_user = k_user;
  h->mlp_layers = n_layers;
  return PRG_OK;
}
