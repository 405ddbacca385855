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
// Fallbacks: the same kernel with PAIR = false (single CTA, M = 128, direct stores) for an odd number of row blocks,
// and the one-tile-per-CTA mlp_layer_kernel (config "mlp_one_tile", A/B measurements).
#include "handle.h"
#include <cuda_bf16.h>
#include <cstring>
#include <vector>

namespace prg {

constexpr int kMlpBM = 128;
constexpr int kMlpBK = 64;       // bf16 elements per k-block = one 128-B swizzle row
#ifndef MLP_STAGES
#define MLP_STAGES 4
#endif
constexpr int kMlpStages = MLP_STAGES;
constexpr int kMlpThreads = 192;

struct MlpLayerParams {
  int M;                 // valid rows
  int K;                 // input width (the A tensor has 2K columns)
  int N;                 // output width
  const float* bias;     // [N]
  uint16_t* out;         // [Mp][2N] bf16 (hi | lo)                 (hidden layers)
  const float* w_last;   // [N] f32 (bf16 values widened)           (FINAL)
  float b_last;          //                                         (FINAL)
  float* logit_out;      // [M]                                     (FINAL)
  // fused score epilogue (FINAL, optional): score = rows[i] == pad ? 0 : (double)(float)sigmoid((fm_logit[i] +) logit)
  const float* fm_logit;  // nullable: logit of the FM part, added first (DeepFM-shaped model)
  const uint32_t* rows;   // candidate rows (0xFFFFFFFF = padding)
  double* score_out;      // nullable: when set, scores are written instead of logits
};

template <int BN>
constexpr size_t mlp_smem_bytes() {
  return (size_t)kMlpStages * (kMlpBM * 128 + BN * 128) + 2 * BN * 4 + (2 * kMlpStages + 1) * 8 + 16;
}

template <int BN, bool FINAL>
__global__ void __launch_bounds__(kMlpThreads, MLP_STAGES <= 2 ? 2 : 1)
mlp_layer_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapW,
                 const MlpLayerParams p) {
  extern __shared__ __align__(1024) uint8_t msm[];
  constexpr int kABytes = kMlpBM * 128, kBBytes = BN * 128, kStageBytes = kABytes + kBBytes;
  constexpr uint32_t kTmemCols = BN <= 32 ? 32 : BN <= 64 ? 64 : BN <= 128 ? 128 : 256;
  float* bias_s = reinterpret_cast<float*>(msm + (size_t)kMlpStages * kStageBytes);
  float* wl_s = bias_s + BN;
  uint64_t* full = reinterpret_cast<uint64_t*>(wl_s + BN);
  uint64_t* empty = full + kMlpStages;
  uint64_t* tmem_full = empty + kMlpStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int m_blk = blockIdx.x, n_blk = blockIdx.y;
  const int num_kb = 2 * p.K / kMlpBK;

  for (int i = tid; i < BN; i += kMlpThreads) {
    bias_s[i] = p.bias ? p.bias[n_blk * BN + i] : 0.f;
    wl_s[i] = FINAL ? p.w_last[n_blk * BN + i] : 0.f;
  }
  if (tid == 0) {
    tma_prefetch_desc(&mapA);
    tma_prefetch_desc(&mapW);
    for (int s = 0; s < kMlpStages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(tmem_full, 1);
    mbar_fence_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "n"(kTmemCols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % kMlpStages;
        const uint32_t ph = (uint32_t)(kb / kMlpStages) & 1u;
        mbar_wait(&empty[s], ph ^ 1u);
        mbar_arrive_expect_tx(&full[s], kStageBytes);
        uint8_t* a_dst = msm + (size_t)s * kStageBytes;
        tma_load_2d(a_dst, &mapA, kb * kMlpBK, m_blk * kMlpBM, &full[s], kEvictNormal);
        tma_load_2d(a_dst + kABytes, &mapW, (kb * kMlpBK) % p.K, n_blk * BN, &full[s], kEvictLast);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // instruction descriptor: D=f32, A=B=bf16, both K-major, N=BN, M=128
      constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(kMlpBM >> 4) << 24);
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % kMlpStages;
        const uint32_t ph = (uint32_t)(kb / kMlpStages) & 1u;
        mbar_wait(&full[s], ph);
        tc_fence_after();
        const uint32_t a_addr = smem_u32(msm + (size_t)s * kStageBytes);
        const uint64_t adesc = umma_desc_k_sw128(a_addr), bdesc = umma_desc_k_sw128(a_addr + kABytes);
#pragma unroll
        for (int k = 0; k < kMlpBK / 16; ++k)  // UMMA_K = 16 bf16 = 32 B: advance the start address by 2 (16-B units)
          umma_bf16(tmem_base, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (kb | k) != 0 ? 1u : 0u);
        umma_commit(&empty[s]);  // frees the stage when these MMAs have read it
      }
      umma_commit(tmem_full);    // accumulator complete
    }
  } else {
    // ---------------------------------------------------------------- epilogue: 4 warps x 32 TMEM lanes
    const int quarter = warp & 3;  // a warp may only touch TMEM lanes [32*(warp%4), +32)
    const int row = m_blk * kMlpBM + quarter * 32 + lane;
    mbar_wait(tmem_full, 0);
    tc_fence_after();
    float logit = 0.f;
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 32) {
      uint32_t v[32];
      tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)c0, v);
      uint32_t hi_w[16], lo_w[16];
#pragma unroll
      for (int c = 0; c < 32; c += 2) {
        float r0 = fmaxf(__fadd_rn(__uint_as_float(v[c]), bias_s[c0 + c]), 0.f);
        float r1 = fmaxf(__fadd_rn(__uint_as_float(v[c + 1]), bias_s[c0 + c + 1]), 0.f);
        const uint16_t h0 = __bfloat16_as_ushort(__float2bfloat16_rn(r0)), h1 = __bfloat16_as_ushort(__float2bfloat16_rn(r1));
        const float h0f = __uint_as_float((uint32_t)h0 << 16), h1f = __uint_as_float((uint32_t)h1 << 16);
        const uint16_t l0 = __bfloat16_as_ushort(__float2bfloat16_rn(__fsub_rn(r0, h0f)));
        const uint16_t l1 = __bfloat16_as_ushort(__float2bfloat16_rn(__fsub_rn(r1, h1f)));
        if (FINAL) {
          const float a0 = __fadd_rn(h0f, __uint_as_float((uint32_t)l0 << 16));
          const float a1 = __fadd_rn(h1f, __uint_as_float((uint32_t)l1 << 16));
          logit = __fmaf_rn(wl_s[c0 + c], a0, logit);
          logit = __fmaf_rn(wl_s[c0 + c + 1], a1, logit);
        } else {
          hi_w[c >> 1] = (uint32_t)h0 | ((uint32_t)h1 << 16);
          lo_w[c >> 1] = (uint32_t)l0 | ((uint32_t)l1 << 16);
        }
      }
      if (!FINAL) {
        uint16_t* o = p.out + (size_t)row * (2 * p.N) + (size_t)n_blk * BN + c0;
        uint4* oh = reinterpret_cast<uint4*>(o);
        uint4* ol = reinterpret_cast<uint4*>(o + p.N);
#pragma unroll
        for (int w = 0; w < 4; ++w) {
          oh[w] = make_uint4(hi_w[4 * w], hi_w[4 * w + 1], hi_w[4 * w + 2], hi_w[4 * w + 3]);
          ol[w] = make_uint4(lo_w[4 * w], lo_w[4 * w + 1], lo_w[4 * w + 2], lo_w[4 * w + 3]);
        }
      }
    }
    if (FINAL && row < p.M) p.logit_out[row] = __fadd_rn(logit, p.b_last);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kTmemCols));
  }
}

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
__device__ __forceinline__ void mlp_write_result(const MlpLayerParams& p, int row, float mlp_logit) {
  if (p.score_out) {
    float l = p.fm_logit ? __fadd_rn(p.fm_logit[row], mlp_logit) : mlp_logit;
    const float sc = (float)(1.0 / (1.0 + exp(-(double)l)));
    p.score_out[row] = (p.rows[row] == 0xFFFFFFFFu) ? 0.0 : (double)sc;
  } else {
    p.logit_out[row] = mlp_logit;
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
template <bool PAIR>
constexpr int mlp_p_stages() { return 3; }
constexpr int kMlpOutStage = 8192;   // per epilogue warp: two buffers of (32 x 32 bf16 hi block | lo block) staged for the TMA stores

template <int BN, bool PAIR>
constexpr size_t mlp_p_smem_bytes() {
  return (size_t)mlp_p_stages<PAIR>() * (2 * kMlpBM * 128 + (PAIR ? BN / 2 : BN) * 128) +
         (PAIR ? (size_t)kMlpEpiWarps * kMlpOutStage : 0) + 2 * 1024 * 4 /*bias, w_last*/ +
         2 * 2 * kMlpBM * 4 /*partials*/ + (2 * mlp_p_stages<PAIR>() + 4) * 8 + 16;
}

// PAIR: a 2-CTA cluster computes a 256 x BN tile with tcgen05.mma.cta_group::2 (M = 256): each CTA stages its own 128
// rows of activations and HALF of the W block (BN/2 rows); the leader's MMA thread issues one instruction for both SMs
// and each SM's accumulator (its 128 rows x BN columns) lands in its own TMEM.  A single-CTA M = 128 MMA reads
// (128 + BN) x 32 B of operands from shared memory per K = 16 step — at BN = 256 that is 12 KiB per 136 tensor-pipe
// cycles, 90 B/clk of the SM's 128 B/clk with the TMA writes of the next stage on top: the tensor pipe idled 60 % of
// the time.  The pair halves the W bytes each SM reads and stages.
// Protocol: every TMA load of both CTAs signals the LEADER's full[s] (.cta_group::2 barrier in the peer); the leader's
// commits are multicast to empty[s] / tfull[b] of both CTAs; the peer's epilogue warps arrive on the leader's tempty[b].
template <int BN, bool FINAL, bool PAIR>
__global__ void __launch_bounds__(kMlpPThreads, 1)
mlp_layer_persistent_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapW,
                            const __grid_constant__ CUtensorMap mapOut, const MlpLayerParams p, const int n_mblk) {
  extern __shared__ __align__(1024) uint8_t msm[];
  constexpr int kABytes = kMlpBM * 128, kBBytes = (PAIR ? BN / 2 : BN) * 128;   // W rows staged by this CTA
  constexpr int kStageBytes = 2 * kABytes + kBBytes;                               // A_hi | A_lo | W
  constexpr int kMlpStages = mlp_p_stages<PAIR>();
  constexpr uint32_t kTmemCols = 2 * BN <= 64 ? 64 : 2 * BN <= 128 ? 128 : 2 * BN <= 256 ? 256 : 512;
  uint8_t* out_stage = msm + (size_t)kMlpStages * kStageBytes;                       // PAIR: [8 warps][2 buffers][hi 2 KiB | lo 2 KiB]
  float* bias_s = reinterpret_cast<float*>(out_stage + (PAIR ? kMlpEpiWarps * kMlpOutStage : 0));  // [1024]
  float* wl_s = bias_s + 1024;                                                       // [1024]
  float* part_s = wl_s + 1024;                                                       // [2 buf][2 halves][128]
  uint64_t* full = reinterpret_cast<uint64_t*>(part_s + 4 * kMlpBM);
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

  for (int i = tid; i < p.N; i += kMlpPThreads) {
    bias_s[i] = p.bias ? p.bias[i] : 0.f;
    wl_s[i] = FINAL ? p.w_last[i] : 0.f;
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
            tma_load_2d_pair(a_dst + kABytes, &mapA, p.K + kb * kMlpBK, m_blk * kMlpBM, lbar, kEvictNormal);
            tma_load_2d_pair(a_dst + 2 * kABytes, &mapW, kb * kMlpBK, n_blk * BN + (int)cta_rank * (BN / 2), lbar,
                             kEvictLast);   // our half of the W block (mapW boxes are BN/2 rows here)
          } else {
            mbar_arrive_expect_tx(&full[s], kStageBytes);
            tma_load_2d(a_dst, &mapA, kb * kMlpBK, m_blk * kMlpBM, &full[s], kEvictNormal);
            tma_load_2d(a_dst + kABytes, &mapA, p.K + kb * kMlpBK, m_blk * kMlpBM, &full[s], kEvictNormal);
            tma_load_2d(a_dst + 2 * kABytes, &mapW, kb * kMlpBK, n_blk * BN, &full[s], kEvictLast);
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
          const uint64_t bdesc = umma_desc_k_sw128(a_addr + 2 * kABytes);
#pragma unroll
          for (int part = 0; part < 2; ++part) {   // hi, then lo, against the same W block
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
      mbar_wait(&tfull[buf], ((uint32_t)i >> 1) & 1u);
      tc_fence_after();
      float logit = 0.f;
#pragma unroll 1
      for (int c0 = chalf * kHalf; c0 < (chalf + 1) * kHalf; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + buf * (uint32_t)BN + (uint32_t)c0, v);
        if (c0 + 32 >= (chalf + 1) * kHalf) {  // last read of this tile by this warp: hand the buffer back
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if constexpr (PAIR) mbar_arrive_cluster(mapa_cluster(smem_u32(&tempty[buf]), 0));   // the leader's barrier
            else mbar_arrive(&tempty[buf]);
          }
        }
        const int nb = n_blk * BN + c0;  // column in the layer's output
        uint32_t hi_w[16], lo_w[16];
#pragma unroll
        for (int c = 0; c < 32; c += 2) {
          float r0 = fmaxf(__fadd_rn(__uint_as_float(v[c]), bias_s[nb + c]), 0.f);
          float r1 = fmaxf(__fadd_rn(__uint_as_float(v[c + 1]), bias_s[nb + c + 1]), 0.f);
          // packed conversions (cvt.rn.bf16x2.f32: one XU instruction per PAIR of values; the XU pipe was 32 % busy)
          const uint32_t hp = pack_bf16x2(r0, r1);                       // low half = bf16(r0), high half = bf16(r1)
          const float h0f = __uint_as_float(hp << 16), h1f = __uint_as_float(hp & 0xFFFF0000u);
          const uint32_t lp = pack_bf16x2(__fsub_rn(r0, h0f), __fsub_rn(r1, h1f));
          if (FINAL) {
            const float a0 = __fadd_rn(h0f, __uint_as_float(lp << 16));
            const float a1 = __fadd_rn(h1f, __uint_as_float(lp & 0xFFFF0000u));
            logit = __fmaf_rn(wl_s[nb + c], a0, logit);
            logit = __fmaf_rn(wl_s[nb + c + 1], a1, logit);
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
        part_s[(buf * 2 + chalf) * kMlpBM + quarter * 32 + lane] = logit;
        asm volatile("bar.sync 1, %0;" ::"n"(kMlpEpiWarps * 32) : "memory");
        if (chalf == 0 && row < p.M)
          mlp_write_result(p, row, __fadd_rn(__fadd_rn(part_s[(buf * 2) * kMlpBM + quarter * 32 + lane],
                                                       part_s[(buf * 2 + 1) * kMlpBM + quarter * 32 + lane]), p.b_last));
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

template <int BN, bool FINAL>
static int launch_layer(prg_handle* h, const CUtensorMap& mapA, const CUtensorMap& mapW, const CUtensorMap& mapWhalf,
                        const CUtensorMap& mapOut, const MlpLayerParams& p, int Mp) {
  if (!h->mlp_one_tile_per_cta && p.N <= 1024) {
    const int n_mblk = Mp / kMlpBM;
    if (!h->mlp_no_pair && n_mblk % 2 == 0 && BN % 32 == 0 && h->sm_count >= 2) {
      // CTA pairs (tcgen05 cta_group::2, 256 x BN tiles): units = pairs of row blocks
      const size_t smem = mlp_p_smem_bytes<BN, true>();
      PRG_CUDA(cudaFuncSetAttribute(mlp_layer_persistent_kernel<BN, FINAL, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)smem));
      const int units = (n_mblk / 2) * (p.N / BN);
      const int pairs = units < h->sm_count / 2 ? units : h->sm_count / 2;
      PRG_CUDA(launch_chained(h, mlp_layer_persistent_kernel<BN, FINAL, true>, dim3((unsigned)(2 * pairs)), dim3(kMlpPThreads),
                              smem, 2, mapA, mapWhalf, mapOut, p, n_mblk));
      count_launch(h);
      return PRG_OK;
    }
    const size_t smem = mlp_p_smem_bytes<BN, false>();
    PRG_CUDA(cudaFuncSetAttribute(mlp_layer_persistent_kernel<BN, FINAL, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)smem));
    const int tiles = n_mblk * (p.N / BN);
    const unsigned grid = (unsigned)(tiles < h->sm_count ? tiles : h->sm_count);
    PRG_CUDA(launch_chained(h, mlp_layer_persistent_kernel<BN, FINAL, false>, dim3(grid), dim3(kMlpPThreads), smem, 1, mapA, mapW,
                            mapOut, p, n_mblk));
    count_launch(h);
    return PRG_OK;
  }
  const size_t smem = mlp_smem_bytes<BN>();
  PRG_CUDA(cudaFuncSetAttribute(mlp_layer_kernel<BN, FINAL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((unsigned)(Mp / kMlpBM), (unsigned)(p.N / BN));
  mlp_layer_kernel<BN, FINAL><<<grid, kMlpThreads, smem, h->stream>>>(mapA, mapW, p);
  PRG_CUDA(cudaGetLastError());
  count_launch(h);
  return PRG_OK;
}

template <bool FINAL>
static int launch_layer_bn(prg_handle* h, int BN, const CUtensorMap& a, const CUtensorMap& w, const CUtensorMap& wh,
                           const CUtensorMap& o, const MlpLayerParams& p, int Mp) {
  switch (BN) {
    case 64: return launch_layer<64, FINAL>(h, a, w, wh, o, p, Mp);
    case 128: return launch_layer<128, FINAL>(h, a, w, wh, o, p, Mp);
    case 192: return launch_layer<192, FINAL>(h, a, w, wh, o, p, Mp);
    case 256: return launch_layer<256, FINAL>(h, a, w, wh, o, p, Mp);
    default: return fail(PRG_EUNSUPPORTED, "MLP hidden width must be 64, 128, 192, 256 or a multiple of 256");
  }
}

// x_dev: [Mp][2*dims[0]] bf16 hi|lo in h->act[0];  logit_dev: [M]
// score_dev != nullptr: the last layer writes rank scores (sigmoid of fm_logit + tower logit, 0 for padding rows)
// instead of logits — only the persistent kernels do that; *fused_score tells the caller whether it happened
int mlp_forward_device(prg_handle* h, const uint16_t* x_dev, int M, float* logit_dev, const float* fm_logit_dev,
                       const uint32_t* rows_dev, double* score_dev, bool* fused_score) {
  if (fused_score) *fused_score = false;
  const int L = h->mlp_layers;
  if (L < 2) return fail(PRG_ESTATE, "MLP weights not set (prg_set_mlp)");
  const int Mp = (M + kMlpBM - 1) / kMlpBM * kMlpBM;
  StageScope span(h, ST_MLP);
  const uint16_t* in = x_dev;
  for (int l = 0; l < L - 1; ++l) {
    const uint32_t K = h->mlp_dims[l], N = h->mlp_dims[l + 1];
    const bool final_layer = (l == L - 2);
    const int BN = tile_n(N);
    CUtensorMap mapA;
    PRG_TRY(encode_bf16_map(&mapA, in, 2ull * K, (uint64_t)Mp, kMlpBM));
    MlpLayerParams p{};
    p.M = M; p.K = (int)K; p.N = (int)N; p.bias = (const float*)h->mlp_b[l].p;
    if (final_layer) {
      p.w_last = (const float*)h->mlp_W[L - 1].p;
      p.b_last = h->mlp_b_last;
      p.logit_out = logit_dev;
      if (score_dev && fused_score && !h->mlp_one_tile_per_cta && N <= 1024) {
        p.fm_logit = fm_logit_dev; p.rows = rows_dev; p.score_out = score_dev;
        *fused_score = true;
      }
      PRG_TRY(launch_layer_bn<true>(h, BN, mapA, h->mlp_Wmap[l], h->mlp_Wmap_half[l], mapA, p, Mp));
    } else {
      uint16_t* out = (uint16_t*)h->act[(l + 1) & 1].p;
      p.out = out;
      CUtensorMap mapOut;   // 32 x 32 blocks of the [Mp][2N] output for the epilogue's TMA stores
      PRG_TRY(encode_bf16_map(&mapOut, out, 2ull * N, (uint64_t)Mp, 32, 32, CU_TENSOR_MAP_SWIZZLE_64B));
      PRG_TRY(launch_layer_bn<false>(h, BN, mapA, h->mlp_Wmap[l], h->mlp_Wmap_half[l], mapOut, p, Mp));
      in = out;
    }
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
  if (dims[n_layers] != 1) return fail(PRG_EUNSUPPORTED, "MLP output width must be 1");
  if (dims[0] % kMlpBK != 0) return fail(PRG_EUNSUPPORTED, "MLP input width must be a multiple of 64");
  for (int l = 1; l < n_layers; ++l) {
    if (tile_n(dims[l]) == 0 || dims[l] % kMlpBK != 0)
      return fail(PRG_EUNSUPPORTED, "MLP hidden widths must be 64, 128, 192, 256 or a multiple of 256");
  }
  if (dims[n_layers - 1] > 256) return fail(PRG_EUNSUPPORTED, "last hidden width must be <= 256 (fused output layer)");
  std::lock_guard<std::mutex> lk(h->mu);
  PRG_CUDA(cudaSetDevice(h->device));
  prg::resolve_pending(h);
  PRG_CUDA(cudaStreamSynchronize(h->stream));
  h->mlp_layers = 0;
  for (int l = 0; l < n_layers; ++l) {
    const size_t K = dims[l], N = dims[l + 1];
    if (!W[l]) return fail(PRG_EINVAL, "null weight matrix");
    if (l < n_layers - 1) {
      PRG_TRY(h->mlp_W[l].ensure(N * K * 2));
      PRG_CUDA(cudaMemcpy(h->mlp_W[l].p, W[l], N * K * 2, cudaMemcpyHostToDevice));
      PRG_TRY(h->mlp_b[l].ensure(N * 4));
      if (bias[l]) PRG_CUDA(cudaMemcpy(h->mlp_b[l].p, bias[l], N * 4, cudaMemcpyHostToDevice));
      else PRG_CUDA(cudaMemset(h->mlp_b[l].p, 0, N * 4));
      PRG_TRY(encode_bf16_map(&h->mlp_Wmap[l], h->mlp_W[l].p, K, N, (uint32_t)tile_n((uint32_t)N)));
      PRG_TRY(encode_bf16_map(&h->mlp_Wmap_half[l], h->mlp_W[l].p, K, N, (uint32_t)tile_n((uint32_t)N) / 2));
    } else {
      std::vector<float> wl(K);
      for (size_t i = 0; i < K; ++i) {
        uint32_t u = (uint32_t)W[l][i] << 16;
        memcpy(&wl[i], &u, 4);
      }
      PRG_TRY(h->mlp_W[l].ensure(K * 4));
      PRG_CUDA(cudaMemcpy(h->mlp_W[l].p, wl.data(), K * 4, cudaMemcpyHostToDevice));
      h->mlp_b_last = bias[l] ? bias[l][0] : 0.f;
    }
  }
  for (int l = 0; l <= n_layers; ++l) h->mlp_dims[l] = dims[l];
  h->mlp_layers = n_layers;
  return PRG_OK;
}
