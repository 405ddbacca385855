// recall_i8.cu — int8 filter index for the recall scan at dim 64, <= 64 queries per pass.
//
// The bf16 filter pass of recall_tc.cu streams rows*dim*2 bytes and runs at 0.89 of the measured HBM rate: the only way
// to make it faster is to stream fewer bytes.  This pass streams an INT8 shadow of the item matrix (rows*dim bytes + 8
// bytes of per-row parameters) through tcgen05.mma kind::i8 (s32 accumulators in tensor memory).  As with the bf16 index
// the fp32 matrix stays the authority: the filter only has to keep a SUPERSET of the rows whose exact key reaches the
// sampled threshold tau; survivors are re-scored exactly (recall.cu rescore_kernel), so every emitted row and score is
// still bit-identical to the oracle.
//
// Quantisation and the bound (design check: tools/int8_filter_check.py).  Row r: s_r = max_d|x_d| / 127,
// x8_d = rint(x_d / s_r), so x_d = s_r (x8_d + e_d), |e_d| <= 1/2.  Query q (scaled by 1/tau_q when the pass uses the
// uniform threshold, as in recall_tc.cu): q_d = t (Q_d + f_d), |f_d| <= 1/2, with ONE scale t for all queries of a
// uniform pass (t_q = max_d|q_d| / 127 per query otherwise).  The tensor core gives I = sum_d x8_d Q_d exactly, and
//     | sum_d x_d q_d  -  s_r t I |  <=  s_r t ( 1/2 sum|x8_d| + 1/2 sum|Q_d| + d/4 ).
// A row survives unless  s_r t I  <  tau - 1.01 s_r t (1/2 L1(x8_r) + 1/2 L1(Q_q) + 17): the 1 % and the 17 instead of 16
// cover the float roundings of the quantisation, of the threshold arithmetic here, of q/tau and of the exact fmaf chain
// itself (all of them <= 5e-4 of the 1/2 L1(Q) term: a non-zero query has an element of magnitude 127).  Uniform pass:
// I >= T_r = (1 - 1e-6) / (s_r t) - hl_r - max_q C_q, ONE integer per row; hl_r = 1.01 / 2 L1(x8_r) and a_r = (1 - 1e-6) / s_r
// are stored per row.  Rows with a NaN / inf element (or a scale below 1e-30) carry a_r = NaN and always survive; all-zero
// rows carry a_r = +inf and survive exactly when tau <= 0; a product s_r t that underflows lets the row survive.  About 1.6 x the rows that reach tau survive (bf16: 1.1 x).
//
// Layout trick for 64-byte rows.  A K-major SWIZZLE_128B operand has 128-byte rows, an int8 row of dim 64 has 64.  The
// index is therefore read as [rows / 2][128]: operand row j holds matrix rows 2j and 2j + 1 side by side, and the B
// operand has 128 columns: column n < 64 is query n against the first 64 bytes (zeros in the second half), column 64 + n
// is query n against the second 64 bytes.  D[j][n] = score(2j, n), D[j][64 + n] = score(2j + 1, n).  Half of the MACs
// multiply zeros; the pass is nowhere near the tensor rate (8 MMAs of 32 cycles per 32-KiB stage against ~1500 cycles
// of HBM time per stage and SM).
//
// Structure (as recall_tc.cu): persistent, one CTA per SM, 10 warps: TMA producer (one 32-KiB box of 512 matrix rows per
// tile + a 4-KiB bulk copy of the rows' parameters, 5-stage ring), one MMA thread (M = 128, N = 128, K = 32, two halves
// per tile, accumulators double buffered in 512 TMEM columns), 16 epilogue warps, one matrix row per thread (tcgen05.ld
// 32x32b.x32; integer maximum per group of 16 queries, compare with T_r, append survivors to the per-CTA, per-query segments).
#include "recall.h"
#include <cuda_bf16.h>
#include <type_traits>

namespace prg {

constexpr int kI8EpiWarps = 16;
constexpr int kI8Threads = (2 + kI8EpiWarps) * 32;   // 576
constexpr int kI8Dim = 64;
constexpr int kI8TileRows = 512;                 // matrix rows per tile = 256 operand rows of 128 B
constexpr int kI8BoxRows = 256;
constexpr int kI8Stages = 5;
constexpr int kI8StageBytes = kI8BoxRows * 128;  // 32 KiB
constexpr int kI8PrmRing = kI8Stages + 3;        // producer up to S tiles ahead of the MMA thread, that 2 ahead of the epilogue
constexpr int kI8QBytes = 128 * 128;             // B operand: 128 columns x 128 B
constexpr float kI8Slack = 1.01f;

constexpr size_t scan_i8_smem_bytes() {
  return (size_t)kI8Stages * kI8StageBytes + kI8QBytes + kQB * 16 /*tq*/ + kQB * 4 /*s_cnt*/ + (2 * kI8Stages + 4) * 8 + 16 +
         (size_t)kI8PrmRing * kI8TileRows * 8;
}

__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}

// float threshold -> the smallest integer accumulator that survives (NaN / -inf: everything, +inf: nothing)
__device__ __forceinline__ int i8_thr(float T) {
  if (!(T > -2.0e9f)) return (int)0x80000000;
  if (T >= 2.0e9f) return 0x7FFFFFFF;
  return __float2int_ru(T);
}
__device__ __forceinline__ int imax3(int a, int b, int c) { return max(max(a, b), c); }
__device__ __forceinline__ float fmax3f(float a, float b, float c) {  // SASS FMNMX3; NaN operands are ignored
  float r;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
  return r;
}

__global__ void __launch_bounds__(kI8Threads, 1)
recall_scan_i8_kernel(const __grid_constant__ CUtensorMap emap, const ScanParams p, const uint32_t n_tiles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* stage_base = smem;
  uint8_t* Qb = smem + (size_t)kI8Stages * kI8StageBytes;
  float4* tq = reinterpret_cast<float4*>(Qb + kI8QBytes);     // [64] {tau_f, t_q, C_q, -}
  uint32_t* s_cnt = reinterpret_cast<uint32_t*>(tq + kQB);
  uint64_t* full = reinterpret_cast<uint64_t*>(s_cnt + kQB);
  uint64_t* empty = full + kI8Stages;
  uint64_t* tfull = empty + kI8Stages;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
  float2* prm = reinterpret_cast<float2*>(tmem_slot + 4);      // [kI8PrmRing][512] {a_r, hl_r}
  __shared__ float s_scale[kQB];
  __shared__ float s_amax[kQB];
  __shared__ uint32_t s_tmax, s_cmax;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    s_tmax = 0u; s_cmax = 0u;
    tma_prefetch_desc(&emap);
    for (int s = 0; s < kI8Stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&tfull[b], 1); mbar_init(&tempty[b], kI8EpiWarps); }
    mbar_fence_init();
  }
  const uint32_t my_tiles = (n_tiles > blockIdx.x) ? (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  auto issue_tile = [&](uint32_t i) {
    const uint32_t t = blockIdx.x + i * gridDim.x;
    const uint32_t s = i % kI8Stages, ph = (i / kI8Stages) & 1u;
    mbar_wait(&empty[s], ph ^ 1u);
    mbar_arrive_expect_tx(&full[s], kI8StageBytes + kI8TileRows * 8);
    // (the parameter array is padded to whole tiles, build_i8_index)
    bulk_load_1d(prm + (size_t)(i % kI8PrmRing) * kI8TileRows, p.row_q8 + (size_t)t * kI8TileRows, kI8TileRows * 8, &full[s]);
    tma_load_2d(stage_base + (size_t)s * kI8StageBytes, &emap, 0, (int)(t * (uint32_t)kI8BoxRows), &full[s], kEvictFirst);
  };
  // the index and its parameters are tables no kernel of the per-batch chain writes: the ring fills while the previous
  // kernel drains (everything below pdl_wait() reads what that kernel wrote: thresholds, queries)
  const uint32_t n_pre = my_tiles < (uint32_t)kI8Stages ? my_tiles : (uint32_t)kI8Stages;
  if (tid == 0)
    for (uint32_t i = 0; i < n_pre; ++i) issue_tile(i);
  for (int i = tid; i < kI8QBytes / 16; i += kI8Threads) reinterpret_cast<uint4*>(Qb)[i] = make_uint4(0u, 0u, 0u, 0u);
  pdl_wait();

  // ---- thresholds: uniform form (queries staged as q / tau_q) when every tau of the pass is positive and well scaled
  int bad = 0;
  for (int q = tid; q < kQB; q += kI8Threads) {
    float sc = 0.f;
    if (q < p.nq) {
      const uint64_t t = p.tau[q];
      const float tf = (t == 0) ? __int_as_float(0xFF800000) : key_score(t);
      if (tf > 0x1p-60f && tf < 0x1p60f) sc = 1.0f / tf; else bad = 1;
    }
    s_scale[q] = sc;
  }
  // ---- queries: 16 lanes per query, one float4 each; kept in registers across the two block-wide decisions
  constexpr int C4 = kI8Dim / 4;                                  // 16
  constexpr int kIters = (C4 * kQB + kI8Threads - 1) / kI8Threads;   // 2
  float4 v[kIters];
#pragma unroll
  for (int itq = 0; itq < kIters; ++itq) {
    const int i4 = tid + itq * kI8Threads;
    const bool in = i4 < C4 * kQB;
    const int q = in ? i4 / C4 : 0, c4 = i4 - q * C4;
    v[itq] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (in && q < p.nq) v[itq] = __ldg(reinterpret_cast<const float4*>(p.Q + (size_t)q * kI8Dim) + c4);
    float am = fmaxf(fmaxf(fabsf(v[itq].x), fabsf(v[itq].y)), fmaxf(fabsf(v[itq].z), fabsf(v[itq].w)));
    float l1 = fabsf(v[itq].x) + fabsf(v[itq].y) + fabsf(v[itq].z) + fabsf(v[itq].w);
#pragma unroll
    for (int off = 1; off < C4; off <<= 1) {
      am = fmaxf(am, __shfl_xor_sync(0xffffffffu, am, off));
      l1 += __shfl_xor_sync(0xffffffffu, l1, off);
    }
    if (!(l1 < __int_as_float(0x7F800000))) { bad = 1; am = __int_as_float(0x7FC00000); }   // NaN / inf element
    if (in && c4 == 0) s_amax[q] = am;
  }
  const bool scaled = __syncthreads_or(bad) == 0;   // (also publishes s_scale, s_amax, the zeroed operand, the barriers)
  if (scaled)
    for (int q = tid; q < p.nq && q < kQB; q += kI8Threads)
      atomicMax(&s_tmax, __float_as_uint(s_amax[q] * s_scale[q]));   // largest |element| of the staged queries (>= 0)
  __syncthreads();
  const float t_pass = __fdiv_rn(__uint_as_float(s_tmax), 127.f);
#pragma unroll
  for (int itq = 0; itq < kIters; ++itq) {
    const int i4 = tid + itq * kI8Threads;
    const bool in = i4 < C4 * kQB;
    const int q = in ? i4 / C4 : 0, c4 = i4 - q * C4;
    float4 w = v[itq];
    float tqs;                                     // this query's scale
    const float am = s_amax[q];
    if (scaled) { const float sc = s_scale[q]; w.x *= sc; w.y *= sc; w.z *= sc; w.w *= sc; tqs = t_pass; }
    else tqs = __fdiv_rn(am, 127.f);
    int e0 = 0, e1 = 0, e2 = 0, e3 = 0;
    const bool finite_q = am == am;                // (NaN marks a query with a NaN / inf element)
    // a scale that underflowed to 0 (or below the range where x / t is safe) stages a zero query: C_q = +inf below
    const bool usable = finite_q && tqs >= 1e-30f;
    if (usable) {
      e0 = max(-127, min(127, __float2int_rn(__fdiv_rn(w.x, tqs))));
      e1 = max(-127, min(127, __float2int_rn(__fdiv_rn(w.y, tqs))));
      e2 = max(-127, min(127, __float2int_rn(__fdiv_rn(w.z, tqs))));
      e3 = max(-127, min(127, __float2int_rn(__fdiv_rn(w.w, tqs))));
    }
    int l1q = abs(e0) + abs(e1) + abs(e2) + abs(e3);
#pragma unroll
    for (int off = 1; off < C4; off <<= 1) l1q += __shfl_xor_sync(0xffffffffu, l1q, off);
    if (in) {
      const uint32_t pk = (uint32_t)(e0 & 255) | ((uint32_t)(e1 & 255) << 8) | ((uint32_t)(e2 & 255) << 16) | ((uint32_t)(e3 & 255) << 24);
      const int dd = c4 * 4, ch = dd >> 4, within = dd & 15;
      // column q: bytes [0, 64) of operand row q; column 64 + q: bytes [64, 128) of operand row 64 + q (SWIZZLE_128B:
      // 16-B chunk index XOR row-in-group)
      *reinterpret_cast<uint32_t*>(Qb + (size_t)q * 128 + ((ch ^ (q & 7)) << 4) + within) = pk;
      *reinterpret_cast<uint32_t*>(Qb + (size_t)(64 + q) * 128 + (((4 + ch) ^ (q & 7)) << 4) + within) = pk;
      if (c4 == 0) {
        float tf = __int_as_float(0x7F800000);     // padded query slots: never reached
        float cq = 0.f;
        if (q < p.nq) {
          const uint64_t t = p.tau[q];
          tf = (t == 0) ? __int_as_float(0xFF800000) : key_score(t);
          // a zero query (am == 0) has exact scores 0 (or NaN) everywhere: Q = 0, C_q finite, handled by the test itself
          cq = (usable || am == 0.f) ? (0.5f * (float)l1q + 17.f) * kI8Slack : __int_as_float(0x7F800000);
          if (scaled) atomicMax(&s_cmax, __float_as_uint(cq));
        }
        tq[q] = make_float4(tf, (usable ? tqs : 0.f), cq, 0.f);
        s_cnt[q] = 0u;
      }
    }
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes of Qb -> tensor core
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      for (uint32_t i = n_pre; i < my_tiles; ++i) issue_tile(i);
      pdl_launch_dependents();
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // D = s32 (2), A = B = signed int8 (1), both K-major, N = 128, M = 128
      constexpr uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      const uint64_t bdesc = umma_desc_k_sw128(smem_u32(Qb));
      for (uint32_t i = 0; i < my_tiles; ++i) {
        const uint32_t buf = i & 1u, use = i >> 1;
        mbar_wait(&tempty[buf], (use & 1u) ^ 1u);
        tc_fence_after();
        const uint32_t s = i % kI8Stages, ph = (i / kI8Stages) & 1u;
        mbar_wait(&full[s], ph);
        tc_fence_after();
        const uint32_t st_addr = smem_u32(stage_base + (size_t)s * kI8StageBytes);
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          const uint32_t d_addr = tmem_base + (buf * 2u + (uint32_t)half) * 128u;
          const uint64_t adesc = umma_desc_k_sw128(st_addr + (uint32_t)half * (128 * 128));
#pragma unroll
          for (int k = 0; k < 4; ++k)   // UMMA_K = 32 B -> +2 in 16-B units
            umma_i8(d_addr, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, k != 0 ? 1u : 0u);
        }
        umma_commit(&empty[s]);
        umma_commit(&tfull[buf]);
      }
    }
  } else {
    // ---------------------------------------------------------------- epilogue: 16 warps, one matrix row per thread
    // (8 warps with two rows each ran the pass at 2900 clk per tile against 1650 clk of HBM time: ~480 dependent
    // instructions per warp and tile, two warps per scheduler, ncu i8b: stalls wait 2.1 + branch 1.8 per issue)
    const int ew = warp - 2;
    const int quarter = warp & 3, half = (ew >> 2) & 1, rr = ew >> 3;   // TMEM lanes [32 * (warp % 4), +32); columns [64 rr, +64)
    const int orow = half * 128 + quarter * 32 + lane;
    const uint32_t seg_stride = gridDim.x * p.seg_cap, seg_base = blockIdx.x * p.seg_cap;
    const float cmax = __uint_as_float(s_cmax);
    const float inv_t = __fdiv_rn(1.0f, t_pass);
    const float inf = __int_as_float(0x7F800000);
    auto run = [&](auto scaled_tag) {
      constexpr bool kScaled = decltype(scaled_tag)::value;
      for (uint32_t i = 0; i < my_tiles; ++i) {
        const uint32_t buf = i & 1u, use = i >> 1;
        const uint64_t lrow = (uint64_t)(blockIdx.x + i * gridDim.x) * kI8TileRows + 2u * (uint32_t)orow + (uint32_t)rr;
        mbar_wait(&tfull[buf], use & 1u);
        tc_fence_after();
        // (the parameters landed before the tile's `full` barrier completed, which the MMA thread observed before the
        // MMAs whose commit completed `tfull`)
        const float2 pr = prm[(size_t)(i % kI8PrmRing) * kI8TileRows + 2 * orow + rr];
        const uint32_t tcol = tmem_base + ((uint32_t)(quarter * 32) << 16) + (buf * 2u + (uint32_t)half) * 128u + (uint32_t)rr * 64u;
        uint32_t v0[32], v1[32];
        tmem_ld32_nowait(tcol, v0);
        tmem_ld32_nowait(tcol + 32u, v1);
        tmem_ld_wait32(v0);
        tmem_ld_wait32(v1);
        tc_fence_before();   // this warp's accumulators are in registers
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty[buf]);
        const float a_r = pr.x, hl_r = pr.y;     // a_r = (1 - 1e-6) / s_r: +inf for an all-zero row, NaN = always survives
        const bool valid = lrow < p.n_rows;
        const uint32_t grow = (uint32_t)(p.row_base + lrow);
        if constexpr (kScaled) {
          // T_r = (1 - 1e-6) / (s_r t) - hl_r - max C_q;  a product that overflows although s_r != 0 (s_r t underflowed)
          // gives no usable bound: the row survives (NaN -> lowest threshold)
          float T = a_r * inv_t;
          T = (a_r < inf && !(T < 1e30f)) ? __int_as_float(0x7FC00000) : T - hl_r - cmax;
          const int Ti = __float2int_ru(fminf(fmaxf(T, -2.0e9f), 2.0e9f));   // (fmaxf drops a NaN: everything survives)
          bool any[4];
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            const uint32_t* x = g < 2 ? v0 : v1;
            const int o = (g & 1) * 16;
            int m = imax3((int)x[o], (int)x[o + 1], (int)x[o + 2]);
#pragma unroll
            for (int j = 3; j < 15; j += 2) m = imax3(m, (int)x[o + j], (int)x[o + j + 1]);
            m = max(m, (int)x[o + 15]);
            any[g] = m >= Ti;
          }
          if (valid && (any[0] || any[1] || any[2] || any[3])) {
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              if (any[g]) {
                uint32_t mask = 0;
#pragma unroll
                for (int j = 0; j < 16; ++j)
                  if ((int)(g < 2 ? v0[(g & 1) * 16 + j] : v1[(g & 1) * 16 + j]) >= Ti) mask |= 1u << j;
                const int left = p.nq - g * 16;
                if (left < 16) mask &= left > 0 ? ((1u << left) - 1u) : 0u;
                while (mask) {
                  const int q = g * 16 + __ffs(mask) - 1;
                  mask &= mask - 1;
                  const uint32_t pos = atomicAdd(&s_cnt[q], 1u);
                  if (pos < p.seg_cap) p.cand_rows[(uint32_t)q * seg_stride + seg_base + pos] = grow;
                }
              }
            }
          }
        } else {
          // per-query form: survive unless  u (I + hl_r + C_q) < tau_q,  u = s_r t_q  (a badly scaled batch; not tuned)
          if (valid) {
            const float s_r = __fdiv_rn(1.0f - 1e-6f, a_r);   // (1e-7 relative: inside the 1 % slack; 0 for a zero row)
#pragma unroll
            for (int q = 0; q < kQB; ++q) {   // (unrolled: the accumulators stay in registers)
              if (q < p.nq) {
                const float4 t4 = tq[q];
                const int I = (int)(q < 32 ? v0[q & 31] : v1[q & 31]);
                const float u = s_r * t4.y;
                const bool under = s_r != 0.f && t4.y != 0.f && u < 1e-30f;
                const bool drop = !under && (u * ((float)I + (hl_r + t4.z)) < t4.x);
                if (!drop) {
                  const uint32_t pos = atomicAdd(&s_cnt[q], 1u);
                  if (pos < p.seg_cap) p.cand_rows[(uint32_t)q * seg_stride + seg_base + pos] = grow;
                }
              }
            }
          }
        }
      }
    };
    if (scaled) run(std::true_type{}); else run(std::false_type{});
    asm volatile("bar.sync 1, %0;" ::"n"(kI8EpiWarps * 32) : "memory");
    for (int q = tid - 64; q < p.nq; q += kI8EpiWarps * 32) p.seg_cnt[(size_t)q * gridDim.x + blockIdx.x] = s_cnt[q];
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512));
  }
}

// ------------------------------------------------------------------------------------------------------------------
// dim 128, up to 256 queries per pass, GROUP output (ScanParams::grp_rows) — the pass of a C5 row shard (12.5 M x 128,
// 1024 queries per step).  Over the bf16 index that pass is tensor-bound (one M 128 x N 256 x K 16 MMA = 128 clk, 2048 clk
// per 256-row tile, 0.92 of the sustained bf16 rate); kind::i8 covers K = 32 per instruction in the same 128 clk — half the
// tensor time per tile (1024 clk), below the HBM time of the tile's 32 KiB + 2 KiB (~1650 clk per SM).  An int8 row of dim
// 128 is one 128-byte swizzle row: no pairing trick.  Stages are HALF tiles (128 rows, 16 KiB); the 2 x 256 accumulator
// columns of a tile fill tensor memory as FOUR buffers (half tile x block of 128 queries, M 128 x N 128 MMAs) with a barrier
// pair each: three are in front of the tensor core while one is read.  16 epilogue warps: (half, lane quarter, query
// block) — a thread reduces 128 accumulators of its row to eight group maxima through two 32-register buffers (the next
// tcgen05.ld in flight while a buffer is reduced), tests them against T_r and records the row once per surviving group of
// 16 queries; rescore_group_kernel (recall.cu) then scores it exactly against the group's queries.
constexpr int kG8Dim = 128;
constexpr int kG8Q = 256;                        // accumulator columns per half tile (queries per pass, zero padded)
constexpr int kG8HalfRows = 128;
constexpr int kG8Stages = 8;                     // half-tile stages
constexpr int kG8StageBytes = kG8HalfRows * 128; // 16 KiB
constexpr int kG8PrmRing = 8;                   // >= kG8Stages / 2 + 3 tiles, a power of two
constexpr int kG8QBytes = kG8Q * 128;            // B operand: 256 queries x 128 B
constexpr int kG8Groups = kG8Q / kGrpQ;          // 16

constexpr size_t scan_i8g_smem_bytes() {
  return (size_t)kG8Stages * kG8StageBytes + kG8QBytes + kG8Q * 16 /*tq*/ + 64 /*s_cnt[16]*/ + (2 * kG8Stages + 8) * 8 + 16 +
         (size_t)kG8PrmRing * kTileRows * 8;
}

// I8 = false: the same kernel over the BF16 index of a dim-64 matrix (a bf16 row of dim 64 is the same 128 bytes): the
// GROUP-mode pass of a c4 row shard (64 G queries per rank).  recall_tc.cu's form of that pass has 8 epilogue warps with
// 256 accumulators per thread and tile and ran at the speed of that serial chain (c4 shard of an 8-GPU run: 33 tiles per
// SM in 63 us per pass, neither HBM- nor MMA-bound); here 16 warps take 128 each.  Filter test as in recall_tc.cu:
// approx' < t_r = (1 - 1e-6) - ||x_r|| max_q(c ||q'_q||), c = 1.05 * 2^-8 (per-query form for badly scaled batches).
constexpr float kG16Margin = 1.05f / 256.f;
template <bool I8>
__global__ void __launch_bounds__(kI8Threads, 1)
recall_scan_grp_kernel(const __grid_constant__ CUtensorMap emap, const ScanParams p, const uint32_t n_tiles) {
  constexpr int kDim = I8 ? kG8Dim : 64;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* stage_base = smem;
  uint8_t* Qb = smem + (size_t)kG8Stages * kG8StageBytes;
  float4* tq = reinterpret_cast<float4*>(Qb + kG8QBytes);      // [256] {tau_f, t_q, C_q, -}
  uint32_t* s_cnt = reinterpret_cast<uint32_t*>(tq + kG8Q);    // [16] records per group
  uint64_t* full = reinterpret_cast<uint64_t*>(s_cnt + 16);
  uint64_t* empty = full + kG8Stages;
  uint64_t* tfull = empty + kG8Stages;     // [4] per (half tile, block of 128 queries): accumulator buffer 2 half + block
  uint64_t* tempty = tfull + 4;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 4);
  float2* prm = reinterpret_cast<float2*>(tmem_slot + 4);      // [kG8PrmRing][256] {a_r, hl_r} (bf16: 256 norm bounds f32)
  __shared__ float s_scale[kG8Q];
  __shared__ float s_amax[kG8Q];
  __shared__ uint32_t s_tmax, s_cmax;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    s_tmax = 0u; s_cmax = 0u;
    tma_prefetch_desc(&emap);
    for (int s = 0; s < kG8Stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int b = 0; b < 4; ++b) { mbar_init(&tfull[b], 1); mbar_init(&tempty[b], kI8EpiWarps / 4); }
    mbar_fence_init();
  }
  const uint32_t my_tiles = (n_tiles > blockIdx.x) ? (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  auto issue_tile = [&](uint32_t i) {
    const uint32_t t = blockIdx.x + i * gridDim.x;
#pragma unroll
    for (uint32_t hf = 0; hf < 2; ++hf) {
      const uint32_t it = 2 * i + hf;
      const uint32_t s = it % kG8Stages, ph = (it / kG8Stages) & 1u;
      mbar_wait(&empty[s], ph ^ 1u);
      constexpr uint32_t kPrmBytes = kTileRows * (I8 ? 8 : 4);
      mbar_arrive_expect_tx(&full[s], kG8StageBytes + (hf == 0 ? kPrmBytes : 0));
      if (hf == 0) {   // (the parameter / norm arrays are padded to whole tiles: build_i8_index, build_row_norms)
        if constexpr (I8) bulk_load_1d(prm + (size_t)(i % kG8PrmRing) * kTileRows, p.row_q8 + (size_t)t * kTileRows, kPrmBytes, &full[s]);
        else bulk_load_1d(prm + (size_t)(i % kG8PrmRing) * kTileRows, p.row_norm + (size_t)t * kTileRows, kPrmBytes, &full[s]);
      }
      tma_load_2d(stage_base + (size_t)s * kG8StageBytes, &emap, 0, (int)(t * (uint32_t)kTileRows + hf * kG8HalfRows), &full[s],
                  kEvictFirst);
    }
  };
  const uint32_t n_pre = my_tiles < (uint32_t)(kG8Stages / 2) ? my_tiles : (uint32_t)(kG8Stages / 2);
  if (tid == 0)
    for (uint32_t i = 0; i < n_pre; ++i) issue_tile(i);
  for (int i = tid; i < kG8QBytes / 16; i += kI8Threads) reinterpret_cast<uint4*>(Qb)[i] = make_uint4(0u, 0u, 0u, 0u);
  if (tid < 16) s_cnt[tid] = 0u;
  pdl_wait();

  // ---- thresholds and queries, as in recall_scan_i8_kernel; a query is 32 float4 pieces = one warp, read twice
  // (maxima first, quantisation second: 256 queries do not fit in registers across the block-wide decision)
  int bad = 0;
  for (int q = tid; q < kG8Q; q += kI8Threads) {
    float sc = 0.f;
    if (q < p.nq) {
      const uint64_t t = p.tau[q];
      const float tf = (t == 0) ? __int_as_float(0xFF800000) : key_score(t);
      if (tf > 0x1p-60f && tf < 0x1p60f) sc = 1.0f / tf; else bad = 1;
    }
    s_scale[q] = sc;
  }
  bool scaled;
  float t_pass = 0.f;
  if constexpr (I8) {
  constexpr int kWarps = kI8Threads / 32;   // 18
  for (int q = warp; q < kG8Q; q += kWarps) {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (q < p.nq) v = __ldg(reinterpret_cast<const float4*>(p.Q + (size_t)q * kG8Dim) + lane);
    float am = fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w)));
    float l1 = fabsf(v.x) + fabsf(v.y) + fabsf(v.z) + fabsf(v.w);
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      am = fmaxf(am, __shfl_xor_sync(0xffffffffu, am, off));
      l1 += __shfl_xor_sync(0xffffffffu, l1, off);
    }
    if (!(l1 < __int_as_float(0x7F800000))) { bad = 1; am = __int_as_float(0x7FC00000); }   // NaN / inf element
    if (lane == 0) s_amax[q] = am;
  }
  scaled = __syncthreads_or(bad) == 0;
  if (scaled)
    for (int q = tid; q < p.nq && q < kG8Q; q += kI8Threads) atomicMax(&s_tmax, __float_as_uint(s_amax[q] * s_scale[q]));
  __syncthreads();
  t_pass = __fdiv_rn(__uint_as_float(s_tmax), 127.f);
  for (int q = warp; q < kG8Q; q += kWarps) {
    float4 w = make_float4(0.f, 0.f, 0.f, 0.f);
    if (q < p.nq) w = __ldg(reinterpret_cast<const float4*>(p.Q + (size_t)q * kG8Dim) + lane);
    const float am = s_amax[q];
    float tqs;
    if (scaled) { const float sc = s_scale[q]; w.x *= sc; w.y *= sc; w.z *= sc; w.w *= sc; tqs = t_pass; }
    else tqs = __fdiv_rn(am, 127.f);
    const bool usable = am == am && tqs >= 1e-30f;
    int e0 = 0, e1 = 0, e2 = 0, e3 = 0;
    if (usable) {
      e0 = max(-127, min(127, __float2int_rn(__fdiv_rn(w.x, tqs))));
      e1 = max(-127, min(127, __float2int_rn(__fdiv_rn(w.y, tqs))));
      e2 = max(-127, min(127, __float2int_rn(__fdiv_rn(w.z, tqs))));
      e3 = max(-127, min(127, __float2int_rn(__fdiv_rn(w.w, tqs))));
    }
    int l1q = abs(e0) + abs(e1) + abs(e2) + abs(e3);
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) l1q += __shfl_xor_sync(0xffffffffu, l1q, off);
    const uint32_t pk = (uint32_t)(e0 & 255) | ((uint32_t)(e1 & 255) << 8) | ((uint32_t)(e2 & 255) << 16) | ((uint32_t)(e3 & 255) << 24);
    const int dd = lane * 4, ch = dd >> 4, within = dd & 15;     // SWIZZLE_128B: 16-B chunk index XOR row-in-group
    *reinterpret_cast<uint32_t*>(Qb + (size_t)q * 128 + ((ch ^ (q & 7)) << 4) + within) = pk;
    if (lane == 0) {
      float tf = __int_as_float(0x7F800000), cq = 0.f;   // padded query slots: never reached
      if (q < p.nq) {
        const uint64_t t = p.tau[q];
        tf = (t == 0) ? __int_as_float(0xFF800000) : key_score(t);
        cq = (usable || am == 0.f) ? (0.5f * (float)l1q + (float)(kG8Dim / 4 + 1)) * kI8Slack : __int_as_float(0x7F800000);
        if (scaled) atomicMax(&s_cmax, __float_as_uint(cq));
      }
      tq[q] = make_float4(tf, (usable ? tqs : 0.f), cq, 0.f);
    }
  }
  } else {
    // bf16 operands, as in recall_tc.cu: the queries (scaled by 1 / tau in a uniform pass) rounded to bf16 into the
    // K-major SWIZZLE_128B B operand, c ||q'|| per query, its maximum for the uniform threshold
    scaled = __syncthreads_or(bad) == 0;
    constexpr int C4 = 16;                                  // 16-B pieces of a 64-d f32 query = lanes per query
    constexpr int kIters = (C4 * kG8Q + kI8Threads - 1) / kI8Threads;
    float* s_ss = s_amax;                                   // squared norms of the staged queries
#pragma unroll 4
    for (int itq = 0; itq < kIters; ++itq) {
      const int i4 = tid + itq * kI8Threads;
      const bool in = i4 < C4 * kG8Q;
      const int q = in ? i4 / C4 : 0, c4 = i4 - q * C4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (in && q < p.nq) v = __ldg(reinterpret_cast<const float4*>(p.Q + (size_t)q * 64) + c4);
      if (scaled) { const float sc = s_scale[q]; v.x *= sc; v.y *= sc; v.z *= sc; v.w *= sc; }
      float ss = fmaf(v.x, v.x, fmaf(v.y, v.y, fmaf(v.z, v.z, v.w * v.w)));
#pragma unroll
      for (int off = 1; off < C4; off <<= 1) ss += __shfl_xor_sync(0xffffffffu, ss, off);   // lanes of one query only
      if (in) {
        const int e = c4 * 4, ch = e >> 3;                  // 4 bf16 = half of a 16-B chunk
        __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
        uint2 pk;
        pk.x = *reinterpret_cast<uint32_t*>(&lo);
        pk.y = *reinterpret_cast<uint32_t*>(&hi);
        *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(Qb) + (size_t)q * 64 + ((ch ^ (q & 7)) << 3) + (e & 7)) = pk;
        if (c4 == 0) s_ss[q] = ss;
      }
    }
    __syncthreads();
    for (int q = tid; q < kG8Q; q += kI8Threads) {
      float tf = __int_as_float(0x7F800000), nq2 = 0.f;     // padded query slots: never reached
      if (q < p.nq) {
        const uint64_t t = p.tau[q];
        tf = (t == 0) ? __int_as_float(0xFF800000) : key_score(t);
        nq2 = (sqrtf(s_ss[q]) * 1.0001f + 1e-30f) * kG16Margin;
        if (scaled) atomicMax(&s_cmax, __float_as_uint(nq2));   // non-negative floats order like their bit patterns
      }
      tq[q] = make_float4(tf, nq2, 0.f, 0.f);
    }
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      for (uint32_t i = n_pre; i < my_tiles; ++i) issue_tile(i);
      pdl_launch_dependents();
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // D = s32, A = B = signed int8, both K-major, N = 128, M = 128.  A half tile's 256 accumulator columns are TWO buffers
      // of 128 queries with a barrier pair each: with one buffer per half (N = 256) the half's next MMAs waited for all of
      // its 256 columns to be read — a load phase of ~700 clk behind every 600 clk of MMAs: 2560 clk per tile, tensor pipe
      // 47 % busy (ncu i8h).  Four buffers in rotation keep three of them in front of the tensor core while one is read.
      // (bf16: D = f32 (1), A = B = bf16 (1), K = 16 per instruction — the same 32 bytes)
      constexpr uint32_t idesc = I8 ? ((2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24))
                                    : ((1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24));
      const uint32_t qb_addr = smem_u32(Qb);
      const int n_qb = p.nq > 128 ? 2 : 1;   // a pass of at most 128 queries leaves the second block's buffers alone
      uint32_t it = 0;
      for (uint32_t i = 0; i < my_tiles; ++i) {
#pragma unroll
        for (int half = 0; half < 2; ++half, ++it) {
          const uint32_t s = it % kG8Stages, ph = (it / kG8Stages) & 1u;
          mbar_wait(&full[s], ph);
          tc_fence_after();
          const uint64_t adesc = umma_desc_k_sw128(smem_u32(stage_base + (size_t)s * kG8StageBytes));
#pragma unroll
          for (int qb = 0; qb < 2; ++qb) {
            if (qb >= n_qb) break;
            const int buf = half * 2 + qb;
            mbar_wait(&tempty[buf], (i & 1u) ^ 1u);   // this buffer's four epilogue warps have drained the previous tile
            tc_fence_after();
            const uint32_t d_addr = tmem_base + (uint32_t)buf * 128u;
            const uint64_t bdesc = umma_desc_k_sw128(qb_addr + (uint32_t)qb * (128u * 128u));   // queries [128 qb, +128)
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              if constexpr (I8) umma_i8(d_addr, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, k != 0 ? 1u : 0u);
              else umma_bf16(d_addr, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, k != 0 ? 1u : 0u);
            }
            umma_commit(&tfull[buf]);
          }
          umma_commit(&empty[s]);
        }
      }
    }
  } else {
    // ---------------------------------------------------------------- epilogue: 16 warps = 2 halves x 4 lane quarters x 2 column halves
    const int ew = warp - 2;
    const int quarter = warp & 3, half = (ew >> 2) & 1, cpart = ew >> 3;
    const int row_local = half * 128 + quarter * 32 + lane;
    const uint32_t grp_stride = gridDim.x * p.grp_cap;
    uint32_t* const grp_base = p.grp_rows + (size_t)blockIdx.x * p.grp_cap + (size_t)(cpart * 8) * grp_stride;   // this warp's first group
    const int n_grp = ((p.nq + kQB - 1) / kQB) * 4;                 // groups that have lists (whole blocks of 64 queries)
    const int my_grp = n_grp - cpart * 8;                           // how many of this warp's eight groups exist
    const uint32_t grp_mask = my_grp >= 8 ? 0xFFu : (my_grp > 0 ? ((1u << my_grp) - 1u) : 0u);
    const float cmax = __uint_as_float(s_cmax);
    const float inv_t = __fdiv_rn(1.0f, t_pass);
    const float inf = __int_as_float(0x7F800000);
    // (row indices in 32 bits: a shard holds fewer than 2^31 rows, recall_topk_device)
    const uint32_t n_rows32 = (uint32_t)p.n_rows, row_base32 = (uint32_t)p.row_base, row_step = gridDim.x * (uint32_t)kTileRows;
    const uint32_t tcol = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)half * (uint32_t)kG8Q + (uint32_t)cpart * 128u;
    uint64_t* const my_tfull = &tfull[half * 2 + cpart];
    uint64_t* const my_tempty = &tempty[half * 2 + cpart];
    uint32_t* const my_cnt = s_cnt + cpart * 8;
    const uint32_t below = (1u << lane) - 1u;
    auto run = [&](auto scaled_tag) {
      constexpr bool kScaled = decltype(scaled_tag)::value;
      uint32_t lrow = blockIdx.x * (uint32_t)kTileRows + (uint32_t)row_local;
      for (uint32_t i = 0; i < my_tiles; ++i, lrow += row_step) {
        mbar_wait(my_tfull, i & 1u);
        tc_fence_after();
        float a_r, hl_r = 0.f;     // bf16: a_r = the row's norm bound
        if constexpr (I8) {
          const float2 pr = prm[(i & (uint32_t)(kG8PrmRing - 1)) * (uint32_t)kTileRows + (uint32_t)row_local];
          a_r = pr.x; hl_r = pr.y;
        } else {
          a_r = reinterpret_cast<const float*>(prm + (i & (uint32_t)(kG8PrmRing - 1)) * (uint32_t)kTileRows)[row_local];
        }
        const bool valid = lrow < n_rows32;
        const uint32_t grow = row_base32 + lrow;
        int Ti = 0;
        float s_r = 0.f, t_r = 0.f;
        if constexpr (!I8) {
          t_r = fmaf(-a_r, cmax, 1.0f - 1e-6f);   // (a row whose norm bound is +inf has t_r = -inf or NaN and always survives)
        } else if constexpr (kScaled) {
          float T = a_r * inv_t;
          T = (a_r < inf && !(T < 1e30f)) ? __int_as_float(0x7FC00000) : T - hl_r - cmax;
          Ti = __float2int_ru(fminf(fmaxf(T, -2.0e9f), 2.0e9f));
        } else {
          s_r = __fdiv_rn(1.0f - 1e-6f, a_r);
        }
        uint32_t hits = 0u;
        uint32_t x[32], y[32];
        tmem_ld32_nowait(tcol, x);
#pragma unroll
        for (int c = 0; c < 4; ++c) {               // 32 columns = two groups per chunk
          uint32_t (&cur)[32] = (c & 1) ? y : x;
          uint32_t (&nxt)[32] = (c & 1) ? x : y;
          tmem_ld_wait32(cur);
          if (c + 1 < 4) tmem_ld32_nowait(tcol + (uint32_t)(c + 1) * 32u, nxt);
          else {             // every accumulator of this warp is in registers
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(my_tempty);
          }
#pragma unroll
          for (int g = 0; g < 2; ++g) {
            const int o = g * 16;
            if constexpr (!I8 && kScaled) {
              const float f0 = fmax3f(__uint_as_float(cur[o]), __uint_as_float(cur[o + 1]), __uint_as_float(cur[o + 2]));
              const float f1 = fmax3f(__uint_as_float(cur[o + 3]), __uint_as_float(cur[o + 4]), __uint_as_float(cur[o + 5]));
              const float f2 = fmax3f(__uint_as_float(cur[o + 6]), __uint_as_float(cur[o + 7]), __uint_as_float(cur[o + 8]));
              const float f3 = fmax3f(__uint_as_float(cur[o + 9]), __uint_as_float(cur[o + 10]), __uint_as_float(cur[o + 11]));
              const float f4 = fmax3f(__uint_as_float(cur[o + 12]), __uint_as_float(cur[o + 13]), __uint_as_float(cur[o + 14]));
              const float f = fmaxf(fmax3f(f0, f1, f2), fmax3f(f3, f4, __uint_as_float(cur[o + 15])));
              hits |= !(f < t_r) ? (1u << (c * 2 + g)) : 0u;   // (NaN accumulators are ignored by max; NaN maxima and t_r survive)
            } else if constexpr (!I8) {
              // per-query form: survive unless approx < tau_q - ||x_r|| c ||q||
              bool any = false;
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                const float4 t4 = tq[cpart * 128 + c * 32 + o + j];
                any |= !(__uint_as_float(cur[o + j]) < fmaf(-a_r, t4.y, t4.x));
              }
              hits |= any ? (1u << (c * 2 + g)) : 0u;
            } else if constexpr (kScaled) {
              // a tree, not a chain: five independent maxima of three, then two levels (the chain's eight dependent
              // VIMNMX3 were a fixed-latency stall each)
              const int m0 = imax3((int)cur[o], (int)cur[o + 1], (int)cur[o + 2]);
              const int m1 = imax3((int)cur[o + 3], (int)cur[o + 4], (int)cur[o + 5]);
              const int m2 = imax3((int)cur[o + 6], (int)cur[o + 7], (int)cur[o + 8]);
              const int m3 = imax3((int)cur[o + 9], (int)cur[o + 10], (int)cur[o + 11]);
              const int m4 = imax3((int)cur[o + 12], (int)cur[o + 13], (int)cur[o + 14]);
              const int m = imax3(imax3(m0, m1, m2), imax3(m3, m4, (int)cur[o + 15]), (int)0x80000000);
              hits |= (m >= Ti) ? (1u << (c * 2 + g)) : 0u;
            } else {
              // per-query form: survive unless  u (I + hl_r + C_q) < tau_q,  u = s_r t_q  (a badly scaled batch; not tuned)
              bool any = false;
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                const float4 t4 = tq[cpart * 128 + c * 32 + o + j];
                const float u = s_r * t4.y;
                const bool under = s_r != 0.f && t4.y != 0.f && u < 1e-30f;
                any |= under || !(u * ((float)(int)cur[o + j] + (hl_r + t4.z)) < t4.x);
              }
              hits |= any ? (1u << (c * 2 + g)) : 0u;
            }
          }
        }
        hits = valid ? hits & grp_mask : 0u;
        // appends, warp-aggregated: one shared-memory atomic per group that has a survivor in this warp's 32 rows (most
        // tiles have none, the rest usually one) instead of eight predicated atomics + stores per tile
        uint32_t todo = __reduce_or_sync(0xffffffffu, hits);
        while (todo) {   // warp-uniform
          const int b = __ffs(todo) - 1;
          todo &= todo - 1;
          const bool mine = (hits >> b) & 1u;
          const uint32_t bal = __ballot_sync(0xffffffffu, mine);
          const int leader = __ffs(bal) - 1;
          uint32_t start = 0;
          if ((int)lane == leader) start = atomicAdd(&my_cnt[b], (uint32_t)__popc(bal));
          start = __shfl_sync(0xffffffffu, start, leader);
          const uint32_t pos = start + __popc(bal & below);
          if (mine && pos < p.grp_cap) grp_base[(uint32_t)b * grp_stride + pos] = grow;
        }
      }
    };
    if (cpart == 1 && p.nq <= 128) { /* the MMA thread leaves the second query block alone */ }
    else if (scaled) run(std::true_type{}); else run(std::false_type{});
    asm volatile("bar.sync 1, %0;" ::"n"(kI8EpiWarps * 32) : "memory");
    const int n_grp_w = ((p.nq + kQB - 1) / kQB) * 4;
    for (int g = tid - 64; g < n_grp_w; g += kI8EpiWarps * 32) p.grp_cnt[(size_t)g * gridDim.x + blockIdx.x] = s_cnt[g];
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512));
  }
}

// One thread per row: scale, int8 row, {a_r, hl_r}.  (Run once per matrix; the strided reads do not matter.)
template <int DIM>
__global__ void i8_index_kernel(const float* __restrict__ E, uint64_t rows, uint8_t* __restrict__ out8, float2* __restrict__ prm) {
  const uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  const float4* x = reinterpret_cast<const float4*>(E + r * DIM);
  float am = 0.f, l1 = 0.f;
#pragma unroll 8
  for (int i = 0; i < DIM / 4; ++i) {
    const float4 v = x[i];
    am = fmaxf(am, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
    l1 += fabsf(v.x) + fabsf(v.y) + fabsf(v.z) + fabsf(v.w);
  }
  const float s = __fdiv_rn(am, 127.f);
  const bool finite = l1 < __int_as_float(0x7F800000) || (am < __int_as_float(0x7F800000) && l1 == l1);   // l1 may overflow for finite rows
  // NaN / inf elements, or a scale that underflowed: no bound, the row always survives (a_r = NaN)
  const bool usable = finite && (am == 0.f || s >= 1e-30f);
  int L1 = 0;
  uint4* o = reinterpret_cast<uint4*>(out8 + r * DIM);
#pragma unroll 2
  for (int i4 = 0; i4 < DIM / 16; ++i4) {     // (the row is read a second time: L1 / L2 hits)
    uint32_t pk[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float4 v = x[i4 * 4 + i];
      int e0 = 0, e1 = 0, e2 = 0, e3 = 0;
      if (usable && am != 0.f) {
        e0 = max(-127, min(127, __float2int_rn(__fdiv_rn(v.x, s))));
        e1 = max(-127, min(127, __float2int_rn(__fdiv_rn(v.y, s))));
        e2 = max(-127, min(127, __float2int_rn(__fdiv_rn(v.z, s))));
        e3 = max(-127, min(127, __float2int_rn(__fdiv_rn(v.w, s))));
      }
      L1 += abs(e0) + abs(e1) + abs(e2) + abs(e3);
      pk[i] = (uint32_t)(e0 & 255) | ((uint32_t)(e1 & 255) << 8) | ((uint32_t)(e2 & 255) << 16) | ((uint32_t)(e3 & 255) << 24);
    }
    o[i4] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
  }
  // {a_r, hl_r}: a_r = (1 - 1e-6) / s_r (+inf for an all-zero row), NaN = no usable bound, the row always survives
  const float nan = __int_as_float(0x7FC00000);
  prm[r] = usable ? make_float2(__fdiv_rn(1.0f - 1e-6f, s), 0.5f * (float)L1 * kI8Slack) : make_float2(nan, nan);
}

// config "scan_int8" (default on): the int8 index exists for matrices large enough for the sampled path — at dim 64 for
// the passes of at most 64 queries, at dim 128 for the GROUP-mode passes of more than 64
bool scan_i8_wanted(const prg_handle* h) {
  return h->scan_int8 && h->scan_filter == SCAN_FILTER_BF16 && !h->scan_ffma2 && (h->E_dim == kI8Dim || h->E_dim == kG8Dim) &&
         h->E_rows >= (uint64_t)kI8TileRows * (uint64_t)h->sm_count;
}
bool scan_i8_available(const prg_handle* h) { return h->E8_map_ok && h->E_dim == kI8Dim; }
bool scan_i8g_available(const prg_handle* h) { return h->E8_map_ok && h->E_dim == kG8Dim; }

int build_i8_index(prg_handle* h) {
  h->E8_map_ok = false;
  if (!scan_i8_wanted(h)) return PRG_OK;
  const size_t padded = (size_t)((h->E_rows + kI8TileRows - 1) / kI8TileRows) * kI8TileRows;
  const size_t dim = h->E_dim;
  PRG_TRY(h->E8.ensure(padded * dim));
  PRG_TRY(h->E8_prm.ensure(padded * 8));
  PRG_CUDA(cudaMemsetAsync(h->E8.p, 0, padded * dim, h->stream));
  PRG_CUDA(cudaMemsetAsync(h->E8_prm.p, 0, padded * 8, h->stream));
  const unsigned grid = (unsigned)((h->E_rows + 127) / 128);
  if (dim == kI8Dim) i8_index_kernel<kI8Dim><<<grid, 128, 0, h->stream>>>(h->E, h->E_rows, (uint8_t*)h->E8.p, (float2*)h->E8_prm.p);
  else i8_index_kernel<kG8Dim><<<grid, 128, 0, h->stream>>>(h->E, h->E_rows, (uint8_t*)h->E8.p, (float2*)h->E8_prm.p);
  PRG_CUDA(cudaGetLastError());
  PRG_CUDA(cudaStreamSynchronize(h->stream));
  count_launch(h);
  PFN_encodeTiled enc = get_encode_tiled();
  if (!enc) return fail(PRG_ECUDA, "cuTensorMapEncodeTiled entry point not available");
  // dim 64: [rows / 2][128], boxes of 256 operand rows (512 matrix rows); dim 128: [rows][128], boxes of half a tile
  cuuint64_t gdim[2] = {128, (cuuint64_t)(dim == kI8Dim ? padded / 2 : padded)};
  cuuint64_t gstride[1] = {128};
  cuuint32_t box[2] = {128, (cuuint32_t)(dim == kI8Dim ? kI8BoxRows : kG8HalfRows)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(&h->E8_map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, h->E8.p, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(PRG_ECUDA, "cuTensorMapEncodeTiled (int8 index) failed: " + std::to_string((int)r));
  h->E8_map_ok = true;
  return PRG_OK;
}

// One pass of <= 64 queries over the int8 index; same outputs as launch_scan_tc (cand_rows / seg_cnt, one segment per CTA
// and query; the grid is the caller's number of segments)
int launch_scan_i8(prg_handle* h, const ScanParams& p_in, uint32_t n_seg) {
  if (!scan_i8_available(h)) return fail(PRG_ESTATE, "int8 filter index not built");
  if (p_in.nq > kQB || p_in.nq <= 0) return fail(PRG_EINVAL, "launch_scan_i8: 1..64 queries per pass");
  const uint32_t n_tiles = (uint32_t)((h->E_rows + kI8TileRows - 1) / kI8TileRows);
  if (n_seg == 0 || n_seg > n_tiles) return fail(PRG_EINVAL, "launch_scan_i8: more segments than tiles");
  ScanParams p = p_in;
  p.row_q8 = (const float2*)h->E8_prm.p;
  constexpr size_t smem = scan_i8_smem_bytes();
  PRG_CUDA(cudaFuncSetAttribute(recall_scan_i8_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  StageScope span(h, ST_SCAN);
  PRG_CUDA(launch_chained(h, recall_scan_i8_kernel, dim3(n_seg), dim3(kI8Threads), smem, 1, h->E8_map, p, n_tiles));
  count_launch(h);
  return PRG_OK;
}

// One GROUP-mode pass of <= 256 queries over the bf16 index of a dim-64 matrix (same kernel, bf16 operands)
int launch_scan_g16(prg_handle* h, const ScanParams& p, uint32_t n_seg) {
  if (h->E_dim != 64 || !h->E16_map_ok) return fail(PRG_ESTATE, "bf16 filter index (dim 64) not built");
  if (p.nq > kG8Q || p.nq <= 0) return fail(PRG_EINVAL, "launch_scan_g16: 1..256 queries per pass");
  if (!p.grp_rows || !p.grp_cnt || !p.row_norm) return fail(PRG_EINVAL, "launch_scan_g16: group outputs / norms missing");
  const uint32_t n_tiles = (uint32_t)((h->E_rows + kTileRows - 1) / kTileRows);
  if (n_seg == 0 || n_seg > n_tiles) return fail(PRG_EINVAL, "launch_scan_g16: more segments than tiles");
  constexpr size_t smem = scan_i8g_smem_bytes();
  PRG_CUDA(cudaFuncSetAttribute(recall_scan_grp_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  StageScope span(h, ST_SCAN);
  PRG_CUDA(launch_chained(h, recall_scan_grp_kernel<false>, dim3(n_seg), dim3(kI8Threads), smem, 1, h->E16_map_h, p, n_tiles));
  count_launch(h);
  return PRG_OK;
}

// One GROUP-mode pass of <= 256 queries over the dim-128 int8 index; same outputs as launch_scan_tc with p.grp_rows set
int launch_scan_i8g(prg_handle* h, const ScanParams& p_in, uint32_t n_seg) {
  if (!scan_i8g_available(h)) return fail(PRG_ESTATE, "int8 filter index (dim 128) not built");
  if (p_in.nq > kG8Q || p_in.nq <= 0) return fail(PRG_EINVAL, "launch_scan_i8g: 1..256 queries per pass");
  if (!p_in.grp_rows || !p_in.grp_cnt) return fail(PRG_EINVAL, "launch_scan_i8g: group outputs missing");
  const uint32_t n_tiles = (uint32_t)((h->E_rows + kTileRows - 1) / kTileRows);
  if (n_seg == 0 || n_seg > n_tiles) return fail(PRG_EINVAL, "launch_scan_i8g: more segments than tiles");
  ScanParams p = p_in;
  p.row_q8 = (const float2*)h->E8_prm.p;
  constexpr size_t smem = scan_i8g_smem_bytes();
  PRG_CUDA(cudaFuncSetAttribute(recall_scan_grp_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  StageScope span(h, ST_SCAN);
  PRG_CUDA(launch_chained(h, recall_scan_grp_kernel<true>, dim3(n_seg), dim3(kI8Threads), smem, 1, h->E8_map, p, n_tiles));
  count_launch(h);
  return PRG_OK;
}

}  // namespace prg
