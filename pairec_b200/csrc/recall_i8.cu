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

// One thread per row: scale, int8 row, {a_r, hl_r}.  (Run once per matrix; the strided reads do not matter.)
__global__ void i8_index_kernel(const float* __restrict__ E, uint64_t rows, uint8_t* __restrict__ out8, float2* __restrict__ prm) {
  const uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  const float4* x = reinterpret_cast<const float4*>(E + r * kI8Dim);
  float4 v[kI8Dim / 4];
  float am = 0.f, l1 = 0.f;
#pragma unroll
  for (int i = 0; i < kI8Dim / 4; ++i) {
    v[i] = x[i];
    am = fmaxf(am, fmaxf(fmaxf(fabsf(v[i].x), fabsf(v[i].y)), fmaxf(fabsf(v[i].z), fabsf(v[i].w))));
    l1 += fabsf(v[i].x) + fabsf(v[i].y) + fabsf(v[i].z) + fabsf(v[i].w);
  }
  float s = __fdiv_rn(am, 127.f);
  const bool finite = l1 < __int_as_float(0x7F800000) || (am < __int_as_float(0x7F800000) && l1 == l1);   // l1 may overflow for finite rows
  // NaN / inf elements, or a scale that underflowed: no bound, the row always survives (a_r = NaN)
  const bool usable = finite && (am == 0.f || s >= 1e-30f);
  uint32_t pk[kI8Dim / 4];
  int L1 = 0;
#pragma unroll
  for (int i = 0; i < kI8Dim / 4; ++i) {
    int e0 = 0, e1 = 0, e2 = 0, e3 = 0;
    if (usable && am != 0.f) {
      e0 = max(-127, min(127, __float2int_rn(__fdiv_rn(v[i].x, s))));
      e1 = max(-127, min(127, __float2int_rn(__fdiv_rn(v[i].y, s))));
      e2 = max(-127, min(127, __float2int_rn(__fdiv_rn(v[i].z, s))));
      e3 = max(-127, min(127, __float2int_rn(__fdiv_rn(v[i].w, s))));
    }
    L1 += abs(e0) + abs(e1) + abs(e2) + abs(e3);
    pk[i] = (uint32_t)(e0 & 255) | ((uint32_t)(e1 & 255) << 8) | ((uint32_t)(e2 & 255) << 16) | ((uint32_t)(e3 & 255) << 24);
  }
  uint4* o = reinterpret_cast<uint4*>(out8 + r * kI8Dim);
#pragma unroll
  for (int i = 0; i < kI8Dim / 16; ++i) o[i] = make_uint4(pk[4 * i], pk[4 * i + 1], pk[4 * i + 2], pk[4 * i + 3]);
  // {a_r, hl_r}: a_r = (1 - 1e-6) / s_r (+inf for an all-zero row), NaN = no usable bound, the row always survives
  const float nan = __int_as_float(0x7FC00000);
  prm[r] = usable ? make_float2(__fdiv_rn(1.0f - 1e-6f, s), 0.5f * (float)L1 * kI8Slack) : make_float2(nan, nan);
}

// config "scan_int8" (default on): the int8 index exists for dim-64 matrices large enough for the sampled path
bool scan_i8_wanted(const prg_handle* h) {
  return h->scan_int8 && h->scan_filter == SCAN_FILTER_BF16 && !h->scan_ffma2 && h->E_dim == kI8Dim &&
         h->E_rows >= (uint64_t)kI8TileRows * (uint64_t)h->sm_count;
}
bool scan_i8_available(const prg_handle* h) { return h->E8_map_ok; }

int build_i8_index(prg_handle* h) {
  h->E8_map_ok = false;
  if (!scan_i8_wanted(h)) return PRG_OK;
  const size_t padded = (size_t)((h->E_rows + kI8TileRows - 1) / kI8TileRows) * kI8TileRows;
  PRG_TRY(h->E8.ensure(padded * kI8Dim));
  PRG_TRY(h->E8_prm.ensure(padded * 8));
  PRG_CUDA(cudaMemsetAsync(h->E8.p, 0, padded * kI8Dim, h->stream));
  PRG_CUDA(cudaMemsetAsync(h->E8_prm.p, 0, padded * 8, h->stream));
  const unsigned grid = (unsigned)((h->E_rows + 127) / 128);
  i8_index_kernel<<<grid, 128, 0, h->stream>>>(h->E, h->E_rows, (uint8_t*)h->E8.p, (float2*)h->E8_prm.p);
  PRG_CUDA(cudaGetLastError());
  PRG_CUDA(cudaStreamSynchronize(h->stream));
  count_launch(h);
  PFN_encodeTiled enc = get_encode_tiled();
  if (!enc) return fail(PRG_ECUDA, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t gdim[2] = {128, (cuuint64_t)(padded / 2)};
  cuuint64_t gstride[1] = {128};
  cuuint32_t box[2] = {128, (cuuint32_t)kI8BoxRows};
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
  if (!h->E8_map_ok) return fail(PRG_ESTATE, "int8 filter index not built");
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

}  // namespace prg
