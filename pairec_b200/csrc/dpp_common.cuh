// dpp_common.cuh — declarations shared by the two fast DPP kernels (dpp_cluster.cu: 4-CTA cluster per request with
// resident features; dpp_lazy.cu: one CTA per request with lazy evaluation).
#pragma once
#include "handle.h"
#include <math_constants.h>

namespace prg {

constexpr int kClCtas = 4;
constexpr int kClItems = 256;                      // candidates per CTA
constexpr int kClThreads = 2 * kClItems;           // two threads per candidate (k blocks split)
constexpr int kClMaxItems = kClCtas * kClItems;    // 1024
constexpr int kClMaxN = 4096;
constexpr int kClCRows = 24;
constexpr double kInvSqrt2c = 0.70710678118654752440;

// A candidate WITHOUT a row in the diversity table: the reference gives it an (unseeded) random unit vector
// (sort/dpp_sort.go:250-262) — not reproducible by construction.  Here it gets row (input position & 1023) of a fixed
// table of 1024 pseudo-random directions built when the diversity matrix is set (dpp_substitute_kernel: 24-bit dyadic
// values from splitmix64, the same on the CPU oracle), so it competes like any other item and the result is
// deterministic.  Row codes: < D_rows a table row, kDppMissBase + s a substitute row, 0xFFFFFFFF not a candidate.
constexpr uint32_t kDppSubRows = 1024;
constexpr uint32_t kDppMissBase = 0xFFFFF000u;
__host__ __device__ __forceinline__ float dpp_substitute_value(uint32_t sub_row, uint32_t d) {
  uint64_t z = (((uint64_t)sub_row << 32) | d) + 0x9E3779B97F4A7C15ull;   // splitmix64
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z = z ^ (z >> 31);
  return (float)((int32_t)(z >> 40) - (1 << 23)) * (1.0f / (float)(1 << 23));   // [-1, 1), exact in f32
}
__device__ __forceinline__ uint32_t dpp_row_code(bool act, uint32_t r, uint64_t D_rows, int in_idx) {
  if (!act) return 0xFFFFFFFFu;
  return (uint64_t)r < D_rows ? r : kDppMissBase + ((uint32_t)in_idx & (kDppSubRows - 1));
}

struct DppClArgs {
  const uint32_t* rows;
  const double* score;
  int n;
  const float* D;
  const double* D_inv;  // 1 / ||row|| per table row (dpp_inv_norm_kernel)
  const float* D_sub;       // [kDppSubRows][D] substitute directions for candidates without a table row
  const double* D_sub_inv;  // their 1 / ||row||
  uint64_t D_rows;
  prg_dpp_params p;
  int32_t* out_idx;
  int32_t* out_n;
  int32_t* status;
  // fused path (nullable): final outputs, written by the kernel itself (pipeline.cu final_gather_kernel semantics)
  uint32_t* fin_row;
  double* fin_score;
  int32_t* fin_n;
};

__device__ __forceinline__ const float* dpp_row_ptr(const DppClArgs& a, uint32_t code, int D) {
  return code >= kDppMissBase ? a.D_sub + (size_t)(code - kDppMissBase) * D : a.D + (size_t)code * D;
}
__device__ __forceinline__ double dpp_row_inv(const DppClArgs& a, uint32_t code) {
  return code >= kDppMissBase ? a.D_sub_inv[code - kDppMissBase] : a.D_inv[code];
}

template <int D>
struct ClCfg {
  static constexpr int kBlocks = (D + 63) / 64;             // gonum 64-wide k blocks that hold embedding features
  static constexpr int kLPC = 4 * kBlocks;                  // lanes per candidate group: one per (block, DotUnitary chain)
  static constexpr int kR = kLPC / 2;                       // candidates per group (and per thread)
  static constexpr int kCL = D >= 64 ? 16 : D / 4;          // chain length: features per (candidate, lane)
  static constexpr int kTR = (D == 128) ? 8 : kCL;          // of those, kept in registers
  static constexpr int kSlots = kR * (kCL - kTR);           // shared-memory feature slots per thread ([slot][512] doubles)
  static constexpr int kFDoubles = (kSlots * 512 > 128 * D) ? kSlots * 512 : 128 * D;  // F region; first holds the f32 staging [256][D]
  static constexpr int kFS = kCL + 2;                       // lane stride (doubles) of a feature record: 16-B accesses of the
                                                            // 4 / 8 lanes of a group fall into distinct banks
  static constexpr bool kConstOwnBlock = (D % 64) == 0;     // the constant feature D opens a block of its own
};

__device__ __forceinline__ uint64_t f64_ord_c(double d) {
  uint64_t u = (uint64_t)__double_as_longlong(d);
  if ((u & 0x7FFFFFFFFFFFFFFFull) > 0x7FF0000000000000ull) return 0ull;
  return (u >> 63) ? ~u : (u | 0x8000000000000000ull);
}
// order key of a d2 value for the first-maximum search: NaN -> 0 (never wins), -0 == +0 (gonum compares with >)
__device__ __forceinline__ uint64_t d2_key(double d) {
  uint64_t u = (uint64_t)__double_as_longlong(d);
  if ((u << 1) == 0) u = 0;
  if ((u & 0x7FFFFFFFFFFFFFFFull) > 0x7FF0000000000000ull) return 0ull;
  return (u >> 63) ? ~u : (u | 0x8000000000000000ull);
}
__device__ __forceinline__ double d2_from_key(uint64_t k) {
  if (k == 0) return CUDART_NAN;
  const uint64_t u = (k >> 63) ? (k ^ 0x8000000000000000ull) : ~k;
  return __longlong_as_double((long long)u);
}
// warp-wide (max key, lowest index among equals) with three REDUX ops instead of five 64-bit shuffle rounds
__device__ __forceinline__ void warp_first_max(uint64_t key, int idx, uint64_t* kout, int* iout) {
  const unsigned hi = (unsigned)(key >> 32), lo = (unsigned)key;
  const unsigned H = __reduce_max_sync(0xffffffffu, hi);
  const unsigned L = __reduce_max_sync(0xffffffffu, hi == H ? lo : 0u);
  const unsigned I = __reduce_min_sync(0xffffffffu, (hi == H && lo == L) ? (unsigned)idx : 0x7FFFFFFFu);
  *kout = ((uint64_t)H << 32) | L;
  *iout = (int)I;
}
__device__ __forceinline__ double shfl_xor_f64(double v, int m) {
  return __shfl_xor_sync(0xffffffffu, v, m);
}

}  // namespace prg
