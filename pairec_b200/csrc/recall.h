// recall.h — constants and parameter blocks shared by recall.cu (FFMA2 exact scan, select) and recall_tc.cu
// (tensor-core candidate filter).
#pragma once
#include "handle.h"

namespace prg {

constexpr int TM = 4, TQ = 16, WR = 2, WQ = 4;
constexpr int kTileRows = WR * 32 * TM;              // 256
constexpr int kQB = WQ * TQ;                         // 64 queries per pass
constexpr int kConsumerWarps = WR * WQ;              // 8
constexpr int kScanThreads = (kConsumerWarps + 1) * 32;
constexpr int kStageFloats = kTileRows * 64;         // one stage = 256 rows x 64 dims
constexpr int kStageBytes = kStageFloats * 4;        // 64 KiB
constexpr int kStages = 3;
constexpr int kSubTileFloats = kTileRows * 32;       // one TMA box: 256 rows x 32 floats (128 B)

enum { SCAN_THRESH = 0, SCAN_DENSE = 1, SCAN_TILEMAX = 2 };   // TILEMAX: recall_tc.cu only (experimental)
#ifndef SCAN_UNROLL
#define SCAN_UNROLL 2
#endif
constexpr int kScanUnroll = SCAN_UNROLL;

struct ScanParams {
  const float* Q;          // [nq][dim]
  int nq;
  uint64_t n_rows;         // local rows in the matrix
  uint64_t row_base;       // global id of local row 0
  uint32_t n_tiles;        // tiles covered by this launch
  uint32_t tile_stride;    // launch tile t reads matrix tile t*tile_stride
  const uint64_t* tau;     // THRESH: [nq]
  uint64_t* cand;          // THRESH: [nq][gridDim.x][seg_cap] — one private segment per CTA and query
  uint32_t seg_cap;
  uint32_t* seg_cnt;       // THRESH: [nq][gridDim.x], written once per CTA at kernel end
  const float* row_norm;   // TC filter: [n_rows] upper bounds of the row norms
  const float2* row_q8;    // int8 filter (recall_i8.cu): [rows padded to 512] {a_r, hl_r}; set by launch_scan_i8
  uint32_t* cand_rows;     // TC filter: [nq][gridDim.x][seg_cap] surviving global rows (re-scored exactly by select)
  uint64_t* dense;         // DENSE: [nq][dense_stride], slot = t*256 + r
  uint64_t dense_stride;
  int q_blocks;            // DENSE: > 1 = nq spans that many blocks of 64 queries, one grid row (blockIdx.y) each
  // TC filter, GROUP mode (grp_rows != nullptr; passes of more than 64 queries): a surviving row is recorded once per
  // GROUP of 16 queries in which it may reach a threshold, not once per query — [group][gridDim.x][grp_cap] global rows,
  // lengths in grp_cnt[group][gridDim.x] (group = query / 16 within the pass).  rescore_group_kernel then scores the row
  // exactly against the group's 16 queries and keeps the keys that reach tau.
  uint32_t* grp_rows;
  uint32_t* grp_cnt;
  uint32_t grp_cap;
};
constexpr int kGrpQ = 16;   // queries per group

struct SelectParams {
  const uint64_t* keys;    // query q reads keys + q*stride
  uint64_t stride;
  const uint32_t* counts;  // per-query list length (nullable -> fixed_m); clamped to cap, overflow flagged
  uint32_t fixed_m;
  uint32_t cap;
  // segmented input (scan<THRESH> output): query q owns n_seg segments of seg_cap keys at keys + (q*n_seg+s)*seg_cap,
  // lengths seg_counts[q*n_seg+s]; they are first packed into compact + q*stride (cap = stride).
  const uint32_t* seg_counts;
  uint32_t n_seg, seg_cap;
  uint64_t* compact;
  // exact re-score of tensor-core survivors while packing: segments hold u32 global rows instead of keys
  const uint32_t* seg_rows;
  const float* E;          // item matrix (local rows)
  const float* Q;          // [nq][dim]
  uint32_t dim;
  uint64_t row_base;
  const uint64_t* tau_check;  // per-query sampled threshold: the k-th exact key must reach it (else flag 2)
  int k;                   // rank wanted
  int k_out;               // row stride of the outputs (== caller's k)
  uint32_t expect;         // MODE_TOPK: number of results that must exist (min(k, total rows)), else flag 2
  uint64_t* out_keys;      // MODE_TOPK: [q][k_out] sorted descending, 0-padded (nullable)
  uint32_t* out_row;       // nullable
  float* out_score;        // nullable
  int32_t* out_n;          // nullable
  uint64_t* tau;           // MODE_KTH: [q] the k-th largest key (0 if fewer than k valid keys)
  int32_t* flags;          // nullable; 0 ok, 1 overflow, 2 underflow
  uint32_t* max_count;     // nullable: atomicMax of the list lengths seen
  const int32_t* skip;     // nullable: query q is skipped when skip[q] != 0 (already served by sample_topr_kernel)
};
enum { SEL_TOPK = 0, SEL_KTH = 1 };


// recall_tc.cu
int launch_scan_tc(prg_handle* h, const ScanParams& p);
int scan_tc_max_queries(const prg_handle* h);
int build_row_norms(prg_handle* h);
bool scan_tc_dense_available(const prg_handle* h);
int launch_scan_tc_dense(prg_handle* h, const ScanParams& p);   // sample scoring (approximate keys) on the tensor cores
int launch_scan_tc_tilemax(prg_handle* h, const ScanParams& p); // sample scoring reduced to one maximum per (tile, query)

// recall_i8.cu — int8 filter index (dim 64, <= 64 queries per pass)
int build_i8_index(prg_handle* h);          // called by build_row_norms; a no-op unless scan_i8_wanted
bool scan_i8_available(const prg_handle* h);
int launch_scan_i8(prg_handle* h, const ScanParams& p, uint32_t n_seg);
bool scan_i8g_available(const prg_handle* h);   // dim 128: GROUP-mode passes of up to 256 queries
int launch_scan_i8g(prg_handle* h, const ScanParams& p, uint32_t n_seg);
int launch_scan_g16(prg_handle* h, const ScanParams& p, uint32_t n_seg);   // dim 64, bf16 index: GROUP-mode passes, 16 epilogue warps

}  // namespace prg
