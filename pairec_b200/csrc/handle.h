// handle.h — the per-GPU state behind a prg_handle.
#pragma once
#include "common.cuh"
#include <nvtx3/nvToolsExt.h>   // header-only NVTX v3: ranges cost nothing unless a tool (nsys / ncu --nvtx) is attached

namespace prg {

struct DevBuf {
  void* p = nullptr;
  size_t bytes = 0;
  int ensure(size_t need);  // grows (never shrinks); contents are NOT preserved
  void release();
};

constexpr int kMaxFields = 64;
constexpr int kMaxLayers = 8;

struct Table {
  const float* factors = nullptr;
  const float* linear = nullptr;
  uint64_t rows = 0;
  bool owned = false;
};

}  // namespace prg

struct prg_handle {
  int device = 0;
  int sm_count = 148;
  cudaStream_t stream = nullptr;
  std::mutex mu;
  uint64_t launches = 0;
  bool defer_check = false;   // config "defer_check": prg_recommend(PRG_MEM_DEVICE) leaves the recall's exactness check to the next call / prg_sync
  bool pdl = true;   // config "pdl": programmatic dependent launch along the per-batch kernel chain (launch_chained)

  // optional per-stage device timing (CUDA events on this handle's stream around each stage's launches)
  int timing = 0;  // 0 off, 1 every stage, 2 only the recall scan (least perturbation of the step)
  struct Span { int stage; cudaEvent_t a, b; };
  std::vector<Span> spans;
  std::vector<cudaEvent_t> ev_pool;
  double stage_ms[8] = {0};
  uint64_t stage_n[8] = {0};

  int max_batch = 64;
  int max_k = 1000;

  // ---- recall: item matrix + scratch
  const float* E = nullptr;
  bool E_owned = false;
  uint64_t E_rows = 0;
  uint32_t E_dim = 0;
  uint64_t E_row_base = 0;
  CUtensorMap E_map;
  bool E_map_ok = false;

  // a second, fully built snapshot of the item matrix (fp32 rows, bf16 filter index, row norms, tensor maps) staged by
  // prg_stage_item_matrix while this one serves; prg_commit_item_matrix swaps the two between batches (capi.cu)
  prg_handle* staged = nullptr;

  prg::DevBuf q_dev;        // B x dim f32 (host-call staging)
  prg::DevBuf sample_keys;  // QB x sample_slots u64
  prg::DevBuf cand_keys;    // QB x cand_cap u64 (packed candidates)
  prg::DevBuf seg_keys;     // QB x n_seg x seg_cap u64 keys (FFMA2 scan) or u32 rows (tensor-core filter)
  prg::DevBuf seg_rows;     // QB x n_seg x seg_cap u32 survivor rows of the tensor-core filter
  prg::DevBuf grp_cnt;      // GROUP mode of the filter: [groups of 16 queries][n_seg] u32 list lengths
  int scan_groups = -1;     // config "scan_groups" (-1 = by dim, see recall.cu scan_groups_on): passes of more than 64 queries record survivors per (row, group of 16 queries) and the exact re-score resolves the queries; 0 = per (row, query) as for <= 64 queries
  prg::DevBuf row_norm;     // rows f32: upper bounds of the item row norms (tensor-core filter margin)
  bool scan128_nqb = true;      // config "scan128_nqb": up to 256 queries per filter pass at dim 128 (2-stage ring); 0 = 64 per pass
  bool recall_tilemax = true;   // config "recall_tilemax": threshold from per-tile maxima of the sample (0 = from sample keys)
  bool scan_ffma2 = false;  // config "scan_ffma2": use the exact FFMA2 scan for the full pass as well
  int scan_filter = 0;      // config "scan_filter": 0 = bf16 shadow index (default), 1 = tf32 on the fp32 rows
  prg::DevBuf E16;          // rows x dim bf16: round-to-nearest shadow of the item matrix (bf16 filter operand)
  CUtensorMap E16_map;
  CUtensorMap E16_map_h;    // the same index with boxes of half a tile (128 rows): stages of the 256-queries-per-pass filter
  bool E16_map_ok = false;
  bool scan_grp16 = true;   // config "scan_grp16": dim-64 GROUP-mode passes run recall_i8.cu's 16-epilogue-warp kernel over the bf16 index (0: recall_tc.cu's)
  bool scan_int8 = true;    // config "scan_int8": dim-64 passes of <= 64 queries stream an int8 index (recall_i8.cu) instead of the bf16 one
  prg::DevBuf E8;           // rows (padded to 512) x 64 int8: per-row-scaled shadow of the item matrix
  prg::DevBuf E8_prm;       // rows (padded) x {s_r, hl_r} f32
  CUtensorMap E8_map;
  bool E8_map_ok = false;
  int i8_backoff = 0;       // recalls that skip the int8 index after one whose candidate lists overflowed (its bound is wider than the
                            // bf16 one: matrices whose scores barely differ, e.g. all-positive rows, overflow it first); reset with the matrix
  prg::DevBuf cand_cnt;     // B u32
  prg::DevBuf tau;          // B u64
  prg::DevBuf dense_keys;   // fallback / small-N: nq x slots u64
  prg::DevBuf topk_keys;    // B x k u64
  prg::DevBuf out_row, out_score, out_n;  // device staging for host calls
  prg::DevBuf flags;        // B i32 per-query status from select
  prg::DevBuf topr_done;    // B i32: sample_topr_kernel served the query (else the generic select runs for it)
  int32_t last_fallback = 0;
  int32_t last_filter = 0;  // PRG_FILTER_* of the last recall's full pass
  int32_t last_max_cand = 0;
  // deferred validation of the sampled recall (fused path): the per-query status is copied to pinned host memory
  // asynchronously and checked later, so the steady state has no host round trip in the middle of a step
  struct Pending {
    bool active = false;        // recall status not checked yet
    bool fused = false;         // downstream (rank/sort/DPP) outputs depend on it: re-run them after a repair
    int B = 0, k = 0, model = 0;
    const float* q_dev = nullptr;
    uint64_t* keys_out = nullptr;
    prg_dpp_params p{};
    prg_user_features user{nullptr, nullptr};   // device pointers of the call's user features
    uint32_t* out_row = nullptr;
    double* out_score = nullptr;
    int32_t* out_n = nullptr;
    // asynchronous host call (recommend_begin): where the results were copied to — a repair copies them again and
    // re-records `done`
    uint32_t* host_row = nullptr;
    double* host_score = nullptr;
    int32_t* host_n = nullptr;
    cudaEvent_t done = nullptr;
    uint64_t seq = 0;
  } pending;
  uint64_t call_seq = 0;            // asynchronous host calls issued so far
  uint64_t repaired_seq[4] = {0};   // the most recent asynchronous calls whose results were repaired (ring)
  uint32_t repaired_pos = 0;
  cudaEvent_t flags_ev = nullptr;
  int32_t* host_flags = nullptr;  // pinned, host_flags_cap + 1 ints
  size_t host_flags_cap = 0;
  int deferred_status = PRG_OK;   // error met while resolving a deferred check (reported by the next call)
  std::string deferred_msg;

  // ---- rank: fields, tables, models
  const uint32_t* fields = nullptr;
  bool fields_owned = false;
  uint64_t fields_rows = 0;
  uint32_t n_fields = 0;
  prg::Table tables[prg::kMaxFields];
  uint32_t fdim = 0;
  float fm_w0 = 0.f;
  prg::DevBuf table_ptrs;  // device array of {factors*, linear*, rows}
  bool table_ptrs_dirty = true;

  int mlp_layers = 0;
  uint32_t mlp_dims[prg::kMaxLayers + 1] = {0};
  prg::DevBuf mlp_W[prg::kMaxLayers];  // bf16, layout chosen by mlp.cu
  prg::DevBuf mlp_b[prg::kMaxLayers];
  CUtensorMap mlp_Wmap[prg::kMaxLayers];
  CUtensorMap mlp_Wmap_half[prg::kMaxLayers];  // boxes of half a W block (one CTA's share of a multicast pair load)
  float mlp_b_last[4] = {0.f, 0.f, 0.f, 0.f};   // biases of the output heads
  double rank_coef[4] = {1.0, 0.0, 0.0, 0.0};   // Item.Score = sum_o coef[o] * score_o (prg_set_rank_score)
  uint32_t mlp_k_item = 0, mlp_k_user = 0;      // input columns from the item fields (tensor cores) / user + context (per request)
  prg::DevBuf mlp_Wu;                           // [k_user][dims[1]] f32: first-layer weights of the user / context columns
  bool mlp_no_pair = false;           // config "mlp_no_pair": single-CTA persistent kernel without W multicast (A/B measurements)

  // ---- user / context features of a request (service/rank/algo_data.go:104-118)
  uint32_t n_user_fields = 0;   // categorical user fields: feature tables n_fields .. n_fields + n_user_fields - 1
  uint32_t n_user_dense = 0;    // numeric context values appended to the tower input
  prg::DevBuf user_ids_dev, user_dense_dev;   // host-call staging: B x U u32, B x n_dense f32
  // the fused path evaluates a batch's user prefix on a side stream while the recall runs (it depends on the request only)
  cudaStream_t side_stream = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  bool prefix_ahead = false;    // user_prefix_kernel of the current batch is in flight on side_stream; ev_join follows it
  prg::DevBuf fm_state;         // B x 36 f32: lin, s[16], ss[16] after the user fields (the gather continues from it)
  prg::DevBuf ubias;            // B x dims[1] f32: b1 + W1[:, user columns] * x_user
  prg::DevBuf act[2];   // activations ping-pong (bf16)
  prg::DevBuf fm_logit; // B*n f32
  prg::DevBuf rank_rows, rank_out, rank_map;   // rank_map: host-call staging of the per-head scores

  // ---- DPP
  const void* D = nullptr;
  bool D_owned = false;
  uint64_t D_rows = 0;
  prg::DevBuf D_inv;        // D_rows f64: 1 / ||row|| (gonum floats.Norm order) of an f32 diversity table, built once at set time
  prg::DevBuf D_sub, D_sub_inv;   // 1024 substitute directions (f32 [1024][D_dim]) + inverse norms for candidates without a table row
  uint32_t D_dim = 0;
  int D_dtype = PRG_F32;
  bool dpp_lazy = false;     // config "dpp_lazy": the lazy-evaluation kernel (dpp_lazy.cu) instead of the cluster kernel
  bool dpp_pair = true;      // config "dpp_pair" (dpp_pair.cu, default for dim-128 tables): 2-CTA clusters, features partly in tensor memory; 0 = 4-CTA cluster kernel
  bool dpp_generic = false;  // config "dpp_generic": force the one-CTA-per-request kernel (A/B measurements)
  prg::DevBuf dpp_scratch, dpp_rows, dpp_score, dpp_idx, dpp_n, dpp_status;
  prg::DevBuf dpp_hook_E, dpp_hook_rows, dpp_hook_in;   // hook path: per-call fp64 vectors [B*n][hook_dim + D_dim], identity rows, staged hooks

  prg::DevBuf ssd_E, ssd_P;  // SSD: mutable fp64 embeddings [B][D][1024], projection history [B][w][1024]

  // ---- sort
  prg::DevBuf sort_in, sort_perm;

  // ---- fused path
  int prerank_model = 0, prerank_keep = 0;   // prg_set_prerank: general (pre-)rank stage of the fused path (0 = none)
  prg::DevBuf pre_rows;                      // [B][keep] rows that survive the pre-rank Action
  prg::DevBuf rec_rows, rec_scores, rec_perm, rec_sorted_rows, rec_sorted_scores;
};

namespace prg {
// launch bookkeeping
inline void count_launch(prg_handle* h, int n = 1) { h->launches += (uint64_t)n; }

#ifdef __CUDACC__
// Launch of a kernel of the per-batch chain.  With h->pdl the launch carries the programmatic-stream-serialization
// attribute: the grid may start while its predecessor in the stream drains and orders itself behind it with
// pdl_wait() (common.cuh).  Only kernels that call pdl_wait() before touching anything an earlier kernel wrote — and
// before they exit — may be launched through here.  cluster_x > 1 adds the cluster dimension.
template <typename... KArgs, typename... Args>
inline cudaError_t launch_chained(prg_handle* h, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem,
                                  unsigned cluster_x, Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = h->stream;
  cudaLaunchAttribute attr[2];
  unsigned n = 0;
  if (cluster_x > 1) {
    attr[n].id = cudaLaunchAttributeClusterDimension;
    attr[n].val.clusterDim.x = cluster_x; attr[n].val.clusterDim.y = 1; attr[n].val.clusterDim.z = 1;
    ++n;
  }
  if (h->pdl) {
    attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  cfg.attrs = attr; cfg.numAttrs = n;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
#endif

enum { SCAN_FILTER_BF16 = 0, SCAN_FILTER_TF32 = 1 };
enum Stage { ST_SCAN = 0, ST_SCAN_DENSE = 1, ST_SELECT = 2, ST_GATHER_FM = 3, ST_MLP = 4, ST_SORT = 5, ST_DPP = 6, ST_OTHER = 7 };
// RAII span: an NVTX range named after the stage around the enclosed launches (host side: where the launches are
// issued), plus an event before and after them when timing is on
inline const char* stage_name(int st) {
  static const char* const names[8] = {"prg:recall_scan", "prg:recall_sample", "prg:recall_select", "prg:gather_fm",
                                       "prg:mlp", "prg:sort", "prg:dpp", "prg:other"};
  return names[st & 7];
}
struct StageScope {
  prg_handle* h;
  int stage;
  cudaEvent_t a = nullptr;
  static cudaEvent_t get(prg_handle* h) {
    cudaEvent_t e = nullptr;
    if (!h->ev_pool.empty()) { e = h->ev_pool.back(); h->ev_pool.pop_back(); }
    else cudaEventCreate(&e);
    return e;
  }
  StageScope(prg_handle* hh, int st) : h(hh), stage(st) {
    nvtxRangePushA(stage_name(st));
    if (h->timing == 1 || (h->timing == 2 && st == ST_SCAN)) { a = get(h); cudaEventRecord(a, h->stream); }
  }
  ~StageScope() {
    if (a) { cudaEvent_t b = get(h); cudaEventRecord(b, h->stream); h->spans.push_back({stage, a, b}); }
    nvtxRangePop();
  }
};

// The fused request path's final outputs (the picks' rows and rank scores, [B][top_n], and the per-request counts):
// a DPP kernel that can write them itself does so and sets `done`; otherwise pipeline.cu's final_gather_kernel runs.
struct DppFinal {
  uint32_t* row = nullptr;
  double* score = nullptr;
  int32_t* n = nullptr;
  bool done = false;
};

// recall.cu
int recall_build_map(prg_handle* h);
int recall_topk_device(prg_handle* h, const float* q_dev, int B, int k, uint64_t* keys_out /*B x k*/, bool defer = false);
int recall_resolve(prg_handle* h, bool* repaired);   // waits for the deferred status, redoes failed queries densely
int resolve_pending(prg_handle* h);                  // pipeline.cu: recall_resolve + re-run of the fused downstream
// pipeline.cu — the fused request path as an asynchronous pair (batcher.cu): `begin` enqueues H2D of the queries, all
// stages and the D2H of the results into pinned buffers, records `done` and returns; `end` waits for `done`, settles
// the recall's deferred check if nobody has yet and returns the call's status.  Several calls may be in flight on a
// handle (stream order); the handle's lock is held only while a call is being enqueued.
int recommend_begin(prg_handle* h, const float* q_pinned, int B, int recall_k, int model, const prg_dpp_params& p,
                    const prg_user_features* user_pinned, uint32_t* out_row, double* out_score, int32_t* out_n,
                    cudaEvent_t done, uint64_t* seq);
int recommend_end(prg_handle* h, uint64_t seq, cudaEvent_t done);
// gather_fm.cu: a batch's user prefix (FM running sums after the user fields + the user share of the tower's first layer);
// ahead = on the handle's side stream, joined by the next rank of the batch (pipeline.cu rank_device)
int user_prefix_device(prg_handle* h, const uint32_t* user_ids_dev, const float* user_dense_dev, int B, bool need_mlp, bool ahead);
int keys_to_outputs(prg_handle* h, const uint64_t* keys_dev, int B, int k, uint32_t* out_row, float* out_score,
                    int32_t* out_n);
int shard_sample_len(int k);
int recall_shard_sample_device(prg_handle* h, const float* q_dev, int Bg, int k, int G, uint64_t* out);
int recall_shard_candidates_device(prg_handle* h, const float* q_dev, int Bg, int k, int G, const uint64_t* all_samples,
                                   uint64_t* out);
// tau (nullable): thresholds of the Bg queries checked (default: the handle's, from recall_shard_candidates_device)
int shard_check_device(prg_handle* h, const uint64_t* gathered, int G, int Bg, int k, int32_t* retry_dev,
                       const uint64_t* tau = nullptr);
int shard_pack_owner_device(prg_handle* h, const uint64_t* in, int Bg, int B, int k, uint64_t* out);
int shard_check_owner_device(prg_handle* h, const uint64_t* received, int G, int B, int k, int q0, int32_t* retry_dev);
int merge_keys_device(prg_handle* h, const uint64_t* keys_dev, int G, uint64_t g_stride, int B, int k,
                      uint64_t* keys_out);
}  // namespace prg
