/*
 * pairec_gpu.h — C ABI of libpairec_gpu.so: the B200 (sm_100a) recall → feature → rank → sort/DPP hot path that
 * drops in behind alibaba/pairec's Recall / IAlgorithm / ISort plugin contracts.
 *
 * Every entry point is `extern "C"`, takes plain pointers and sizes, returns an int status (PRG_OK == 0) and
 * records a thread-local message readable through prg_last_error().  Nothing here aborts or throws across the
 * boundary: the reference's callers treat an error as "log and keep the items unchanged"
 * (sort/dpp_sort.go:304-307, service/rank/rank_service.go:274-277, service/recall/vector_recall.go:89-92).
 *
 * Reference interface each group replaces (paths relative to alibaba/pairec):
 *   prg_recall_*    algorithm/faiss/vector_client.go:32-42 (VectorRetrieval.Search, vectorretrieval.proto:11-20),
 *                   reached from service/recall/vector_recall.go:88 through algorithm.Run (algorithm/algorithm.go:126)
 *   prg_rank        service/rank/rank_service.go:264-289 → algorithm.Run → eas/tfserving remote models
 *                   (algorithm/eas/fm_response.go:28-34, algorithm/tfserving/response.go:51-63) and the item-side
 *                   module.FeatureDao.FeatureFetch (module/feature_dao.go:24-26)
 *   prg_sort_*      sort/item_rank_score.go:26-32, sort/item_score.go:15-18, sort/algo_score_sort.go:38-66
 *   prg_dpp         sort/dpp_sort.go:271-351 (doSort), :372-475 (KernelMatrix), :477-551 (DPPWithWindow, DPP)
 *   prg_lookup      algorithm/lookup.go:37-51
 *   prg_set_*       the table loaders behind module/vector_*_dao.go, module/feature_*_dao.go and
 *                   sort/dpp_sort.go:169-269 (loadEmbeddingCache): tables become HBM resident.
 *
 * Memory kinds: every data pointer is accompanied (per call) by `mem`: PRG_MEM_HOST = caller host memory, the
 * call copies host<->device itself and returns when the outputs are valid; PRG_MEM_DEVICE = device pointers on
 * this handle's device, the work is enqueued on the handle's stream and the caller orders against it with
 * prg_sync() / prg_stream().  No pointer is retained after return, except tables adopted with PRG_MEM_DEVICE,
 * which the caller must keep alive until prg_destroy or the next prg_set_* of the same table.
 *
 * Threading: any entry point may be called from any OS thread (cgo); calls on one handle are serialised by an
 * internal mutex, different handles (one per GPU) run concurrently.
 */
#ifndef PAIREC_GPU_H
#define PAIREC_GPU_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct prg_handle prg_handle;

enum {
  PRG_OK = 0,
  PRG_EINVAL = 1,       /* bad argument */
  PRG_ECUDA = 2,        /* CUDA runtime / driver failure (message carries the CUDA error string) */
  PRG_ENOMEM = 3,
  PRG_ESTATE = 4,       /* a table or model the call needs has not been set */
  PRG_EUNSUPPORTED = 5, /* shape outside what the kernels are built for */
  PRG_ENODEVICE = 6     /* no CUDA device: there is NO CPU fallback in this library */
};

enum { PRG_MEM_HOST = 0, PRG_MEM_DEVICE = 1 };
enum { PRG_F32 = 0, PRG_F64 = 1 };

/* rank model selector for prg_rank (what the replaced remote processor computed) */
enum {
  PRG_MODEL_FM = 0,     /* ALINK_FM-shaped: sigmoid(w0 + sum w + 1/2 sum_k((sum v)^2 - sum v^2)) */
  PRG_MODEL_MLP = 1,    /* EasyRec/TF-Serving-shaped DNN: concat(field factors) -> dense layers (bf16) -> sigmoid */
  PRG_MODEL_FM_MLP = 2  /* DeepFM-shaped: sigmoid(fm_logit + mlp_logit) */
};

/* ---------------------------------------------------------------- lifecycle */

/* json_cfg: {"device":0,"max_batch":64,"max_k":1000,"max_candidates":1024} — all keys optional.
 * Kernel-variant switches (integers; results never depend on them, only speed): "scan_int8" (1: int8 filter index for
 * dim-64 passes of at most 64 queries, default), "scan_tf32" (1: filter over the fp32 rows, no shadow index), "scan_ffma2"
 * (1: exact fp32 scan, no filter), "scan_groups", "scan128_nqb", "recall_tilemax", "dpp_pair", "dpp_generic", "dpp_lazy",
 * "mlp_no_pair", "defer_check", "pdl", "sm_limit".  DESIGN.md names what each selects. */
int prg_init(const char* json_cfg, prg_handle** out);
void prg_destroy(prg_handle* h);
const char* prg_last_error(void);
const char* prg_version(void);
/* Blocks until everything enqueued on the handle's stream has finished. */
int prg_sync(prg_handle* h);
/* The handle's cudaStream_t (as void*), for callers that order their own device work against ours. */
void* prg_stream(prg_handle* h);

/* ---------------------------------------------------------------- HBM-resident tables */

/* Item embedding matrix scanned by recall: rows x dim f32 row-major, dim in {64,128}.  row_base is the global row
 * id of local row 0 (non-zero on a row shard).  Replaces the index held by the remote faiss server. */
int prg_set_item_matrix(prg_handle* h, const float* data, uint64_t rows, uint32_t dim, uint64_t row_base, int mem);

/* Snapshot swap of the item matrix without stopping the service — the reference's vector DAO switches to the newest
 * partition table in the background, once a minute (module/vector_hologres_dao.go:40-61).
 * prg_stage_item_matrix uploads the new matrix and builds everything recall needs for it (bf16 filter index, row
 * norms, tensor maps) on a side stream WITHOUT taking the handle's lock: calls on the handle keep being served from the
 * live snapshot (both are resident meanwhile).  prg_commit_item_matrix swaps the snapshots between two batches and
 * frees the old one; PRG_ESTATE if nothing is staged; PRG_EINVAL if the staged dim differs from the live one (callers
 * and batchers hold query buffers of the live dim: use prg_set_item_matrix for that).  Staging again before a commit
 * drops the earlier staged snapshot.  A PRG_MEM_DEVICE matrix is adopted, not copied, and must stay valid while it is staged or live. */
/* dim of the live item matrix (0 before prg_set_item_matrix).  The recall entry points take no vector length: a host
 * glue layer checks len(vector) == prg_item_dim() before passing &vector[0] (the reference parses user vectors from
 * 'i:v i:v' text and silently skips malformed pairs, service/recall/vector_recall.go:72-82). */
uint32_t prg_item_dim(prg_handle* h);
int prg_stage_item_matrix(prg_handle* h, const float* data, uint64_t rows, uint32_t dim, uint64_t row_base, int mem);
int prg_commit_item_matrix(prg_handle* h);

/* Per-item categorical field ids, rows x n_fields u32 row-major (what FeatureDao.FeatureFetch would have written
 * into item.Properties, module/feature_hologres_dao.go:644-675, already id-encoded). */
int prg_set_item_fields(prg_handle* h, const uint32_t* ids, uint64_t rows, uint32_t n_fields, int mem);

/* Feature table `table` (0 <= table < n_fields): factors rows x fdim f32 (fdim == 16) and linear weights rows f32
 * (may be NULL = zeros).  These are the embedding variables of the replaced remote model. */
int prg_set_feature_table(prg_handle* h, int table, const float* factors, const float* linear, uint64_t rows,
                          uint32_t fdim, int mem);
int prg_set_fm_bias(prg_handle* h, float w0);

/* User / context side of the rank features.  The reference merges, per candidate, features = userFeatures with the
 * item's features on top (service/rank/algo_data.go:104-118; user map from user.MakeUserFeatures(),
 * service/rank/rank_service.go:175-183) and sends the merged map to the remote model.  Id-encoded here like the item
 * side: n_user_fields categorical user fields, whose embedding tables are feature tables n_fields ..
 * n_fields + n_user_fields - 1 (prg_set_feature_table), plus n_user_dense numeric context values that only the dense
 * tower sees.  FM sums run over the user fields first, then the item fields; the tower input row is
 * [item factors | user factors | dense values], so dims[0] of prg_set_mlp == (n_fields + n_user_fields) * 16 +
 * n_user_dense.  Call before prg_set_mlp (changing the counts drops the tower). */
int prg_set_user_fields(prg_handle* h, uint32_t n_user_fields, uint32_t n_user_dense);
typedef struct prg_user_features {
  const uint32_t* ids;   /* B x n_user_fields u32; 0xFFFFFFFF (or >= table rows) = the user has no such feature; NULL = none */
  const float* dense;    /* B x n_user_dense f32; NULL = zeros */
} prg_user_features;

/* Dense tower: n_layers weight matrices, dims[n_layers+1]; dims[0] == (n_fields + n_user_fields) * fdim + n_user_dense,
 * dims[n_layers] == number of output heads, 1..4 (a multi-target EasyRec model returns one score per head,
 * algorithm/eas/easyrec_response.go:35-70; TF-Serving Outputs rows, algorithm/tfserving/response.go:51-63).
 * W[l] is bf16 (raw uint16 bit patterns) [dims[l+1]][dims[l]] row-major (out-major), bias[l] f32 [dims[l+1]].
 * Arithmetic: the input is rounded to bf16 once; computed activations are carried as bf16 pairs (hi + lo); fp32
 * accumulation; see oracle/oracle.c orc_mlp_forward.  Host pointers only. */
int prg_set_mlp(prg_handle* h, int n_layers, const uint32_t* dims, const uint16_t* const* W, const float* const* bias);
/* Item.Score = sum_o coef[o] * score_o over the tower's heads, evaluated left to right in fp64 — the RankScore
 * expression (service/rank/rank_service.go:339-363, utils/ast) for the sums of products the device path folds into the
 * last layer; any other expression is evaluated by the host from the score map of prg_rank_ex.  Default {1}. */
int prg_set_rank_score(prg_handle* h, const double* coef, int n);

/* Diversity embeddings read by DPP: rows x dim, f32 or f64 row-major (the table behind
 * sort/dpp_sort.go:169-269; rows are L2-normalised in fp64 at use when normalize != 0, :234-237). */
int prg_set_diversity_matrix(prg_handle* h, const void* data, uint64_t rows, uint32_t dim, int dtype, int mem);

/* ---------------------------------------------------------------- recall (faiss VectorRetrieval.Search replacement) */

/* Exact inner-product top-k.  q: B x dim f32.  out_row: B x k u32 global row ids, out_score: B x k f32,
 * out_n: B i32 (= min(k, rows)); rows beyond out_n are filled with 0xFFFFFFFF / -inf.
 * Score(row,q) = fmaf chain over dims 0..dim-1 starting from +0 (one accumulator, IEEE RN).
 * Order: score descending, ties by ascending row; NaN scores rank below -inf. */
int prg_recall_topk(prg_handle* h, const float* q, int B, int k, uint32_t* out_row, float* out_score,
                    int32_t* out_n, int mem);

/* Row-sharded recall: local top-k as 64-bit order keys, device pointers only.
 * key = (ordered_bits(score) << 32) | (0xFFFFFFFF - global_row); 0 = empty slot.  out_keys: B x k u64. */
int prg_recall_local_keys(prg_handle* h, const float* q_dev, int B, int k, uint64_t* out_keys_dev);
/* Row-sharded recall with ONE global threshold per query: the refine work of a rank no longer grows with the shard
 * count G (with prg_recall_local_keys every shard refines ~4k survivors per query to a local top-k for all B*G queries).
 * All pointers are device pointers; Bg = B*G queries, identical on every rank; the two exchanges are the host's
 * (ncclAllGather / torch.distributed.all_gather_into_tensor).
 *   1. prg_shard_sample:     out_sample[Bg][r] = the r = prg_shard_sample_len(k) best keys of this shard's 1/128 row
 *                            sample per query (sorted, 0-padded)                              -> all-gather #1 (G x Bg x r)
 *   2. prg_shard_candidates: tau[q] = r-th largest of the G*r gathered sample keys; out[Bg*k + Bg] = per query the
 *                            exact keys of this shard's rows that reach tau (sorted, at most k, 0-padded), then Bg
 *                            status words (non-zero: the shard's candidate list overflowed)    -> all-gather #2
 *   3. prg_shard_check:      retry[0] |= 1 and retry[1] += 1 for every query that needs the exact protocol (a shard
 *                            overflowed, or fewer than k gathered keys reach tau: both < 1e-10 per query unless the
 *                            row order defeats the strided sample).  Every rank computes the same answer from the
 *                            same gathered buffer, so all ranks redo the batch together with prg_recall_local_keys.
 *                            Then prg_recommend_from_keys with g_stride = Bg*k + Bg merges as before.
 * The union of the gathered lists holds every row whose exact key reaches tau; with >= k of them the merged top-k is the
 * global top-k, bit-identical to prg_recall_topk over the unsharded matrix. */
int prg_shard_sample_len(int k);
int prg_shard_sample(prg_handle* h, const float* q_dev, int Bg, int k, int G, uint64_t* out_sample_dev);
int prg_shard_candidates(prg_handle* h, const float* q_dev, int Bg, int k, int G, const uint64_t* all_samples_dev,
                         uint64_t* out_keys_dev);
int prg_shard_check(prg_handle* h, const uint64_t* gathered_dev, int G, int Bg, int k, int32_t* retry_dev);
/* All-to-all form of exchange #2 (what bench.py uses at N > 1): a rank only ranks its OWN B = Bg / G queries
 * [rank*B, rank*B + B), so it only needs their lists.  prg_shard_pack_owner regroups prg_shard_candidates' output into G
 * chunks of [B*k keys | B status words] (chunk o = the part rank o needs); ONE ncclAllToAll (chunk size B*k + B) replaces
 * the all-gather (G x less data received); prg_shard_check_owner runs the check on the received G chunks for the owned
 * queries (q0 = rank*B) — the retry flag is then per rank, so the ranks agree on a redo through one integer all-reduce
 * (or per batch of batches); prg_recommend_from_keys(_ex) merges the received chunks with g_stride = B*k + B. */
int prg_shard_pack_owner(prg_handle* h, const uint64_t* cand_dev, int Bg, int B, int k, uint64_t* out_dev);
int prg_shard_check_owner(prg_handle* h, const uint64_t* received_dev, int G, int B, int k, int q0, int32_t* retry_dev);
/* Merge G gathered key lists (keys_dev: G x B x k, the all-gather output) into the global top-k. */
int prg_merge_keys(prg_handle* h, const uint64_t* keys_dev, int G, int B, int k, uint32_t* out_row,
                   float* out_score, int32_t* out_n, int mem);

/* ---------------------------------------------------------------- rank (feature gather + FM / MLP forward) */

/* rows: B x n u32 item rows (0xFFFFFFFF = padding, scored 0).  out_score: B x n f64 (the AlgoResponse.GetScore()
 * values, algorithm/response/resonse.go:3-7), f32 arithmetic widened exactly. */
int prg_rank(prg_handle* h, int model, const uint32_t* rows, int B, int n, double* out_score, int mem);
/* prg_rank with the requests' user / context features (request b's candidates are rows[b*n .. b*n+n-1]) and,
 * optionally, every head's score: out_score_map B x n x heads f64 (AlgoResponse.GetScoreMap(); head 0 carries the FM
 * part of PRG_MODEL_FM_MLP).  out_score (nullable when out_score_map is given) = sum_o coef_o * score_o.
 * user == NULL: no user features (their tables contribute nothing). */
int prg_rank_ex(prg_handle* h, int model, const uint32_t* rows, int B, int n, const prg_user_features* user,
                double* out_score, double* out_score_map, int mem);

/* ---------------------------------------------------------------- sort */

/* Order of sort.Sort(sort.Reverse(ItemScoreSlice)) (sort/item_rank_score.go:29): descending score.  Host-side
 * emulation of Go's pdqsort so that tie order follows the reference for n <= a few thousand.  out_perm[i] = input
 * index of the item at output position i.  Pure host code (no GPU involved). */
int prg_sort_desc_host(const double* score, int n, int32_t* out_perm);
/* Device batched variant: B independent lists of n scores; stable total order (score desc, input index asc). */
int prg_sort_desc(prg_handle* h, const double* score, int B, int n, int32_t* out_perm, int mem);

/* ---------------------------------------------------------------- DPP re-rank (sort/dpp_sort.go) */

typedef struct prg_dpp_params {
  double alpha;             /* DPPConf.Alpha / abtest dpp_alpha (dpp_sort.go:374) */
  int32_t top_n;            /* ctx.Size */
  int32_t window_size;      /* DPPConf.WindowSize, <=0 -> 10 (dpp_sort.go:89-91) */
  int32_t norm_mode;        /* dpp_norm_relevance_score 0/1/2 (dpp_sort.go:382-405) */
  int32_t normalize_emb;    /* NormalizeEmb (dpp_sort.go:234-237) */
  int32_t candidate_count;  /* DPPConf.CandidateCount (dpp_sort.go:280-287) */
  double min_score_percent; /* DPPConf.MinScorePercent (dpp_sort.go:288-298) */
  int32_t no_positive_sim;  /* 1 = DPPConf.EnsurePositiveSim == "false" (hook-only embeddings, dpp_sort.go:440-445): the
                               feature row is [v ; 0] without the 1/sqrt2 scaling.  0 (default) = [v ; 1] / sqrt2 */
  int32_t reserved;
} prg_dpp_params;

/* rows: B x n item rows into the diversity matrix; score: B x n f64 relevance (Item.Score) in the order the items
 * reach doSort.  out_idx: B x top_n i32 indices into the request's input list, out_n: B i32 number written
 * (min(top_n, n_after_truncation)), status: B i32 per-request (0 ok; 1 = reference would have returned the input
 * unchanged, e.g. "all item score is zero", dpp_sort.go:385-388,397-400).  0xFFFFFFFF rows are padding. */
int prg_dpp(prg_handle* h, const uint32_t* rows, const double* score, int B, int n, const prg_dpp_params* p,
            int32_t* out_idx, int32_t* out_n, int32_t* status, int mem);
/* A candidate whose row is outside the diversity table (>= rows, e.g. 0xFFFFFFFE for an id the host could not map)
 * has no embedding: the reference gives it a random unit vector (dpp_sort.go:250-262, unseeded); here it takes row
 * (position in the request's list & 1023) of a fixed table of pseudo-random directions, so it competes like any other
 * item and results are reproducible.
 *
 * prg_dpp_ex: embeddings from registered hooks (sort/dpp_sort.go:56-58, :362-370).  hook: B x n x hook_dim f64, the
 * concatenated GenerateEmbedding output per candidate.  use_table != 0: concat(hook, table row) re-normalised as a whole
 * (:416-421; rows indexes the f32 diversity table); use_table == 0: hook only (:432-447; rows may be NULL; NormalizeEmb
 * and EnsurePositiveSim as in p).  hook == NULL is prg_dpp. */
int prg_dpp_ex(prg_handle* h, const uint32_t* rows, const double* score, const double* hook, int hook_dim, int use_table,
               int B, int n, const prg_dpp_params* p, int32_t* out_idx, int32_t* out_n, int32_t* status, int mem);

/* ---------------------------------------------------------------- SSD re-rank (sort/ssd_sort.go) */

typedef struct prg_ssd_params {
  double gamma;             /* SSDConf.Gamma (<=0 in the config -> 0.25, ssd_sort.go:62,81-83); abtest ssd_gamma; 0 = skip */
  int32_t top_n;            /* ctx.Size */
  int32_t window_size;      /* SSDConf.WindowSize, <=1 -> 5 (ssd_sort.go:358-361) */
  int32_t norm_mode;        /* ssd_norm_quality_score 0/1/2 (ssd_sort.go:368-391) */
  int32_t normalize_emb;    /* NormalizeEmb */
  int32_t use_ssd_star;     /* SSDConf.UseSSDStar */
  int32_t candidate_count;  /* SSDConf.CandidateCount */
  double min_score_percent; /* SSDConf.MinScorePercent */
} prg_ssd_params;

/* Same calling convention as prg_dpp.  doSort sorts the candidates by score first (ssd_sort.go:301; stable order on
 * the device), then SSDWithSlidingWindow picks min(n', top_n) items.  status 1 = upstream returned the sorted
 * (truncated) list without re-ranking (gamma == 0, "all item score are zeros"): out_idx holds its first top_n. */
int prg_ssd(prg_handle* h, const uint32_t* rows, const double* score, int B, int n, const prg_ssd_params* p,
            int32_t* out_idx, int32_t* out_n, int32_t* status, int mem);

/* ---------------------------------------------------------------- fused request path */

/* recall -> gather+rank -> score sort -> DPP for B requests, everything device resident in between.
 * q: B x dim f32.  out_row: B x top_n u32 item rows in final order, out_score: B x top_n f64 rank scores,
 * out_n: B i32.  This is what one /api/recommend costs below the four plugin call sites. */
int prg_recommend(prg_handle* h, const float* q, int B, int recall_k, int model, const prg_dpp_params* p,
                  uint32_t* out_row, double* out_score, int32_t* out_n, int mem);
/* The same with each request's user / context features (prg_set_user_fields).
 * PRG_MEM_DEVICE: the outputs are final when the work enqueued by the call has completed (prg_sync, or an event on
 * prg_stream()); the call itself waits for the recall stage's exactness check and repairs a failed query before it
 * returns.  A handle created with "defer_check":1 skips that wait: then a (rare) repair is enqueued by the NEXT call
 * on the handle or by prg_sync, and the output / user buffers must stay valid until one of those has returned. */
int prg_recommend_ex(prg_handle* h, const float* q, int B, int recall_k, int model, const prg_dpp_params* p,
                     const prg_user_features* user, uint32_t* out_row, double* out_score, int32_t* out_n, int mem);

/* General (pre-)rank inside the fused path (service/general_rank/base_general_rank.go:66-109: GeneralRankConfs[scene] =
 * {RankConf, ActionConfs}): the whole recall set is scored by `model` (normally the cheap PRG_MODEL_FM), sorted, and the
 * Action keeps the best `keep` candidates per request (:183); only those reach the rank model of the call, the score
 * sort and DPP.  Everything stays on the device.  keep == 0 (default) or keep >= recall_k: no pre-rank stage.  Applies to
 * prg_recommend*, prg_recommend_from_keys*, the batcher and prg_group_recommend on this handle. */
int prg_set_prerank(prg_handle* h, int model, int keep);

/* Row-sharded variant: keys_dev points at this rank's slice of the all-gathered per-shard top-k keys —
 * G lists of B x k keys, list g at keys_dev + g*g_stride (u64 elements) — which are merged (same total order, so the
 * result is replica-identical) and then ranked / sorted / DPP-re-ranked as in prg_recommend.  Feature, field and
 * diversity tables must hold every global row (replicated per GPU). */
int prg_recommend_from_keys(prg_handle* h, const uint64_t* keys_dev, int G, uint64_t g_stride, int B, int k, int model,
                            const prg_dpp_params* p, uint32_t* out_row, double* out_score, int32_t* out_n, int mem);
int prg_recommend_from_keys_ex(prg_handle* h, const uint64_t* keys_dev, int G, uint64_t g_stride, int B, int k, int model,
                               const prg_dpp_params* p, const prg_user_features* user, uint32_t* out_row,
                               double* out_score, int32_t* out_n, int mem);

/* ---------------------------------------------------------------- all GPUs of a box behind one call */

/* The reference is ONE Go process (pairec.Run, pairec.go:61-86).  A prg_group owns G handles of that process — handle g
 * holds row shard g of the item matrix (prg_set_item_matrix with row_base = first global row of the shard) and replicas of
 * the field / feature / diversity tables — and runs the row-sharded request path without NCCL and without a host round
 * trip inside a batch: sample keys and candidate lists travel as stores into the peers' buffers over NVLink (P2P), the
 * GPUs order themselves with events.  Request i is ranked / re-ranked by GPU i / ceil(n_requests / G).  Results are
 * bit-identical to prg_recommend_ex over the unsharded matrix.  *out_redone (nullable) = 1 when a query failed the
 * global-threshold check and the batch was redone with exact per-shard lists.  Host buffers only; one batch at a time
 * per group; the member handles must not be destroyed before the group.  Members on the same device are allowed. */
typedef struct prg_group prg_group;
int prg_group_create(prg_handle* const* handles, int G, prg_group** out);
int prg_group_size(prg_group* grp);
int prg_group_recommend(prg_group* grp, const float* q, int n_requests, int recall_k, int model, const prg_dpp_params* p,
                        const prg_user_features* user, uint32_t* out_row, double* out_score, int32_t* out_n,
                        int32_t* out_redone);
void prg_group_destroy(prg_group* grp);

/* ---------------------------------------------------------------- cross-call request batcher */

/* The reference handles ONE request per goroutine (service/user_recommend.go:46; the rank stage then fans out into
 * BatchCount-sized RPCs, service/rank/rank_service.go:163-166, :264-289); the kernels serve up to 256 queries with one
 * pass over the item matrix.  The batcher coalesces concurrent single-request calls into prg_recommend batches:
 * it dispatches as soon as the GPU is free and a request waits (no added delay on an idle server); while a batch
 * runs, the next one fills, and a FULL batch is enqueued at once behind the running one (at most two in flight), so
 * the device does not idle between batches.  max_wait_us > 0 lets a non-full batch wait that long after its first
 * request. */
typedef struct prg_batcher prg_batcher;
typedef struct prg_batcher_config {
  int32_t max_batch;    /* 1..256 requests per prg_recommend call */
  int32_t max_wait_us;  /* 0 = dispatch immediately when the GPU is free */
  int32_t recall_k;     /* RecallConfs[].RecallCount */
  int32_t model;        /* PRG_MODEL_* */
  prg_dpp_params dpp;   /* shared by all requests of the scene */
} prg_batcher_config;
int prg_batcher_start(prg_handle* h, const prg_batcher_config* cfg, prg_batcher** out);
/* Blocks the calling thread until its request has been served.  q: dim f32 (host); out_row / out_score: top_n
 * entries, out_n: 1 entry (host).  Callable concurrently from any number of threads.  Results are those of
 * prg_recommend for the same query, whatever batch it lands in. */
int prg_batcher_recommend(prg_batcher* b, const float* q, uint32_t* out_row, double* out_score, int32_t* out_n);
/* The same with the request's user features: user_ids n_user_fields u32, user_dense n_user_dense f32 (either may be NULL). */
int prg_batcher_recommend_ex(prg_batcher* b, const float* q, const uint32_t* user_ids, const float* user_dense,
                             uint32_t* out_row, double* out_score, int32_t* out_n);
/* size_hist9: batches by size 1, 2, 3-4, 5-8, 9-16, 17-32, 33-64, 65-128, 129+ (any pointer may be NULL). */
int prg_batcher_stats(prg_batcher* b, uint64_t* n_requests, uint64_t* n_batches, uint64_t* size_hist9);
/* Closed-loop load generator (measurement aid; what a pool of request goroutines does): n_threads host threads each
 * issue per_thread blocking prg_batcher_recommend calls, request i = thread*per_thread + j taking query i % n_pool of
 * q_pool (n_pool x dim f32).  latency_us[n_threads*per_thread]: per-request latency; wall_s: the whole run;
 * rows_out[n_pool x top_n] / n_out[n_pool]: the answers of requests 0..n_pool-1 (any output may be NULL). */
int prg_batcher_drive(prg_batcher* b, const float* q_pool, int n_pool, int n_threads, int per_thread, float* latency_us,
                      double* wall_s, uint32_t* rows_out, int32_t* n_out);
/* Serves what is queued, makes late callers return PRG_ESTATE, joins the worker and frees the batcher. */
void prg_batcher_stop(prg_batcher* b);

/* ---------------------------------------------------------------- LOOKUP algorithm (algorithm/lookup.go:37-51) */

/* present[i] != 0 -> out[i] = value[i], else 0.5.  Pure host code. */
int prg_lookup(const double* value, const uint8_t* present, int n, double* out);

/* ---------------------------------------------------------------- instrumentation */

/* Number of kernels launched by this handle since init (bench.py's gpu_launches). */
uint64_t prg_launch_count(prg_handle* h);
/* Per-stage device time: enable = 1 turns on CUDA-event spans around each stage's launches (on the handle's
 * stream), enable = 2 only around the recall scan (the dominant kernel; least perturbation), 0 turns them off.  When ms_out / n_out are non-NULL the call synchronises, writes the accumulated milliseconds and span
 * counts per stage (8 entries: 0 scan, 1 scan-dense/sample, 2 select, 3 gather+FM, 4 MLP, 5 sort, 6 DPP, 7 other)
 * and resets the accumulators. */
int prg_timing(prg_handle* h, int enable, double* ms_out, uint64_t* n_out);
/* Last recall: how many queries took the dense re-scan path, and candidates collected per query (max). */
int prg_recall_stats(prg_handle* h, int32_t* n_fallback, int32_t* max_candidates);
/* Last recall: which filter the full pass over the item matrix used — PRG_FILTER_NONE (dense path: every key
 * materialised), _FFMA2 (exact fp32 scan), _TF32 (tensor cores over the fp32 rows), _BF16 (bf16 index), _INT8 (int8
 * index: dim 64, at most 64 queries per pass; config "scan_int8", default on).  Results do not depend on it. */
enum { PRG_FILTER_NONE = 0, PRG_FILTER_FFMA2 = 1, PRG_FILTER_TF32 = 2, PRG_FILTER_BF16 = 3, PRG_FILTER_INT8 = 4 };
int prg_recall_filter(prg_handle* h, int32_t* kind);

#ifdef __cplusplus
}
#endif
#endif /* PAIREC_GPU_H */
