// pipeline.cu — C-ABI entry points for the rank / DPP stages and the fused request path
// (recall -> gather+rank -> score sort -> DPP, everything device resident in between).
#include "handle.h"

namespace prg {
// gather_fm.cu
int gather_fm_device(prg_handle* h, const uint32_t* rows_dev, int M, float* logit_dev, uint16_t* x_dev,
                     const uint64_t* keys_dev, uint32_t* rows_out, int rows_per_req);
int user_prefix_device(prg_handle* h, const uint32_t* user_ids_dev, const float* user_dense_dev, int B, bool need_mlp, bool ahead);
int logits_to_scores_device(prg_handle* h, const float* a, const float* b, const uint32_t* rows_dev, int M, double* out);
// mlp.cu
int mlp_forward_device(prg_handle* h, const uint16_t* x_dev, int M, float* logit_dev, const float* fm_logit_dev,
                       const uint32_t* rows_dev, double* score_dev, double* score_map, const float* ubias, int rows_per_req);
size_t mlp_act_bytes(const prg_handle* h, int M);
// sort.cu
int sort_desc_device(prg_handle* h, const double* score_dev, int B, int n, int32_t* perm_dev, const uint32_t* rows_dev,
                     uint32_t* rows_sorted, double* scores_sorted);
// dpp.cu
int dpp_device(prg_handle* h, const uint32_t* rows_dev, const double* score_dev, int B, int n, const prg_dpp_params& p,
               int32_t* out_idx, int32_t* out_n, int32_t* status, DppFinal* fin = nullptr);

int dpp_hook_device(prg_handle* h, const uint32_t* rows_dev, const double* score_dev, const double* hook_dev, int hook_dim,
                    int use_table, int B, int n, const prg_dpp_params& p, int32_t* out_idx, int32_t* out_n, int32_t* status);
// dpp_cluster.cu
int dpp_cluster_prepare(prg_handle* h);
// ssd.cu
int ssd_device(prg_handle* h, const uint32_t* rows_dev, const double* score_dev, int B, int n, const prg_ssd_params& p,
               int32_t* out_idx, int32_t* out_n, int32_t* status);

struct DevGuard {
  std::unique_lock<std::mutex> lk;
  explicit DevGuard(prg_handle* h) : lk(h->mu) {
    cudaSetDevice(h->device);
    resolve_pending(h);  // a deferred recall check of the previous fused call (errors surface through deferred_status)
  }
};

static int adopt(const void* src, size_t bytes, int mem, const void** dst, bool* owned) {
  if (*owned && *dst) cudaFree(const_cast<void*>(*dst));
  *dst = nullptr;
  *owned = false;
  if (mem == PRG_MEM_DEVICE) { *dst = src; return PRG_OK; }
  void* d = nullptr;
  cudaError_t e = cudaMalloc(&d, bytes ? bytes : 16);
  if (e != cudaSuccess) return fail(PRG_ENOMEM, std::string("cudaMalloc table: ") + cudaGetErrorString(e));
  e = cudaMemcpy(d, src, bytes, cudaMemcpyHostToDevice);
  if (e != cudaSuccess) { cudaFree(d); return fail(PRG_ECUDA, std::string("cudaMemcpy table: ") + cudaGetErrorString(e)); }
  *dst = d;
  *owned = true;
  return PRG_OK;
}

// rank on device buffers: rows_dev [B*n] -> score_dev [B*n] f64 (Item.Score = sum_o coef_o * score_o) and, optionally,
// map_dev [B*n][heads].  keys_dev != nullptr (fused path): the candidates come as the recall's order keys; the gather
// unpacks the rows into rows_dev (which is then an OUTPUT).  user: device pointers (ids [B][U], dense [B][n_dense]) or
// nulls = every user feature absent (service/rank/algo_data.go:104-118 with an empty user map).
static int rank_device(prg_handle* h, int model, const uint32_t* rows_dev, int B, int n, double* score_dev,
                       const uint64_t* keys_dev = nullptr, const prg_user_features* user = nullptr,
                       double* map_dev = nullptr, bool reuse_prefix = false) {
  if (model != PRG_MODEL_FM && model != PRG_MODEL_MLP && model != PRG_MODEL_FM_MLP)
    return fail(PRG_EINVAL, "unknown rank model");
  const int M = B * n;
  const bool need_mlp = model != PRG_MODEL_FM;
  const bool need_fm = model != PRG_MODEL_MLP;
  if (need_mlp && h->mlp_layers == 0) return fail(PRG_ESTATE, "MLP weights not set (prg_set_mlp)");
  if (map_dev && !need_mlp) return fail(PRG_EINVAL, "per-head scores need a tower model");
  PRG_TRY(h->fm_logit.ensure((size_t)M * 4 * (1 + 4)));
  float* fm_logit = (float*)h->fm_logit.p;
  float* mlp_logit = fm_logit + M;
  uint16_t* x = nullptr;
  const bool has_user = h->n_user_fields + h->n_user_dense > 0;
  if (need_mlp) {
    if ((size_t)h->n_fields * 16 != h->mlp_k_item || h->mlp_k_user != h->n_user_fields * 16 + h->n_user_dense)
      return fail(PRG_ESTATE, "MLP input width != (n_fields + n_user_fields) * 16 + n_user_dense");
    PRG_TRY(h->act[0].ensure(mlp_act_bytes(h, M)));
    PRG_TRY(h->act[1].ensure(mlp_act_bytes(h, M)));
    x = (uint16_t*)h->act[0].p;
  }
  if (has_user && !reuse_prefix) {   // reuse_prefix: an earlier rank of the same batch left fm_state / ubias in place
    if (h->prefix_ahead) {   // the fused path launched it on the side stream before the recall: join here
      h->prefix_ahead = false;
      PRG_CUDA(cudaStreamWaitEvent(h->stream, h->ev_join, 0));
    } else {
      PRG_TRY(user_prefix_device(h, user ? user->ids : nullptr, user ? user->dense : nullptr, B, need_mlp, /*ahead=*/false));
    }
  }
  PRG_TRY(gather_fm_device(h, rows_dev, M, fm_logit, x, keys_dev, keys_dev ? const_cast<uint32_t*>(rows_dev) : nullptr,
                           has_user ? n : 0));
  if (need_mlp)
    return mlp_forward_device(h, x, M, mlp_logit, need_fm ? fm_logit : nullptr, rows_dev, score_dev, map_dev,
                              has_user ? (const float*)h->ubias.p : nullptr, n);
  return logits_to_scores_device(h, fm_logit, nullptr, rows_dev, M, score_dev);
}

// host-call staging of a batch's user features: *dev receives device pointers (nulls stay null)
static int stage_user(prg_handle* h, const prg_user_features* user, int B, int mem, prg_user_features* dev) {
  dev->ids = nullptr; dev->dense = nullptr;
  if (!user) return PRG_OK;
  if (mem == PRG_MEM_DEVICE) { *dev = *user; return PRG_OK; }
  if (user->ids && h->n_user_fields) {
    const size_t bytes = (size_t)B * h->n_user_fields * 4;
    PRG_TRY(h->user_ids_dev.ensure(bytes));
    PRG_CUDA(cudaMemcpyAsync(h->user_ids_dev.p, user->ids, bytes, cudaMemcpyHostToDevice, h->stream));
    dev->ids = (const uint32_t*)h->user_ids_dev.p;
  }
  if (user->dense && h->n_user_dense) {
    const size_t bytes = (size_t)B * h->n_user_dense * 4;
    PRG_TRY(h->user_dense_dev.ensure(bytes));
    PRG_CUDA(cudaMemcpyAsync(h->user_dense_dev.p, user->dense, bytes, cudaMemcpyHostToDevice, h->stream));
    dev->dense = (const float*)h->user_dense_dev.p;
  }
  return PRG_OK;
}

__global__ void final_gather_kernel(const uint32_t* rows, const double* scores, const int32_t* idx, const int32_t* cnt,
                                    const int32_t* status, int B, int n, int T, uint32_t* out_row, double* out_score,
                                    int32_t* out_n) {
  pdl_wait();                 // chained launch: the predecessor's writes are visible from here on
  pdl_launch_dependents();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * T) return;
  const int b = i / T, t = i - b * T;
  // status 1: the reference returns the items unchanged (sort/dpp_sort.go:317-320) -> first T of the sorted list
  const bool unchanged = status[b] != 0;
  int c = unchanged ? (n < T ? n : T) : cnt[b];
  if (t == 0) out_n[b] = c;
  if (t < c) {
    const int src = b * n + (unchanged ? t : idx[i]);
    out_row[i] = rows[src];
    out_score[i] = scores[src];
  } else {
    out_row[i] = 0xFFFFFFFFu;
    out_score[i] = 0.0;
  }
}

// stages after recall: topk_keys [B][k] (sorted, merged) -> rank -> sort -> DPP -> outputs (also called by group.cu)
int post_recall_device(prg_handle* h, int B, int k, int model, const prg_dpp_params& p, uint32_t* out_row,
                       double* out_score, int32_t* out_n, const prg_user_features& user);

// The recall's per-query status is validated AFTER the downstream stages have been enqueued (resolve_pending): in the
// steady state the host never waits in the middle of a step; a failed query (adversarial row order) is redone densely
// and the downstream stages are simply run again.
static int recommend_device(prg_handle* h, const float* q_dev, int B, int k, int model, const prg_dpp_params& p,
                            uint32_t* out_row, double* out_score, int32_t* out_n, bool resolve_now,
                            const prg_user_features& user) {
  PRG_TRY(h->topk_keys.ensure((size_t)B * k * 8));
  // the user prefix depends on the request only: it runs on the side stream while the recall streams the item index
  if (h->n_user_fields + h->n_user_dense > 0 && (model == PRG_MODEL_FM || h->mlp_layers > 0) &&
      (model == PRG_MODEL_FM || h->mlp_k_user == h->n_user_fields * 16 + h->n_user_dense))
    PRG_TRY(user_prefix_device(h, user.ids, user.dense, B,
                               model != PRG_MODEL_FM || (h->prerank_keep > 0 && h->prerank_model != PRG_MODEL_FM), /*ahead=*/true));
  {
    const int rc = recall_topk_device(h, q_dev, B, k, (uint64_t*)h->topk_keys.p, /*defer=*/true);
    if (rc != PRG_OK) { h->prefix_ahead = false; return rc; }   // nobody will join the side stream for this batch
  }
  PRG_TRY(post_recall_device(h, B, k, model, p, out_row, out_score, out_n, user));
  if (h->pending.active) {
    h->pending.fused = true;
    h->pending.model = model; h->pending.p = p; h->pending.user = user;
    h->pending.out_row = out_row; h->pending.out_score = out_score; h->pending.out_n = out_n;
  }
  return resolve_now ? resolve_pending(h) : PRG_OK;
}

int post_recall_device(prg_handle* h, int B, int k, int model, const prg_dpp_params& p, uint32_t* out_row,
                       double* out_score, int32_t* out_n, const prg_user_features& user) {
  const int M = B * k;
  PRG_TRY(h->rec_rows.ensure((size_t)M * 4));
  PRG_TRY(h->out_score.ensure((size_t)M * 4));
  PRG_TRY(h->out_n.ensure((size_t)B * 4));
  PRG_TRY(h->rec_scores.ensure((size_t)M * 8));
  PRG_TRY(h->rec_perm.ensure((size_t)M * 4));
  PRG_TRY(h->rec_sorted_rows.ensure((size_t)M * 4));
  PRG_TRY(h->rec_sorted_scores.ensure((size_t)M * 8));
  PRG_TRY(h->dpp_idx.ensure((size_t)B * p.top_n * 4));
  PRG_TRY(h->dpp_n.ensure((size_t)B * 4));
  PRG_TRY(h->dpp_status.ensure((size_t)B * 4));
  int n = k;   // candidates per request that reach rank / sort / DPP
  if (h->prerank_keep > 0 && h->prerank_keep < k) {
    // General (pre-)rank, device resident (service/general_rank/base_general_rank.go:66-109): a cheaper model scores the
    // whole recall set, the Action keeps its best `keep` (:183, ActionType sort + truncation), and only those reach the
    // full rank model — no host round trip between the two ranks.  Both ranks need the user prefix: tower share included
    // when either model has a tower.
    const int keep = h->prerank_keep;
    PRG_TRY(h->pre_rows.ensure((size_t)B * keep * 4));
    const bool need_mlp_any = h->prerank_model != PRG_MODEL_FM || model != PRG_MODEL_FM;
    if (h->n_user_fields + h->n_user_dense > 0 && !h->prefix_ahead)
      PRG_TRY(user_prefix_device(h, user.ids, user.dense, B, need_mlp_any && h->mlp_layers > 0, /*ahead=*/false));
    else if (h->prefix_ahead) { h->prefix_ahead = false; PRG_CUDA(cudaStreamWaitEvent(h->stream, h->ev_join, 0)); }
    PRG_TRY(rank_device(h, h->prerank_model, (const uint32_t*)h->rec_rows.p, B, k, (double*)h->rec_scores.p,
                        (const uint64_t*)h->topk_keys.p, &user, nullptr, /*reuse_prefix=*/true));
    PRG_TRY(sort_desc_device(h, (const double*)h->rec_scores.p, B, k, (int32_t*)h->rec_perm.p, (const uint32_t*)h->rec_rows.p,
                             (uint32_t*)h->rec_sorted_rows.p, (double*)h->rec_sorted_scores.p));
    PRG_CUDA(cudaMemcpy2DAsync(h->pre_rows.p, (size_t)keep * 4, h->rec_sorted_rows.p, (size_t)k * 4, (size_t)keep * 4, (size_t)B,
                               cudaMemcpyDeviceToDevice, h->stream));
    n = keep;
    PRG_TRY(rank_device(h, model, (const uint32_t*)h->pre_rows.p, B, n, (double*)h->rec_scores.p, nullptr, &user, nullptr,
                        /*reuse_prefix=*/true));
    PRG_TRY(sort_desc_device(h, (const double*)h->rec_scores.p, B, n, (int32_t*)h->rec_perm.p, (const uint32_t*)h->pre_rows.p,
                             (uint32_t*)h->rec_sorted_rows.p, (double*)h->rec_sorted_scores.p));
  } else {
    // 2. gather + rank (the gather unpacks the recall keys into rec_rows on the way)
    PRG_TRY(rank_device(h, model, (const uint32_t*)h->rec_rows.p, B, k, (double*)h->rec_scores.p, (const uint64_t*)h->topk_keys.p,
                        &user));
    // 3. ItemRankScore sort; the sort writes the sorted list for DPP itself
    PRG_TRY(sort_desc_device(h, (const double*)h->rec_scores.p, B, k, (int32_t*)h->rec_perm.p, (const uint32_t*)h->rec_rows.p,
                             (uint32_t*)h->rec_sorted_rows.p, (double*)h->rec_sorted_scores.p));
  }
  // 4. DPP; the cluster kernel writes the final outputs itself, the other kernels leave them to final_gather_kernel
  DppFinal fin;
  fin.row = out_row; fin.score = out_score; fin.n = out_n;
  PRG_TRY(dpp_device(h, (const uint32_t*)h->rec_sorted_rows.p, (const double*)h->rec_sorted_scores.p, B, n, p,
                     (int32_t*)h->dpp_idx.p, (int32_t*)h->dpp_n.p, (int32_t*)h->dpp_status.p, &fin));
  if (fin.done) return PRG_OK;
  const int tot = B * p.top_n;
  PRG_CUDA(launch_chained(h, final_gather_kernel, dim3((tot + 255) / 256), dim3(256), 0, 1,
                          (const uint32_t*)h->rec_sorted_rows.p, (const double*)h->rec_sorted_scores.p,
                          (const int32_t*)h->dpp_idx.p, (const int32_t*)h->dpp_n.p, (const int32_t*)h->dpp_status.p, B, n,
                          p.top_n, out_row, out_score, out_n));
  count_launch(h);
  return PRG_OK;
}

// D2H of a fused call's results (device staging -> the caller's pinned buffers), then `done`
static int copy_results_to_host(prg_handle* h, int B, int top_n, const uint32_t* row_dev, const double* score_dev,
                                const int32_t* n_dev, uint32_t* row, double* score, int32_t* n, cudaEvent_t done) {
  const size_t TT = (size_t)B * top_n;
  PRG_CUDA(cudaMemcpyAsync(row, row_dev, TT * 4, cudaMemcpyDeviceToHost, h->stream));
  PRG_CUDA(cudaMemcpyAsync(score, score_dev, TT * 8, cudaMemcpyDeviceToHost, h->stream));
  PRG_CUDA(cudaMemcpyAsync(n, n_dev, (size_t)B * 4, cudaMemcpyDeviceToHost, h->stream));
  if (done) PRG_CUDA(cudaEventRecord(done, h->stream));
  return PRG_OK;
}

int resolve_pending(prg_handle* h) {
  if (!h->pending.active) return PRG_OK;
  const bool fused = h->pending.fused;
  bool repaired = false;
  int rc = recall_resolve(h, &repaired);
  if (rc == PRG_OK && repaired && fused) {
    rc = post_recall_device(h, h->pending.B, h->pending.k, h->pending.model, h->pending.p, h->pending.out_row,
                            h->pending.out_score, h->pending.out_n, h->pending.user);
    if (rc == PRG_OK && h->pending.host_row) {   // an asynchronous call: its first copies carried the unrepaired results
      rc = copy_results_to_host(h, h->pending.B, h->pending.p.top_n, h->pending.out_row, h->pending.out_score,
                                h->pending.out_n, h->pending.host_row, h->pending.host_score, h->pending.host_n,
                                h->pending.done);
      h->repaired_seq[h->repaired_pos++ & 3u] = h->pending.seq;
    }
  }
  if (rc != PRG_OK) { h->deferred_status = rc; h->deferred_msg = prg_last_error(); }
  return rc;
}

int recommend_begin(prg_handle* h, const float* q, int B, int recall_k, int model, const prg_dpp_params& p,
                    const prg_user_features* user, uint32_t* out_row, double* out_score, int32_t* out_n, cudaEvent_t done,
                    uint64_t* seq) {
  DevGuard g(h);   // settles the previous call's deferred check first (its scratch is about to be reused)
  if (!h->E) return fail(PRG_ESTATE, "item matrix not set (prg_set_item_matrix)");
  prg_user_features udev;
  PRG_TRY(stage_user(h, user, B, PRG_MEM_HOST, &udev));
  if (h->deferred_status != PRG_OK) { const int rc = h->deferred_status; h->deferred_status = PRG_OK; return fail(rc, "deferred: " + h->deferred_msg); }
  const size_t TT = (size_t)B * p.top_n;
  PRG_TRY(h->q_dev.ensure((size_t)B * h->E_dim * 4));
  PRG_TRY(h->out_row.ensure(TT * 4 > (size_t)B * recall_k * 4 ? TT * 4 : (size_t)B * recall_k * 4));
  PRG_TRY(h->rank_out.ensure(TT * 8));
  PRG_TRY(h->flags.ensure((size_t)(B > 64 ? B : 64) * 4));
  PRG_TRY(h->sort_perm.ensure((size_t)B * 4));
  PRG_CUDA(cudaMemcpyAsync(h->q_dev.p, q, (size_t)B * h->E_dim * 4, cudaMemcpyHostToDevice, h->stream));
  PRG_TRY(recommend_device(h, (const float*)h->q_dev.p, B, recall_k, model, p, (uint32_t*)h->out_row.p,
                           (double*)h->rank_out.p, (int32_t*)h->sort_perm.p, /*resolve_now=*/false, udev));
  PRG_TRY(copy_results_to_host(h, B, p.top_n, (const uint32_t*)h->out_row.p, (const double*)h->rank_out.p,
                               (const int32_t*)h->sort_perm.p, out_row, out_score, out_n, done));
  *seq = ++h->call_seq;
  if (h->pending.active) {
    h->pending.host_row = out_row; h->pending.host_score = out_score; h->pending.host_n = out_n;
    h->pending.done = done; h->pending.seq = *seq;
  }
  return PRG_OK;
}

int recommend_end(prg_handle* h, uint64_t seq, cudaEvent_t done) {
  cudaSetDevice(h->device);
  PRG_CUDA(cudaEventSynchronize(done));
  int rc = PRG_OK;
  bool repaired = false;
  {
    // the plain lock, not DevGuard: a LATER call's deferred check is not ours to wait for
    std::unique_lock<std::mutex> lk(h->mu);
    if (h->pending.active && h->pending.seq == seq) rc = resolve_pending(h);
    for (uint64_t r : h->repaired_seq) repaired |= (r == seq);
    if (rc == PRG_OK && h->deferred_status != PRG_OK) {
      rc = h->deferred_status;
      h->deferred_status = PRG_OK;
      fail(rc, "deferred: " + h->deferred_msg);
    }
  }
  if (rc != PRG_OK) return rc;
  if (repaired) PRG_CUDA(cudaEventSynchronize(done));   // re-recorded behind the repaired results' copies
  return PRG_OK;
}

}  // namespace prg

using namespace prg;

extern "C" {

int prg_set_item_fields(prg_handle* h, const uint32_t* ids, uint64_t rows, uint32_t n_fields, int mem) {
  if (!h) return fail(PRG_EINVAL, "null handle");
  if (!ids || rows == 0 || n_fields == 0 || n_fields > (uint32_t)kMaxFields)
    return fail(PRG_EINVAL, "bad item fields (1 <= n_fields <= 64)");
  DevGuard g(h);
  PRG_CUDA(cudaStreamSynchronize(h->stream));
  const void* d = h->fields;
  PRG_TRY(adopt(ids, (size_t)rows * n_fields * 4, mem, &d, &h->fields_owned));
  h->fields = (const uint32_t*)d;
  h->fields_rows = rows;
  h->n_fields = n_fields;
  return PRG_OK;
}

int prg_set_feature_table(prg_handle* h, int table, const float* factors, const float* linear, uint64_t rows,
                          uint32_t fdim, int mem) {
  if (!h) return fail(PRG_EINVAL, "null handle");
  if (table < 0 || table >= kMaxFields) return fail(PRG_EINVAL, "table index out of range");
  if (!factors || rows == 0) return fail(PRG_EINVAL, "empty feature table");
  if (fdim != 16) return fail(PRG_EUNSUPPORTED, "feature tables must have fdim == 16");
  if (rows > 0xFFFFFFFFull) return fail(PRG_EUNSUPPORTED, "feature table rows must fit u32");
  DevGuard g(h);
  PRG_CUDA(cudaStreamSynchronize(h->stream));
  Table& t = h->tables[table];
  bool own_f = t.owned, own_l = t.owned && t.linear;
  const void* f = t.factors;
  const void* l = t.linear;
  PRG_TRY(adopt(factors, (size_t)rows * fdim * 4, mem, &f, &own_f));
  if (linear) {
    PRG_TRY(adopt(linear, (size_t)rows * 4, mem, &l, &own_l));
  } else {
    if (own_l && l) cudaFree(const_cast<void*>(l));
    l = nullptr;
  }
  t.factors = (const float*)f;
  t.linear = (const float*)l;
  t.rows = rows;
  t.owned = own_f;
  h->fdim = fdim;
  return PRG_OK;
}

int prg_set_fm_bias(prg_handle* h, float w0) {
  if (!h) return fail(PRG_EINVAL, "null handle");
  DevGuard g(h);
  h->fm_w0 = w0;
  return PRG_OK;
}

int prg_set_diversity_matrix(prg_handle* h, const void* data, uint64_t rows, uint32_t dim, int dtype, int mem) {
  if (!h) return fail(PRG_EINVAL, "null handle");
  if (!data || rows == 0 || dim == 0) return fail(PRG_EINVAL, "empty diversity matrix");
  if (dtype != PRG_F32 && dtype != PRG_F64) return fail(PRG_EINVAL, "dtype must be PRG_F32 or PRG_F64");
  if (dim > 512) return fail(PRG_EUNSUPPORTED, "diversity dim > 512");
  DevGuard g(h);
  PRG_CUDA(cudaStreamSynchronize(h->stream));
  const void* d = h->D;
  PRG_TRY(adopt(data, (size_t)rows * dim * (dtype == PRG_F64 ? 8 : 4), mem, &d, &h->D_owned));
  h->D = d;
  h->D_rows = rows;
  h->D_dim = dim;
  h->D_dtype = dtype;
  PRG_TRY(dpp_cluster_prepare(h));
  return PRG_OK;
}

int prg_set_user_fields(prg_handle* h, uint32_t n_user_fields, uint32_t n_user_dense) {
  if (!h) return fail(PRG_EINVAL, "null handle");
  if (n_user_fields > (uint32_t)kMaxFields || n_user_dense > 64) return fail(PRG_EINVAL, "n_user_fields <= 64, n_user_dense <= 64");
  DevGuard g(h);
  PRG_CUDA(cudaStreamSynchronize(h->stream));
  if (n_user_fields != h->n_user_fields || n_user_dense != h->n_user_dense) h->mlp_layers = 0;   // prg_set_mlp splits W1 by these
  h->n_user_fields = n_user_fields;
  h->n_user_dense = n_user_dense;
  return PRG_OK;
}

int prg_set_prerank(prg_handle* h, int model, int keep) {
  if (!h) return fail(PRG_EINVAL, "null handle");
  if (keep < 0 || (keep > 0 && model != PRG_MODEL_FM && model != PRG_MODEL_MLP && model != PRG_MODEL_FM_MLP))
    return fail(PRG_EINVAL, "bad pre-rank model / keep");
  DevGuard g(h);
  h->prerank_model = model;
  h->prerank_keep = keep;
  return PRG_OK;
}

int prg_set_rank_score(prg_handle* h, const double* coef, int n) {
  if (!h) return fail(PRG_EINVAL, "null handle");
  if (!coef || n < 1 || n > 4) return fail(PRG_EINVAL, "1..4 coefficients");
  DevGuard g(h);
  for (int o = 0; o < 4; ++o) h->rank_coef[o] = o < n ? coef[o] : 0.0;
  return PRG_OK;
}

int prg_rank_ex(prg_handle* h, int model, const uint32_t* rows, int B, int n, const prg_user_features* user,
                double* out_score, double* out_score_map, int mem) {
  if (!h) return fail(PRG_EINVAL, "null handle");
  if (!rows || (!out_score && !out_score_map) || B <= 0 || n <= 0) return fail(PRG_EINVAL, "bad arguments");
  DevGuard g(h);
  const int M = B * n;
  const int heads = h->mlp_layers ? (int)h->mlp_dims[h->mlp_layers] : 1;
  prg_user_features udev;
  PRG_TRY(stage_user(h, user, B, mem, &udev));
  if (mem == PRG_MEM_DEVICE) {
    if (!out_score) { PRG_TRY(h->rank_out.ensure((size_t)M * 8)); out_score = (double*)h->rank_out.p; }
    return rank_device(h, model, rows, B, n, out_score, nullptr, &udev, out_score_map);
  }
  PRG_TRY(h->rank_rows.ensure((size_t)M * 4));
  PRG_TRY(h->rank_out.ensure((size_t)M * 8));
  if (out_score_map) PRG_TRY(h->rank_map.ensure((size_t)M * heads * 8));
  PRG_CUDA(cudaMemcpyAsync(h->rank_rows.p, rows, (size_t)M * 4, cudaMemcpyHostToDevice, h->stream));
  PRG_TRY(rank_device(h, model, (const uint32_t*)h->rank_rows.p, B, n, (double*)h->rank_out.p, nullptr, &udev,
                      out_score_map ? (double*)h->rank_map.p : nullptr));
  if (out_score) PRG_CUDA(cudaMemcpyAsync(out_score, h->rank_out.p, (size_t)M * 8, cudaMemcpyDeviceToHost, h->stream));
  if (out_score_map)
    PRG_CUDA(cudaMemcpyAsync(out_score_map, h->rank_map.p, (size_t)M * heads * 8, cudaMemcpyDeviceToHost, h->stream));
  PRG_CUDA(cudaStreamSynchronize(h->stream));
  return PRG_OK;
}

int prg_rank(prg_handle* h, int model, const uint32_t* rows, int B, int n, double* out_score, int mem) {
  return prg_rank_ex(h, model, rows, B, n, nullptr, out_score, nullptr, mem);
}

int prg_dpp_ex(prg_handle* h, const uint32_t* rows, const double* score, const double* hook, int hook_dim, int use_table,
               int B, int n, const prg_dpp_params* p, int32_t* out_idx, int32_t* out_n, int32_t* status, int mem) {
  if (!h) return fail(PRG_EINVAL, "null handle");
  const bool hooks = hook != nullptr && hook_dim > 0;
  if (!hooks) use_table = 1;
  if ((use_table && !rows) || !score || !p || !out_idx || !out_n || !status) return fail(PRG_EINVAL, "null buffer");
  if (B <= 0 || n <= 0 || p->top_n <= 0) return fail(PRG_EINVAL, "B, n, top_n must be positive");
  DevGuard g(h);
  if (use_table && !h->D) return fail(PRG_ESTATE, "diversity matrix not set (prg_set_diversity_matrix)");
  auto run = [&](const uint32_t* r, const double* s, const double* hk, int32_t* oi, int32_t* on, int32_t* st) {
    if (hooks) return dpp_hook_device(h, r, s, hk, hook_dim, use_table, B, n, *p, oi, on, st);
    return dpp_device(h, r, s, B, n, *p, oi, on, st);
  };
  if (mem == PRG_MEM_DEVICE) return run(rows, score, hook, out_idx, out_n, status);
  const size_t M = (size_t)B * n, TT = (size_t)B * p->top_n;
  PRG_TRY(h->dpp_rows.ensure(M * 4));
  PRG_TRY(h->dpp_score.ensure(M * 8));
  PRG_TRY(h->dpp_idx.ensure(TT * 4));
  PRG_TRY(h->dpp_n.ensure((size_t)B * 4));
  PRG_TRY(h->dpp_status.ensure((size_t)B * 4));
  if (rows) PRG_CUDA(cudaMemcpyAsync(h->dpp_rows.p, rows, M * 4, cudaMemcpyHostToDevice, h->stream));
  PRG_CUDA(cudaMemcpyAsync(h->dpp_score.p, score, M * 8, cudaMemcpyHostToDevice, h->stream));
  if (hooks) {
    PRG_TRY(h->dpp_hook_in.ensure(M * hook_dim * 8));
    PRG_CUDA(cudaMemcpyAsync(h->dpp_hook_in.p, hook, M * hook_dim * 8, cudaMemcpyHostToDevice, h->stream));
  }
  PRG_CUDA(cudaMemsetAsync(h->dpp_idx.p, 0xFF, TT * 4, h->stream));
  PRG_TRY(run(rows ? (const uint32_t*)h->dpp_rows.p : nullptr, (const double*)h->dpp_score.p,
              hooks ? (const double*)h->dpp_hook_in.p : nullptr, (int32_t*)h->dpp_idx.p, (int32_t*)h->dpp_n.p,
              (int32_t*)h->dpp_status.p));
  PRG_CUDA(cudaMemcpyAsync(out_idx, h->dpp_idx.p, TT * 4, cudaMemcpyDeviceToHost, h->stream));
  PRG_CUDA(cudaMemcpyAsync(out_n, h->dpp_n.p, (size_t)B * 4, cudaMemcpyDeviceToHost, h->stream));
  PRG_CUDA(cudaMemcpyAsync(status, h->dpp_status.p, (size_t)B * 4, cudaMemcpyDeviceToHost, h->stream));
  PRG_CUDA(cudaStreamSynchronize(h->stream));
  return PRG_OK;
}

int prg_dpp(prg_handle* h, const uint32_t* rows, const double* score, int B, int n, const prg_dpp_params* p,
            int32_t* out_idx, int32_t* out_n, int32_t* status, int mem) {
  return prg_dpp_ex(h, rows, score, nullptr, 0, 1, B, n, p, out_idx, out_n, status, mem);
}

int prg_ssd(prg_handle* h, const uint32_t* rows, const double* score, int B, int n, const prg_ssd_params* p,
            int32_t* out_idx, int32_t* out_n, int32_t* status, int mem) {
  if (!h) return fail(PRG_EINVAL, "null handle");
  if (!rows || !score || !p || !out_idx || !out_n || !status) return fail(PRG_EINVAL, "null buffer");
  if (B <= 0 || n <= 0 || p->top_n <= 0) return fail(PRG_EINVAL, "B, n, top_n must be positive");
  DevGuard g(h);
  if (mem == PRG_MEM_DEVICE) return ssd_device(h, rows, score, B, n, *p, out_idx, out_n, status);
  const size_t M = (size_t)B * n, TT = (size_t)B * p->top_n;
  PRG_TRY(h->dpp_rows.ensure(M * 4));
  PRG_TRY(h->dpp_score.ensure(M * 8));
  PRG_TRY(h->dpp_idx.ensure(TT * 4));
  PRG_TRY(h->dpp_n.ensure((size_t)B * 4));
  PRG_TRY(h->dpp_status.ensure((size_t)B * 4));
  PRG_CUDA(cudaMemcpyAsync(h->dpp_rows.p, rows, M * 4, cudaMemcpyHostToDevice, h->stream));
  PRG_CUDA(cudaMemcpyAsync(h->dpp_score.p, score, M * 8, cudaMemcpyHostToDevice, h->stream));
  PRG_CUDA(cudaMemsetAsync(h->dpp_idx.p, 0xFF, TT * 4, h->stream));
  PRG_TRY(ssd_device(h, (const uint32_t*)h->dpp_rows.p, (const double*)h->dpp_score.p, B, n, *p, (int32_t*)h->dpp_idx.p,
                     (int32_t*)h->dpp_n.p, (int32_t*)h->dpp_status.p));
  PRG_CUDA(cudaMemcpyAsync(out_idx, h->dpp_idx.p, TT * 4, cudaMemcpyDeviceToHost, h->stream));
  PRG_CUDA(cudaMemcpyAsync(out_n, h->dpp_n.p, (size_t)B * 4, cudaMemcpyDeviceToHost, h->stream));
  PRG_CUDA(cudaMemcpyAsync(status, h->dpp_status.p, (size_t)B * 4, cudaMemcpyDeviceToHost, h->stream));
  PRG_CUDA(cudaStreamSynchronize(h->stream));
  return PRG_OK;
}

int prg_recommend_from_keys_ex(prg_handle* h, const uint64_t* keys_dev, int G, uint64_t g_stride, int B, int k, int model,
                               const prg_dpp_params* p, const prg_user_features* user, uint32_t* out_row,
                               double* out_score, int32_t* out_n, int mem) {
  if (!h) return fail(PRG_EINVAL, "null handle");
  if (!keys_dev || !p || !out_row || !out_score || !out_n) return fail(PRG_EINVAL, "null buffer");
  if (G <= 0 || B <= 0 || k <= 0 || p->top_n <= 0) return fail(PRG_EINVAL, "G, B, k, top_n must be positive");
  DevGuard g(h);
  prg_user_features udev;
  PRG_TRY(stage_user(h, user, B, mem, &udev));
  PRG_TRY(h->topk_keys.ensure((size_t)B * k * 8));
  // the user prefix runs on the side stream beside the merge of the G lists (it depends on the request only)
  if (h->n_user_fields + h->n_user_dense > 0 && (model == PRG_MODEL_FM || h->mlp_layers > 0) &&
      (model == PRG_MODEL_FM || h->mlp_k_user == h->n_user_fields * 16 + h->n_user_dense))
    PRG_TRY(user_prefix_device(h, udev.ids, udev.dense, B,
                               model != PRG_MODEL_FM || (h->prerank_keep > 0 && h->prerank_model != PRG_MODEL_FM), /*ahead=*/true));
  {
    const int rc = merge_keys_device(h, keys_dev, G, g_stride, B, k, (uint64_t*)h->topk_keys.p);
    if (rc != PRG_OK) { h->prefix_ahead = false; return rc; }
  }
  if (mem == PRG_MEM_DEVICE) return post_recall_device(h, B, k, model, *p, out_row, out_score, out_n, udev);
  const size_t TT = (size_t)B * p->top_n;
  PRG_TRY(h->out_row.ensure(TT * 4 > (size_t)B * k * 4 ? TT * 4 : (size_t)B * k * 4));
  PRG_TRY(h->rank_out.ensure(TT * 8));
  PRG_TRY(h->sort_perm.ensure((size_t)B * 4));
  PRG_TRY(post_recall_device(h, B, k, model, *p, (uint32_t*)h->out_row.p, (double*)h->rank_out.p, (int32_t*)h->sort_perm.p, udev));
  PRG_CUDA(cudaMemcpyAsync(out_row, h->out_row.p, TT * 4, cudaMemcpyDeviceToHost, h->stream));
  PRG_CUDA(cudaMemcpyAsync(out_score, h->rank_out.p, TT * 8, cudaMemcpyDeviceToHost, h->stream));
  PRG_CUDA(cudaMemcpyAsync(out_n, h->sort_perm.p, (size_t)B * 4, cudaMemcpyDeviceToHost, h->stream));
  PRG_CUDA(cudaStreamSynchronize(h->stream));
  return PRG_OK;
}

int prg_recommend_from_keys(prg_handle* h, const uint64_t* keys_dev, int G, uint64_t g_stride, int B, int k, int model,
                            const prg_dpp_params* p, uint32_t* out_row, double* out_score, int32_t* out_n, int mem) {
  return prg_recommend_from_keys_ex(h, keys_dev, G, g_stride, B, k, model, p, nullptr, out_row, out_score, out_n, mem);
}

int prg_recommend_ex(prg_handle* h, const float* q, int B, int recall_k, int model, const prg_dpp_params* p,
                     const prg_user_features* user, uint32_t* out_row, double* out_score, int32_t* out_n, int mem) {
  if (!h) return fail(PRG_EINVAL, "null handle");
  if (!q || !p || !out_row || !out_score || !out_n) return fail(PRG_EINVAL, "null buffer");
  if (B <= 0 || recall_k <= 0 || p->top_n <= 0) return fail(PRG_EINVAL, "B, recall_k, top_n must be positive");
  DevGuard g(h);
  if (!h->E) return fail(PRG_ESTATE, "item matrix not set (prg_set_item_matrix)");
  if (h->deferred_status != PRG_OK) { const int rc = h->deferred_status; h->deferred_status = PRG_OK; return fail(rc, "deferred: " + h->deferred_msg); }
  prg_user_features udev;
  PRG_TRY(stage_user(h, user, B, mem, &udev));
  // PRG_MEM_DEVICE: the recall's exactness check is settled before the call returns (the host waits for the recall
  // part only; everything downstream is already enqueued) unless the handle was created with "defer_check":1
  if (mem == PRG_MEM_DEVICE)
    return recommend_device(h, q, B, recall_k, model, *p, out_row, out_score, out_n, /*resolve_now=*/!h->defer_check, udev);
  const size_t TT = (size_t)B * p->top_n;
  PRG_TRY(h->q_dev.ensure((size_t)B * h->E_dim * 4));
  PRG_TRY(h->out_row.ensure(TT * 4 > (size_t)B * recall_k * 4 ? TT * 4 : (size_t)B * recall_k * 4));
  PRG_TRY(h->rank_out.ensure(TT * 8));
  PRG_TRY(h->flags.ensure((size_t)(B > 64 ? B : 64) * 4));
  PRG_CUDA(cudaMemcpyAsync(h->q_dev.p, q, (size_t)B * h->E_dim * 4, cudaMemcpyHostToDevice, h->stream));
  PRG_TRY(h->sort_perm.ensure((size_t)B * 4));
  PRG_TRY(recommend_device(h, (const float*)h->q_dev.p, B, recall_k, model, *p, (uint32_t*)h->out_row.p,
                           (double*)h->rank_out.p, (int32_t*)h->sort_perm.p, /*resolve_now=*/true, udev));
  PRG_CUDA(cudaMemcpyAsync(out_row, h->out_row.p, TT * 4, cudaMemcpyDeviceToHost, h->stream));
  PRG_CUDA(cudaMemcpyAsync(out_score, h->rank_out.p, TT * 8, cudaMemcpyDeviceToHost, h->stream));
  PRG_CUDA(cudaMemcpyAsync(out_n, h->sort_perm.p, (size_t)B * 4, cudaMemcpyDeviceToHost, h->stream));
  PRG_CUDA(cudaStreamSynchronize(h->stream));
  return PRG_OK;
}

int prg_recommend(prg_handle* h, const float* q, int B, int recall_k, int model, const prg_dpp_params* p,
                  uint32_t* out_row, double* out_score, int32_t* out_n, int mem) {
  return prg_recommend_ex(h, q, B, recall_k, model, p, nullptr, out_row, out_score, out_n, mem);
}

}  // extern "C"
