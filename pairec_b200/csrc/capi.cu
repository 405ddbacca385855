// capi.cu — the extern "C" surface declared in include/pairec_gpu.h: handle lifecycle, table residency, and the
// host-buffer / device-buffer wrappers around the kernels.  No CPU compute path exists here: without a CUDA device
// prg_init fails with PRG_ENODEVICE.
#include "handle.h"
#include "recall.h"
#include <cstring>
#include <cstdlib>
#include <cmath>
#include <new>
#include <utility>

namespace prg {

static thread_local std::string g_err;
void set_error(const std::string& msg) { g_err = msg; }
int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}

int DevBuf::ensure(size_t need) {
  if (need <= bytes && p) return PRG_OK;
  if (p) { cudaFree(p); p = nullptr; bytes = 0; }
  if (need == 0) need = 16;
  cudaError_t e = cudaMalloc(&p, need);
  if (e != cudaSuccess) {
    p = nullptr;
    return fail(PRG_ENOMEM, std::string("cudaMalloc(") + std::to_string(need) + "): " + cudaGetErrorString(e));
  }
  bytes = need;
  return PRG_OK;
}
void DevBuf::release() {
  if (p) cudaFree(p);
  p = nullptr;
  bytes = 0;
}

PFN_encodeTiled get_encode_tiled() {
  // resolved once; a function-local static's initialisation is thread-safe (snapshots are staged from other threads)
  static const PFN_encodeTiled fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr) == cudaSuccess &&
        qr == cudaDriverEntryPointSuccess)
      return reinterpret_cast<PFN_encodeTiled>(p);
    return static_cast<PFN_encodeTiled>(nullptr);
  }();
  return fn;
}

// minimal "key": integer lookup in a flat JSON object (prg_init's config is a handful of integers)
static bool json_int(const char* js, const char* key, long* out) {
  if (!js) return false;
  std::string pat = std::string("\"") + key + "\"";
  const char* p = strstr(js, pat.c_str());
  if (!p) return false;
  p += pat.size();
  while (*p == ' ' || *p == '\t' || *p == '\n' || *p == ':') ++p;
  char* end = nullptr;
  long v = strtol(p, &end, 10);
  if (end == p) return false;
  *out = v;
  return true;
}

struct Guard {
  prg_handle* h;
  std::unique_lock<std::mutex> lk;
  int prev = -1;
  explicit Guard(prg_handle* hh) : h(hh), lk(hh->mu) {
    cudaGetDevice(&prev);
    if (prev != h->device) cudaSetDevice(h->device);
    resolve_pending(h);  // deferred recall check of a previous fused call
  }
  ~Guard() {
    if (prev >= 0 && prev != h->device) cudaSetDevice(prev);
  }
};

// copy-or-adopt helper for tables
static int take_table(const void* src, size_t bytes, int mem, const void** dst, bool* owned) {
  if (*owned && *dst) cudaFree(const_cast<void*>(*dst));
  *dst = nullptr;
  *owned = false;
  if (mem == PRG_MEM_DEVICE) {
    *dst = src;
    return PRG_OK;
  }
  void* d = nullptr;
  cudaError_t e = cudaMalloc(&d, bytes ? bytes : 16);
  if (e != cudaSuccess) return fail(PRG_ENOMEM, std::string("cudaMalloc table: ") + cudaGetErrorString(e));
  e = cudaMemcpy(d, src, bytes, cudaMemcpyHostToDevice);
  if (e != cudaSuccess) {
    cudaFree(d);
    return fail(PRG_ECUDA, std::string("cudaMemcpy table: ") + cudaGetErrorString(e));
  }
  *dst = d;
  *owned = true;
  return PRG_OK;
}

// A staged item-matrix snapshot lives in a bare prg_handle of its own (only the item-matrix fields and a side stream are
// used): the index builders of recall.cu / recall_tc.cu run on it unchanged while the live handle keeps serving.
static void free_snapshot(prg_handle* s) {
  if (!s) return;
  if (s->stream) cudaStreamSynchronize(s->stream);
  if (s->E_owned && s->E) cudaFree(const_cast<float*>(s->E));
  s->E16.release();
  s->E8.release();
  s->E8_prm.release();
  s->row_norm.release();
  if (s->stream) cudaStreamDestroy(s->stream);
  delete s;
}

}  // namespace prg

using namespace prg;

#define CHECK_H(h) \
  if (!(h)) return prg::fail(PRG_EINVAL, "null handle")

extern "C" {

const char* prg_last_error(void) { return prg::g_err.c_str(); }
const char* prg_version(void) { return "pairec_b200 0.1 (sm_100a)"; }

int prg_init(const char* json_cfg, prg_handle** out) {
  if (!out) return fail(PRG_EINVAL, "out is null");
  *out = nullptr;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return fail(PRG_ENODEVICE, std::string("no CUDA device (") + cudaGetErrorString(e) +
                                   "); libpairec_gpu has no CPU fallback");
  long v;
  int dev = 0;
  if (json_int(json_cfg, "device", &v)) dev = (int)v;
  if (dev < 0 || dev >= ndev) return fail(PRG_EINVAL, "device index out of range");
  PRG_CUDA(cudaSetDevice(dev));
  cudaDeviceProp prop;
  PRG_CUDA(cudaGetDeviceProperties(&prop, dev));
  if (prop.major != 10)
    return fail(PRG_EUNSUPPORTED, std::string("device is sm_") + std::to_string(prop.major) + std::to_string(prop.minor) +
                                      ", this library is built for sm_100a only");
  prg_handle* h = new (std::nothrow) prg_handle();
  if (!h) return fail(PRG_ENOMEM, "handle allocation failed");
  h->device = dev;
  h->sm_count = prop.multiProcessorCount;
  if (json_int(json_cfg, "max_batch", &v) && v > 0) h->max_batch = (int)v;
  if (json_int(json_cfg, "max_k", &v) && v > 0) h->max_k = (int)v;
  if (json_int(json_cfg, "sm_limit", &v) && v > 0 && v < h->sm_count) h->sm_count = (int)v;
  if (json_int(json_cfg, "scan_ffma2", &v) && v != 0) h->scan_ffma2 = true;
  if (json_int(json_cfg, "scan_tf32", &v) && v != 0) h->scan_filter = SCAN_FILTER_TF32;
  if (json_int(json_cfg, "dpp_generic", &v) && v != 0) h->dpp_generic = true;
  if (json_int(json_cfg, "dpp_lazy", &v) && v != 0) h->dpp_lazy = true;
  if (json_int(json_cfg, "scan128_nqb", &v)) h->scan128_nqb = v != 0;
  if (const char* ev = getenv("PRG_SCAN128_NQB")) h->scan128_nqb = atoi(ev) != 0;
  if (json_int(json_cfg, "recall_tilemax", &v)) h->recall_tilemax = v != 0;
  if (const char* ev = getenv("PRG_RECALL_TILEMAX")) h->recall_tilemax = atoi(ev) != 0;
  if (json_int(json_cfg, "scan_int8", &v)) h->scan_int8 = v != 0;
  if (json_int(json_cfg, "scan_grp16", &v)) h->scan_grp16 = v != 0;
  if (const char* ev = getenv("PRG_SCAN_GRP16")) h->scan_grp16 = atoi(ev) != 0;
  if (const char* ev = getenv("PRG_SCAN_INT8")) h->scan_int8 = atoi(ev) != 0;
  if (json_int(json_cfg, "scan_groups", &v)) h->scan_groups = v < 0 ? -1 : (v != 0);
  if (const char* ev = getenv("PRG_SCAN_GROUPS")) h->scan_groups = atoi(ev) < 0 ? -1 : (atoi(ev) != 0);
  if (json_int(json_cfg, "dpp_pair", &v)) h->dpp_pair = v != 0;
  if (const char* ev = getenv("PRG_DPP_PAIR")) h->dpp_pair = atoi(ev) != 0;
  if (json_int(json_cfg, "mlp_no_pair", &v) && v != 0) h->mlp_no_pair = true;
  if (json_int(json_cfg, "defer_check", &v)) h->defer_check = v != 0;
  if (json_int(json_cfg, "pdl", &v)) h->pdl = v != 0;
  if (const char* ev = getenv("PRG_PDL")) h->pdl = atoi(ev) != 0;   // A/B measurements without a config change
  e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking);
  if (e != cudaSuccess) {
    delete h;
    return fail(PRG_ECUDA, std::string("cudaStreamCreate: ") + cudaGetErrorString(e));
  }
  *out = h;
  return PRG_OK;
}

void prg_destroy(prg_handle* h) {
  if (!h) return;
  {
    Guard g(h);
    cudaStreamSynchronize(h->stream);
    free_snapshot(h->staged);
    h->staged = nullptr;
    if (h->E_owned && h->E) cudaFree(const_cast<float*>(h->E));
    if (h->fields_owned && h->fields) cudaFree(const_cast<uint32_t*>(h->fields));
    if (h->D_owned && h->D) cudaFree(const_cast<void*>(h->D));
    for (int t = 0; t < kMaxFields; ++t) {
      if (h->tables[t].owned) {
        if (h->tables[t].factors) cudaFree(const_cast<float*>(h->tables[t].factors));
        if (h->tables[t].linear) cudaFree(const_cast<float*>(h->tables[t].linear));
      }
    }
    DevBuf* bufs[] = {&h->q_dev, &h->sample_keys, &h->cand_keys, &h->seg_keys, &h->seg_rows, &h->grp_cnt, &h->row_norm, &h->E16, &h->E8, &h->E8_prm, &h->cand_cnt, &h->tau, &h->dense_keys, &h->topk_keys,
                      &h->out_row, &h->out_score, &h->out_n, &h->flags, &h->table_ptrs, &h->act[0], &h->act[1],
                      &h->fm_logit, &h->rank_rows, &h->rank_out, &h->mlp_Wu, &h->user_ids_dev, &h->user_dense_dev, &h->fm_state, &h->ubias, &h->rank_map, &h->pre_rows, &h->D_sub, &h->D_sub_inv, &h->dpp_hook_E, &h->dpp_hook_rows, &h->dpp_hook_in, &h->dpp_scratch, &h->dpp_rows, &h->dpp_score,
                      &h->dpp_idx, &h->dpp_n, &h->dpp_status, &h->ssd_E, &h->ssd_P, &h->sort_in, &h->sort_perm, &h->rec_rows,
                      &h->rec_scores, &h->rec_perm, &h->rec_sorted_rows, &h->rec_sorted_scores};
    for (DevBuf* b : bufs) b->release();
    for (int l = 0; l < kMaxLayers; ++l) { h->mlp_W[l].release(); h->mlp_b[l].release(); }
    if (h->flags_ev) cudaEventDestroy(h->flags_ev);
    if (h->side_stream) { cudaStreamSynchronize(h->side_stream); cudaStreamDestroy(h->side_stream); }
    if (h->ev_fork) cudaEventDestroy(h->ev_fork);
    if (h->ev_join) cudaEventDestroy(h->ev_join);
    if (h->host_flags) cudaFreeHost(h->host_flags);
    for (auto& sp : h->spans) { cudaEventDestroy(sp.a); cudaEventDestroy(sp.b); }
    for (auto e : h->ev_pool) cudaEventDestroy(e);
    cudaStreamDestroy(h->stream);
  }
  delete h;
}

int prg_sync(prg_handle* h) {
  CHECK_H(h);
  Guard g(h);  // resolves a deferred recall check (and re-runs the fused downstream if a query had to be repaired)
  PRG_CUDA(cudaStreamSynchronize(h->stream));
  if (h->deferred_status != PRG_OK) {
    const int rc = h->deferred_status;
    h->deferred_status = PRG_OK;
    return fail(rc, "deferred: " + h->deferred_msg);
  }
  return PRG_OK;
}
void* prg_stream(prg_handle* h) { return h ? (void*)h->stream : nullptr; }
uint32_t prg_item_dim(prg_handle* h) { return (h && h->E) ? h->E_dim : 0; }
uint64_t prg_launch_count(prg_handle* h) { return h ? h->launches : 0; }

int prg_timing(prg_handle* h, int enable, double* ms_out, uint64_t* n_out) {
  CHECK_H(h);
  Guard g(h);
  if (ms_out || n_out) {
    PRG_CUDA(cudaStreamSynchronize(h->stream));
    for (auto& sp : h->spans) {
      float ms = 0.f;
      if (cudaEventElapsedTime(&ms, sp.a, sp.b) == cudaSuccess) { h->stage_ms[sp.stage] += ms; h->stage_n[sp.stage] += 1; }
      h->ev_pool.push_back(sp.a);
      h->ev_pool.push_back(sp.b);
    }
    h->spans.clear();
    for (int i = 0; i < 8; ++i) {
      if (ms_out) ms_out[i] = h->stage_ms[i];
      if (n_out) n_out[i] = h->stage_n[i];
      h->stage_ms[i] = 0;
      h->stage_n[i] = 0;
    }
  }
  h->timing = enable;
  return PRG_OK;
}

int prg_recall_stats(prg_handle* h, int32_t* n_fallback, int32_t* max_candidates) {
  CHECK_H(h);
  if (n_fallback) *n_fallback = h->last_fallback;
  if (max_candidates) *max_candidates = h->last_max_cand;
  return PRG_OK;
}

int prg_recall_filter(prg_handle* h, int32_t* kind) {
  CHECK_H(h);
  if (!kind) return fail(PRG_EINVAL, "null buffer");
  *kind = h->last_filter;
  return PRG_OK;
}

// ------------------------------------------------------------------ tables
int prg_set_item_matrix(prg_handle* h, const float* data, uint64_t rows, uint32_t dim, uint64_t row_base, int mem) {
  CHECK_H(h);
  if (!data || rows == 0) return fail(PRG_EINVAL, "empty item matrix");
  if (dim != 64 && dim != 128) return fail(PRG_EUNSUPPORTED, "item matrix dim must be 64 or 128");
  if ((reinterpret_cast<uintptr_t>(data) & 15) && mem == PRG_MEM_DEVICE)
    return fail(PRG_EINVAL, "device item matrix must be 16-byte aligned");
  Guard g(h);
  PRG_CUDA(cudaStreamSynchronize(h->stream));
  const void* d = h->E;
  PRG_TRY(take_table(data, (size_t)rows * dim * 4, mem, &d, &h->E_owned));
  h->E = (const float*)d;
  h->E_rows = rows;
  h->E_dim = dim;
  h->E_row_base = row_base;
  h->row_norm.release();  // norms and the bf16 filter index are rebuilt lazily for the new matrix
  h->E16_map_ok = false;
  h->E8_map_ok = false;
  h->i8_backoff = 0;
  return recall_build_map(h);
}

// Snapshot swap (the reference's VectorHologresDao switches to a new partition table once a minute
// without stopping the service, module/vector_hologres_dao.go:40-61).  Stage: upload + bf16 filter index + row norms +
// tensor maps of the NEW matrix on a side stream, without the handle's lock — requests keep running on the live
// snapshot (both are resident meanwhile: size for 2 x (6 B per element + 4 B per row)).  Commit: between two batches,
// under the lock, the two snapshots change places and the old one is freed; the first request after the commit pays
// nothing extra (prg_set_item_matrix rebuilds the index lazily inside the first recall instead).
int prg_stage_item_matrix(prg_handle* h, const float* data, uint64_t rows, uint32_t dim, uint64_t row_base, int mem) {
  CHECK_H(h);
  if (!data || rows == 0) return fail(PRG_EINVAL, "empty item matrix");
  if (dim != 64 && dim != 128) return fail(PRG_EUNSUPPORTED, "item matrix dim must be 64 or 128");
  if ((reinterpret_cast<uintptr_t>(data) & 15) && mem == PRG_MEM_DEVICE)
    return fail(PRG_EINVAL, "device item matrix must be 16-byte aligned");
  int prev = -1;
  cudaGetDevice(&prev);
  if (prev != h->device) cudaSetDevice(h->device);
  struct Restore { int prev, dev; ~Restore() { if (prev >= 0 && prev != dev) cudaSetDevice(prev); } } restore{prev, h->device};
  prg_handle* s = new (std::nothrow) prg_handle();
  if (!s) return fail(PRG_ENOMEM, "snapshot allocation failed");
  s->device = h->device; s->sm_count = h->sm_count;
  s->scan_filter = h->scan_filter; s->scan_ffma2 = h->scan_ffma2; s->scan_int8 = h->scan_int8; s->pdl = false;
  cudaError_t e = cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking);
  if (e != cudaSuccess) { delete s; return fail(PRG_ECUDA, std::string("cudaStreamCreate: ") + cudaGetErrorString(e)); }
  const void* d = nullptr;
  int rc = take_table(data, (size_t)rows * dim * 4, mem, &d, &s->E_owned);
  if (rc == PRG_OK) {
    s->E = (const float*)d; s->E_rows = rows; s->E_dim = dim; s->E_row_base = row_base;
    rc = recall_build_map(s);
  }
  if (rc == PRG_OK && !s->scan_ffma2) rc = build_row_norms(s);   // norms + bf16 index + its map; waits for s->stream
  if (rc != PRG_OK) { free_snapshot(s); return rc; }
  prg_handle* old = nullptr;
  {
    std::lock_guard<std::mutex> lk(h->mu);
    old = h->staged;         // a snapshot staged earlier and never committed is dropped
    h->staged = s;
  }
  free_snapshot(old);
  return PRG_OK;
}

int prg_commit_item_matrix(prg_handle* h) {
  CHECK_H(h);
  Guard g(h);                // also settles a deferred recall check — against the snapshot it ran on
  if (!h->staged) return fail(PRG_ESTATE, "no staged item matrix (prg_stage_item_matrix)");
  if (h->E && h->staged->E_dim != h->E_dim)
    return fail(PRG_EINVAL, "staged item matrix has dim " + std::to_string(h->staged->E_dim) + ", the live one " +
                                std::to_string(h->E_dim) + ": a snapshot swap keeps the dim (use prg_set_item_matrix)");
  PRG_CUDA(cudaStreamSynchronize(h->stream));
  prg_handle* s = h->staged;
  h->staged = nullptr;
  std::swap(h->E, s->E); std::swap(h->E_owned, s->E_owned); std::swap(h->E_rows, s->E_rows);
  std::swap(h->E_dim, s->E_dim); std::swap(h->E_row_base, s->E_row_base);
  std::swap(h->E_map, s->E_map); std::swap(h->E_map_ok, s->E_map_ok);
  std::swap(h->E16, s->E16); std::swap(h->E16_map, s->E16_map); std::swap(h->E16_map_h, s->E16_map_h); std::swap(h->E16_map_ok, s->E16_map_ok);
  std::swap(h->row_norm, s->row_norm);
  std::swap(h->E8, s->E8); std::swap(h->E8_prm, s->E8_prm); std::swap(h->E8_map, s->E8_map); std::swap(h->E8_map_ok, s->E8_map_ok);
  h->i8_backoff = 0;
  free_snapshot(s);          // the previous snapshot
  return PRG_OK;
}

// ------------------------------------------------------------------ recall
int prg_recall_topk(prg_handle* h, const float* q, int B, int k, uint32_t* out_row, float* out_score, int32_t* out_n,
                    int mem) {
  CHECK_H(h);
  if (!q || !out_row || !out_score || !out_n) return fail(PRG_EINVAL, "null buffer");
  if (B <= 0 || k <= 0) return fail(PRG_EINVAL, "B and k must be positive");
  Guard g(h);
  if (!h->E) return fail(PRG_ESTATE, "item matrix not set (prg_set_item_matrix)");
  const size_t nk = (size_t)B * k;
  PRG_TRY(h->topk_keys.ensure(nk * 8));
  if (mem == PRG_MEM_DEVICE) {
    PRG_TRY(recall_topk_device(h, q, B, k, (uint64_t*)h->topk_keys.p));
    return keys_to_outputs(h, (const uint64_t*)h->topk_keys.p, B, k, out_row, out_score, out_n);
  }
  PRG_TRY(h->q_dev.ensure((size_t)B * h->E_dim * 4));
  PRG_TRY(h->out_row.ensure(nk * 4));
  PRG_TRY(h->out_score.ensure(nk * 4));
  PRG_TRY(h->out_n.ensure((size_t)B * 4));
  PRG_CUDA(cudaMemcpyAsync(h->q_dev.p, q, (size_t)B * h->E_dim * 4, cudaMemcpyHostToDevice, h->stream));
  PRG_TRY(recall_topk_device(h, (const float*)h->q_dev.p, B, k, (uint64_t*)h->topk_keys.p));
  PRG_TRY(keys_to_outputs(h, (const uint64_t*)h->topk_keys.p, B, k, (uint32_t*)h->out_row.p, (float*)h->out_score.p,
                          (int32_t*)h->out_n.p));
  PRG_CUDA(cudaMemcpyAsync(out_row, h->out_row.p, nk * 4, cudaMemcpyDeviceToHost, h->stream));
  PRG_CUDA(cudaMemcpyAsync(out_score, h->out_score.p, nk * 4, cudaMemcpyDeviceToHost, h->stream));
  PRG_CUDA(cudaMemcpyAsync(out_n, h->out_n.p, (size_t)B * 4, cudaMemcpyDeviceToHost, h->stream));
  PRG_CUDA(cudaStreamSynchronize(h->stream));
  return PRG_OK;
}

int prg_recall_local_keys(prg_handle* h, const float* q_dev, int B, int k, uint64_t* out_keys_dev) {
  CHECK_H(h);
  if (!q_dev || !out_keys_dev) return fail(PRG_EINVAL, "null buffer");
  Guard g(h);
  return recall_topk_device(h, q_dev, B, k, out_keys_dev);
}

int prg_shard_sample_len(int k) { return k > 0 ? shard_sample_len(k) : 0; }

int prg_shard_sample(prg_handle* h, const float* q_dev, int Bg, int k, int G, uint64_t* out_sample_dev) {
  CHECK_H(h);
  if (!q_dev || !out_sample_dev) return fail(PRG_EINVAL, "null buffer");
  Guard g(h);
  return recall_shard_sample_device(h, q_dev, Bg, k, G, out_sample_dev);
}

int prg_shard_candidates(prg_handle* h, const float* q_dev, int Bg, int k, int G, const uint64_t* all_samples_dev,
                         uint64_t* out_keys_dev) {
  CHECK_H(h);
  if (!q_dev || !all_samples_dev || !out_keys_dev) return fail(PRG_EINVAL, "null buffer");
  Guard g(h);
  return recall_shard_candidates_device(h, q_dev, Bg, k, G, all_samples_dev, out_keys_dev);
}

int prg_shard_check(prg_handle* h, const uint64_t* gathered_dev, int G, int Bg, int k, int32_t* retry_dev) {
  CHECK_H(h);
  if (!gathered_dev || !retry_dev) return fail(PRG_EINVAL, "null buffer");
  if (G <= 0 || Bg <= 0 || k <= 0) return fail(PRG_EINVAL, "G, Bg, k must be positive");
  Guard g(h);
  return shard_check_device(h, gathered_dev, G, Bg, k, retry_dev);
}

int prg_shard_pack_owner(prg_handle* h, const uint64_t* cand_dev, int Bg, int B, int k, uint64_t* out_dev) {
  CHECK_H(h);
  if (!cand_dev || !out_dev || Bg <= 0 || k <= 0) return fail(PRG_EINVAL, "bad arguments");
  Guard g(h);
  return shard_pack_owner_device(h, cand_dev, Bg, B, k, out_dev);
}

int prg_shard_check_owner(prg_handle* h, const uint64_t* received_dev, int G, int B, int k, int q0, int32_t* retry_dev) {
  CHECK_H(h);
  if (!received_dev || !retry_dev || G <= 0 || B <= 0 || k <= 0 || q0 < 0) return fail(PRG_EINVAL, "bad arguments");
  Guard g(h);
  return shard_check_owner_device(h, received_dev, G, B, k, q0, retry_dev);
}

int prg_merge_keys(prg_handle* h, const uint64_t* keys_dev, int G, int B, int k, uint32_t* out_row, float* out_score,
                   int32_t* out_n, int mem) {
  CHECK_H(h);
  if (!keys_dev || !out_row || !out_score || !out_n) return fail(PRG_EINVAL, "null buffer");
  Guard g(h);
  const size_t nk = (size_t)B * k;
  PRG_TRY(h->topk_keys.ensure(nk * 8));
  PRG_TRY(merge_keys_device(h, keys_dev, G, (uint64_t)B * k, B, k, (uint64_t*)h->topk_keys.p));
  if (mem == PRG_MEM_DEVICE) return keys_to_outputs(h, (const uint64_t*)h->topk_keys.p, B, k, out_row, out_score, out_n);
  PRG_TRY(h->out_row.ensure(nk * 4));
  PRG_TRY(h->out_score.ensure(nk * 4));
  PRG_TRY(h->out_n.ensure((size_t)B * 4));
  PRG_TRY(keys_to_outputs(h, (const uint64_t*)h->topk_keys.p, B, k, (uint32_t*)h->out_row.p, (float*)h->out_score.p,
                          (int32_t*)h->out_n.p));
  PRG_CUDA(cudaMemcpyAsync(out_row, h->out_row.p, nk * 4, cudaMemcpyDeviceToHost, h->stream));
  PRG_CUDA(cudaMemcpyAsync(out_score, h->out_score.p, nk * 4, cudaMemcpyDeviceToHost, h->stream));
  PRG_CUDA(cudaMemcpyAsync(out_n, h->out_n.p, (size_t)B * 4, cudaMemcpyDeviceToHost, h->stream));
  PRG_CUDA(cudaStreamSynchronize(h->stream));
  return PRG_OK;
}

// ------------------------------------------------------------------ host-only entries
int prg_lookup(const double* value, const uint8_t* present, int n, double* out) {
  if (n < 0 || (n > 0 && (!value || !present || !out))) return fail(PRG_EINVAL, "null buffer");
  for (int i = 0; i < n; ++i) out[i] = present[i] ? value[i] : 0.5;  // algorithm/lookup.go:44-49
  return PRG_OK;
}

}  // extern "C"
