/*
 * oracle.c — CPU restatement of the pairec recall -> feature -> rank -> sort/DPP hot path.
 *
 * THIS IS TEST INFRASTRUCTURE.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load it; the product (libpairec_gpu.so) never does and has no CPU path of its own.
 *
 * PARITY STATUS
 *   - DPP / KernelMatrix / DPPWithWindow, score sorts, LOOKUP restate code that IS in the reference
 *     (sort/dpp_sort.go, sort/item_rank_score.go, sort/algo_score_sort.go, algorithm/lookup.go) line by line, but the
 *     reference holds no test, golden vector or fixture for any of them and neither a Go toolchain
 *     nor gonum v0.12.0 / the Go stdlib sources exist in the build container: "parity unpinned".  The gonum and
 *     pdqsort internals below (summation orders, pivot rules) are restated from the published algorithms from
 *     memory and are marked [UNVERIFIED-UPSTREAM] where results could differ in the last bit / in tie order.
 *   - recall / gather+FM / MLP arithmetic is NOT in the reference (remote faiss / EAS / TF-Serving); the reference
 *     pins only the wire contract.  The semantics are DEFINED here (and in DESIGN.md) and the CUDA path must
 *     reproduce them: bit-exact for recall keys and FM logits, 1e-5 relative for MLP scores.
 *   - The one known-answer test of the path the reference does hold, utils/ast/ast_test.go:12-27
 *     (${ctr}+${click}+${price} with 0.1,0.3,0.1 -> 0.5), is checked against orc_rank_score_expr in tests/.
 *
 * Build: see oracle/Makefile (gcc -O2 -mavx2 -mfma -ffp-contract=off, pthreads).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <unistd.h>

/* minimal fork-join over pthreads (libgomp is not in the image) */
typedef void (*orc_job_fn)(void* arg, int tid, int n_threads);
typedef struct { orc_job_fn fn; void* arg; int tid, nt; } orc_job;
static void* orc_job_main(void* p) { orc_job* j = (orc_job*)p; j->fn(j->arg, j->tid, j->nt); return NULL; }
static int orc_hw_threads(void) { long n = sysconf(_SC_NPROCESSORS_ONLN); return n > 0 ? (int)n : 1; }
static void orc_parallel(int nt, orc_job_fn fn, void* arg) {
  if (nt <= 1) { fn(arg, 0, 1); return; }
  pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * (size_t)nt);
  orc_job* jobs = (orc_job*)malloc(sizeof(orc_job) * (size_t)nt);
  for (int t = 0; t < nt; ++t) { jobs[t].fn = fn; jobs[t].arg = arg; jobs[t].tid = t; jobs[t].nt = nt; }
  for (int t = 1; t < nt; ++t) pthread_create(&th[t], NULL, orc_job_main, &jobs[t]);
  orc_job_main(&jobs[0]);
  for (int t = 1; t < nt; ++t) pthread_join(th[t], NULL);
  free(th); free(jobs);
}

#define ORC_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------------------------------ order keys */
/* Same total order as pairec_b200/csrc/common.cuh: key = ord(score)<<32 | (0xFFFFFFFF-row); NaN -> ord 1. */
static inline uint32_t f32_ord(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  if ((u & 0x7FFFFFFFu) > 0x7F800000u) return 1u;
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
static inline float ord_f32(uint32_t o) {
  uint32_t u = (o == 1u) ? 0x7FC00000u : ((o & 0x80000000u) ? (o & 0x7FFFFFFFu) : ~o);
  float f;
  memcpy(&f, &u, 4);
  return f;
}
static inline uint64_t make_key(float s, uint32_t row) { return ((uint64_t)f32_ord(s) << 32) | (uint64_t)(0xFFFFFFFFu - row); }

ORC_API uint64_t orc_make_key(float s, uint32_t row) { return make_key(s, row); }
ORC_API float orc_key_score(uint64_t key) { return ord_f32((uint32_t)(key >> 32)); }
ORC_API uint32_t orc_key_row(uint64_t key) { return 0xFFFFFFFFu - (uint32_t)key; }

static int cmp_key_desc(const void* a, const void* b) {
  uint64_t x = *(const uint64_t*)a, y = *(const uint64_t*)b;
  return (x < y) - (x > y);
}

/* ------------------------------------------------------------------------------------------------ recall */
/*
 * Exact inner-product top-k (the arithmetic the remote faiss VectorRetrieval.Search server performs for
 * service/recall/vector_recall.go:88; metric/tie rule unspecified upstream, defined here):
 *   score(row,q) = fmaf(E[row][d-1],Q[q][d-1], ... fmaf(E[row][0],Q[q][0], +0)) — one accumulator, dims ascending.
 *   order: score descending, ties by ascending global row; NaN below -inf.
 * out_keys: B x k, descending, 0-padded.  Threads: pthreads over row chunks with per-thread bounded buffers.
 */
typedef struct { uint64_t* v; int n, cap, k; uint64_t thr; } keybuf;

static void kb_init(keybuf* b, int k) {
  b->k = k; b->cap = 4 * k + 64; b->n = 0; b->thr = 0;
  b->v = (uint64_t*)malloc(sizeof(uint64_t) * (size_t)b->cap);
}
static void kb_shrink(keybuf* b) {
  qsort(b->v, (size_t)b->n, sizeof(uint64_t), cmp_key_desc);
  if (b->n > b->k) { b->n = b->k; b->thr = b->v[b->k - 1]; }
}
static inline void kb_push(keybuf* b, uint64_t key) {
  if (key < b->thr) return;
  b->v[b->n++] = key;
  if (b->n == b->cap) kb_shrink(b);
}

typedef struct { const float* E; uint64_t rows; uint32_t dim; uint64_t row_base; const float* Qt; int nq; keybuf* bufs; } recall_job;
static void recall_worker(void* arg, int tid, int nt) {
  const recall_job* J = (const recall_job*)arg;
  keybuf* mine = J->bufs + (size_t)tid * 64;
  const uint64_t r0 = J->rows * (uint64_t)tid / (uint64_t)nt, r1 = J->rows * (uint64_t)(tid + 1) / (uint64_t)nt;
  const uint32_t dim = J->dim;
  for (uint64_t r = r0; r < r1; ++r) {
    const float* x = J->E + (size_t)r * dim;
    float acc[64] __attribute__((aligned(64)));
    for (int q = 0; q < 64; ++q) acc[q] = 0.0f;
    for (uint32_t d = 0; d < dim; ++d) {
      const float xv = x[d];
      const float* qt = J->Qt + (size_t)d * 64;
      for (int q = 0; q < 64; ++q) acc[q] = __builtin_fmaf(xv, qt[q], acc[q]);
    }
    const uint32_t grow = (uint32_t)(J->row_base + r);
    for (int q = 0; q < J->nq; ++q) kb_push(&mine[q], make_key(acc[q], grow));
  }
}

ORC_API int orc_recall_topk(const float* E, uint64_t rows, uint32_t dim, uint64_t row_base, const float* Q, int B, int k,
                            uint64_t* out_keys, int n_threads) {
  if (!E || !Q || !out_keys || B <= 0 || k <= 0 || dim == 0) return 1;
  if (n_threads <= 0) n_threads = orc_hw_threads();
  for (int q0 = 0; q0 < B; q0 += 64) {
    const int nq = (B - q0 < 64) ? (B - q0) : 64;
    /* Qt[dd][q] so that the q loop vectorises; every lane is still one exact fmaf per dim */
    float* Qt = (float*)aligned_alloc(64, sizeof(float) * (size_t)dim * 64);
    memset(Qt, 0, sizeof(float) * (size_t)dim * 64);
    for (int q = 0; q < nq; ++q)
      for (uint32_t d = 0; d < dim; ++d) Qt[(size_t)d * 64 + q] = Q[(size_t)(q0 + q) * dim + d];
    keybuf* bufs = (keybuf*)malloc(sizeof(keybuf) * (size_t)n_threads * 64);
    for (int i = 0; i < n_threads * 64; ++i) kb_init(&bufs[i], k);
    recall_job job = {E, rows, dim, row_base, Qt, nq, bufs};
    orc_parallel(n_threads, recall_worker, &job);
    for (int q = 0; q < nq; ++q) {
      keybuf all;
      kb_init(&all, k);
      for (int t = 0; t < n_threads; ++t) {
        keybuf* b = &bufs[(size_t)t * 64 + q];
        for (int i = 0; i < b->n; ++i) kb_push(&all, b->v[i]);
      }
      kb_shrink(&all);
      uint64_t* o = out_keys + (size_t)(q0 + q) * k;
      for (int i = 0; i < k; ++i) o[i] = (i < all.n) ? all.v[i] : 0ull;
      free(all.v);
    }
    for (int i = 0; i < n_threads * 64; ++i) free(bufs[i].v);
    free(bufs);
    free(Qt);
  }
  return 0;
}

/* scores only (for property checks): out[q*rows + r] */
ORC_API int orc_recall_scores(const float* E, uint64_t rows, uint32_t dim, const float* Q, int B, float* out) {
  for (int q = 0; q < B; ++q)
    for (uint64_t r = 0; r < rows; ++r) {
      float acc = 0.0f;
      for (uint32_t d = 0; d < dim; ++d) acc = __builtin_fmaf(E[r * dim + d], Q[(size_t)q * dim + d], acc);
      out[(size_t)q * rows + r] = acc;
    }
  return 0;
}

/* merge of G sorted shard lists: keys [G][B][k] -> [B][k] */
ORC_API int orc_merge_keys(const uint64_t* keys, int G, int B, int k, uint64_t* out) {
  uint64_t* tmp = (uint64_t*)malloc(sizeof(uint64_t) * (size_t)G * k);
  for (int b = 0; b < B; ++b) {
    for (int g = 0; g < G; ++g) memcpy(tmp + (size_t)g * k, keys + ((size_t)g * B + b) * k, sizeof(uint64_t) * (size_t)k);
    qsort(tmp, (size_t)G * k, sizeof(uint64_t), cmp_key_desc);
    memcpy(out + (size_t)b * k, tmp, sizeof(uint64_t) * (size_t)k);
  }
  free(tmp);
  return 0;
}

/* ------------------------------------------------------------------------------------------------ gather + FM */
/*
 * Item-side feature gather + FM second-order forward, the work module/feature_*_dao.go (FeatureFetch) and the remote
 * ALINK_FM processor (algorithm/eas/fm_request.go:29-79, fm_response.go:28-34) do between them.  Defined semantics
 * (f32, fixed order, every step one IEEE operation):
 *   id_f = fields[row][f];  id_f >= table_rows[f]  -> the field contributes zeros
 *   lin   = w0;            lin   = lin + w_f[id_f]                       f = 0..F-1
 *   s_k   = 0;             s_k   = s_k + v_f[id_f][k]                    f = 0..F-1
 *   ss_k  = 0;             ss_k  = fmaf(v, v, ss_k)
 *   inter = 0;             inter = inter + fmaf(s_k, s_k, -ss_k)         k = 0..fdim-1
 *   logit = fmaf(0.5f, inter, lin)
 * x_out (optional): the F*fdim concatenated factors (MLP input), row-major [n][F*fdim].
 * rows == 0xFFFFFFFF is padding: logit 0, x zeros.
 */
typedef struct { const uint32_t* fields; uint64_t field_rows; uint32_t F; const float* const* factors;
                 const float* const* linear; const uint64_t* table_rows; uint32_t fdim; float w0; const uint32_t* rows;
                 int n; float* logit_out; float* x_out;
                 uint32_t U; const uint32_t* user_ids; uint32_t n_dense; const float* user_dense; } gfm_job;
static void gfm_worker(void* arg, int tid, int nt) {
  const gfm_job* J = (const gfm_job*)arg;
  const uint32_t F = J->F, fdim = J->fdim, U = J->U;
  const size_t xw = (size_t)(F + U) * fdim + J->n_dense;
  const int i0 = (int)((int64_t)J->n * tid / nt), i1 = (int)((int64_t)J->n * (tid + 1) / nt);
  /* user prefix (service/rank/algo_data.go:110-112: the user's features enter the map first): tables F..F+U-1 */
  float lin0 = J->w0, s0[64], ss0[64];
  for (uint32_t k = 0; k < fdim; ++k) { s0[k] = 0.0f; ss0[k] = 0.0f; }
  for (uint32_t u = 0; u < U; ++u) {
    const uint32_t id = J->user_ids ? J->user_ids[u] : 0xFFFFFFFFu;
    const int ok = id < J->table_rows[F + u];
    const float w = (ok && J->linear[F + u]) ? J->linear[F + u][id] : 0.0f;
    lin0 = lin0 + w;
    for (uint32_t k = 0; k < fdim; ++k) {
      const float v = ok ? J->factors[F + u][(size_t)id * fdim + k] : 0.0f;
      s0[k] = s0[k] + v;
      ss0[k] = __builtin_fmaf(v, v, ss0[k]);
    }
  }
  for (int i = i0; i < i1; ++i) {
    const uint32_t row = J->rows[i];
    float* x = J->x_out ? J->x_out + (size_t)i * xw : NULL;
    if (row == 0xFFFFFFFFu || row >= J->field_rows) {
      if (J->logit_out) J->logit_out[i] = 0.0f;
      if (x) memset(x, 0, sizeof(float) * xw);
      continue;
    }
    float lin = lin0, s[64], ss[64];
    for (uint32_t k = 0; k < fdim; ++k) { s[k] = s0[k]; ss[k] = ss0[k]; }
    for (uint32_t f = 0; f < F; ++f) {
      const uint32_t id = J->fields[(size_t)row * F + f];
      const int ok = id < J->table_rows[f];
      const float w = (ok && J->linear[f]) ? J->linear[f][id] : 0.0f;
      lin = lin + w;
      for (uint32_t k = 0; k < fdim; ++k) {
        const float v = ok ? J->factors[f][(size_t)id * fdim + k] : 0.0f;
        s[k] = s[k] + v;
        ss[k] = __builtin_fmaf(v, v, ss[k]);
        if (x) x[f * fdim + k] = v;
      }
    }
    if (x) {
      for (uint32_t u = 0; u < U; ++u) {
        const uint32_t id = J->user_ids ? J->user_ids[u] : 0xFFFFFFFFu;
        const int ok = id < J->table_rows[F + u];
        for (uint32_t k = 0; k < fdim; ++k)
          x[(F + u) * fdim + k] = ok ? J->factors[F + u][(size_t)id * fdim + k] : 0.0f;
      }
      for (uint32_t c = 0; c < J->n_dense; ++c) x[(F + U) * fdim + c] = J->user_dense ? J->user_dense[c] : 0.0f;
    }
    float inter = 0.0f;
    for (uint32_t k = 0; k < fdim; ++k) inter = inter + __builtin_fmaf(s[k], s[k], -ss[k]);
    if (J->logit_out) J->logit_out[i] = __builtin_fmaf(0.5f, inter, lin);
  }
}
ORC_API int orc_gather_fm(const uint32_t* fields, uint64_t field_rows, uint32_t F, const float* const* factors,
                          const float* const* linear, const uint64_t* table_rows, uint32_t fdim, float w0,
                          const uint32_t* rows, int n, float* logit_out, float* x_out) {
  if (fdim > 64) return 1;
  gfm_job job = {fields, field_rows, F, factors, linear, table_rows, fdim, w0, rows, n, logit_out, x_out, 0, NULL, 0, NULL};
  orc_parallel(n > 1024 ? orc_hw_threads() : 1, gfm_worker, &job);
  return 0;
}
/* One request's candidates with that request's user / context features merged in.  The reference builds
 *   features = userFeatures, then itemFeatures on top   (service/rank/algo_data.go:104-118)
 * from user.MakeUserFeatures() (service/rank/rank_service.go:175-183) and item.GetFeatures() (module/item.go:229-248).
 * Id-encoded here: U categorical user fields (tables F..F+U-1 of factors / linear / table_rows, ids user_ids[U],
 * 0xFFFFFFFF = the user has no such feature) and n_dense numeric context values that only the tower sees.
 * FM sums run over the user fields first, then the item fields, each in index order.
 * x_out: [n][(F+U)*fdim + n_dense] = item factors | user factors | dense values. */
ORC_API int orc_gather_fm_user(const uint32_t* fields, uint64_t field_rows, uint32_t F, const float* const* factors,
                               const float* const* linear, const uint64_t* table_rows, uint32_t fdim, float w0,
                               const uint32_t* rows, int n, uint32_t U, const uint32_t* user_ids, uint32_t n_dense,
                               const float* user_dense, float* logit_out, float* x_out) {
  if (fdim > 64) return 1;
  gfm_job job = {fields, field_rows, F, factors, linear, table_rows, fdim, w0, rows, n, logit_out, x_out,
                 U, user_ids, n_dense, user_dense};
  orc_parallel(n > 1024 ? orc_hw_threads() : 1, gfm_worker, &job);
  return 0;
}

/* score = (float)(1 / (1 + exp(-(double)logit))): the f32 AlgoResponse score, evaluated through fp64 so that the
 * CPU and GPU libm differences (<= 1 ulp of fp64) vanish in the f32 rounding. */
/* libm exp as this oracle calls it (Go's math.Exp has an assembly kernel of its own on amd64 / arm64 / s390x and can differ
 * from it in the last bit: baseline/go injects these values where it pins gonum's summation order) */
ORC_API double orc_exp(double x) { return exp(x); }
ORC_API float orc_sigmoid(float logit) { return (float)(1.0 / (1.0 + exp(-(double)logit))); }

/* ------------------------------------------------------------------------------------------------ MLP */
static inline uint16_t f32_to_bf16(float f) { /* round to nearest even; NaN kept quiet */
  uint32_t u;
  memcpy(&u, &f, 4);
  if ((u & 0x7FFFFFFFu) > 0x7F800000u) return (uint16_t)((u >> 16) | 0x40);
  u += 0x7FFFu + ((u >> 16) & 1u);
  return (uint16_t)(u >> 16);
}
static inline float bf16_to_f32(uint16_t h) {
  uint32_t u = (uint32_t)h << 16;
  float f;
  memcpy(&f, &u, 4);
  return f;
}
ORC_API uint16_t orc_f32_to_bf16(float f) { return f32_to_bf16(f); }
ORC_API float orc_bf16_to_f32(uint16_t h) { return bf16_to_f32(h); }

/*
 * Dense tower of the replaced remote DNN (algorithm/eas easyrec / algorithm/tfserving; wire contract
 * easyrec_response.go:35-70, tfserving/response.go:51-63).  Defined semantics ("bf16x2" activations):
 *   an activation a (f32) is carried as hi = bf16(a), lo = bf16(a - hi): the tensor cores see two bf16 operands,
 *   the value carried is hi + lo (>= 16 significant bits), which keeps the result continuous under the 1e-6-level
 *   accumulation-order differences between implementations (plain bf16 activations flip whole bf16 ulps).
 *   z_j   = b_j + sum_i W[j][i] * (hi_i + lo_i)      W bf16, accumulation wide (oracle: fp64), z rounded to f32
 *   hidden: a = max(z, 0) -> split again;   last layer (width O >= 1, one column per model output:
 *   easyrec_response.go:35-70 / tfserving/response.go:51-63): logit_o = z_o (f32)
 *   The tower INPUT is carried as bf16(x) alone (lo = 0): gathered table values carry no accumulation-order noise, so
 *   the single rounding is reproducible bit for bit; only COMPUTED activations need the second term.
 * x: [n][dims[0]] f32.  W[l]: [dims[l+1]][dims[l]] bf16 bits.  logit_out: [n][O].
 */
typedef struct { const float* x; int n, n_layers; const uint32_t* dims; double* const* Wt; const float* const* bias; float* logit_out; uint32_t maxd; } mlp_job;
static void mlp_worker(void* arg, int tid, int nt) {
  const mlp_job* J = (const mlp_job*)arg;
  const uint32_t* dims = J->dims;
  double* a = (double*)malloc(sizeof(double) * J->maxd);
  double* acc = (double*)malloc(sizeof(double) * J->maxd);
  const int i0 = (int)((int64_t)J->n * tid / nt), i1 = (int)((int64_t)J->n * (tid + 1) / nt);
  for (int i = i0; i < i1; ++i) {
    for (uint32_t c = 0; c < dims[0]; ++c) {
      const float v = J->x[(size_t)i * dims[0] + c];
      a[c] = (double)bf16_to_f32(f32_to_bf16(v));
    }
    for (int l = 0; l < J->n_layers; ++l) {
      const uint32_t K = dims[l], N = dims[l + 1];
      /* acc_j = sum_c W[j][c] * a[c], c ascending, one accumulator per j (Wt is [c][j] so the j loop vectorises;
       * the per-j summation order is unchanged) */
      for (uint32_t j = 0; j < N; ++j) acc[j] = 0.0;
      for (uint32_t c = 0; c < K; ++c) {
        const double ac = a[c];
        const double* w = J->Wt[l] + (size_t)c * N;
        for (uint32_t j = 0; j < N; ++j) acc[j] += w[j] * ac;
      }
      if (l == J->n_layers - 1) {
        for (uint32_t j = 0; j < N; ++j)
          J->logit_out[(size_t)i * N + j] = (float)(acc[j] + (double)(J->bias[l] ? J->bias[l][j] : 0.0f));
      } else {
        for (uint32_t j = 0; j < N; ++j) {
          const float z = (float)(acc[j] + (double)(J->bias[l] ? J->bias[l][j] : 0.0f));
          const float r = z > 0.0f ? z : 0.0f;
          const float hi = bf16_to_f32(f32_to_bf16(r));
          const float lo = bf16_to_f32(f32_to_bf16(r - hi));
          a[j] = (double)hi + (double)lo;
        }
      }
    }
  }
  free(a);
  free(acc);
}

ORC_API int orc_mlp_forward(const float* x, int n, int n_layers, const uint32_t* dims, const uint16_t* const* W,
                            const float* const* bias, float* logit_out) {
  if (n_layers < 1 || n_layers > 64) return 1;
  uint32_t maxd = 0;
  for (int l = 0; l <= n_layers; ++l)
    if (dims[l] > maxd) maxd = dims[l];
  if (dims[n_layers] < 1) return 1;
  double** Wt = (double**)malloc(sizeof(double*) * (size_t)n_layers);
  for (int l = 0; l < n_layers; ++l) {
    const uint32_t K = dims[l], N = dims[l + 1];
    Wt[l] = (double*)malloc(sizeof(double) * (size_t)K * N);
    for (uint32_t j = 0; j < N; ++j)
      for (uint32_t c = 0; c < K; ++c) Wt[l][(size_t)c * N + j] = (double)bf16_to_f32(W[l][(size_t)j * K + c]);
  }
  mlp_job job = {x, n, n_layers, dims, Wt, bias, logit_out, maxd};
  orc_parallel(n > 256 ? orc_hw_threads() : 1, mlp_worker, &job);
  for (int l = 0; l < n_layers; ++l) free(Wt[l]);
  free(Wt);
  return 0;
}

/* ------------------------------------------------------------------------------------------------ LOOKUP */
/* algorithm/lookup.go:37-51: score = features[FieldName].(float64) when the key is present, else 0.5. */
ORC_API void orc_lookup(const double* value, const uint8_t* present, int n, double* out) {
  for (int i = 0; i < n; ++i) out[i] = present[i] ? value[i] : 0.5;
}

/* utils/ast: the simple "${a} + ${b} * w" RankScore expressions (service/rank/rank_service.go:339-363) reduce, for
 * the sums-of-products the GPU path accepts, to left-to-right fp64 evaluation.  terms[i] = coef[i] * value[i]. */
ORC_API double orc_rank_score_expr(const double* value, const double* coef, int n) {
  double acc = 0.0;
  for (int i = 0; i < n; ++i) acc = (i == 0) ? coef[i] * value[i] : acc + coef[i] * value[i];
  return acc;
}

/* ------------------------------------------------------------------------------------------------ Go sort (pdqsort) */
/*
 * sort.Sort / sort.Slice of Go >= 1.19 (pattern-defeating quicksort, src/sort/zsortinterface.go) driving a
 * permutation, so that the tie order of sort/item_rank_score.go:29, sort/algo_score_sort.go:51 and
 * sort/dpp_sort.go:119,281 can be reproduced.  [UNVERIFIED-UPSTREAM]: restated from the published algorithm; the
 * stdlib source is not in the container.  less(i,j) is evaluated on the CURRENT positions i,j like sort.Interface.
 */
typedef struct { const double* score; int32_t* perm; int descending; } sortctx;

static inline int s_less(const sortctx* c, int i, int j) {
  const double a = c->score[c->perm[i]], b = c->score[c->perm[j]];
  return c->descending ? (b < a) : (a < b); /* sort.Reverse: Less(i,j) = inner.Less(j,i) */
}
static inline void s_swap(sortctx* c, int i, int j) { int32_t t = c->perm[i]; c->perm[i] = c->perm[j]; c->perm[j] = t; }

static void s_insertion(sortctx* c, int a, int b) {
  for (int i = a + 1; i < b; ++i)
    for (int j = i; j > a && s_less(c, j, j - 1); --j) s_swap(c, j, j - 1);
}
static void s_sift_down(sortctx* c, int lo, int hi, int first) {
  int root = lo;
  for (;;) {
    int child = 2 * root + 1;
    if (child >= hi) return;
    if (child + 1 < hi && s_less(c, first + child, first + child + 1)) child++;
    if (!s_less(c, first + root, first + child)) return;
    s_swap(c, first + root, first + child);
    root = child;
  }
}
static void s_heapsort(sortctx* c, int a, int b) {
  const int first = a, lo = 0, hi = b - a;
  for (int i = (hi - 1) / 2; i >= 0; --i) s_sift_down(c, i, hi, first);
  for (int i = hi - 1; i >= 0; --i) { s_swap(c, first, first + i); s_sift_down(c, lo, i, first); }
}
static int bits_len(unsigned x) { int n = 0; while (x) { ++n; x >>= 1; } return n; }
static void s_break_patterns(sortctx* c, int a, int b) {
  const int length = b - a;
  if (length >= 8) {
    uint64_t r = (uint64_t)length;
    const unsigned modulus = 1u << bits_len((unsigned)length);
    const int idx = a + (length / 4) * 2 - 1;
    for (int i = 0; i < 3; ++i) {
      r ^= r << 13; r ^= r >> 7; r ^= r << 17;
      int other = (int)((unsigned)r & (modulus - 1));
      if (other >= length) other -= length;
      s_swap(c, idx - 1 + i, a + other);
    }
  }
}
static void s_order2(sortctx* c, int* a, int* b, int* swaps) {
  if (s_less(c, *b, *a)) { int t = *a; *a = *b; *b = t; (*swaps)++; }
}
static int s_median(sortctx* c, int a, int b, int cc, int* swaps) {
  s_order2(c, &a, &b, swaps);
  s_order2(c, &b, &cc, swaps);
  s_order2(c, &a, &b, swaps);
  return b;
}
enum { HINT_UNKNOWN = 0, HINT_INCREASING = 1, HINT_DECREASING = 2 };
static int s_choose_pivot(sortctx* c, int a, int b, int* hint) {
  const int l = b - a;
  int swaps = 0, i = a + l / 4 * 1, j = a + l / 4 * 2, k = a + l / 4 * 3;
  if (l >= 8) {
    if (l >= 50) {
      i = s_median(c, i - 1, i, i + 1, &swaps);
      j = s_median(c, j - 1, j, j + 1, &swaps);
      k = s_median(c, k - 1, k, k + 1, &swaps);
    }
    j = s_median(c, i, j, k, &swaps);
  }
  *hint = (swaps == 0) ? HINT_INCREASING : (swaps == 12) ? HINT_DECREASING : HINT_UNKNOWN;
  return j;
}
static void s_reverse(sortctx* c, int a, int b) {
  int i = a, j = b - 1;
  while (i < j) { s_swap(c, i, j); ++i; --j; }
}
static int s_partial_insertion(sortctx* c, int a, int b) {
  int i = a + 1;
  for (int step = 0; step < 5; ++step) {
    while (i < b && !s_less(c, i, i - 1)) ++i;
    if (i == b) return 1;
    if (b - a < 50) return 0;
    s_swap(c, i, i - 1);
    if (i - a >= 2)
      for (int j = i - 1; j >= 1; --j) { if (!s_less(c, j, j - 1)) break; s_swap(c, j, j - 1); }
    if (b - i >= 2)
      for (int j = i + 1; j < b; ++j) { if (!s_less(c, j, j - 1)) break; s_swap(c, j, j - 1); }
  }
  return 0;
}
static int s_partition_equal(sortctx* c, int a, int b, int pivot) {
  s_swap(c, a, pivot);
  int i = a + 1, j = b - 1;
  for (;;) {
    while (i <= j && !s_less(c, a, i)) ++i;
    while (i <= j && s_less(c, a, j)) --j;
    if (i > j) break;
    s_swap(c, i, j); ++i; --j;
  }
  return i;
}
static int s_partition(sortctx* c, int a, int b, int pivot, int* already) {
  s_swap(c, a, pivot);
  int i = a + 1, j = b - 1;
  while (i <= j && s_less(c, i, a)) ++i;
  while (i <= j && !s_less(c, j, a)) --j;
  if (i > j) { s_swap(c, j, a); *already = 1; return j; }
  s_swap(c, i, j); ++i; --j;
  for (;;) {
    while (i <= j && s_less(c, i, a)) ++i;
    while (i <= j && !s_less(c, j, a)) --j;
    if (i > j) break;
    s_swap(c, i, j); ++i; --j;
  }
  s_swap(c, j, a);
  *already = 0;
  return j;
}
static void s_pdqsort(sortctx* c, int a, int b, int limit) {
  int was_balanced = 1, was_partitioned = 1;
  for (;;) {
    const int length = b - a;
    if (length <= 12) { s_insertion(c, a, b); return; }
    if (limit == 0) { s_heapsort(c, a, b); return; }
    if (!was_balanced) { s_break_patterns(c, a, b); --limit; }
    int hint;
    int pivot = s_choose_pivot(c, a, b, &hint);
    if (hint == HINT_DECREASING) {
      s_reverse(c, a, b);
      pivot = (b - 1) - (pivot - a);
      hint = HINT_INCREASING;
    }
    if (was_balanced && was_partitioned && hint == HINT_INCREASING) {
      if (s_partial_insertion(c, a, b)) return;
    }
    if (a > 0 && !s_less(c, a - 1, pivot)) { a = s_partition_equal(c, a, b, pivot); continue; }
    int already;
    const int mid = s_partition(c, a, b, pivot, &already);
    was_partitioned = already;
    const int left = mid - a, right = b - mid, thr = length / 8;
    if (left < right) {
      was_balanced = left >= thr;
      s_pdqsort(c, a, mid, limit);
      a = mid + 1;
    } else {
      was_balanced = right >= thr;
      s_pdqsort(c, mid + 1, b, limit);
      b = mid;
    }
  }
}
/* perm_out[i] = input index placed at position i.  descending != 0: sort.Sort(sort.Reverse(ItemScoreSlice)). */
ORC_API void orc_go_sort(const double* score, int n, int descending, int32_t* perm_out) {
  for (int i = 0; i < n; ++i) perm_out[i] = i;
  if (n <= 1) return;
  sortctx c = {score, perm_out, descending};
  s_pdqsort(&c, 0, n, bits_len((unsigned)n));
}
/* The stable total order the GPU sort defines: score descending, then input index ascending. */
ORC_API void orc_stable_sort_desc(const double* score, int n, int32_t* perm_out) {
  for (int i = 0; i < n; ++i) perm_out[i] = i;
  for (int i = 1; i < n; ++i) { /* insertion sort: n is a few thousand at most */
    const int32_t p = perm_out[i];
    int j = i - 1;
    while (j >= 0 && score[perm_out[j]] < score[p]) { perm_out[j + 1] = perm_out[j]; --j; }
    perm_out[j + 1] = p;
  }
}
/* sort/algo_score_sort.go:28-66: max(Score) > SwitchThreshold -> sort by current score, else by the field;
 * sort.Slice(less = iScore > jScore).  (The :59 bug only triggers on a missing field, not modelled.) */
ORC_API void orc_algo_score_sort(const double* score, const double* field, int n, double switch_threshold,
                                 int32_t* perm_out) {
  double mx = -1e300;
  for (int i = 0; i < n; ++i)
    if (score[i] > mx) mx = score[i];
  orc_go_sort(mx > switch_threshold ? score : field, n, 1, perm_out);
}

/* ------------------------------------------------------------------------------------------------ DPP */
/* gonum floats.Norm(v, 2) -> f64.L2NormUnitary, scaled form (internal/asm/f64 l2norm noasm variant).
 * [UNVERIFIED-UPSTREAM]: the amd64 assembly variant may sum in a different order (last-bit differences). */
static double g_norm2(const double* x, int n) {
  double scale = 0.0, sumsq = 1.0;
  for (int i = 0; i < n; ++i) {
    const double v = x[i];
    if (v == 0) continue;
    const double a = fabs(v);
    if (isnan(a)) return NAN;
    if (scale < a) { const double s = scale / a; sumsq = 1 + sumsq * s * s; scale = a; }
    else { const double s = a / scale; sumsq += s * s; }
  }
  if (isinf(scale)) return INFINITY;
  return scale * sqrt(sumsq);
}
/* gonum f64.DotUnitary (amd64 SSE2): four partial sums by index mod 4, tail into lane 0, (s0+s2)+(s1+s3); separate
 * multiply and add roundings.  [UNVERIFIED-UPSTREAM] */
static double g_dot_unitary(const double* x, const double* y, int n) {
  double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
  int i = 0;
  for (; i + 4 <= n; i += 4) {
    s0 += x[i] * y[i];
    s1 += x[i + 1] * y[i + 1];
    s2 += x[i + 2] * y[i + 2];
    s3 += x[i + 3] * y[i + 3];
  }
  for (; i < n; ++i) s0 += x[i] * y[i];
  return (s0 + s2) + (s1 + s3);
}
/* gonum blas Dgemm(NoTrans, Trans) element of an (items x k) * (items x k)^T product.  dgemmParallel
 * (blas/gonum/dgemm.go) cuts C into 64 x 64 blocks and walks k in blocks of 64 — each k block one DotUnitary, block
 * results added to C in order — but when the product has fewer than minParBlock = 4 blocks of C, i.e. items <= 64, it
 * calls dgemmSerial on the whole problem: ONE DotUnitary over all of k.  The two differ in the last bit once k > 64.
 * [UNVERIFIED-UPSTREAM] */
static int g_gemm_serial(int items) { return ((items + 63) / 64) * ((items + 63) / 64) < 4; }
static double g_gemm_nt_elem(const double* a, const double* b, int k, int serial) {
  if (serial) return 0.0 + g_dot_unitary(a, b, k);
  double c = 0.0;
  for (int k0 = 0; k0 < k; k0 += 64) {
    const int len = (k - k0 < 64) ? (k - k0) : 64;
    c += g_dot_unitary(a + k0, b + k0, len);
  }
  return c;
}
static int g_max_idx(const double* s, int n) { /* floats.MaxIdx: NaN skipped, first maximum wins, 0 if all NaN */
  double mx = NAN;
  int ind = 0;
  for (int i = 0; i < n; ++i) {
    const double v = s[i];
    if (isnan(v)) continue;
    if (v > mx || isnan(mx)) { mx = v; ind = i; }
  }
  return ind;
}
static int index_of(const int32_t* a, int n, int e) {
  for (int i = 0; i < n; ++i)
    if (a[i] == e) return i;
  return -1;
}

typedef struct {
  double alpha;
  int32_t top_n, window_size, norm_mode, normalize_emb, candidate_count;
  double min_score_percent;
} orc_dpp_params;

/* sort/dpp_sort.go:493-551.  L accessed through rows computed on demand (Lrow) — mathematically the dense kernel
 * matrix of :372-475, element for element.  F: n x (D+1) features, r: n quality terms. */
static void dpp_L_row(const double* F, const double* r, int n, int D1, int j, double* out) {
  for (int i = 0; i < n; ++i) {
    const double s = g_gemm_nt_elem(F + (size_t)j * D1, F + (size_t)i * D1, D1, g_gemm_serial(n)); /* S[j][i] */
    out[i] = (r[j] * s) * r[i];                                                   /* diag(r) S diag(r), :466-472 */
  }
}
static int dpp_once(const double* F, const double* r, int n, int D1, int top_n, const int32_t* existed, int n_existed,
                    int32_t* Y) {
  const double eps = 1e-10;
  if (top_n > n) top_n = n;
  double* d2 = (double*)malloc(sizeof(double) * (size_t)n);
  double* C = (double*)calloc((size_t)(top_n > 0 ? top_n : 1) * n, sizeof(double));
  double* Lj = (double*)malloc(sizeof(double) * (size_t)n);
  double* e = (double*)malloc(sizeof(double) * (size_t)n);
  int ny = 0;
  for (int i = 0; i < n; ++i) {
    if (index_of(existed, n_existed, i) < 0) {
      const double s = g_gemm_nt_elem(F + (size_t)i * D1, F + (size_t)i * D1, D1, g_gemm_serial(n));
      d2[i] = (r[i] * s) * r[i];
    } else d2[i] = NAN;
  }
  int j = g_max_idx(d2, n);
  Y[ny++] = j;
  while (ny < top_n) {
    double dj = d2[j];
    if (dj < eps) break;
    dj = sqrt(dj);
    const int k = ny - 1;
    const double inv = 1 / dj;
    dpp_L_row(F, r, n, D1, j, Lj);
    if (k == 0) {
      for (int i = 0; i < n; ++i) e[i] = inv * Lj[i]; /* e.Scale(1/dj, Lj) */
    } else {
      /* ss = cj^T * C[0:k] : Dgemm(Trans,NoTrans), l ascending, AxpyUnitary (mul then add), skipped when tmp == 0 */
      for (int i = 0; i < n; ++i) e[i] = 0.0;
      for (int l = 0; l < k; ++l) {
        const double tmp = C[(size_t)l * n + j];
        if (tmp != 0)
          for (int i = 0; i < n; ++i) e[i] += tmp * C[(size_t)l * n + i];
      }
      for (int i = 0; i < n; ++i) e[i] = inv * (Lj[i] - e[i]); /* e.Sub(Lj, ss); e.Scale(1/dj, e) */
    }
    memcpy(C + (size_t)k * n, e, sizeof(double) * (size_t)n);
    for (int i = 0; i < n; ++i) d2[i] = d2[i] - e[i] * e[i]; /* MulElem then SubVec */
    d2[j] = NAN;
    j = g_max_idx(d2, n);
    Y[ny++] = j;
  }
  if (ny < top_n) {
    for (int i = 0; i < n; ++i) {
      if (index_of(existed, n_existed, i) < 0 && index_of(Y, ny, i) < 0) {
        Y[ny++] = i;
        if (ny == top_n) break;
      }
    }
  }
  free(d2); free(C); free(Lj); free(e);
  return ny;
}

/*
 * One request through DPPSort.doSort (sort/dpp_sort.go:271-351), table path (hasTable, no hooks).
 *   emb: n x D embeddings of the candidates in input order (f64; an f32 table row widened exactly).
 *   score: Item.Score in input order.
 * out_idx: indices into the INPUT list, in output order; returns the count (ctx.Size entries in windowed mode, even
 * when that repeats index 0 — the reference does, :477-491 with :497-499).  *status = 1 when the reference logs an
 * error and returns the items unchanged (out_idx = identity over the truncated list).
 */
/* the fixed pseudo-random direction a candidate WITHOUT an embedding receives (the reference draws an unseeded random
 * unit vector, dpp_sort.go:250-262): splitmix64 of (substitute row, dimension) -> 24-bit dyadic value in [-1, 1).
 * Same function as pairec_b200/csrc/dpp_common.cuh dpp_substitute_value. */
ORC_API float orc_dpp_substitute(uint32_t sub_row, uint32_t d) {
  uint64_t z = (((uint64_t)sub_row << 32) | d) + 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z = z ^ (z >> 31);
  return (float)((int32_t)(z >> 40) - (1 << 23)) * (1.0f / (float)(1 << 23));
}

/*
 * General form.  present: n bytes (NULL = every candidate has a table embedding); a candidate without one takes
 * substitute row (input position & 1023).  hook: n x hook_dim (NULL = no hooks), the concatenated GenerateEmbedding
 * output (:362-370).  use_table = 0: hook-only path (:432-447) with p->normalize_emb and no_positive_sim
 * (EnsurePositiveSim == "false": append 0, no 1/sqrt2).  hooks + table (:416-421): concat(hook, table row as cached),
 * re-normalised as a whole whatever NormalizeEmb says.  The cached table row is normalised at load when
 * p->normalize_emb (:234-237).  L_diag / L_row0 (nullable, n_trunc values each): diag(L) and row 0 of L.
 */
ORC_API int orc_dpp_request_ex(const double* emb, const uint8_t* present, const double* hook, int hook_dim, int use_table,
                               int no_positive_sim, const double* score, int n, int D, const orc_dpp_params* p,
                               int32_t* out_idx, int32_t* status, double* L_diag, double* L_row0) {
  *status = 0;
  if (n == 0) return 0;
  int window = p->window_size > 0 ? p->window_size : 10;
  const int T = p->top_n;
  int32_t* order = (int32_t*)malloc(sizeof(int32_t) * (size_t)n);
  int m = n;
  for (int i = 0; i < n; ++i) order[i] = i;
  /* :280-300 presort + truncation */
  if ((p->candidate_count > 0 || p->min_score_percent > 0) && n > T) {
    orc_go_sort(score, n, 1, order);
    if (p->candidate_count > 0) {
      const int cnt = T > p->candidate_count ? T : p->candidate_count;
      if (cnt < m) m = cnt;
    }
    if (p->min_score_percent > 0 && m > T) {
      int idx = T;
      const double mx = score[order[0]];
      for (; idx < m; ++idx)
        if (score[order[idx]] / mx < p->min_score_percent) break;
      m = idx;
    }
  }
  /* KernelMatrix :372-475 */
  double* rel = (double*)malloc(sizeof(double) * (size_t)m);
  for (int i = 0; i < m; ++i) rel[i] = score[order[i]];
  int err = 0;
  if (p->norm_mode == 1) { /* stat.PopMeanVariance + StdScore */
    /* [UNVERIFIED-UPSTREAM] the mean is stat.Mean -> floats.Sum -> f64.Sum, which on amd64 is an SSE2 kernel with eight
     * partial sums that first peels one element when the slice is not 16-byte aligned: its last bit depends on the
     * address of the slice and cannot be pinned by any fixture.  Restated in the plain sequential (noasm) order; the
     * z-scores that follow can differ from an amd64 run in the last bits (abtest mode dpp_norm_relevance_score = 1 only). */
    double mean = 0;
    for (int i = 0; i < m; ++i) mean += rel[i];
    mean /= (double)m;
    double ss = 0, comp = 0; /* gonum PopMeanVariance -> MeanVariance two-pass with compensation, / n */
    for (int i = 0; i < m; ++i) { const double d = rel[i] - mean; ss += d * d; comp += d; }
    const double var = (ss - comp * comp / (double)m) / (double)m;
    if (mean == 0 || var == 0) err = 1;
    else { const double sd = sqrt(var); for (int i = 0; i < m; ++i) rel[i] = (rel[i] - mean) / sd; }
  } else if (p->norm_mode == 2) {
    const double mx = rel[0], mn = rel[m - 1], span = mx - mn;
    if (span == 0) err = 1;
    else for (int i = 0; i < m; ++i) rel[i] = ((rel[i] - mn) / span) * (1 - 1e-6) + 1e-6;
  }
  int ny = 0;
  if (err) {
    *status = 1;
    for (int i = 0; i < m; ++i) out_idx[i] = order[i];
    ny = m;
  } else {
    const int Dh = hook ? hook_dim : 0, Dt = use_table ? D : 0;
    const int W = Dh + Dt, D1 = W + 1;
    double* F = (double*)malloc(sizeof(double) * (size_t)m * D1);
    double* r = (double*)malloc(sizeof(double) * (size_t)m);
    const double inv_sqrt2 = 0.70710678118654752440; /* 1/math.Sqrt2 as a Go constant expression */
    for (int i = 0; i < m; ++i) {
      double* f = F + (size_t)i * D1;
      double* t = f + Dh; /* the table part */
      if (Dh) memcpy(f, hook + (size_t)order[i] * Dh, sizeof(double) * (size_t)Dh);
      if (Dt) {
        if (!present || present[order[i]]) memcpy(t, emb + (size_t)order[i] * D, sizeof(double) * (size_t)D);
        else for (int d = 0; d < D; ++d) t[d] = (double)orc_dpp_substitute((uint32_t)order[i] & 1023u, (uint32_t)d);
        if (p->normalize_emb) { /* :234-237 floats.Scale(1/normV, vector) at load */
          const double s = 1 / g_norm2(t, D);
          for (int d = 0; d < D; ++d) t[d] *= s;
        }
      }
      int pos = 1;
      if (Dt && Dh) {                 /* :416-421 concat re-normalised */
        const double s = 1 / g_norm2(f, W);
        for (int d = 0; d < W; ++d) f[d] *= s;
      } else if (!Dt) {               /* :432-447 hook only */
        if (p->normalize_emb) {
          const double s = 1 / g_norm2(f, W);
          for (int d = 0; d < W; ++d) f[d] *= s;
        }
        pos = !no_positive_sim;
      }
      if (pos) {
        f[W] = 1;
        for (int d = 0; d < D1; ++d) f[d] *= inv_sqrt2; /* :428-429 / :441-442 */
      } else {
        f[W] = 0;                                       /* :444 */
      }
      r[i] = exp(p->alpha * rel[i]);                  /* :431 */
    }
    if (L_diag || L_row0) {
      double* row = (double*)malloc(sizeof(double) * (size_t)m);
      if (L_row0) { dpp_L_row(F, r, m, D1, 0, row); memcpy(L_row0, row, sizeof(double) * (size_t)m); }
      if (L_diag)
        for (int i = 0; i < m; ++i) {
          const double sii = g_gemm_nt_elem(F + (size_t)i * D1, F + (size_t)i * D1, D1, g_gemm_serial(m));
          L_diag[i] = (r[i] * sii) * r[i];
        }
      free(row);
    }
    int32_t* res = (int32_t*)malloc(sizeof(int32_t) * (size_t)(T > 0 ? T : 1));
    if (T <= window) {
      ny = dpp_once(F, r, m, D1, T, res, 0, res);
    } else {
      int32_t* sub = (int32_t*)malloc(sizeof(int32_t) * (size_t)window);
      for (int w = 0; w < T / window; ++w) {
        const int c = dpp_once(F, r, m, D1, window, res, ny, sub);
        memcpy(res + ny, sub, sizeof(int32_t) * (size_t)c);
        ny += c;
      }
      if (T % window > 0) {
        const int c = dpp_once(F, r, m, D1, T % window, res, ny, sub);
        memcpy(res + ny, sub, sizeof(int32_t) * (size_t)c);
        ny += c;
      }
      free(sub);
    }
    for (int i = 0; i < ny; ++i) out_idx[i] = order[res[i]];
    free(res); free(F); free(r);
  }
  free(rel); free(order);
  return ny;
}

ORC_API int orc_dpp_request(const double* emb, const double* score, int n, int D, const orc_dpp_params* p,
                            int32_t* out_idx, int32_t* status) {
  return orc_dpp_request_ex(emb, NULL, NULL, 0, 1, 0, score, n, D, p, out_idx, status, NULL, NULL);
}

/* Dense kernel matrix exactly as KernelMatrix materialises it (for tests of the row-on-demand form). */
ORC_API void orc_dpp_kernel_matrix(const double* emb, const double* rel, int n, int D, double alpha, int normalize,
                                   double* L) {
  const int D1 = D + 1;
  double* F = (double*)malloc(sizeof(double) * (size_t)n * D1);
  double* r = (double*)malloc(sizeof(double) * (size_t)n);
  for (int i = 0; i < n; ++i) {
    double* f = F + (size_t)i * D1;
    memcpy(f, emb + (size_t)i * D, sizeof(double) * (size_t)D);
    if (normalize) { const double s = 1 / g_norm2(f, D); for (int d = 0; d < D; ++d) f[d] *= s; }
    f[D] = 1;
    for (int d = 0; d < D1; ++d) f[d] *= 0.70710678118654752440;
    r[i] = exp(alpha * rel[i]);
  }
  for (int j = 0; j < n; ++j) dpp_L_row(F, r, n, D1, j, L + (size_t)j * n);
  free(F); free(r);
}

/* ------------------------------------------------------------------------------------------------ SSD */
/*
 * SSDSort.doSort + SSDWithSlidingWindow (sort/ssd_sort.go:297-343, :346-486), table path, one request with FRESH
 * embeddings (upstream mutates the cached slices in place, :423-431,447-449, so its result depends on what earlier
 * requests did to the cache; the first-request behaviour is the one restated).
 *   emb: n x D (f64), score: Item.Score in input order.  doSort always sorts descending first (:301, Go pdqsort).
 * gonum pieces: floats.Dot = f64.DotUnitary (4 partial sums), floats.Norm (scaled), floats.Add/Sub elementwise,
 * VecDense.ScaleVec (dst = alpha * x).  [UNVERIFIED-UPSTREAM] as for DPP.
 * out_idx: indices into the INPUT list.  *status: 0 = re-ranked (T = min(n', ctx.Size) entries), 1 = upstream returned
 * the (sorted, truncated) items unchanged (gamma == 0, or all-zero scores with normalisation) -> identity over them.
 */
typedef struct {
  double gamma;
  int32_t top_n, window_size, norm_mode, normalize_emb, use_ssd_star, candidate_count;
  double min_score_percent;
} orc_ssd_params;

ORC_API int orc_ssd_request(const double* emb_in, const double* score, int n, int D, const orc_ssd_params* p, int32_t* out_idx,
                            int32_t* status) {
  *status = 0;
  if (n == 0) return 0;
  const int Tsz = p->top_n;
  int32_t* order = (int32_t*)malloc(sizeof(int32_t) * (size_t)n);
  orc_go_sort(score, n, 1, order); /* :301 */
  int m = n;
  if (p->gamma == 0) { /* :304-307 */
    for (int i = 0; i < n; ++i) out_idx[i] = order[i];
    *status = 1;
    free(order);
    return n;
  }
  if ((p->candidate_count > 0 || p->min_score_percent > 0) && n > Tsz) { /* :311-330 */
    if (p->candidate_count > 0) {
      const int cnt = Tsz > p->candidate_count ? Tsz : p->candidate_count;
      if (cnt < m) m = cnt;
    }
    if (p->min_score_percent > 0 && m > Tsz) {
      int idx = Tsz;
      const double mx = score[order[0]];
      for (; idx < m; ++idx)
        if (score[order[idx]] / mx < p->min_score_percent) break;
      m = idx;
    }
  }
  int window = p->window_size;
  if (window <= 1) window = 5; /* :358-361 (NewSSDSort defaults <=0 to 5 as well, :89-91) */
  double* rel = (double*)malloc(sizeof(double) * (size_t)m);
  for (int i = 0; i < m; ++i) rel[i] = score[order[i]];
  if (p->norm_mode == 1) {
    double mean = 0;
    for (int i = 0; i < m; ++i) mean += rel[i];
    mean /= (double)m;
    double ss = 0, comp = 0;
    for (int i = 0; i < m; ++i) { const double d = rel[i] - mean; ss += d * d; comp += d; }
    const double var = (ss - comp * comp / (double)m) / (double)m;
    if (mean == 0 || var == 0) *status = 1;
    else { const double sd = sqrt(var); for (int i = 0; i < m; ++i) rel[i] = (rel[i] - mean) / sd; }
  } else if (p->norm_mode == 2) {
    const double mx = rel[0], mn = rel[m - 1], span = mx - mn;
    if (span == 0) *status = 1;
    else for (int i = 0; i < m; ++i) rel[i] = ((rel[i] - mn) / span) * (1 - 1e-6) + 1e-6;
  }
  if (*status) {
    for (int i = 0; i < m; ++i) out_idx[i] = order[i];
    free(rel); free(order);
    return m;
  }
  double* E = (double*)malloc(sizeof(double) * (size_t)m * D);
  for (int i = 0; i < m; ++i) {
    double* e = E + (size_t)i * D;
    memcpy(e, emb_in + (size_t)order[i] * D, sizeof(double) * (size_t)D);
    if (p->normalize_emb) { const double s = 1 / g_norm2(e, D); for (int d = 0; d < D; ++d) e[d] *= s; }
  }
  const int T = m < Tsz ? m : Tsz;
  uint8_t* selected = (uint8_t*)calloc((size_t)m, 1);
  int32_t* indices = (int32_t*)malloc(sizeof(int32_t) * (size_t)(T > 0 ? T : 1));
  int* qB = (int*)malloc(sizeof(int) * (size_t)window);            /* CycleQueue of capacity window */
  double* qP = (double*)malloc(sizeof(double) * (size_t)window * m);
  int q_front = 0, q_len = 0;
  double* qual = (double*)malloc(sizeof(double) * (size_t)m);
  int t = 1, ni = 0;
  int idx = g_max_idx(rel, m);
  selected[idx] = 1;
  indices[ni++] = idx;
  double volume = p->gamma;
  if (!p->use_ssd_star) {
    const double l2 = g_norm2(E + (size_t)idx * D, D);
    if (!(isnan(l2) || isinf(l2))) volume *= l2;
  }
  while (t < T) {
    if (t > window) { /* :415-432: give back the projection on the item leaving the window */
      const int i = qB[q_front];
      const double* proj = qP + (size_t)q_front * m;
      q_front = (q_front + 1) % window; --q_len;
      const double* ei = E + (size_t)i * D;
      for (int j = 0; j < m; ++j) {
        if (selected[j]) continue;
        double* ej = E + (size_t)j * D;
        for (int d = 0; d < D; ++d) ej[d] = ej[d] + proj[j] * ei[d]; /* ScaleVec then floats.Add */
      }
    }
    const int slot = (q_front + q_len) % window; /* B.Push(idx), P.Push(projections) */
    qB[slot] = idx; ++q_len;
    double* proj = qP + (size_t)slot * m;
    const double* ei = E + (size_t)idx * D;
    for (int j = 0; j < m; ++j) {
      proj[j] = 0.0;
      if (selected[j]) continue;
      double* ej = E + (size_t)j * D;
      double pj = g_dot_unitary(ej, ei, D);
      pj /= g_dot_unitary(ei, ei, D);
      if (isnan(pj) || isinf(pj)) pj = 1.0;
      proj[j] = pj;
      for (int d = 0; d < D; ++d) ej[d] = ej[d] - pj * ei[d]; /* ScaleVec then floats.Sub */
    }
    ++t;
    for (int i = 0; i < m; ++i) {
      if (selected[i]) qual[i] = -1.7976931348623157e308;
      else {
        const double l2 = g_norm2(E + (size_t)i * D, D);
        qual[i] = (isnan(l2) || isinf(l2)) ? rel[i] + volume * 0.5 : rel[i] + volume * l2;
      }
    }
    idx = g_max_idx(qual, m);
    selected[idx] = 1;
    indices[ni++] = idx;
    if (!p->use_ssd_star) {
      const double l2 = g_norm2(E + (size_t)idx * D, D);
      if (!(isnan(l2) || isinf(l2))) volume *= l2;
    }
  }
  for (int i = 0; i < ni; ++i) out_idx[i] = order[indices[i]];
  free(rel); free(order); free(E); free(selected); free(indices); free(qB); free(qP); free(qual);
  return ni;
}

ORC_API int orc_num_threads(void) { return orc_hw_threads(); }
