// gather_fm.cu — multi-table categorical feature gather + FM second-order forward.
//
// Replaces, for the item side, module.FeatureDao.FeatureFetch (module/feature_dao.go:24-26; e.g.
// module/feature_hologres_dao.go:489-696: ceil(n/600) SQL IN() queries per request) and the remote ALINK_FM
// processor behind algorithm/eas/fm_request.go:29-79 / fm_response.go:28-34.  The per-item field ids and the
// per-field factor/linear tables are HBM resident; one launch gathers F rows of 64 B per candidate, evaluates the
// FM logit in a fixed f32 order (bit-identical to oracle/oracle.c orc_gather_fm) and, for the dense tower, emits the
// concatenated factors as bf16 (round to nearest even; the tower input carries no second term, see mlp.cu) in the
// layout the MMA reads.
//
// User / context features (service/rank/algo_data.go:104-118: features = userFeatures, then itemFeatures on top;
// service/rank/rank_service.go:175-183 user.MakeUserFeatures()): a request's U categorical user fields are the same for
// all of its candidates, so user_prefix_kernel evaluates them ONCE per request — the FM running sums after the user
// fields (the gather continues from that state: user fields first, then item fields, the oracle's order) and the user
// share of the tower's first layer, ubias[b][j] = b1[j] + sum_c W1[j][user column c] * bf16(x_user[c]).
//
// Mapping: 4 lanes per candidate, lane s owns factor dims 4s..4s+3 (one LDG.128 per field -> the 4 lanes of a
// candidate read one 64-B row in a single request); fields are walked in order, 8 independent loads in flight per
// lane.  HBM-bound: algorithmic bytes per candidate = F*4 (ids) + F*fdim*4 (rows) + F*4 (linear) + 4 (logit).
#include "handle.h"
#include <cstdlib>
#include <cuda_bf16.h>

namespace prg {

struct TableSet {
  const float* factors[kMaxFields];
  const float* linear[kMaxFields];
  uint32_t rows[kMaxFields];
};

__device__ __forceinline__ uint16_t bf16_bits(float f) { return __bfloat16_as_ushort(__float2bfloat16_rn(f)); }
__device__ __forceinline__ float bf16_val(uint16_t h) { return __uint_as_float((uint32_t)h << 16); }

// X layout: [M][F*16] bf16: bf16(factor k of field f) at column f*16+k.
constexpr int kFmState = 36;   // per request: lin, pad[3], s[16], ss[16]

// One CTA per request.  ids: [B][U] (0xFFFFFFFF or out of range = feature absent), dense: [B][n_dense].
// Wu: [U*16 + n_dense][N1] f32 (bf16 weights widened), b1: [N1].  state: [B][kFmState], ubias: [B][N1] (nullable).
__global__ void __launch_bounds__(256)
user_prefix_kernel(const uint32_t* __restrict__ ids, const float* __restrict__ dense, int U, int n_dense, int F,
                   const __grid_constant__ TableSet ts, float w0, float* __restrict__ state,
                   const float* __restrict__ Wu, const float* __restrict__ b1, int N1, float* __restrict__ ubias) {
  pdl_wait();
  pdl_launch_dependents();
  __shared__ float xv[kMaxFields * 16];        // the user's factor values [u][k]
  __shared__ float xw[kMaxFields];             // the user's linear weights [u]
  __shared__ float xu[kMaxFields * 16 + 64];   // bf16-rounded user input columns of the tower
  const int b = blockIdx.x, tid = threadIdx.x;
  const uint32_t* idp = ids ? ids + (size_t)b * U : nullptr;
  // all loads of the request at once (one thread per factor value / linear weight / dense value), the sequential sums after
  for (int i = tid; i < U * 16; i += 256) {
    const int u = i >> 4, k = i & 15;
    const uint32_t id = idp ? idp[u] : 0xFFFFFFFFu;
    const float v = id < ts.rows[F + u] ? ts.factors[F + u][(size_t)id * 16 + k] : 0.f;
    xv[i] = v;
    xu[i] = bf16_val(bf16_bits(v));
  }
  for (int u = tid; u < U; u += 256) {
    const uint32_t id = idp ? idp[u] : 0xFFFFFFFFu;
    xw[u] = (id < ts.rows[F + u] && ts.linear[F + u]) ? ts.linear[F + u][id] : 0.f;
  }
  for (int c = tid; c < n_dense; c += 256) xu[U * 16 + c] = bf16_val(bf16_bits(dense ? dense[(size_t)b * n_dense + c] : 0.f));
  __syncthreads();
  if (tid < 16) {          // factor dim k: running sums over the user fields, in field order
    float s = 0.f, ss = 0.f;
    for (int u = 0; u < U; ++u) {
      const float v = xv[u * 16 + tid];
      s = __fadd_rn(s, v);
      ss = __fmaf_rn(v, v, ss);
    }
    state[(size_t)b * kFmState + 4 + tid] = s;
    state[(size_t)b * kFmState + 20 + tid] = ss;
  } else if (tid == 16) {
    float lin = w0;
    for (int u = 0; u < U; ++u) lin = __fadd_rn(lin, xw[u]);
    state[(size_t)b * kFmState] = lin;
  }
  if (!ubias) return;
  const int Ku = U * 16 + n_dense;
  // fp64 accumulation, rounded once at the end.  Two output columns per thread and 16 rows of Wu per batch: 32 independent
  // L2 loads in flight per thread (the kernel is pure load latency: 52 us with one load in flight, r2d launch list)
  for (int j0 = tid; j0 < N1; j0 += 512) {
    const int j1 = j0 + 256;
    const bool two = j1 < N1;
    double a0 = 0.0, a1 = 0.0;
    int c = 0;
    for (; c + 16 <= Ku; c += 16) {
      float w0v[16], w1v[16];
#pragma unroll
      for (int t = 0; t < 16; ++t) {
        w0v[t] = __ldg(Wu + (size_t)(c + t) * N1 + j0);
        w1v[t] = two ? __ldg(Wu + (size_t)(c + t) * N1 + j1) : 0.f;
      }
#pragma unroll
      for (int t = 0; t < 16; ++t) {
        const double x = (double)xu[c + t];
        a0 = fma((double)w0v[t], x, a0);
        a1 = fma((double)w1v[t], x, a1);
      }
    }
    for (; c < Ku; ++c) {
      const double x = (double)xu[c];
      a0 = fma((double)__ldg(Wu + (size_t)c * N1 + j0), x, a0);
      if (two) a1 = fma((double)__ldg(Wu + (size_t)c * N1 + j1), x, a1);
    }
    ubias[(size_t)b * N1 + j0] = (float)(a0 + (double)b1[j0]);
    if (two) ubias[(size_t)b * N1 + j1] = (float)(a1 + (double)b1[j1]);
  }
}

template <int F_UNROLL, int MIN_CTAS>
__global__ void __launch_bounds__(256, MIN_CTAS)
gather_fm_kernel(const uint32_t* __restrict__ rows, const uint64_t* __restrict__ keys, uint32_t* __restrict__ rows_out,
                 int M, const uint32_t* __restrict__ fields, uint64_t field_rows,
                 int F, const __grid_constant__ TableSet ts, float w0, float* __restrict__ logit_out,
                 uint16_t* __restrict__ x_out, const float* __restrict__ fm_state, int rows_per_req) {
  pdl_wait();                 // chained launch: the predecessor's writes are visible from here on
  pdl_launch_dependents();
  const int gid = blockIdx.x * blockDim.x + threadIdx.x;
  const int item = gid >> 2, sub = gid & 3, lane = threadIdx.x & 31;
  const bool in_range = item < M;
  // candidates arrive as item rows, or (fused path) as the recall's order keys: the row is unpacked here and written
  // out for the later stages instead of by a launch of its own
  uint32_t row = 0xFFFFFFFFu;
  if (in_range) {
    if (keys) {
      const uint64_t key = keys[item];
      row = key ? key_row(key) : 0xFFFFFFFFu;
      if (sub == 0) rows_out[item] = row;
    } else {
      row = rows[item];
    }
  }
  const bool live = row != 0xFFFFFFFFu && (uint64_t)row < field_rows;
  const uint32_t* idp = fields + (size_t)(live ? row : 0) * F;
  const int K = F * 16;

  float lin = w0;
  float s[4] = {0.f, 0.f, 0.f, 0.f}, ss[4] = {0.f, 0.f, 0.f, 0.f};
  if (fm_state) {   // the request's running sums after its user fields (user_prefix_kernel)
    const float* st = fm_state + (size_t)((in_range ? item : M - 1) / rows_per_req) * kFmState;
    lin = st[0];
    const float4 a = *reinterpret_cast<const float4*>(st + 4 + sub * 4), b = *reinterpret_cast<const float4*>(st + 20 + sub * 4);
    s[0] = a.x; s[1] = a.y; s[2] = a.z; s[3] = a.w;
    ss[0] = b.x; ss[1] = b.y; ss[2] = b.z; ss[3] = b.w;
  }
  for (int f0 = 0; f0 < F; f0 += F_UNROLL) {
    uint32_t id[F_UNROLL];
    float4 v[F_UNROLL];
    float w[F_UNROLL];
#pragma unroll
    for (int u = 0; u < F_UNROLL; ++u) id[u] = (live && f0 + u < F) ? idp[f0 + u] : 0xFFFFFFFFu;
#pragma unroll
    for (int u = 0; u < F_UNROLL; ++u) {
      const int f = f0 + u;
      const bool ok = f < F && id[u] < ts.rows[f < F ? f : 0];
      v[u] = ok ? *reinterpret_cast<const float4*>(ts.factors[f] + (size_t)id[u] * 16 + sub * 4)
                : make_float4(0.f, 0.f, 0.f, 0.f);
      w[u] = (ok && ts.linear[f]) ? ts.linear[f][id[u]] : 0.f;
    }
#pragma unroll
    for (int u = 0; u < F_UNROLL; ++u) {
      const int f = f0 + u;
      if (f < F) {
        lin = __fadd_rn(lin, w[u]);
        s[0] = __fadd_rn(s[0], v[u].x); ss[0] = __fmaf_rn(v[u].x, v[u].x, ss[0]);
        s[1] = __fadd_rn(s[1], v[u].y); ss[1] = __fmaf_rn(v[u].y, v[u].y, ss[1]);
        s[2] = __fadd_rn(s[2], v[u].z); ss[2] = __fmaf_rn(v[u].z, v[u].z, ss[2]);
        s[3] = __fadd_rn(s[3], v[u].w); ss[3] = __fmaf_rn(v[u].w, v[u].w, ss[3]);
        if (x_out && in_range) {
          const __nv_bfloat162 p0 = __floats2bfloat162_rn(v[u].x, v[u].y), p1 = __floats2bfloat162_rn(v[u].z, v[u].w);
          *reinterpret_cast<uint2*>(x_out + (size_t)item * K + f * 16 + sub * 4) =
              make_uint2(*reinterpret_cast<const uint32_t*>(&p0), *reinterpret_cast<const uint32_t*>(&p1));
        }
      }
    }
  }
  // t_k = s_k^2 - ss_k for this lane's 4 dims; inter = sequential sum over k = 0..15 across the 4 lanes
  float t[4];
#pragma unroll
  for (int c = 0; c < 4; ++c) t[c] = __fmaf_rn(s[c], s[c], -ss[c]);
  float inter = 0.f;
  const int base = lane & ~3;
#pragma unroll
  for (int src = 0; src < 4; ++src) {
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const float tv = __shfl_sync(0xffffffffu, t[c], base + src);
      inter = __fadd_rn(inter, tv);
    }
  }
  if (in_range && sub == 0 && logit_out) logit_out[item] = live ? __fmaf_rn(0.5f, inter, lin) : 0.f;
}

// score = (float)(1/(1+exp(-(double)logit))) widened to f64 (AlgoResponse.GetScore() is float64).
__global__ void logit_to_score_kernel(const float* a, const float* b, const uint32_t* rows, int M, double* out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M) return;
  float l = a ? a[i] : 0.f;
  if (b) l = __fadd_rn(l, b[i]);
  const float sc = (float)(1.0 / (1.0 + exp(-(double)l)));
  out[i] = (rows[i] == 0xFFFFFFFFu) ? 0.0 : (double)sc;
}

static int fill_tables(prg_handle* h, TableSet* ts, uint32_t n_tables) {
  for (uint32_t f = 0; f < n_tables; ++f) {
    if (!h->tables[f].factors) return fail(PRG_ESTATE, "feature table " + std::to_string(f) + " not set");
    ts->factors[f] = h->tables[f].factors;
    ts->linear[f] = h->tables[f].linear;
    ts->rows[f] = (uint32_t)h->tables[f].rows;
  }
  return PRG_OK;
}

// user_ids_dev [B][U] / user_dense_dev [B][n_dense] (nullable = every user feature absent) -> h->fm_state, h->ubias
// ahead: launch on the handle's side stream, forked from / joined to the main stream by events, so that the kernel (pure
// load latency, a handful of CTAs) runs beside the recall instead of in front of the gather; the caller makes the main
// stream wait for h->ev_join before the gather (rank_device)
int user_prefix_device(prg_handle* h, const uint32_t* user_ids_dev, const float* user_dense_dev, int B, bool need_mlp, bool ahead) {
  const uint32_t U = h->n_user_fields, nd = h->n_user_dense;
  if (h->n_fields + U > (uint32_t)kMaxFields) return fail(PRG_EUNSUPPORTED, "item + user fields > 64");
  TableSet ts{};
  PRG_TRY(fill_tables(h, &ts, h->n_fields + U));
  PRG_TRY(h->fm_state.ensure((size_t)B * kFmState * 4));
  float* ubias = nullptr;
  int N1 = 0;
  if (need_mlp) {
    N1 = (int)h->mlp_dims[1];
    if (h->mlp_k_user != U * 16 + nd) return fail(PRG_ESTATE, "prg_set_mlp must follow prg_set_user_fields");
    PRG_TRY(h->ubias.ensure((size_t)B * N1 * 4));
    ubias = (float*)h->ubias.p;
  }
  if (ahead) {
    if (!h->side_stream) {
      PRG_CUDA(cudaStreamCreateWithFlags(&h->side_stream, cudaStreamNonBlocking));
      PRG_CUDA(cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming));
      PRG_CUDA(cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming));
    }
    PRG_CUDA(cudaEventRecord(h->ev_fork, h->stream));            // behind the H2D of the user ids / the previous batch
    PRG_CUDA(cudaStreamWaitEvent(h->side_stream, h->ev_fork, 0));
    user_prefix_kernel<<<(unsigned)B, 256, 0, h->side_stream>>>(user_ids_dev, user_dense_dev, (int)U, (int)nd, (int)h->n_fields, ts,
                                                                h->fm_w0, (float*)h->fm_state.p, (const float*)h->mlp_Wu.p,
                                                                (const float*)h->mlp_b[0].p, N1, ubias);
    PRG_CUDA(cudaGetLastError());
    PRG_CUDA(cudaEventRecord(h->ev_join, h->side_stream));
    h->prefix_ahead = true;
    count_launch(h);
    return PRG_OK;
  }
  StageScope span(h, ST_GATHER_FM);
  PRG_CUDA(launch_chained(h, user_prefix_kernel, dim3((unsigned)B), dim3(256), 0, 1, user_ids_dev, user_dense_dev, (int)U, (int)nd,
                          (int)h->n_fields, ts, h->fm_w0, (float*)h->fm_state.p, (const float*)h->mlp_Wu.p,
                          (const float*)h->mlp_b[0].p, N1, ubias));
  count_launch(h);
  return PRG_OK;
}

// rows_per_req > 0: the candidates of request b are items [b * rows_per_req, (b+1) * rows_per_req) and the FM sums
// continue from h->fm_state[b] (user_prefix_device ran before)
int gather_fm_device(prg_handle* h, const uint32_t* rows_dev, int M, float* logit_dev, uint16_t* x_dev,
                     const uint64_t* keys_dev, uint32_t* rows_out, int rows_per_req) {
  if (!h->fields) return fail(PRG_ESTATE, "item fields not set (prg_set_item_fields)");
  if (h->fdim != 16) return fail(PRG_EUNSUPPORTED, "feature tables must have fdim == 16");
  TableSet ts{};
  PRG_TRY(fill_tables(h, &ts, h->n_fields));
  const float* state = rows_per_req > 0 ? (const float*)h->fm_state.p : nullptr;
  StageScope span(h, ST_GATHER_FM);
  const int threads = 256;
  const long long total = (long long)M * 4;
  const unsigned grid = (unsigned)((total + threads - 1) / threads);
  // 4 CTAs per SM (64 registers) instead of 3: the kernel is bound by the row loads it keeps in flight (0.097 -> 0.090 ms)
  PRG_CUDA(launch_chained(h, gather_fm_kernel<8, 4>, dim3(grid), dim3(threads), 0, 1, rows_dev, keys_dev, rows_out, M, h->fields,
                          h->fields_rows, (int)h->n_fields, ts, h->fm_w0, logit_dev, x_dev, state, rows_per_req > 0 ? rows_per_req : 1));
  count_launch(h);
  return PRG_OK;
}

int logits_to_scores_device(prg_handle* h, const float* a, const float* b, const uint32_t* rows_dev, int M, double* out) {
  logit_to_score_kernel<<<(M + 255) / 256, 256, 0, h->stream>>>(a, b, rows_dev, M, out);
  PRG_CUDA(cudaGetLastError());
  count_launch(h);
  return PRG_OK;
}

}  // namespace prg
