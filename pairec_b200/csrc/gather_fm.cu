// gather_fm.cu — multi-table categorical feature gather + FM second-order forward.
//
// Replaces, for the item side, module.FeatureDao.FeatureFetch (module/feature_dao.go:24-26; e.g.
// module/feature_hologres_dao.go:489-696: ceil(n/600) SQL IN() queries per request) and the remote ALINK_FM
// processor behind algorithm/eas/fm_request.go:29-79 / fm_response.go:28-34.  The per-item field ids and the
// per-field factor/linear tables are HBM resident; one launch gathers F rows of 64 B per candidate, evaluates the
// FM logit in a fixed f32 order (bit-identical to oracle/oracle.c orc_gather_fm) and, for the dense tower, emits the
// concatenated factors as bf16 hi/lo pairs ("bf16x2" activations, see mlp.cu) in the layout the MMA reads.
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

// X layout: [M][2*F*16] bf16: hi at column f*16+k, lo at F*16 + f*16+k.
template <int F_UNROLL, int MIN_CTAS>
__global__ void __launch_bounds__(256, MIN_CTAS)
gather_fm_kernel(const uint32_t* __restrict__ rows, const uint64_t* __restrict__ keys, uint32_t* __restrict__ rows_out,
                 int M, const uint32_t* __restrict__ fields, uint64_t field_rows,
                 int F, const __grid_constant__ TableSet ts, float w0, float* __restrict__ logit_out,
                 uint16_t* __restrict__ x_out) {
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
          const float e[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
          uint16_t hi[4], lo[4];
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            hi[c] = bf16_bits(e[c]);
            lo[c] = bf16_bits(__fsub_rn(e[c], bf16_val(hi[c])));
          }
          uint16_t* xr = x_out + (size_t)item * (2 * K) + f * 16 + sub * 4;
          *reinterpret_cast<uint2*>(xr) = make_uint2((uint32_t)hi[0] | ((uint32_t)hi[1] << 16), (uint32_t)hi[2] | ((uint32_t)hi[3] << 16));
          *reinterpret_cast<uint2*>(xr + K) = make_uint2((uint32_t)lo[0] | ((uint32_t)lo[1] << 16), (uint32_t)lo[2] | ((uint32_t)lo[3] << 16));
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

int gather_fm_device(prg_handle* h, const uint32_t* rows_dev, int M, float* logit_dev, uint16_t* x_dev,
                     const uint64_t* keys_dev, uint32_t* rows_out) {
  if (!h->fields) return fail(PRG_ESTATE, "item fields not set (prg_set_item_fields)");
  if (h->fdim != 16) return fail(PRG_EUNSUPPORTED, "feature tables must have fdim == 16");
  TableSet ts{};
  for (uint32_t f = 0; f < h->n_fields; ++f) {
    if (!h->tables[f].factors) return fail(PRG_ESTATE, "feature table " + std::to_string(f) + " not set");
    ts.factors[f] = h->tables[f].factors;
    ts.linear[f] = h->tables[f].linear;
    ts.rows[f] = (uint32_t)h->tables[f].rows;
  }
  StageScope span(h, ST_GATHER_FM);
  const int threads = 256;
  const long long total = (long long)M * 4;
  const unsigned grid = (unsigned)((total + threads - 1) / threads);
  static const int min_ctas = getenv("PRG_GATHER_MINB") ? atoi(getenv("PRG_GATHER_MINB")) : 4;   // A/B measurements
  // 4 CTAs per SM (64 registers) instead of 3: the kernel is bound by the row loads it keeps in flight (0.097 -> 0.090 ms)
  if (min_ctas == 4)
    PRG_CUDA(launch_chained(h, gather_fm_kernel<8, 4>, dim3(grid), dim3(threads), 0, 1, rows_dev, keys_dev, rows_out, M, h->fields,
                            h->fields_rows, (int)h->n_fields, ts, h->fm_w0, logit_dev, x_dev));
  else
    PRG_CUDA(launch_chained(h, gather_fm_kernel<8, 3>, dim3(grid), dim3(threads), 0, 1, rows_dev, keys_dev, rows_out, M, h->fields,
                            h->fields_rows, (int)h->n_fields, ts, h->fm_w0, logit_dev, x_dev));
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
