// ssd.cu — SSDSort.doSort + SSDWithSlidingWindow on the GPU (sort/ssd_sort.go:297-343, :346-486).
//
// Sliding spectrum decomposition: T greedy picks; after each pick every remaining candidate's embedding loses its
// projection on the picked item (Gram-Schmidt step, fp64, in place), the pick's "volume" scales the diversity term,
// and once the window is full the projection on the item leaving the window is given back.  The embeddings are
// MUTABLE fp64 state (n x D x 8 B = 1 MB per request at n=1000, D=128), so unlike DPP they cannot stay in shared
// memory or registers; they live in a transposed global scratch E[d][i] (coalesced, L2 resident: 64 MB per 64-request
// batch).  One CTA per request, one thread per candidate.  Two passes over the dims per pick: (A) give-back + dot with
// the picked item, (B) projection removal fused with the scaled-norm recurrence of the quality term.
//
// Arithmetic follows oracle/oracle.c orc_ssd_request (gonum floats.Dot = DotUnitary with four partial sums,
// floats.Norm scaled form, ScaleVec then Add/Sub as separate roundings, MaxIdx first maximum); upstream mutates cached
// slices across requests — the first-request behaviour (fresh embeddings) is what is implemented.
#include "handle.h"
#include <math_constants.h>

namespace prg {

constexpr int kSsdMaxItems = 1024;
constexpr int kSsdMaxN = 4096;

struct SsdArgs {
  const uint32_t* rows;
  const double* score;
  int n;
  const void* D;
  uint64_t D_rows;
  int D_dim;
  prg_ssd_params p;
  double* Es;   // [B][D_dim][1024]
  double* Ps;   // [B][window][1024]
  int32_t* out_idx;
  int32_t* out_n;
  int32_t* status;
  int window;
};

__device__ __forceinline__ uint64_t f64_ord_s(double d) {
  uint64_t u = (uint64_t)__double_as_longlong(d);
  if ((u & 0x7FFFFFFFFFFFFFFFull) > 0x7FF0000000000000ull) return 0ull;
  return (u >> 63) ? ~u : (u | 0x8000000000000000ull);
}
struct AmS { double v; int i; };
__device__ __forceinline__ AmS ams(AmS a, AmS b) {
  if (isnan(b.v)) return a;
  if (isnan(a.v)) return b;
  if (b.v > a.v || (b.v == a.v && b.i < a.i)) return b;
  return a;
}

template <typename T>
__global__ void __launch_bounds__(kSsdMaxItems, 1) ssd_kernel(const SsdArgs a) {
  extern __shared__ __align__(16) uint8_t ssm[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int b = blockIdx.x, n = a.n, D = a.D_dim, Tsz = a.p.top_n, window = a.window;

  double* ei = reinterpret_cast<double*>(ssm);          // [512] picked item's embedding
  double* eo = ei + 512;                                // [512] embedding of the item leaving the window
  double* red_v = eo + 512;                             // [32]
  int32_t* red_i = reinterpret_cast<int32_t*>(red_v + 32);  // [32]
  int32_t* order = red_i + 32;                          // [1024]
  int32_t* indices = order + kSsdMaxItems;              // [Tsz]
  int32_t* qB = indices + ((Tsz + 3) & ~3);             // [window]
  uint64_t* skey = reinterpret_cast<uint64_t*>(qB + ((window + 3) & ~3) + 2);  // presort staging: [4096] keys + idx
  __shared__ double s_vol, s_p0, s_p1, s_l2;
  __shared__ int s_j, s_m, s_err;

  const uint32_t* rows = a.rows + (size_t)b * n;
  const double* score = a.score + (size_t)b * n;
  double* Es = a.Es + (size_t)b * D * kSsdMaxItems;
  double* Ps = a.Ps + (size_t)b * window * kSsdMaxItems;

  if (tid == 0) { s_m = 0; s_err = 0; }
  __syncthreads();
  {
    int cnt = 0;
    for (int i = tid; i < n; i += blockDim.x) cnt += (rows[i] != 0xFFFFFFFFu);
    if (cnt) atomicAdd(&s_m, cnt);
  }
  __syncthreads();
  const int nv = s_m;
  __syncthreads();
  if (nv == 0) {
    if (tid == 0) { a.out_n[b] = 0; a.status[b] = 0; }
    return;
  }
  // doSort always sorts descending first (:301): stable (score desc, input position asc) on the device
  {
    uint32_t P2 = 32;
    while (P2 < (uint32_t)nv) P2 <<= 1;
    int32_t* sidx = reinterpret_cast<int32_t*>(skey + P2);
    for (uint32_t i = tid; i < P2; i += blockDim.x) {
      skey[i] = (i < (uint32_t)nv) ? f64_ord_s(score[i]) : 0ull;
      sidx[i] = (i < (uint32_t)nv) ? (int32_t)i : 0x7FFFFFFF;
    }
    __syncthreads();
    for (uint32_t size = 2; size <= P2; size <<= 1) {
      for (uint32_t stride = size >> 1; stride > 0; stride >>= 1) {
        for (uint32_t i = tid; i < (P2 >> 1); i += blockDim.x) {
          const uint32_t pos = 2 * i - (i & (stride - 1));
          const uint64_t ka = skey[pos], kb = skey[pos + stride];
          const int32_t ia = sidx[pos], ib = sidx[pos + stride];
          const bool a_after_b = (ka < kb) || (ka == kb && ia > ib);
          if (a_after_b == ((pos & size) == 0)) { skey[pos] = kb; skey[pos + stride] = ka; sidx[pos] = ib; sidx[pos + stride] = ia; }
        }
        __syncthreads();
      }
    }
    if (tid == 0) {
      int mm = nv;
      if (a.p.gamma != 0 && (a.p.candidate_count > 0 || a.p.min_score_percent > 0) && nv > Tsz) {  // :311-330
        if (a.p.candidate_count > 0) {
          const int cnt = Tsz > a.p.candidate_count ? Tsz : a.p.candidate_count;
          if (cnt < mm) mm = cnt;
        }
        if (a.p.min_score_percent > 0 && mm > Tsz) {
          int id = Tsz;
          const double mx = score[sidx[0]];
          for (; id < mm; ++id)
            if (score[sidx[id]] / mx < a.p.min_score_percent) break;
          mm = id;
        }
      }
      s_m = mm;
    }
    __syncthreads();
    if (a.p.gamma == 0) {  // :304-307 — upstream returns the sorted items; report the first top_n of them
      const int c = nv < Tsz ? nv : Tsz;
      for (int i = tid; i < c; i += blockDim.x) a.out_idx[(size_t)b * Tsz + i] = sidx[i];
      if (tid == 0) { a.out_n[b] = c; a.status[b] = 1; }
      return;
    }
    const int mm = s_m;
    if (mm <= kSsdMaxItems)
      for (int i = tid; i < mm; i += blockDim.x) order[i] = sidx[i];
    __syncthreads();
  }
  const int m = s_m;
  if (m > kSsdMaxItems) {
    if (tid == 0) { a.out_n[b] = 0; a.status[b] = 2; }
    return;
  }
  const bool active = tid < m;
  const int my_in = active ? order[tid] : 0;

  double rel = active ? score[my_in] : 0.0;
  if (a.p.norm_mode == 1 || a.p.norm_mode == 2) {  // :368-391
    if (tid == 0) {
      if (a.p.norm_mode == 1) {
        double sum = 0.0;
        for (int i = 0; i < m; ++i) sum = __dadd_rn(sum, score[order[i]]);
        const double mean = sum / (double)m;
        double ssq = 0.0, comp = 0.0;
        for (int i = 0; i < m; ++i) {
          const double d = __dsub_rn(score[order[i]], mean);
          ssq = __dadd_rn(ssq, __dmul_rn(d, d));
          comp = __dadd_rn(comp, d);
        }
        const double var = __dsub_rn(ssq, __dmul_rn(comp, comp) / (double)m) / (double)m;
        if (mean == 0 || var == 0) s_err = 1;
        s_p0 = mean;
        s_p1 = sqrt(var);
      } else {
        const double r0 = score[order[0]], r1 = score[order[m - 1]];
        if (__dsub_rn(r0, r1) == 0) s_err = 1;
        s_p0 = r1;
        s_p1 = __dsub_rn(r0, r1);
      }
    }
    __syncthreads();
    if (a.p.norm_mode == 1) rel = __dsub_rn(rel, s_p0) / s_p1;
    else rel = __dadd_rn(__dmul_rn(__dsub_rn(rel, s_p0) / s_p1, 1 - 1e-6), 1e-6);
  }
  __syncthreads();
  if (s_err) {  // "all item score are zeros": upstream returns the (sorted, truncated) items unchanged
    const int c = m < Tsz ? m : Tsz;
    for (int i = tid; i < c; i += blockDim.x) a.out_idx[(size_t)b * Tsz + i] = order[i];
    if (tid == 0) { a.out_n[b] = c; a.status[b] = 1; }
    return;
  }

  // gonum floats.Norm(v, 2) over this thread's column of Es
  auto col_norm = [&]() -> double {
    double scale = 0.0, sumsq = 1.0;
    for (int d = 0; d < D; ++d) {
      const double v = Es[(size_t)d * kSsdMaxItems + tid];
      if (v != 0.0) {
        const double av = fabs(v);
        if (isnan(av)) return CUDART_NAN;
        if (scale < av) {
          const double s = scale / av;
          sumsq = __dadd_rn(1.0, __dmul_rn(__dmul_rn(sumsq, s), s));
          scale = av;
        } else {
          const double s = av / scale;
          sumsq = __dadd_rn(sumsq, __dmul_rn(s, s));
        }
      }
    }
    if (isinf(scale)) return CUDART_INF;
    return __dmul_rn(scale, sqrt(sumsq));
  };

  // ---- embeddings -> fp64 scratch (transposed), L2-normalised as loadEmbeddingCache does
  if (active) {
    const uint32_t row = rows[my_in];
    const bool have = (uint64_t)row < a.D_rows;
    const T* src = reinterpret_cast<const T*>(a.D) + (size_t)(have ? row : 0) * D;
    for (int d = 0; d < D; ++d) Es[(size_t)d * kSsdMaxItems + tid] = have ? (double)src[d] : 0.0;
    if (a.p.normalize_emb) {
      const double s = 1.0 / col_norm();
      for (int d = 0; d < D; ++d) Es[(size_t)d * kSsdMaxItems + tid] = __dmul_rn(Es[(size_t)d * kSsdMaxItems + tid], s);
    }
  }
  __syncthreads();

  auto argmax = [&](double v) -> int {
    AmS am{v, tid};
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      AmS o;
      o.v = __shfl_xor_sync(0xffffffffu, am.v, off);
      o.i = __shfl_xor_sync(0xffffffffu, am.i, off);
      am = ams(am, o);
    }
    if (lane == 0) { red_v[warp] = am.v; red_i[warp] = am.i; }
    __syncthreads();
    if (warp == 0) {
      AmS x{red_v[lane], red_i[lane]};
      if (lane >= (int)(blockDim.x >> 5)) x.v = CUDART_NAN;
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) {
        AmS o;
        o.v = __shfl_xor_sync(0xffffffffu, x.v, off);
        o.i = __shfl_xor_sync(0xffffffffu, x.i, off);
        x = ams(x, o);
      }
      if (lane == 0) s_j = isnan(x.v) ? 0 : x.i;
    }
    __syncthreads();
    return s_j;
  };

  const int Tn = m < Tsz ? m : Tsz;
  bool selected = !active;  // padding threads never compete
  int t = 1, q_front = 0, q_len = 0;
  int idx = argmax(active ? rel : CUDART_NAN);
  if (tid == idx) selected = true;
  if (tid == 0) indices[0] = idx;
  if (tid == idx) {
    double vol = a.p.gamma;
    if (!a.p.use_ssd_star) {
      const double l2 = col_norm();
      if (!(isnan(l2) || isinf(l2))) vol = __dmul_rn(vol, l2);
    }
    s_vol = vol;
  }
  __syncthreads();
  int ni = 1;
  while (t < Tn) {
    const bool give_back = t > window;
    int slot_old = 0;
    if (give_back) {  // :415-432
      slot_old = q_front;
      const int i_old = qB[q_front];
      for (int d = tid; d < D; d += blockDim.x) eo[d] = Es[(size_t)d * kSsdMaxItems + i_old];
      q_front = (q_front + 1) % window;
      --q_len;
    }
    const int slot = (q_front + q_len) % window;
    ++q_len;
    for (int d = tid; d < D; d += blockDim.x) ei[d] = Es[(size_t)d * kSsdMaxItems + idx];
    __syncthreads();
    if (tid == 0) qB[slot] = idx;
    double l2 = 0.0;
    if (!selected) {
      const double po = give_back ? Ps[(size_t)slot_old * kSsdMaxItems + tid] : 0.0;
      // pass A: give-back fused with floats.Dot(e_j, e_idx) (DotUnitary: 4 partial sums, tail into s0)
      double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
      double d0 = 0.0, d1 = 0.0, d2 = 0.0, d3 = 0.0;  // Dot(e_idx, e_idx), same order
      int d = 0;
      for (; d + 4 <= D; d += 4) {
        double e[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          double v = Es[(size_t)(d + c) * kSsdMaxItems + tid];
          if (give_back) {
            v = __dadd_rn(v, __dmul_rn(po, eo[d + c]));
            Es[(size_t)(d + c) * kSsdMaxItems + tid] = v;
          }
          e[c] = v;
        }
        s0 = __dadd_rn(s0, __dmul_rn(e[0], ei[d]));
        s1 = __dadd_rn(s1, __dmul_rn(e[1], ei[d + 1]));
        s2 = __dadd_rn(s2, __dmul_rn(e[2], ei[d + 2]));
        s3 = __dadd_rn(s3, __dmul_rn(e[3], ei[d + 3]));
        d0 = __dadd_rn(d0, __dmul_rn(ei[d], ei[d]));
        d1 = __dadd_rn(d1, __dmul_rn(ei[d + 1], ei[d + 1]));
        d2 = __dadd_rn(d2, __dmul_rn(ei[d + 2], ei[d + 2]));
        d3 = __dadd_rn(d3, __dmul_rn(ei[d + 3], ei[d + 3]));
      }
      for (; d < D; ++d) {
        double v = Es[(size_t)d * kSsdMaxItems + tid];
        if (give_back) {
          v = __dadd_rn(v, __dmul_rn(po, eo[d]));
          Es[(size_t)d * kSsdMaxItems + tid] = v;
        }
        s0 = __dadd_rn(s0, __dmul_rn(v, ei[d]));
        d0 = __dadd_rn(d0, __dmul_rn(ei[d], ei[d]));
      }
      double pj = __dadd_rn(__dadd_rn(s0, s2), __dadd_rn(s1, s3)) / __dadd_rn(__dadd_rn(d0, d2), __dadd_rn(d1, d3));
      if (isnan(pj) || isinf(pj)) pj = 1.0;  // :440-444
      Ps[(size_t)slot * kSsdMaxItems + tid] = pj;
      // pass B: e_j -= pj * e_idx fused with the scaled norm of the updated vector
      double scale = 0.0, sumsq = 1.0;
      bool nan_seen = false;
      for (d = 0; d < D; ++d) {
        const double v = __dsub_rn(Es[(size_t)d * kSsdMaxItems + tid], __dmul_rn(pj, ei[d]));
        Es[(size_t)d * kSsdMaxItems + tid] = v;
        if (v != 0.0 && !nan_seen) {
          const double av = fabs(v);
          if (isnan(av)) nan_seen = true;
          else if (scale < av) {
            const double s = scale / av;
            sumsq = __dadd_rn(1.0, __dmul_rn(__dmul_rn(sumsq, s), s));
            scale = av;
          } else {
            const double s = av / scale;
            sumsq = __dadd_rn(sumsq, __dmul_rn(s, s));
          }
        }
      }
      l2 = nan_seen ? CUDART_NAN : (isinf(scale) ? CUDART_INF : __dmul_rn(scale, sqrt(sumsq)));
    } else if (tid < m) {
      Ps[(size_t)slot * kSsdMaxItems + tid] = 0.0;
    }
    ++t;
    const double vol = s_vol;
    double q;
    if (selected) q = active ? -1.7976931348623157e308 : CUDART_NAN;
    else q = (isnan(l2) || isinf(l2)) ? __dadd_rn(rel, __dmul_rn(vol, 0.5)) : __dadd_rn(rel, __dmul_rn(vol, l2));
    __syncthreads();  // all reads of ei / eo / s_vol done before the next pick rewrites them
    idx = argmax(q);
    if (tid == idx) {
      selected = true;
      if (!a.p.use_ssd_star && !(isnan(l2) || isinf(l2))) s_vol = __dmul_rn(vol, l2);  // :471-478
    }
    if (tid == 0) indices[ni] = idx;
    ++ni;
    __syncthreads();
  }
  for (int i = tid; i < ni; i += blockDim.x) a.out_idx[(size_t)b * Tsz + i] = order[indices[i]];
  if (tid == 0) { a.out_n[b] = ni; a.status[b] = 0; }
}

static size_t ssd_smem_bytes(int top_n, int window) {
  return 1024 * 8 + 32 * 8 + 32 * 4 + kSsdMaxItems * 4 + (size_t)((top_n + 3) & ~3) * 4 + (size_t)((window + 3) & ~3) * 4 + 16 +
         (size_t)kSsdMaxN * 12 + 64;
}

int ssd_device(prg_handle* h, const uint32_t* rows_dev, const double* score_dev, int B, int n, const prg_ssd_params& p,
               int32_t* out_idx, int32_t* out_n, int32_t* status) {
  if (!h->D) return fail(PRG_ESTATE, "diversity matrix not set (prg_set_diversity_matrix)");
  if (B <= 0 || n <= 0 || p.top_n <= 0) return fail(PRG_EINVAL, "B, n, top_n must be positive");
  if (n > kSsdMaxN) return fail(PRG_EUNSUPPORTED, "prg_ssd: n > 4096");
  if (h->D_dim > 512) return fail(PRG_EUNSUPPORTED, "prg_ssd: embedding dim > 512");
  if (p.top_n > 4096) return fail(PRG_EUNSUPPORTED, "prg_ssd: top_n > 4096");
  if (n > kSsdMaxItems && !(p.candidate_count > 0 && (p.candidate_count > p.top_n ? p.candidate_count : p.top_n) <= kSsdMaxItems))
    return fail(PRG_EUNSUPPORTED, "prg_ssd: more than 1024 candidates reach the kernel (set CandidateCount <= 1024)");
  int window = p.window_size;
  if (window <= 1) window = 5;  // ssd_sort.go:358-361
  if (window > 64) return fail(PRG_EUNSUPPORTED, "prg_ssd: window > 64");
  PRG_TRY(h->ssd_E.ensure((size_t)B * h->D_dim * kSsdMaxItems * 8));
  PRG_TRY(h->ssd_P.ensure((size_t)B * window * kSsdMaxItems * 8));
  SsdArgs a{};
  a.rows = rows_dev; a.score = score_dev; a.n = n; a.D = h->D; a.D_rows = h->D_rows; a.D_dim = (int)h->D_dim; a.p = p;
  a.Es = (double*)h->ssd_E.p; a.Ps = (double*)h->ssd_P.p; a.out_idx = out_idx; a.out_n = out_n; a.status = status;
  a.window = window;
  const size_t smem = ssd_smem_bytes(p.top_n, window);
  StageScope span(h, ST_DPP);
  if (h->D_dtype == PRG_F64) {
    PRG_CUDA(cudaFuncSetAttribute(ssd_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ssd_kernel<double><<<B, kSsdMaxItems, smem, h->stream>>>(a);
  } else {
    PRG_CUDA(cudaFuncSetAttribute(ssd_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ssd_kernel<float><<<B, kSsdMaxItems, smem, h->stream>>>(a);
  }
  PRG_CUDA(cudaGetLastError());
  count_launch(h);
  return PRG_OK;
}

}  // namespace prg
