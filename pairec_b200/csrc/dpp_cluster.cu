// dpp_cluster.cu — the fast path of DPPSort.doSort (sort/dpp_sort.go:271-351, :372-475, :477-551) for f32 diversity
// tables with dim 32 / 64 / 128: one thread-block CLUSTER of 4 CTAs per request, one thread per candidate, the
// candidate's embedding held in REGISTERS.
//
// Why: the selection loop is 50 dependent steps of "row of L for the chosen item against every candidate" — n x (D+1)
// fp64 multiply-adds per step with nothing reused between steps.  With one CTA per request (dpp.cu) every step
// re-reads the request's embeddings through L2 and stalls on that latency (ncu: 62 % long-scoreboard, 1.56 ms per
// 64-request batch).  Spreading a request over 4 SMs lets 1024 candidates keep their 128 floats in registers, so a
// step is pure fp64 issue + one cluster barrier; all 148 SMs work on 37 requests at a time instead of 64 SMs on 64.
//
// Arithmetic is exactly dpp.cu's (and oracle/oracle.c's): fp64, gonum operation order, separate multiply/add
// roundings, first-maximum argmax (rank order == index order), NaN masking, 1e-10 early stop + lowest-index fill.
// Per step each CTA publishes its local arg-max candidate TOGETHER with everything the others need about it
// (d2, 1/norm, quality, table row, its column of C) into every CTA's shared memory (DSMEM), so one cluster barrier
// per step suffices; the records are double buffered so a fast CTA never overwrites what a slow one still reads.
#include "handle.h"
#include <cooperative_groups.h>
#include <math_constants.h>

namespace cg = cooperative_groups;

namespace prg {

constexpr int kClCtas = 4;
constexpr int kClThreads = 256;
constexpr int kClMaxItems = kClCtas * kClThreads;  // 1024
constexpr int kClMaxN = 4096;
constexpr int kClCRows = 24;
constexpr double kInvSqrt2c = 0.70710678118654752440;

struct DppClArgs {
  const uint32_t* rows;
  const double* score;
  int n;
  const float* D;
  uint64_t D_rows;
  prg_dpp_params p;
  int32_t* out_idx;
  int32_t* out_n;
  int32_t* status;
};

struct __align__(16) CandRec {
  double v;       // d2 of the candidate (NaN if the CTA has none)
  double inv;     // 1 / ||e||
  double q;       // exp(alpha * rel)
  int32_t idx;    // index in the truncated list
  uint32_t row;   // diversity-table row
  double cj[kClCRows];
};

__device__ __forceinline__ uint64_t f64_ord_c(double d) {
  uint64_t u = (uint64_t)__double_as_longlong(d);
  if ((u & 0x7FFFFFFFFFFFFFFFull) > 0x7FF0000000000000ull) return 0ull;
  return (u >> 63) ? ~u : (u | 0x8000000000000000ull);
}
struct AmC { double v; int i; };
__device__ __forceinline__ AmC amc(AmC a, AmC b) {
  if (isnan(b.v)) return a;
  if (isnan(a.v)) return b;
  if (b.v > a.v || (b.v == a.v && b.i < a.i)) return b;
  return a;
}

// S[j][i] in gonum Dgemm(NoTrans,Trans) order over D+1 features; x in registers, `other` = f_j in shared memory or
// nullptr for the diagonal.  Feature d < D is (x*inv)*c (or x*c), feature D is c.
template <int D>
__device__ __forceinline__ double gram_reg(const float (&x)[D], double inv, bool do_norm, const double* other) {
  // (interleaving the 64-wide blocks for more ILP was tried and measured slower: 0.507 vs 0.449 ms per batch)
  constexpr int D1 = D + 1;
  double acc = 0.0;
#pragma unroll
  for (int k0 = 0; k0 < D1; k0 += 64) {
    const int len = (D1 - k0 < 64) ? (D1 - k0) : 64;
    double s[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
    for (int t = 0; t < 64; ++t) {
      if (t < len) {
        const int d = k0 + t;
        double f;
        if (d == D) f = kInvSqrt2c;
        else {
          const double xv = (double)x[d < D ? d : 0];
          f = do_norm ? __dmul_rn(__dmul_rn(xv, inv), kInvSqrt2c) : __dmul_rn(xv, kInvSqrt2c);
        }
        const double g = other ? other[d] : f;
        const int lane4 = (t < (len & ~3)) ? (t & 3) : 0;  // full groups of 4 -> 4 partial sums, tail -> s0
        s[lane4] = __dadd_rn(s[lane4], __dmul_rn(g, f));
      }
    }
    acc = __dadd_rn(acc, __dadd_rn(__dadd_rn(s[0], s[2]), __dadd_rn(s[1], s[3])));
  }
  return acc;
}

template <int D>
__global__ void __launch_bounds__(kClThreads, 1) dpp_cluster_kernel(const DppClArgs a) {
  cg::cluster_group cluster = cg::this_cluster();
  const unsigned rank = cluster.block_rank();
  const int b = blockIdx.x / kClCtas;
  extern __shared__ __align__(16) uint8_t csm[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n = a.n, T_out = a.p.top_n;
  const int window = a.p.window_size > 0 ? a.p.window_size : 10;

  double* C = reinterpret_cast<double*>(csm);                                   // [24][256]  (48 KiB, also presort staging)
  CandRec* pub = reinterpret_cast<CandRec*>(C + kClCRows * kClThreads);         // [2][4]
  double* fj = reinterpret_cast<double*>(pub + 2 * kClCtas);                    // [D+1] padded to 136
  double* red_v = fj + 136;                                                     // [8]
  int32_t* red_i = reinterpret_cast<int32_t*>(red_v + 8);                       // [8]
  int32_t* order = red_i + 8;                                                   // [1024]
  int32_t* res = order + kClMaxItems;                                           // [T_out]
  uint8_t* existed = reinterpret_cast<uint8_t*>(res + ((T_out + 3) & ~3));      // [1024]
  __shared__ int s_m, s_err, s_ny, s_li;
  __shared__ double s_p0, s_p1;

  const uint32_t* rows = a.rows + (size_t)b * n;
  const double* score = a.score + (size_t)b * n;

  // ---- 0. valid count, optional presort + truncation (:280-300); done redundantly by every CTA of the cluster
  if (tid == 0) { s_m = 0; s_err = 0; }
  __syncthreads();
  {
    int cnt = 0;
    for (int i = tid; i < n; i += kClThreads) cnt += (rows[i] != 0xFFFFFFFFu);
    if (cnt) atomicAdd(&s_m, cnt);
  }
  __syncthreads();
  const int nv = s_m;
  __syncthreads();
  int m = nv;
  const bool presort = (a.p.candidate_count > 0 || a.p.min_score_percent > 0) && nv > T_out;
  if (nv > 0 && presort) {
    uint32_t P2 = 32;
    while (P2 < (uint32_t)nv) P2 <<= 1;
    uint64_t* key = reinterpret_cast<uint64_t*>(csm);
    int32_t* idx = reinterpret_cast<int32_t*>(key + P2);
    for (uint32_t i = tid; i < P2; i += kClThreads) {
      key[i] = (i < (uint32_t)nv) ? f64_ord_c(score[i]) : 0ull;
      idx[i] = (i < (uint32_t)nv) ? (int32_t)i : 0x7FFFFFFF;
    }
    __syncthreads();
    for (uint32_t size = 2; size <= P2; size <<= 1) {
      for (uint32_t stride = size >> 1; stride > 0; stride >>= 1) {
        for (uint32_t i = tid; i < (P2 >> 1); i += kClThreads) {
          const uint32_t pos = 2 * i - (i & (stride - 1));
          const uint64_t ka = key[pos], kb = key[pos + stride];
          const int32_t ia = idx[pos], ib = idx[pos + stride];
          const bool a_after_b = (ka < kb) || (ka == kb && ia > ib);
          if (a_after_b == ((pos & size) == 0)) { key[pos] = kb; key[pos + stride] = ka; idx[pos] = ib; idx[pos + stride] = ia; }
        }
        __syncthreads();
      }
    }
    if (tid == 0) {
      int mm = nv;
      if (a.p.candidate_count > 0) {
        const int cnt = T_out > a.p.candidate_count ? T_out : a.p.candidate_count;
        if (cnt < mm) mm = cnt;
      }
      if (a.p.min_score_percent > 0 && mm > T_out) {
        int id = T_out;
        const double mx = score[idx[0]];
        for (; id < mm; ++id)
          if (score[idx[id]] / mx < a.p.min_score_percent) break;
        mm = id;
      }
      s_m = mm;
    }
    __syncthreads();
    m = s_m;
    if (m <= kClMaxItems)
      for (int i = tid; i < m; i += kClThreads) order[i] = idx[i];
    __syncthreads();
  } else if (m <= kClMaxItems) {
    for (int i = tid; i < m; i += kClThreads) order[i] = i;
  }
  if (nv == 0 || m > kClMaxItems) {  // uniform across the cluster: every CTA sees the same request
    if (rank == 0 && tid == 0) { a.out_n[b] = 0; a.status[b] = (nv == 0) ? 0 : 2; }
    return;
  }
  __syncthreads();

  const int gi = (int)rank * kClThreads + tid;  // this thread's candidate (index in the truncated list)
  const bool active = gi < m;
  const int my_in = active ? order[gi] : 0;

  // ---- 1. relevance + abtest normalisation modes (:382-405), redundantly per CTA
  double rel = active ? score[my_in] : 0.0;
  if (a.p.norm_mode == 1 || a.p.norm_mode == 2) {
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
        const double span = __dsub_rn(r0, r1);
        if (span == 0) s_err = 1;
        s_p0 = r1;
        s_p1 = span;
      }
    }
    __syncthreads();
    if (a.p.norm_mode == 1) rel = __dsub_rn(rel, s_p0) / s_p1;
    else rel = __dadd_rn(__dmul_rn(__dsub_rn(rel, s_p0) / s_p1, 1 - 1e-6), 1e-6);
  }
  __syncthreads();
  if (s_err) {
    if (rank == 0 && tid == 0) { a.out_n[b] = 0; a.status[b] = 1; }
    return;
  }

  // ---- 2. the candidate's embedding -> registers; norm (gonum floats.Norm scaled form), quality
  float x[D];
  uint32_t my_row = 0;
  double inv = 1.0, qi = 0.0;
  {
    my_row = active ? rows[my_in] : 0u;
    const bool have = active && (uint64_t)my_row < a.D_rows;
    if (!have) my_row = 0;
    const float4* src = reinterpret_cast<const float4*>(a.D + (size_t)my_row * D);
#pragma unroll
    for (int d4 = 0; d4 < D / 4; ++d4) {
      const float4 v = have ? src[d4] : make_float4(0.f, 0.f, 0.f, 0.f);
      x[4 * d4] = v.x; x[4 * d4 + 1] = v.y; x[4 * d4 + 2] = v.z; x[4 * d4 + 3] = v.w;
    }
    double scale = 0.0, sumsq = 1.0;
#pragma unroll
    for (int d = 0; d < D; ++d) {
      const double v = (double)x[d];
      if (v != 0.0) {
        const double av = fabs(v);
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
    if (a.p.normalize_emb) inv = 1.0 / __dmul_rn(scale, sqrt(sumsq));
    if (active) qi = exp(__dmul_rn(a.p.alpha, rel));
  }
  for (int i = tid; i < kClMaxItems; i += kClThreads) existed[i] = 0;
  const bool do_norm = a.p.normalize_emb != 0;
  const double diag = active ? __dmul_rn(__dmul_rn(qi, gram_reg<D>(x, inv, do_norm, nullptr)), qi) : CUDART_NAN;
  __syncthreads();

  // cluster-wide first-maximum argmax; every CTA ends up with the winner's record in pub[par][w]
  int par = 0;
  auto cluster_argmax = [&](double v, int krows) -> int {
    AmC am{v, tid};
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      AmC o;
      o.v = __shfl_xor_sync(0xffffffffu, am.v, off);
      o.i = __shfl_xor_sync(0xffffffffu, am.i, off);
      am = amc(am, o);
    }
    if (lane == 0) { red_v[warp] = am.v; red_i[warp] = am.i; }
    __syncthreads();
    if (warp == 0) {
      AmC xx{lane < (kClThreads / 32) ? red_v[lane] : CUDART_NAN, lane < (kClThreads / 32) ? red_i[lane] : 0};
#pragma unroll
      for (int off = 4; off > 0; off >>= 1) {
        AmC o;
        o.v = __shfl_xor_sync(0xffffffffu, xx.v, off);
        o.i = __shfl_xor_sync(0xffffffffu, xx.i, off);
        xx = amc(xx, o);
      }
      if (lane == 0) s_li = isnan(xx.v) ? 0 : xx.i;
    }
    __syncthreads();
    const int li = s_li;
    // publish this CTA's candidate into every CTA's pub[par][rank]
    if (tid == li) {
#pragma unroll
      for (int r = 0; r < kClCtas; ++r) {
        CandRec* dst = cluster.map_shared_rank(&pub[par * kClCtas + rank], r);
        dst->v = v;
        dst->inv = inv;
        dst->q = qi;
        dst->idx = gi;
        dst->row = my_row;
      }
    }
    if (tid < krows) {
      const double cv = C[tid * kClThreads + li];
#pragma unroll
      for (int r = 0; r < kClCtas; ++r) cluster.map_shared_rank(&pub[par * kClCtas + rank], r)->cj[tid] = cv;
    }
    cluster.sync();
    int w = 0;
    AmC best{pub[par * kClCtas].v, 0};
#pragma unroll
    for (int r = 1; r < kClCtas; ++r) {
      const double rv = pub[par * kClCtas + r].v;
      if (!isnan(rv) && (isnan(best.v) || rv > best.v)) { best.v = rv; best.i = r; }  // lower rank == lower index wins ties
    }
    w = best.i;
    const int used = par * kClCtas + w;
    par ^= 1;
    return used;
  };

  // ---- 3. DPPWithWindow (:477-491) over DPP (:493-551)
  int total = 0;
  const int n_calls = (T_out <= window) ? 1 : (T_out / window + (T_out % window > 0 ? 1 : 0));
  for (int call = 0; call < n_calls; ++call) {
    int top = (T_out <= window) ? T_out : ((call < T_out / window) ? window : T_out % window);
    if (top > m) top = m;
    double d2 = (active && !existed[gi]) ? diag : CUDART_NAN;
    int wrec = cluster_argmax(d2, 0);
    int j = isnan(pub[wrec].v) ? 0 : pub[wrec].idx;
    if (tid == 0) res[total] = j;
    int ny = 1;
    bool broke = false;
    while (ny < top) {
      const CandRec& W = pub[wrec];
      double dj = W.v;  // == d2[j]; NaN when every candidate is used up (the reference then repeats index 0)
      if (dj < 1e-10) { broke = true; break; }
      dj = sqrt(dj);
      const double inv_dj = 1.0 / dj;
      const int k = ny - 1;
      const double inv_j = W.inv, q_j = W.q;
      {
        const float* rj = a.D + (size_t)W.row * D;
        if (tid < D) {
          const double xv = (double)rj[tid];
          fj[tid] = do_norm ? __dmul_rn(__dmul_rn(xv, inv_j), kInvSqrt2c) : __dmul_rn(xv, kInvSqrt2c);
        }
        if (tid == D % kClThreads && D < kClThreads) fj[D] = kInvSqrt2c;
        if (D >= kClThreads && tid == 0) fj[D] = kInvSqrt2c;
      }
      __syncthreads();
      double e = CUDART_NAN;
      if (active) {
        const double Lji = __dmul_rn(__dmul_rn(q_j, gram_reg<D>(x, inv, do_norm, fj)), qi);
        if (k == 0) {
          e = __dmul_rn(inv_dj, Lji);
        } else {
          double ss = 0.0;
          for (int l = 0; l < k; ++l) {
            const double tmp = W.cj[l];
            if (tmp != 0) ss = __dadd_rn(ss, __dmul_rn(tmp, C[l * kClThreads + tid]));
          }
          e = __dmul_rn(inv_dj, __dsub_rn(Lji, ss));
        }
        C[k * kClThreads + tid] = e;
        d2 = __dsub_rn(d2, __dmul_rn(e, e));
      }
      if (gi == j) d2 = CUDART_NAN;
      __syncthreads();  // C row k complete before the next candidate's column is read; fj free for reuse
      wrec = cluster_argmax(d2, ny);
      j = isnan(pub[wrec].v) ? 0 : pub[wrec].idx;
      if (tid == 0) res[total + ny] = j;
      ++ny;
    }
    __syncthreads();
    if (broke && ny < top) {  // :539-548 lowest unused indices (identical in every CTA)
      if (tid == 0) {
        int c = ny;
        for (int i = 0; i < m && c < top; ++i) {
          if (existed[i]) continue;
          bool in_y = false;
          for (int t = 0; t < c; ++t) in_y |= (res[total + t] == i);
          if (!in_y) res[total + c++] = i;
        }
        s_ny = c;
      }
      __syncthreads();
      ny = s_ny;
    }
    __syncthreads();
    if (tid < ny) existed[res[total + tid]] = 1;
    total += ny;
    __syncthreads();
  }
  if (rank == 0) {
    for (int t = tid; t < total; t += kClThreads) a.out_idx[(size_t)b * T_out + t] = order[res[t]];
    if (tid == 0) { a.out_n[b] = total; a.status[b] = 0; }
  }
  cluster.sync();  // no CTA may exit while peers can still write into its shared memory
}

static size_t dpp_cluster_smem(int top_n) {
  return (size_t)kClCRows * kClThreads * 8 + 2 * kClCtas * sizeof(CandRec) + 136 * 8 + 8 * 8 + 8 * 4 + kClMaxItems * 4 +
         (size_t)((top_n + 3) & ~3) * 4 + kClMaxItems + 64;
}

template <int D>
static int launch_cluster(prg_handle* h, const DppClArgs& a, int B) {
  const size_t smem = dpp_cluster_smem(a.p.top_n);
  PRG_CUDA(cudaFuncSetAttribute(dpp_cluster_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)(B * kClCtas));
  cfg.blockDim = dim3(kClThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = h->stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = kClCtas;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  PRG_CUDA(cudaLaunchKernelEx(&cfg, dpp_cluster_kernel<D>, a));
  count_launch(h);
  return PRG_OK;
}

// returns PRG_OK and sets *handled when the request shape is served by the cluster kernel
int dpp_cluster_device(prg_handle* h, const uint32_t* rows_dev, const double* score_dev, int B, int n,
                       const prg_dpp_params& p, int32_t* out_idx, int32_t* out_n, int32_t* status, bool* handled) {
  *handled = false;
  if (h->D_dtype != PRG_F32) return PRG_OK;
  if (h->D_dim != 32 && h->D_dim != 64 && h->D_dim != 128) return PRG_OK;
  if ((reinterpret_cast<uintptr_t>(h->D) & 15) != 0) return PRG_OK;
  if (n > kClMaxN || p.top_n > 2048) return PRG_OK;
  const int window = p.window_size > 0 ? p.window_size : 10;
  const int c_rows = p.top_n <= window ? p.top_n : window;
  if (c_rows > kClCRows) return PRG_OK;
  DppClArgs a{};
  a.rows = rows_dev; a.score = score_dev; a.n = n; a.D = (const float*)h->D; a.D_rows = h->D_rows; a.p = p;
  a.out_idx = out_idx; a.out_n = out_n; a.status = status;
  StageScope span(h, ST_DPP);
  int rc = PRG_OK;
  if (h->D_dim == 32) rc = launch_cluster<32>(h, a, B);
  else if (h->D_dim == 64) rc = launch_cluster<64>(h, a, B);
  else rc = launch_cluster<128>(h, a, B);
  if (rc == PRG_OK) *handled = true;
  return rc;
}

}  // namespace prg
