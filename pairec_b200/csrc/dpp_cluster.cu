// dpp_cluster.cu — the fast path of DPPSort.doSort (sort/dpp_sort.go:271-351, :372-475, :477-551) for f32 diversity
// tables with dim 32 / 64 / 128: one thread-block CLUSTER of 4 CTAs per request, 256 candidates per CTA, the
// candidates' embeddings resident in SHARED MEMORY ([dim][item], conflict-free), two threads per candidate.
//
// Why: the selection loop is 50 dependent steps of "row of L for the chosen item against every candidate" — n x (D+1)
// fp64 multiply-adds per step with nothing reused between steps.  With one CTA per request (dpp.cu) every step
// re-reads the request's embeddings through L2 and stalls on that latency (ncu r1: 62 % long-scoreboard, 1.56 ms per
// 64-request batch).  Spreading a request over 4 SMs keeps the embeddings on chip; the two 64-wide k blocks of the
// gonum dot product go to two threads (their partial sums are independent by construction), so a step is one pass of
// fp64 issue over 16 warps + ONE cluster barrier.  v2 of this kernel kept the embedding in 128 registers per thread
// (0.45 ms per batch; ncu r2: 63 % of issue slots lost to instruction fetch of the fully unrolled 129-term loops and
// ~45 % of the step spent in barriers and the winner's row fetch); v3 loops over shared memory instead and ships
// the winner's embedding inside the published record, so no global load sits on the critical path.
//
// Arithmetic is exactly dpp.cu's (and oracle/oracle.c's): fp64, gonum operation order, separate multiply/add
// roundings, first-maximum argmax (rank order == index order), NaN masking, 1e-10 early stop + lowest-index fill.
// Per step each CTA publishes its local arg-max candidate TOGETHER with everything the others need about it
// (d2, 1/norm, quality, its embedding, its column of C) into every CTA's shared memory (DSMEM); the records are
// double buffered so a fast CTA never overwrites what a slow one still reads.
#include "handle.h"
#include <cooperative_groups.h>
#include <math_constants.h>

namespace cg = cooperative_groups;

namespace prg {

constexpr int kClCtas = 4;
constexpr int kClItems = 256;                      // candidates per CTA
constexpr int kClThreads = 2 * kClItems;           // two threads per candidate (k blocks split)
constexpr int kClMaxItems = kClCtas * kClItems;    // 1024
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

template <int D>
struct __align__(16) CandRecT {
  double v;       // d2 of the candidate (NaN if the CTA has none)
  double inv;     // 1 / ||e||
  double q;       // exp(alpha * rel)
  int32_t idx;    // index in the truncated list
  uint32_t row;   // diversity-table row (diagnostics)
  double cj[kClCRows];
  float x[D];     // the candidate's embedding
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

// One 64-wide k block of gonum's DotUnitary for candidate `it`: positions [k0, k0+len) of the D+1 features, four
// partial sums by position mod 4 over the full groups, the tail (always the constant feature D) into s0, then
// (s0+s2)+(s1+s3).  xs = embeddings [d][256] in shared memory; other = f_j (nullptr: the diagonal, g == f).
template <int D>
__device__ __forceinline__ double block_dot(const float* xs, int it, int k0, int len, double inv, bool do_norm,
                                            const double* other) {
  double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
  const int full = len & ~3;
  auto feat = [&](int d) -> double {
    const double xv = (double)xs[d * kClItems + it];
    return do_norm ? __dmul_rn(__dmul_rn(xv, inv), kInvSqrt2c) : __dmul_rn(xv, kInvSqrt2c);
  };
#pragma unroll 2
  for (int t = 0; t < full; t += 4) {
    const int d = k0 + t;
    const double f0 = feat(d), f1 = feat(d + 1), f2 = feat(d + 2), f3 = feat(d + 3);
    const double g0 = other ? other[d] : f0, g1 = other ? other[d + 1] : f1, g2 = other ? other[d + 2] : f2,
                 g3 = other ? other[d + 3] : f3;
    s0 = __dadd_rn(s0, __dmul_rn(g0, f0));
    s1 = __dadd_rn(s1, __dmul_rn(g1, f1));
    s2 = __dadd_rn(s2, __dmul_rn(g2, f2));
    s3 = __dadd_rn(s3, __dmul_rn(g3, f3));
  }
  for (int t = full; t < len; ++t) {  // at most one element: the constant feature (position D)
    const int d = k0 + t;
    const double f = (d == D) ? kInvSqrt2c : feat(d);
    const double g = other ? other[d] : f;
    s0 = __dadd_rn(s0, __dmul_rn(g, f));
  }
  return __dadd_rn(__dadd_rn(s0, s2), __dadd_rn(s1, s3));
}

template <int D>
__global__ void __launch_bounds__(kClThreads, 1) dpp_cluster_kernel(const DppClArgs a) {
  using CandRec = CandRecT<D>;
  constexpr int D1 = D + 1;
  constexpr int NB = (D1 + 63) / 64;  // k blocks: part 0 owns block 0, part 1 the rest
  cg::cluster_group cluster = cg::this_cluster();
  const unsigned rank = cluster.block_rank();
  const int b = blockIdx.x / kClCtas;
  extern __shared__ __align__(16) uint8_t csm[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int it = tid & (kClItems - 1), part = tid >> 8;
  const int n = a.n, T_out = a.p.top_n;
  const int window = a.p.window_size > 0 ? a.p.window_size : 10;

  double* C = reinterpret_cast<double*>(csm);                                   // [24][256]  (48 KiB, also presort staging)
  float* xs = reinterpret_cast<float*>(C + kClCRows * kClItems);                // [D][256]
  CandRec* pub = reinterpret_cast<CandRec*>(xs + D * kClItems);                 // [2][4]
  double* fj = reinterpret_cast<double*>(pub + 2 * kClCtas);                    // [D+1] padded to 136
  double* blk_s = fj + 136;                                                     // [2][256] partial block sums of part 1
  double* inv_s = blk_s + 2 * kClItems;                                         // [256]
  double* red_v = inv_s + kClItems;                                             // [8]
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

  const int gi = (int)rank * kClItems + it;  // this thread's candidate (index in the truncated list)
  const bool active = gi < m;
  const bool owner = part == 0;               // part 0 owns d2, C, the publication; part 1 only adds block sums
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

  // ---- 2. embeddings -> shared memory [d][item]; norm (gonum floats.Norm scaled form) and quality by the owner
  uint32_t my_row = 0;
  double inv = 1.0, qi = 0.0;
  if (owner) {
    my_row = active ? rows[my_in] : 0u;
    const bool have = active && (uint64_t)my_row < a.D_rows;
    if (!have) my_row = 0;
    const float4* src = reinterpret_cast<const float4*>(a.D + (size_t)my_row * D);
    double scale = 0.0, sumsq = 1.0;
#pragma unroll 4
    for (int d4 = 0; d4 < D / 4; ++d4) {
      const float4 v4 = have ? src[d4] : make_float4(0.f, 0.f, 0.f, 0.f);
      const float e4[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        xs[(4 * d4 + c) * kClItems + it] = e4[c];
        const double v = (double)e4[c];
        if (v != 0.0) {
          const double av = fabs(v);
          if (scale < av) {
            const double sc = scale / av;
            sumsq = __dadd_rn(1.0, __dmul_rn(__dmul_rn(sumsq, sc), sc));
            scale = av;
          } else {
            const double sc = av / scale;
            sumsq = __dadd_rn(sumsq, __dmul_rn(sc, sc));
          }
        }
      }
    }
    if (a.p.normalize_emb) inv = 1.0 / __dmul_rn(scale, sqrt(sumsq));
    if (active) qi = exp(__dmul_rn(a.p.alpha, rel));
    inv_s[it] = inv;
  }
  for (int i = tid; i < kClMaxItems; i += kClThreads) existed[i] = 0;
  const bool do_norm = a.p.normalize_emb != 0;
  __syncthreads();
  inv = inv_s[it];

  // S[j][i] in gonum Dgemm(NoTrans,Trans) order: block sums added to C in block order (0 + b0) + b1 + b2.
  // part 0 computes block 0, part 1 the remaining blocks -> blk_s; one barrier; the owner finishes.
  auto gram = [&](const double* other) -> double {
    double acc = 0.0;
    if (owner) {
      acc = __dadd_rn(0.0, block_dot<D>(xs, it, 0, D1 < 64 ? D1 : 64, inv, do_norm, other));
    } else {
#pragma unroll
      for (int bb = 1; bb < NB; ++bb) {
        const int k0 = bb * 64;
        blk_s[(bb - 1) * kClItems + it] = block_dot<D>(xs, it, k0, (D1 - k0 < 64) ? (D1 - k0) : 64, inv, do_norm, other);
      }
    }
    if (NB > 1) {
      __syncthreads();
      if (owner) {
#pragma unroll
        for (int bb = 1; bb < NB; ++bb) acc = __dadd_rn(acc, blk_s[(bb - 1) * kClItems + it]);
      }
    }
    return acc;
  };

  const double g0 = gram(nullptr);
  const double diag = (owner && active) ? __dmul_rn(__dmul_rn(qi, g0), qi) : CUDART_NAN;
  __syncthreads();

  // cluster-wide first-maximum argmax over the owners; every CTA ends up with the winner's record in pub[par][w]
  int par = 0;
  auto cluster_argmax = [&](double v, int krows) -> int {
    if (warp < kClItems / 32) {  // owner warps
      AmC am{v, it};
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) {
        AmC o;
        o.v = __shfl_xor_sync(0xffffffffu, am.v, off);
        o.i = __shfl_xor_sync(0xffffffffu, am.i, off);
        am = amc(am, o);
      }
      if (lane == 0) { red_v[warp] = am.v; red_i[warp] = am.i; }
    }
    __syncthreads();  // also: row k of C (written by the owners just before) is complete
    if (warp == 0) {
      AmC xx{lane < (kClItems / 32) ? red_v[lane & 7] : CUDART_NAN, lane < (kClItems / 32) ? red_i[lane & 7] : 0};
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
    // publish this CTA's candidate into every CTA's pub[par][rank]: scalars by its owner thread, its column of C by
    // threads 0..krows-1, its embedding by threads 256..256+D-1 (part 1 is otherwise idle here)
    if (owner && it == li) {
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
      const double cv = C[tid * kClItems + li];
#pragma unroll
      for (int r = 0; r < kClCtas; ++r) cluster.map_shared_rank(&pub[par * kClCtas + rank], r)->cj[tid] = cv;
    }
    if (tid >= kClItems && tid - kClItems < D) {
      const float xv = xs[(tid - kClItems) * kClItems + li];
#pragma unroll
      for (int r = 0; r < kClCtas; ++r) cluster.map_shared_rank(&pub[par * kClCtas + rank], r)->x[tid - kClItems] = xv;
    }
    cluster.sync();
    AmC best{pub[par * kClCtas].v, 0};
#pragma unroll
    for (int r = 1; r < kClCtas; ++r) {
      const double rv = pub[par * kClCtas + r].v;
      if (!isnan(rv) && (isnan(best.v) || rv > best.v)) { best.v = rv; best.i = r; }  // lower rank == lower index wins ties
    }
    const int used = par * kClCtas + best.i;
    par ^= 1;
    return used;
  };

  // ---- 3. DPPWithWindow (:477-491) over DPP (:493-551)
  int total = 0;
  const int n_calls = (T_out <= window) ? 1 : (T_out / window + (T_out % window > 0 ? 1 : 0));
  for (int call = 0; call < n_calls; ++call) {
    int top = (T_out <= window) ? T_out : ((call < T_out / window) ? window : T_out % window);
    if (top > m) top = m;
    double d2 = (owner && active && !existed[gi]) ? diag : CUDART_NAN;
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
      if (tid < D) {
        const double xv = (double)W.x[tid];
        fj[tid] = do_norm ? __dmul_rn(__dmul_rn(xv, inv_j), kInvSqrt2c) : __dmul_rn(xv, kInvSqrt2c);
      }
      if (tid == D) fj[D] = kInvSqrt2c;
      __syncthreads();
      const double S = gram(fj);
      if (owner) {
        if (active) {
          const double Lji = __dmul_rn(__dmul_rn(q_j, S), qi);
          double e;
          if (k == 0) {
            e = __dmul_rn(inv_dj, Lji);
          } else {
            double ss = 0.0;
            for (int l = 0; l < k; ++l) {
              const double tmp = W.cj[l];
              if (tmp != 0) ss = __dadd_rn(ss, __dmul_rn(tmp, C[l * kClItems + it]));
            }
            e = __dmul_rn(inv_dj, __dsub_rn(Lji, ss));
          }
          C[k * kClItems + it] = e;
          d2 = __dsub_rn(d2, __dmul_rn(e, e));
        }
        if (gi == j) d2 = CUDART_NAN;
      }
      wrec = cluster_argmax(d2, ny);  // its first barrier also orders the C[k] writes before the column reads
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

template <int D>
static size_t dpp_cluster_smem(int top_n) {
  return (size_t)kClCRows * kClItems * 8 + (size_t)D * kClItems * 4 + 2 * kClCtas * sizeof(CandRecT<D>) + 136 * 8 +
         2 * kClItems * 8 + kClItems * 8 + 8 * 8 + 8 * 4 + kClMaxItems * 4 + (size_t)((top_n + 3) & ~3) * 4 + kClMaxItems + 64;
}

template <int D>
static int launch_cluster(prg_handle* h, const DppClArgs& a, int B) {
  const size_t smem = dpp_cluster_smem<D>(a.p.top_n);
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
