// dpp.cu — DPPSort.doSort on the GPU (sort/dpp_sort.go:271-351, :372-475, :477-551).
//
// One CTA per request, one thread per candidate (<= 1024 after truncation).  The reference materialises
// S = F*F^T and L = diag(r)*S*diag(r) densely (0.26 + 4 GFLOP of fp64 Dgemm per request); only diag(L) and the rows
// of the selected items are ever read, so rows are produced on demand: 12.9 MFLOP per request.  All arithmetic is
// fp64 with the reference's operation order (gonum Dgemm/DotUnitary/AxpyUnitary/MaxIdx as restated in
// oracle/oracle.c): separate multiply and add roundings (no FMA contraction), 64-wide k blocks with four partial
// sums, reciprocal-multiply scaling, NaN masking, first-maximum argmax, 1e-10 early stop + lowest-index fill.
//
// Embeddings are staged once per request into a transposed scratch Et[d][i] (coalesced per-dim reads in the
// selection loop); normalisation (1/||e||, 1/sqrt2) is applied on the fly so the scratch stays at the table's width.
#include "dpp_common.cuh"

namespace prg {

constexpr int kDppMaxItems = 1024;  // candidates per request after truncation (one thread each)
constexpr int kDppMaxN = 4096;      // candidates per request before truncation (presort in shared memory)
constexpr double kInvSqrt2 = 0.70710678118654752440;

struct DppArgs {
  const uint32_t* rows;  // [B][n]
  const double* score;   // [B][n]
  int n;
  const void* D;
  uint64_t D_rows;
  int D_dim;
  prg_dpp_params p;
  void* Et;  // scratch [B][D_dim][kDppMaxItems] of T
  int32_t* out_idx;  // [B][top_n]
  int32_t* out_n;    // [B]
  int32_t* status;   // [B]
  int c_rows;        // rows of C held in shared memory
  const float* D_sub;  // substitute directions for candidates without a table row (dpp_common.cuh)
  int force_norm;      // hook + table path: the concatenated vector is always re-normalised (dpp_sort.go:418-421)
  int no_pos;          // EnsurePositiveSim == "false" (hook-only path, :440-445): append 0, no 1/sqrt2 scaling
};

__device__ __forceinline__ uint64_t f64_ord_dev(double d) {
  uint64_t u = (uint64_t)__double_as_longlong(d);
  if ((u & 0x7FFFFFFFFFFFFFFFull) > 0x7FF0000000000000ull) return 0ull;
  return (u >> 63) ? ~u : (u | 0x8000000000000000ull);
}

struct ArgMax { double v; int i; };
__device__ __forceinline__ ArgMax am_combine(ArgMax a, ArgMax b) {
  // floats.MaxIdx: NaN skipped, first (lowest index) maximum wins
  if (isnan(b.v)) return a;
  if (isnan(a.v)) return b;
  if (b.v > a.v || (b.v == a.v && b.i < a.i)) return b;
  return a;
}

template <typename T>
__global__ void __launch_bounds__(kDppMaxItems, 1) dpp_kernel(const DppArgs a) {
  extern __shared__ __align__(16) uint8_t dsm[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int b = blockIdx.x;
  const int n = a.n, D = a.D_dim, D1 = D + 1;
  const int T_out = a.p.top_n;
  const int window = a.p.window_size > 0 ? a.p.window_size : 10;

  // shared layout
  double* C = reinterpret_cast<double*>(dsm);                   // [c_rows][1024]
  double* inv_s = C + (size_t)a.c_rows * kDppMaxItems;          // [1024]
  double* q_s = inv_s + kDppMaxItems;                           // [1024]
  double* fj = q_s + kDppMaxItems;                              // [D1] (padded to 520)
  double* red_v = fj + 520;                                     // [32]
  int32_t* red_i = reinterpret_cast<int32_t*>(red_v + 32);      // [32]
  int32_t* order = red_i + 32;                                  // [1024] truncated list -> input index
  int32_t* res = order + kDppMaxItems;                          // [T_out] selected (indices into truncated list)
  uint8_t* existed = reinterpret_cast<uint8_t*>(res + ((T_out + 3) & ~3));  // [1024]
  __shared__ double s_dj;
  __shared__ int s_j, s_m, s_err, s_ny;

  const uint32_t* rows = a.rows + (size_t)b * n;
  const double* score = a.score + (size_t)b * n;

  // ---- 0. valid count (padding rows 0xFFFFFFFF sit at the tail), optional presort + truncation (:280-300)
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
  int m = nv;
  if (nv == 0) {
    if (tid == 0) { a.out_n[b] = 0; a.status[b] = 0; }
    return;
  }
  const bool presort = (a.p.candidate_count > 0 || a.p.min_score_percent > 0) && nv > T_out;
  if (presort) {
    // stable descending sort of the nv candidates, keys staged in the C region (n <= 4096 -> 48 KiB)
    uint32_t P2 = 32;
    while (P2 < (uint32_t)nv) P2 <<= 1;
    uint64_t* key = reinterpret_cast<uint64_t*>(dsm);
    int32_t* idx = reinterpret_cast<int32_t*>(key + P2);
    for (uint32_t i = tid; i < P2; i += blockDim.x) {
      key[i] = (i < (uint32_t)nv) ? f64_ord_dev(score[i]) : 0ull;
      idx[i] = (i < (uint32_t)nv) ? (int32_t)i : 0x7FFFFFFF;
    }
    __syncthreads();
    for (uint32_t size = 2; size <= P2; size <<= 1) {
      for (uint32_t stride = size >> 1; stride > 0; stride >>= 1) {
        for (uint32_t i = tid; i < (P2 >> 1); i += blockDim.x) {
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
    if (m <= kDppMaxItems)
      for (int i = tid; i < m; i += blockDim.x) order[i] = idx[i];
    __syncthreads();
  } else {
    if (m <= kDppMaxItems)
      for (int i = tid; i < m; i += blockDim.x) order[i] = i;
  }
  if (m > kDppMaxItems) {  // host checks this too; never silently wrong
    if (tid == 0) { a.out_n[b] = 0; a.status[b] = 2; }
    return;
  }
  __syncthreads();

  const bool active = tid < m;
  const int my_in = active ? order[tid] : 0;

  // ---- 1. relevance (+ abtest normalisation modes, :382-405)
  double rel = active ? score[my_in] : 0.0;
  if (a.p.norm_mode == 1 || a.p.norm_mode == 2) {
    double* rel_s = C;  // staging, C is not live yet
    if (active) rel_s[tid] = rel;
    __syncthreads();
    if (tid == 0) {
      if (a.p.norm_mode == 1) {
        double sum = 0.0;
        for (int i = 0; i < m; ++i) sum = __dadd_rn(sum, rel_s[i]);
        const double mean = sum / (double)m;
        double ssq = 0.0, comp = 0.0;
        for (int i = 0; i < m; ++i) {
          const double d = __dsub_rn(rel_s[i], mean);
          ssq = __dadd_rn(ssq, __dmul_rn(d, d));
          comp = __dadd_rn(comp, d);
        }
        const double var = __dsub_rn(ssq, __dmul_rn(comp, comp) / (double)m) / (double)m;
        if (mean == 0 || var == 0) s_err = 1;
        rel_s[kDppMaxItems] = mean;
        rel_s[kDppMaxItems + 1] = sqrt(var);
      } else {
        const double span = __dsub_rn(rel_s[0], rel_s[m - 1]);
        if (span == 0) s_err = 1;
        rel_s[kDppMaxItems] = rel_s[m - 1];
        rel_s[kDppMaxItems + 1] = span;
      }
    }
    __syncthreads();
    const double p0 = rel_s[kDppMaxItems], p1 = rel_s[kDppMaxItems + 1];
    if (a.p.norm_mode == 1) rel = __dsub_rn(rel, p0) / p1;
    else rel = __dadd_rn(__dmul_rn(__dsub_rn(rel, p0) / p1, 1 - 1e-6), 1e-6);
    __syncthreads();
  }
  if (s_err) {
    if (tid == 0) { a.out_n[b] = 0; a.status[b] = 1; }
    return;
  }

  // ---- 2. stage embeddings transposed, norms, quality terms
  T* Et = reinterpret_cast<T*>(a.Et) + (size_t)b * D * kDppMaxItems;
  double inv = 1.0;
  if (active) {
    const uint32_t row = rows[my_in];
    const bool have = (uint64_t)row < a.D_rows;
    const T* src = reinterpret_cast<const T*>(a.D) + (size_t)(have ? row : 0) * D;
    const float* sub = a.D_sub + (size_t)((uint32_t)my_in & (kDppSubRows - 1)) * D;   // no table row: substitute direction
    // gonum floats.Norm(v, 2): scaled sum of squares, sequential
    double scale = 0.0, sumsq = 1.0;
    for (int d = 0; d < D; ++d) {
      const T xv = have ? src[d] : (T)sub[d];
      Et[(size_t)d * kDppMaxItems + tid] = xv;
      const double v = (double)xv;
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
    if (a.p.normalize_emb || a.force_norm) inv = 1.0 / __dmul_rn(scale, sqrt(sumsq));
    inv_s[tid] = inv;
    q_s[tid] = exp(__dmul_rn(a.p.alpha, rel));
  } else {
    inv_s[tid] = 1.0;
    q_s[tid] = 0.0;
  }
  existed[tid] = 0;
  __syncthreads();
  const double qi = q_s[tid];
  const bool do_norm = a.p.normalize_emb != 0 || a.force_norm != 0;
  const bool pos = a.no_pos == 0;

  // f_i[d] as the reference rounds it: (x * inv) * (1/sqrt2), or x * (1/sqrt2) without normalisation; f_i[D] = 1/sqrt2.
  // EnsurePositiveSim == false: no 1/sqrt2 scaling and f_i[D] = 0 (:440-445)
  auto feat = [&](int d, int i, double invn) -> double {
    if (d == D) return pos ? kInvSqrt2 : 0.0;
    const double x = (double)Et[(size_t)d * kDppMaxItems + i];
    const double t = do_norm ? __dmul_rn(x, invn) : x;
    return pos ? __dmul_rn(t, kInvSqrt2) : t;
  };
  // S[j][i] in gonum Dgemm(NoTrans,Trans) order: 64-wide k blocks, DotUnitary = 4 partial sums, (s0+s2)+(s1+s3).
  // other[] holds f_j (shared memory) or nullptr for the diagonal (f_j == f_i).
  // At most 64 items: dgemmParallel (blas/gonum/dgemm.go) sees fewer than minParBlock = 4 blocks of C and calls dgemmSerial
  // on the whole product — ONE DotUnitary over all of k, no k blocks (oracle.c g_gemm_serial).
  const int kblk = (m <= 64) ? D1 : 64;
  auto gram = [&](const double* other) -> double {
    double acc = 0.0;
    for (int k0 = 0; k0 < D1; k0 += kblk) {
      const int len = (D1 - k0 < kblk) ? (D1 - k0) : kblk;
      double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
      int t = 0;
      for (; t + 4 <= len; t += 4) {
        const int d = k0 + t;
        const double f0 = feat(d, tid, inv), f1 = feat(d + 1, tid, inv), f2 = feat(d + 2, tid, inv), f3 = feat(d + 3, tid, inv);
        const double g0 = other ? other[d] : f0, g1 = other ? other[d + 1] : f1, g2 = other ? other[d + 2] : f2,
                     g3 = other ? other[d + 3] : f3;
        s0 = __dadd_rn(s0, __dmul_rn(g0, f0));
        s1 = __dadd_rn(s1, __dmul_rn(g1, f1));
        s2 = __dadd_rn(s2, __dmul_rn(g2, f2));
        s3 = __dadd_rn(s3, __dmul_rn(g3, f3));
      }
      for (; t < len; ++t) {
        const int d = k0 + t;
        const double f0 = feat(d, tid, inv);
        const double g0 = other ? other[d] : f0;
        s0 = __dadd_rn(s0, __dmul_rn(g0, f0));
      }
      acc = __dadd_rn(acc, __dadd_rn(__dadd_rn(s0, s2), __dadd_rn(s1, s3)));
    }
    return acc;
  };

  const double diag = active ? __dmul_rn(__dmul_rn(qi, gram(nullptr)), qi) : CUDART_NAN;

  auto argmax = [&](double v) -> int {
    ArgMax am{v, tid};
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      ArgMax o;
      o.v = __shfl_xor_sync(0xffffffffu, am.v, off);
      o.i = __shfl_xor_sync(0xffffffffu, am.i, off);
      am = am_combine(am, o);
    }
    if (lane == 0) { red_v[warp] = am.v; red_i[warp] = am.i; }
    __syncthreads();
    if (warp == 0) {
      ArgMax x{red_v[lane], red_i[lane]};
      if (lane >= (int)(blockDim.x >> 5)) x.v = CUDART_NAN;
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) {
        ArgMax o;
        o.v = __shfl_xor_sync(0xffffffffu, x.v, off);
        o.i = __shfl_xor_sync(0xffffffffu, x.i, off);
        x = am_combine(x, o);
      }
      if (lane == 0) s_j = isnan(x.v) ? 0 : x.i;
    }
    __syncthreads();
    return s_j;
  };

  // ---- 3. DPPWithWindow (:477-491) over DPP (:493-551)
  int total = 0;
  const int n_calls = (T_out <= window) ? 1 : (T_out / window + (T_out % window > 0 ? 1 : 0));
  for (int call = 0; call < n_calls; ++call) {
    int top = (T_out <= window) ? T_out : ((call < T_out / window) ? window : T_out % window);
    if (top > m) top = m;
    double d2 = (active && !existed[tid]) ? diag : CUDART_NAN;
    int j = argmax(d2);
    if (tid == 0) res[total] = j;
    int ny = 1;
    bool broke = false;
    while (ny < top) {
      if (tid == j) s_dj = d2;
      __syncthreads();
      double dj = s_dj;
      if (dj < 1e-10) { broke = true; break; }
      dj = sqrt(dj);
      const double inv_dj = 1.0 / dj;
      const int k = ny - 1;
      if (tid < D1) fj[tid] = feat(tid, j, inv_s[j]);
      if (blockDim.x < (unsigned)D1)
        for (int d = tid + blockDim.x; d < D1; d += blockDim.x) fj[d] = feat(d, j, inv_s[j]);
      __syncthreads();
      double e = CUDART_NAN;
      if (active) {
        const double Lji = __dmul_rn(__dmul_rn(q_s[j], gram(fj)), qi);
        if (k == 0) {
          e = __dmul_rn(inv_dj, Lji);
        } else {
          double ss = 0.0;
          for (int l = 0; l < k; ++l) {
            const double tmp = C[(size_t)l * kDppMaxItems + j];
            if (tmp != 0) ss = __dadd_rn(ss, __dmul_rn(tmp, C[(size_t)l * kDppMaxItems + tid]));
          }
          e = __dmul_rn(inv_dj, __dsub_rn(Lji, ss));
        }
      }
      __syncthreads();  // every thread has read column j of C before row k is written
      if (active) {
        C[(size_t)k * kDppMaxItems + tid] = e;
        d2 = __dsub_rn(d2, __dmul_rn(e, e));
      }
      if (tid == j) d2 = CUDART_NAN;
      j = argmax(d2);
      if (tid == 0) res[total + ny] = j;
      ++ny;
    }
    __syncthreads();
    if (broke && ny < top) {  // :539-548 lowest unused indices
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
    // mark this call's picks as existed for the following windows
    if (tid < ny) existed[res[total + tid]] = 1;
    total += ny;
    __syncthreads();
  }
  for (int t = tid; t < total; t += blockDim.x) a.out_idx[(size_t)b * T_out + t] = order[res[t]];
  if (tid == 0) { a.out_n[b] = total; a.status[b] = 0; }
}

int dpp_lazy_device(prg_handle* h, const uint32_t* rows_dev, const double* score_dev, int B, int n, const prg_dpp_params& p,
                    int32_t* out_idx, int32_t* out_n, int32_t* status, bool* handled);
int dpp_cluster_device(prg_handle* h, const uint32_t* rows_dev, const double* score_dev, int B, int n,
                       const prg_dpp_params& p, int32_t* out_idx, int32_t* out_n, int32_t* status, bool* handled,
                       DppFinal* fin);

int dpp_pair_device(prg_handle* h, const uint32_t* rows_dev, const double* score_dev, int B, int n,
                    const prg_dpp_params& p, int32_t* out_idx, int32_t* out_n, int32_t* status, bool* handled,
                    DppFinal* fin);

int dpp_generic_launch(prg_handle* h, const uint32_t* rows_dev, const double* score_dev, int B, int n, const prg_dpp_params& p,
                       int32_t* out_idx, int32_t* out_n, int32_t* status, int c_rows, const void* D, uint64_t D_rows,
                       int D_dim, int dtype, int force_norm, int no_pos);

static size_t dpp_smem_bytes(int c_rows, int top_n) {
  return (size_t)c_rows * kDppMaxItems * 8 + 2 * kDppMaxItems * 8 + 520 * 8 + 32 * 8 + 32 * 4 + kDppMaxItems * 4 +
         (size_t)((top_n + 3) & ~3) * 4 + kDppMaxItems + 64;
}

int dpp_device(prg_handle* h, const uint32_t* rows_dev, const double* score_dev, int B, int n, const prg_dpp_params& p,
               int32_t* out_idx, int32_t* out_n, int32_t* status, DppFinal* fin) {
  if (fin) fin->done = false;
  if (!h->D) return fail(PRG_ESTATE, "diversity matrix not set (prg_set_diversity_matrix)");
  if (B <= 0 || n <= 0 || p.top_n <= 0) return fail(PRG_EINVAL, "B, n, top_n must be positive");
  if (n > kDppMaxN) return fail(PRG_EUNSUPPORTED, "prg_dpp: n > 4096");
  if (h->D_dim > 512) return fail(PRG_EUNSUPPORTED, "prg_dpp: embedding dim > 512");
  if (p.top_n > 4096) return fail(PRG_EUNSUPPORTED, "prg_dpp: top_n > 4096");
  const bool truncates = (p.candidate_count > 0 || p.min_score_percent > 0);
  if (n > kDppMaxItems && !(truncates && p.candidate_count > 0 &&
                            (p.candidate_count > p.top_n ? p.candidate_count : p.top_n) <= kDppMaxItems))
    return fail(PRG_EUNSUPPORTED, "prg_dpp: more than 1024 candidates reach the kernel (set CandidateCount <= 1024)");
  const int window = p.window_size > 0 ? p.window_size : 10;
  int c_rows = p.top_n <= window ? p.top_n : window;
  if (c_rows < 6) c_rows = 6;  // the region doubles as presort staging (4096 x 12 B)
  if (c_rows > 24) return fail(PRG_EUNSUPPORTED, "prg_dpp: window (or top_n when <= window) > 24");
  // The fast kernels below build S = F F^T in gonum's 64-wide k blocks (dgemmParallel).  With at most 64 items the
  // reference takes dgemmSerial instead (one dot product over all D + 1 features, dpp_kernel's kblk): a different last
  // bit once D + 1 > 64.  Where the host can see that no request of the call keeps more than 64 candidates, the generic
  // kernel serves it; a request that falls to <= 64 only through MinScorePercent or padding rows stays on the blocked order.
  int m_upper = n;
  if (p.candidate_count > 0 && n > p.top_n) {
    const int cnt = p.candidate_count > p.top_n ? p.candidate_count : p.top_n;
    if (cnt < m_upper) m_upper = cnt;
  }
  const bool serial_gemm = (int)h->D_dim + 1 > 64 && m_upper <= 64;
  if (serial_gemm)
    return dpp_generic_launch(h, rows_dev, score_dev, B, n, p, out_idx, out_n, status, c_rows, h->D, h->D_rows, (int)h->D_dim,
                              h->D_dtype, 0, 0);
  if (!h->dpp_generic && h->dpp_lazy) {  // config "dpp_lazy": one CTA per request, lazy evaluation of the greedy step (dpp_lazy.cu)
    bool handled = false;
    PRG_TRY(dpp_lazy_device(h, rows_dev, score_dev, B, n, p, out_idx, out_n, status, &handled));
    if (handled) return PRG_OK;
  }
  if (!h->dpp_generic && h->dpp_pair) {  // config "dpp_pair" (experimental): 2-CTA cluster per request, one wave per batch (dpp_pair.cu)
    bool handled = false;
    PRG_TRY(dpp_pair_device(h, rows_dev, score_dev, B, n, p, out_idx, out_n, status, &handled, fin));
    if (handled) return PRG_OK;
  }
  if (!h->dpp_generic) {  // default: 4-CTA cluster per request, features resident on chip (dpp_cluster.cu)
    bool handled = false;
    PRG_TRY(dpp_cluster_device(h, rows_dev, score_dev, B, n, p, out_idx, out_n, status, &handled, fin));
    if (handled) return PRG_OK;
  }
  return dpp_generic_launch(h, rows_dev, score_dev, B, n, p, out_idx, out_n, status, c_rows, h->D, h->D_rows, (int)h->D_dim,
                            h->D_dtype, 0, 0);
}

// the one-CTA-per-request kernel over any table (D, D_rows, D_dim, dtype)
int dpp_generic_launch(prg_handle* h, const uint32_t* rows_dev, const double* score_dev, int B, int n, const prg_dpp_params& p,
                       int32_t* out_idx, int32_t* out_n, int32_t* status, int c_rows, const void* D, uint64_t D_rows,
                       int D_dim, int dtype, int force_norm, int no_pos) {
  const size_t esz = dtype == PRG_F64 ? 8 : 4;
  PRG_TRY(h->dpp_scratch.ensure((size_t)B * D_dim * kDppMaxItems * esz));
  DppArgs a{};
  a.rows = rows_dev; a.score = score_dev; a.n = n; a.D = D; a.D_rows = D_rows; a.D_dim = D_dim;
  a.p = p; a.Et = h->dpp_scratch.p; a.out_idx = out_idx; a.out_n = out_n; a.status = status; a.c_rows = c_rows;
  a.D_sub = (const float*)h->D_sub.p; a.force_norm = force_norm; a.no_pos = no_pos;
  const size_t smem = dpp_smem_bytes(c_rows, p.top_n);
  StageScope span(h, ST_DPP);
  if (dtype == PRG_F64) {
    PRG_CUDA(cudaFuncSetAttribute(dpp_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dpp_kernel<double><<<B, kDppMaxItems, smem, h->stream>>>(a);
  } else {
    PRG_CUDA(cudaFuncSetAttribute(dpp_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dpp_kernel<float><<<B, kDppMaxItems, smem, h->stream>>>(a);
  }
  PRG_CUDA(cudaGetLastError());
  count_launch(h);
  return PRG_OK;
}

// ---------------------------------------------------------------- hook embeddings (sort/dpp_sort.go:362-370, :412-447)
// RegisterEmbeddingHook functions return a []float64 per item; the host passes their concatenation per candidate as
// hook[B][n][hook_dim] f64.  Three shapes of KernelMatrix:
//   table only             (:422-431)  f = [e_table ; 1] / sqrt2                        -> the fast kernels above
//   hooks + table          (:416-421)  v = concat(hook, e_table); v /= ||v||; f = [v ; 1] / sqrt2
//   hooks only             (:432-447)  v = hook; v /= ||v|| if NormalizeEmb; f = [v ; 1] / sqrt2, or [v ; 0] when
//                                      EnsurePositiveSim == "false"
// e_table is the row as the cache holds it: normalised at load when NormalizeEmb (:234-237).  The per-call vectors are
// assembled in fp64 by dpp_hook_concat_kernel and the generic kernel runs over them.
__global__ void dpp_hook_concat_kernel(const uint32_t* __restrict__ rows, const double* __restrict__ hook, int hook_dim, int M,
                                       int n, const float* __restrict__ D, const double* __restrict__ D_inv, uint64_t D_rows,
                                       int D_dim, const float* __restrict__ D_sub, const double* __restrict__ D_sub_inv,
                                       int normalize_emb, double* __restrict__ out, uint32_t* __restrict__ rows_out) {
  const int i = blockIdx.x;            // candidate (b * n + position)
  if (i >= M) return;
  const int W = hook_dim + D_dim;
  const uint32_t row = rows ? rows[i] : (uint32_t)i;
  if (threadIdx.x == 0) rows_out[i] = (rows && row == 0xFFFFFFFFu) ? 0xFFFFFFFFu : (uint32_t)i;
  for (int d = threadIdx.x; d < hook_dim; d += blockDim.x) out[(size_t)i * W + d] = hook[(size_t)i * hook_dim + d];
  if (D_dim > 0) {
    const bool have = (uint64_t)row < D_rows;
    const uint32_t srow = (uint32_t)(i % n) & (kDppSubRows - 1);
    const float* src = have ? D + (size_t)row * D_dim : D_sub + (size_t)srow * D_dim;
    const double inv = !normalize_emb ? 1.0 : (have ? D_inv[row] : D_sub_inv[srow]);
    for (int d = threadIdx.x; d < D_dim; d += blockDim.x) {
      const double x = (double)src[d];
      out[(size_t)i * W + hook_dim + d] = normalize_emb ? __dmul_rn(x, inv) : x;   // floats.Scale(1/normV, vector) at load
    }
  }
}

int dpp_hook_device(prg_handle* h, const uint32_t* rows_dev, const double* score_dev, const double* hook_dev, int hook_dim,
                    int use_table, int B, int n, const prg_dpp_params& p, int32_t* out_idx, int32_t* out_n, int32_t* status) {
  if (hook_dim <= 0 || !hook_dev) return fail(PRG_EINVAL, "hook embeddings missing");
  if (use_table && (!h->D || h->D_dtype != PRG_F32)) return fail(PRG_EUNSUPPORTED, "hooks + table need an f32 diversity table");
  const int Dt = use_table ? (int)h->D_dim : 0;
  const int W = hook_dim + Dt;
  if (W > 512) return fail(PRG_EUNSUPPORTED, "prg_dpp_ex: hook_dim + table dim > 512");
  if (n > kDppMaxN) return fail(PRG_EUNSUPPORTED, "prg_dpp: n > 4096");
  const int window = p.window_size > 0 ? p.window_size : 10;
  int c_rows = p.top_n <= window ? p.top_n : window;
  if (c_rows < 6) c_rows = 6;
  if (c_rows > 24) return fail(PRG_EUNSUPPORTED, "prg_dpp: window (or top_n when <= window) > 24");
  const int M = B * n;
  if (!h->D_sub.p) {   // no table was ever set: the substitutes are only touched when use_table
    PRG_TRY(h->D_sub.ensure(16));
  }
  PRG_TRY(h->dpp_hook_E.ensure((size_t)M * W * 8));
  PRG_TRY(h->dpp_hook_rows.ensure((size_t)M * 4));
  dpp_hook_concat_kernel<<<M, 128, 0, h->stream>>>(rows_dev, hook_dev, hook_dim, M, n, (const float*)h->D,
                                                   (const double*)h->D_inv.p, h->D_rows, Dt, (const float*)h->D_sub.p,
                                                   (const double*)h->D_sub_inv.p, p.normalize_emb,
                                                   (double*)h->dpp_hook_E.p, (uint32_t*)h->dpp_hook_rows.p);
  PRG_CUDA(cudaGetLastError());
  count_launch(h);
  prg_dpp_params q = p;
  int force_norm = 0;
  if (use_table) { force_norm = 1; q.no_positive_sim = 0; }   // :416-431: always re-normalised, always [v ; 1] / sqrt2
  return dpp_generic_launch(h, (const uint32_t*)h->dpp_hook_rows.p, score_dev, B, n, q, out_idx, out_n, status, c_rows,
                            h->dpp_hook_E.p, (uint64_t)M, W, PRG_F64, force_norm, use_table ? 0 : (p.no_positive_sim != 0));
}

}  // namespace prg
