// dpp_lazy.cu — DPPSort.doSort (sort/dpp_sort.go:271-351, :372-475, :477-551) with LAZY evaluation of the greedy
// step: one CTA per request, no cluster, nothing resident but scalars.
//
// The reference's step k computes, for EVERY candidate i, e_i = (L_ji - <c_j, c_i>) / d_j and d2_i -= e_i^2, then takes
// the first maximum of d2.  Two facts make most of that work unnecessary for finding the same maximum:
//   * d2_i never increases (a non-negative number is subtracted, also in floating point), so a value computed some
//     steps ago is an upper bound of the current one;
//   * e_i and d2_i depend only on candidate i's own history and on the picks, never on other candidates.
// So each candidate carries (d2 as of step t_i, t_i, its rows of C) and is brought up to date only when its stale
// bound could still win: a step refreshes candidates in passes of P (one group of 4 / 8 lanes per candidate) until no
// stale candidate has a bound above the best up-to-date value (ties: lower list index, exactly the reference's
// first-maximum rule).  A refresh replays the missed steps in order with the arithmetic of dpp_cluster.cu / oracle.c
// (gonum Dgemm / DotUnitary summation order, separate multiply and add roundings), so every d2 that takes part in a
// decision is bit-identical to the reference's, and so is the selection sequence.
//
// Status: OPT-IN (config "dpp_lazy"); the default stays dpp_cluster.cu.  The arithmetic shrinks as predicted — on the
// bench's score distribution 7.7 k - 10 k candidate refreshes per request instead of 45 k candidate-steps (all-equal
// scores: 21 k), ~2.3 passes of 128 candidates per step — and every test passes bit for bit, but the kernel is SLOWER
// than the cluster kernel: 0.67 ms per launch for any batch up to 128 requests (one wave, 1.28 M cycles per request)
// against 0.14 ms per wave of 37.  ncu (profiles/r01_dpp_lazy_ncu_summary.txt): 36 % of the samples are barrier stalls —
// the passes after the first hold a handful of candidates that each replay up to 9 missed steps (a serial chain of
// ~500 cycles per step) while 25 of 32 warps wait — and 20 % wait on the L2 loads that rebuild features from the f32
// rows.  What would have to change for it to win: register-cached features for the hot prefix, the Gram values of all
// missed steps computed side by side before the serial e / d2 chain, and passes without block-wide barriers.
// Candidates are visited in descending-score order, which approximates descending-bound order within a window.
#include "dpp_common.cuh"
#include <type_traits>

namespace prg {

constexpr int kLzThreads = 1024;
constexpr int kLzCRows = 10;      // rows of C kept (window size, or top_n when it is smaller)
constexpr int kLzMaxItems = 1024;

template <int D>
struct __align__(16) LzPick {     // what a refresh needs to know about the pick of one window step
  double inv_dj;                  // 1 / sqrt(d2_j)
  double q;                       // exp(alpha * rel_j)
  double cj[kLzCRows];            // column j of C (rows < step)
  double f[ClCfg<D>::kLPC * ClCfg<D>::kFS];   // its features by (lane of the group, chain position)
};

template <int D>
__global__ void __launch_bounds__(kLzThreads, 1) dpp_lazy_kernel(const DppClArgs a) {
  using Cfg = ClCfg<D>;
  using Pick = LzPick<D>;
  constexpr int LPC = Cfg::kLPC, CL = Cfg::kCL, FS = Cfg::kFS;
  constexpr int GROUPS = kLzThreads / LPC;      // candidates refreshed per pass (128 at D = 128, 256 at D = 32 / 64)
  const int b = blockIdx.x;
  extern __shared__ __align__(16) uint8_t lsm[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int grp = tid / LPC, lam = tid % LPC;
  const int q = lam & 3, bb = lam >> 2;
  const int n = a.n, T_out = a.p.top_n;
  const int window = a.p.window_size > 0 ? a.p.window_size : 10;
  const int c_rows = T_out <= window ? T_out : window;

  // per-candidate state, indexed by position p in descending-score ("static") order
  double* C = reinterpret_cast<double*>(lsm);                          // [kLzCRows][1024]
  double* d2_s = C + (size_t)kLzCRows * kLzMaxItems;                   // [1024] value as of step tstep[p]
  double* q_s = d2_s + kLzMaxItems;                                    // [1024]
  double* inv_s = q_s + kLzMaxItems;                                   // [1024] 1 / ||row||
  double* diag_s = inv_s + kLzMaxItems;                                // [1024] L_pp
  Pick* picks = reinterpret_cast<Pick*>(diag_s + kLzMaxItems);         // [kLzCRows]
  uint64_t* red_k = reinterpret_cast<uint64_t*>(picks + kLzCRows);     // [32]
  int32_t* red_i = reinterpret_cast<int32_t*>(red_k + 32);             // [32] list index of the warp's maximum
  int32_t* red_p = red_i + 32;                                         // [32] its position
  uint32_t* row_s = reinterpret_cast<uint32_t*>(red_p + 32);           // [1024] diversity-table row
  int32_t* lidx = reinterpret_cast<int32_t*>(row_s + kLzMaxItems);     // [1024] position -> index in the (truncated) list
  int32_t* order = lidx + kLzMaxItems;                                 // [1024] list index -> index in the request's input
  int32_t* work = order + kLzMaxItems;                                 // [1024] positions to refresh in this round
  int32_t* wcnt = work + kLzMaxItems;                                  // [32] per-warp counts of the compaction
  int32_t* res = wcnt + 32;                                            // [T_out] picks (list indices)
  uint8_t* tstep = reinterpret_cast<uint8_t*>(res + ((T_out + 3) & ~3));  // [1024] window steps applied
  uint8_t* existed = tstep + kLzMaxItems;                              // [1024] by LIST index
  __shared__ int s_m, s_err, s_ny, s_sorted, s_work;
  __shared__ double s_p0, s_p1;

  const uint32_t* rows = a.rows + (size_t)b * n;
  const double* score = a.score + (size_t)b * n;

  // ---- 0. valid count, optional presort + truncation (:280-300) -> order[], m
  if (tid == 0) { s_m = 0; s_err = 0; s_sorted = 1; }
  __syncthreads();
  {
    int cnt = 0;
    for (int i = tid; i < n; i += kLzThreads) cnt += (rows[i] != 0xFFFFFFFFu);
    if (cnt) atomicAdd(&s_m, cnt);
  }
  __syncthreads();
  const int nv = s_m;
  __syncthreads();
  int m = nv;
  const bool presort = (a.p.candidate_count > 0 || a.p.min_score_percent > 0) && nv > T_out;
  // one bitonic sort serves both the reference's presort and the static visiting order (descending score, stable)
  auto sort_desc = [&](int count, auto score_of, int32_t* out_idx) {   // out_idx[rank] = index, count <= 4096
    uint32_t P2 = 32;
    while (P2 < (uint32_t)count) P2 <<= 1;
    uint64_t* key = reinterpret_cast<uint64_t*>(lsm);   // staging over C (not live yet): 12 B x P2 <= 48 KiB
    int32_t* idx = reinterpret_cast<int32_t*>(key + P2);
    for (uint32_t i = tid; i < P2; i += kLzThreads) {
      key[i] = (i < (uint32_t)count) ? f64_ord_c(score_of((int)i)) : 0ull;
      idx[i] = (i < (uint32_t)count) ? (int32_t)i : 0x7FFFFFFF;
    }
    __syncthreads();
    for (uint32_t size = 2; size <= P2; size <<= 1) {
      for (uint32_t stride = size >> 1; stride > 0; stride >>= 1) {
        for (uint32_t i = tid; i < (P2 >> 1); i += kLzThreads) {
          const uint32_t pos = 2 * i - (i & (stride - 1));
          const uint64_t ka = key[pos], kb = key[pos + stride];
          const int32_t ia = idx[pos], ib = idx[pos + stride];
          const bool a_after_b = (ka < kb) || (ka == kb && ia > ib);
          if (a_after_b == ((pos & size) == 0)) { key[pos] = kb; key[pos + stride] = ka; idx[pos] = ib; idx[pos + stride] = ia; }
        }
        __syncthreads();
      }
    }
    for (int i = tid; i < count && i < kLzMaxItems; i += kLzThreads) out_idx[i] = idx[i];
    __syncthreads();
    return idx;   // still valid until C is written
  };
  if (nv > 0 && presort) {
    int32_t* idx = sort_desc(nv, [&](int i) { return score[i]; }, order);
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
  } else if (m <= kLzMaxItems) {
    for (int i = tid; i < m; i += kLzThreads) order[i] = i;
  }
  if (nv == 0 || m > kLzMaxItems) {
    if (tid == 0) { a.out_n[b] = 0; a.status[b] = (nv == 0) ? 0 : 2; }
    return;
  }
  __syncthreads();

  // ---- 1. abtest normalisation parameters (:382-405)
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
  }
  if (s_err) {
    if (tid == 0) { a.out_n[b] = 0; a.status[b] = 1; }
    return;
  }

  // ---- 2. static order: list indices by descending score (identity when the list already is sorted: the fused path)
  for (int i = tid; i + 1 < m; i += kLzThreads) {
    const uint64_t k0 = f64_ord_c(score[order[i]]), k1 = f64_ord_c(score[order[i + 1]]);
    if (k0 < k1) s_sorted = 0;
  }
  __syncthreads();
  if (s_sorted) {
    for (int i = tid; i < m; i += kLzThreads) lidx[i] = i;
    __syncthreads();
  } else {
    // `order` is only read through the lambda until the sort has copied its result out
    sort_desc(m, [&](int i) { return score[order[i]]; }, lidx);
  }

  // per-position scalars
  for (int p = tid; p < kLzMaxItems; p += kLzThreads) {
    const bool act = p < m;
    const int li = act ? lidx[p] : 0;
    const uint32_t r = act ? rows[order[li]] : 0xFFFFFFFFu;
    const uint32_t rw = dpp_row_code(act, r, a.D_rows, act ? order[li] : 0);
    row_s[p] = rw;
    double rel = act ? score[order[li]] : 0.0;
    if (a.p.norm_mode == 1) rel = __dsub_rn(rel, s_p0) / s_p1;
    else if (a.p.norm_mode == 2) rel = __dadd_rn(__dmul_rn(__dsub_rn(rel, s_p0) / s_p1, 1 - 1e-6), 1e-6);
    inv_s[p] = !a.p.normalize_emb ? 1.0 : (rw != 0xFFFFFFFFu ? dpp_row_inv(a, rw) : 1.0);
    q_s[p] = act ? exp(__dmul_rn(a.p.alpha, rel)) : 0.0;
    existed[p] = 0;
  }
  __syncthreads();

  const double cc = __dmul_rn(kInvSqrt2c, kInvSqrt2c);
  const bool do_norm = a.p.normalize_emb != 0;
  // features of the candidate at position p for this lane: chain q of k block bb, elements 64*bb + 4*t + q
  auto build = [&](int p, double (&f)[CL]) {
    const uint32_t rw = row_s[p];
    const double inv = inv_s[p];
    const float* src = dpp_row_ptr(a, rw != 0xFFFFFFFFu ? rw : 0u, D) + 64 * bb + q;
#pragma unroll
    for (int t = 0; t < CL; ++t) {
      const double x = (rw != 0xFFFFFFFFu) ? (double)__ldg(src + 4 * t) : 0.0;
      f[t] = do_norm ? __dmul_rn(__dmul_rn(x, inv), kInvSqrt2c) : __dmul_rn(x, kInvSqrt2c);
    }
  };
  // <g, f> in gonum Dgemm(NoTrans,Trans) order (see dpp_cluster.cu); g == nullptr: <f, f>.  Every lane of the group
  // returns the total.
  auto gram = [&](const double* g, const double (&f)[CL]) -> double {
    double acc = 0.0;
#pragma unroll
    for (int t = 0; t < CL; t += 2) {
      double gx = f[t], gy = f[t + 1];
      if (g) { const double2 g2 = *reinterpret_cast<const double2*>(g + lam * FS + t); gx = g2.x; gy = g2.y; }
      acc = __dadd_rn(acc, __dmul_rn(gx, f[t]));
      acc = __dadd_rn(acc, __dmul_rn(gy, f[t + 1]));
    }
    double u = acc;
    if (!Cfg::kConstOwnBlock && q == 0 && bb == Cfg::kBlocks - 1) u = __dadd_rn(u, cc);
    double o = shfl_xor_f64(u, 2);
    const double pr = (q & 2) ? __dadd_rn(o, u) : __dadd_rn(u, o);
    o = shfl_xor_f64(pr, 1);
    const double bs = (q & 1) ? __dadd_rn(o, pr) : __dadd_rn(pr, o);
    double tot;
    if (Cfg::kBlocks == 2) {
      o = shfl_xor_f64(bs, 4);
      tot = __dadd_rn(__dadd_rn(0.0, bb ? o : bs), bb ? bs : o);
    } else {
      tot = __dadd_rn(0.0, bs);
    }
    if (Cfg::kConstOwnBlock) tot = __dadd_rn(tot, cc);
    return tot;
  };

  // ---- 3. diag(L) once per request: L_pp = (q_p * <f_p, f_p>) * q_p
  for (int p0 = 0; p0 < m; p0 += GROUPS) {
    const int p = p0 + grp;
    const bool act = p < m;
    double f[CL];
    build(act ? p : 0, f);
    const double s = gram(nullptr, f);
    if (act && lam == 0) diag_s[p] = __dmul_rn(__dmul_rn(q_s[p], s), q_s[p]);
  }
  __syncthreads();

  // block-wide first maximum over the up-to-date candidates (want_step), by LIST index among equal values
  auto block_argmax = [&](int want_step, uint64_t* kout, int* iout, int* pout) {
    uint64_t key = 0ull;
    int li = 0x7FFFFFFF;
    if (tid < m && tstep[tid] == (uint8_t)want_step) { key = d2_key(d2_s[tid]); li = lidx[tid]; }
    uint64_t wk;
    int wi;
    warp_first_max(key, li, &wk, &wi);
    const unsigned who = __ballot_sync(0xffffffffu, key == wk && li == wi);
    if (lane == 0) { red_k[warp] = wk; red_i[warp] = wi; red_p[warp] = (warp << 5) + (__ffs(who) - 1); }
    __syncthreads();
    uint64_t bk;
    int bi;
    warp_first_max(red_k[lane], red_i[lane], &bk, &bi);
    const unsigned w2 = __ballot_sync(0xffffffffu, red_k[lane] == bk && red_i[lane] == bi);
    *kout = bk;
    *iout = bi;
    *pout = red_p[__ffs(w2) - 1];
    __syncthreads();   // red_* may be rewritten by the next call
  };

  // ---- 4. DPPWithWindow (:477-491) over DPP (:493-551)
  int total = 0;
  const int n_calls = (T_out <= window) ? 1 : (T_out / window + (T_out % window > 0 ? 1 : 0));
  for (int call = 0; call < n_calls; ++call) {
    int top = (T_out <= window) ? T_out : ((call < T_out / window) ? window : T_out % window);
    if (top > m) top = m;
    for (int p = tid; p < kLzMaxItems; p += kLzThreads) {
      d2_s[p] = (p < m && !existed[lidx[p]]) ? diag_s[p] : CUDART_NAN;
      tstep[p] = 0;
    }
    __syncthreads();
    uint64_t bkey;
    int j_li, j_p;
    block_argmax(0, &bkey, &j_li, &j_p);
    if (bkey == 0) { j_li = 0; j_p = -1; }   // every candidate used up: the reference's MaxIdx returns index 0
    if (tid == 0) res[total] = j_li;
    int ny = 1;
    bool broke = false;
    while (ny < top) {
      const int k = ny - 1;   // window step being applied; picks[k] describes its pick
      const double dj = (j_p >= 0) ? d2_s[j_p] : CUDART_NAN;
      if (dj < 1e-10) { broke = true; break; }
      // publish the pick: scalars, its column of C (rows < k are up to date: it was up to date at step k), features
      if (j_p >= 0) {
        if (grp == 0) {
          double f[CL];
          build(j_p, f);
#pragma unroll
          for (int t = 0; t < CL; ++t) picks[k].f[lam * FS + t] = f[t];
        } else if (tid >= 64 && tid < 64 + kLzCRows) {
          const int l = tid - 64;
          picks[k].cj[l] = (l < k) ? C[l * kLzMaxItems + j_p] : 0.0;
        } else if (tid == 96) {
          picks[k].inv_dj = 1.0 / sqrt(dj);
          picks[k].q = q_s[j_p];
        }
      } else {  // no candidate left (NaN pick): every refresh would produce NaN; nothing is up to date any more
        if (tid < LPC * FS) picks[k].f[tid] = CUDART_NAN;
        if (tid >= 64 && tid < 64 + kLzCRows) picks[k].cj[tid - 64] = CUDART_NAN;
        if (tid == 96) { picks[k].inv_dj = CUDART_NAN; picks[k].q = CUDART_NAN; }
      }
      __syncthreads();
      if (j_p >= 0 && tid == 0) d2_s[j_p] = CUDART_NAN;   // picked: never a candidate again in this window
      __syncthreads();

      // passes: refresh stale candidates whose bound can still win, in static order, GROUPS at a time
      uint64_t best_k = 0ull;
      int best_li = 0x7FFFFFFF, best_p = -1;
      for (;;) {
        // compaction of the positions to refresh (order preserving)
        bool need = false;
        if (tid < m && tstep[tid] != (uint8_t)ny) {
          const uint64_t key = d2_key(d2_s[tid]);
          need = key != 0ull && (key > best_k || (key == best_k && lidx[tid] < best_li));
        }
        const unsigned bal = __ballot_sync(0xffffffffu, need);
        if (lane == 0) wcnt[warp] = __popc(bal);
        __syncthreads();
        {
          const int c = wcnt[lane];
          int run = c;
#pragma unroll
          for (int off = 1; off < 32; off <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, run, off);
            if (lane >= off) run += v;
          }
          const int base = __shfl_sync(0xffffffffu, run - c, warp);
          const int tot = __shfl_sync(0xffffffffu, run, 31);
          if (need) work[base + __popc(bal & ((1u << lane) - 1u))] = tid;
          if (tid == 0) s_work = tot;
        }
        __syncthreads();
        const int n_work = s_work;
        if (n_work == 0) break;
        // refresh the first GROUPS of them
        {
          const bool act = grp < n_work;
          const int p = act ? work[grp] : 0;
          double f[CL];
          build(p, f);
          const int t0 = act ? (int)tstep[p] : ny;
          double d2 = d2_s[p];
          const double qi = q_s[p];
          // every group of the warp runs the same number of loop steps (shuffles inside gram are warp-wide)
          int steps = ny - t0;
#pragma unroll
          for (int off = LPC; off < 32; off <<= 1) { const int o = __shfl_xor_sync(0xffffffffu, steps, off); steps = o > steps ? o : steps; }
          for (int it = 0; it < steps; ++it) {
            const int s = t0 + it;
            const bool live = act && s < ny;
            const Pick& W = picks[live ? s : 0];
            const double S = gram(W.f, f);
            if (live && lam == 0) {
              double ss = 0.0;
              for (int l = 0; l < s; ++l) {
                const double tm = W.cj[l];
                if (tm != 0) ss = __dadd_rn(ss, __dmul_rn(tm, C[l * kLzMaxItems + p]));
              }
              const double Lji = __dmul_rn(__dmul_rn(W.q, S), qi);
              const double e = (s == 0) ? __dmul_rn(W.inv_dj, Lji) : __dmul_rn(W.inv_dj, __dsub_rn(Lji, ss));
              C[s * kLzMaxItems + p] = e;
              d2 = __dsub_rn(d2, __dmul_rn(e, e));
            }
          }
          if (act && lam == 0) { d2_s[p] = d2; tstep[p] = (uint8_t)ny; }
        }
        __syncthreads();
        block_argmax(ny, &best_k, &best_li, &best_p);
        if (best_k == 0) { best_li = 0x7FFFFFFF; best_p = -1; }
      }
      // winner of this step
      j_li = (best_k == 0) ? 0 : best_li;
      j_p = best_p;
      if (tid == 0) res[total + ny] = j_li;
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
    if (tid < ny) existed[res[total + tid]] = 1;
    total += ny;
    __syncthreads();
  }
  for (int t = tid; t < total; t += kLzThreads) a.out_idx[(size_t)b * T_out + t] = order[res[t]];
  if (tid == 0) { a.out_n[b] = total; a.status[b] = 0; }
}

template <int D>
static size_t dpp_lazy_smem(int top_n) {
  size_t live = (size_t)kLzCRows * kLzMaxItems * 8 + 4 * (size_t)kLzMaxItems * 8 + kLzCRows * sizeof(LzPick<D>) + 32 * 8 +
                2 * 32 * 4 + 4 * (size_t)kLzMaxItems * 4 + 32 * 4 + (size_t)((top_n + 3) & ~3) * 4 + 2 * kLzMaxItems + 64;
  const size_t sort_staging = (size_t)kClMaxN * 12;
  return live > sort_staging ? live : sort_staging;
}

template <int D>
static int launch_lazy(prg_handle* h, const DppClArgs& a, int B) {
  const size_t smem = dpp_lazy_smem<D>(a.p.top_n);
  PRG_CUDA(cudaFuncSetAttribute(dpp_lazy_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dpp_lazy_kernel<D><<<B, kLzThreads, smem, h->stream>>>(a);
  PRG_CUDA(cudaGetLastError());
  count_launch(h);
  return PRG_OK;
}

// returns PRG_OK and sets *handled when the request shape is served by the lazy kernel
int dpp_lazy_device(prg_handle* h, const uint32_t* rows_dev, const double* score_dev, int B, int n, const prg_dpp_params& p,
                    int32_t* out_idx, int32_t* out_n, int32_t* status, bool* handled) {
  *handled = false;
  if (h->D_dtype != PRG_F32 || !h->D_inv.p) return PRG_OK;
  if (h->D_dim != 32 && h->D_dim != 64 && h->D_dim != 128) return PRG_OK;
  if (n > kClMaxN || p.top_n > 2048) return PRG_OK;
  const int window = p.window_size > 0 ? p.window_size : 10;
  const int c_rows = p.top_n <= window ? p.top_n : window;
  if (c_rows > kLzCRows) return PRG_OK;
  DppClArgs a{};
  a.rows = rows_dev; a.score = score_dev; a.n = n; a.D = (const float*)h->D; a.D_inv = (const double*)h->D_inv.p; a.D_rows = h->D_rows; a.p = p;
  a.D_sub = (const float*)h->D_sub.p; a.D_sub_inv = (const double*)h->D_sub_inv.p;
  a.out_idx = out_idx; a.out_n = out_n; a.status = status;
  StageScope span(h, ST_DPP);
  int rc = PRG_OK;
  if (h->D_dim == 32) rc = launch_lazy<32>(h, a, B);
  else if (h->D_dim == 64) rc = launch_lazy<64>(h, a, B);
  else rc = launch_lazy<128>(h, a, B);
  if (rc == PRG_OK) *handled = true;
  return rc;
}

}  // namespace prg
