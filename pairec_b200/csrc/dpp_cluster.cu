// dpp_cluster.cu — the fast path of DPPSort.doSort (sort/dpp_sort.go:271-351, :372-475, :477-551) for f32 diversity
// tables with dim 32 / 64 / 128: one thread-block CLUSTER of 4 CTAs per request, 256 candidates per CTA, two threads
// (a lane pair) per candidate, the candidates' fp64 FEATURES resident on chip (registers + shared memory).
//
// Why: the selection loop is 50 dependent steps of "row of L for the chosen item against every candidate" — n x (D+1)
// fp64 multiply-adds per step with nothing reused between steps.  With one CTA per request (dpp.cu) every step
// re-reads the request's embeddings through L2 and stalls on that latency (ncu r1: 62 % long-scoreboard, 1.56 ms per
// 64-request batch).  History of this kernel (64 requests x 1000 candidates x 128-d, top 50, window 10):
//   v2  embeddings as f32 in 128 registers per thread: 0.45 ms; 63 % of issue slots lost to instruction fetch.
//   v3  embeddings as f32 in shared memory, features rebuilt every step (cvt + 2 dmul per element before the
//       multiply-add): 0.46 ms; ncu (profiles/r01_dpp_v3_ncu_full_summary.txt): fp64 pipe 33 %, 5 fp64 ops + one
//       F2F.F64.F32 per element, 20 % of samples in CTA barriers (unbalanced k-block split, 4 barriers per step).
//   v4  (this file) the rounded fp64 feature f = rn(rn(x*inv)*2^-1/2) is computed ONCE and kept — the winner's record
//       carries its features, so a step is exactly one dmul + one dadd per element; gonum's four DotUnitary chains
//       are split by parity over the lane pair (lane 0: s0,s2; lane 1: s1,s3; one shuffle per 64-wide block), which
//       balances the two lanes for every dim and needs no barrier; one CTA barrier + one cluster barrier per step.
//
// Arithmetic is exactly dpp.cu's (and oracle/oracle.c's): fp64, gonum operation order, separate multiply/add
// roundings, first-maximum argmax (rank order == index order), NaN masking, 1e-10 early stop + lowest-index fill.
// Per step each CTA publishes its local arg-max candidate TOGETHER with everything the others need about it
// (d2, quality, its features, its column of C) into every CTA's shared memory (DSMEM); the records are
// double buffered so a fast CTA never overwrites what a slow one still reads.
#include "dpp_common.cuh"
#include <cooperative_groups.h>
#include <type_traits>

namespace cg = cooperative_groups;

namespace prg {

template <int D>
struct __align__(16) CandRecT {
  double v;       // d2 of the candidate (NaN if the CTA has none)
  uint64_t key;   // its order key (0 = none)
  double q;       // exp(alpha * rel)
  double inv_dj;  // 1 / sqrt(d2): computed by the publishing thread while the winner's warp ships the features
  int32_t idx;    // index in the truncated list
  uint32_t row;   // diversity-table row (diagnostics)
  double pad_;
  double cj[kClCRows];
  double f[ClCfg<D>::kLPC * ClCfg<D>::kFS];  // features by (lane of the group, chain position); the constant feature is implied
};

// cluster publication primitives: smem -> peer smem bulk copy (async proxy) completing on the PEER's mbarrier
__device__ __forceinline__ uint32_t mapa_u32(uint32_t cta_smem_addr, uint32_t cta_rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(cta_smem_addr), "r"(cta_rank));
  return r;
}
__device__ __forceinline__ void bulk_copy_to_peer(uint32_t dst_cluster_addr, uint32_t src_cta_addr, uint32_t bytes,
                                                  uint32_t mbar_cluster_addr) {
  asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   dst_cluster_addr),
               "r"(src_cta_addr), "r"(bytes), "r"(mbar_cluster_addr)
               : "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// bounded wait: a protocol bug must trap, not hang the GPU
__device__ __forceinline__ void mbar_wait_bounded(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t done = 0;
  for (long long spins = 0; !done; ++spins) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
    if (!done && spins > (1ll << 26)) __trap();
  }
}

#ifdef PRG_DPP_PROFILE
__device__ long long g_dpp_clk[16];
// per-segment cycle sums of thread 0 of CTA 0, accumulated in shared memory (a global read-modify-write per probe would
// add its own L2 round trip to every segment) and flushed once at the end
#define DPP_CLK(slot) do { if (tid == 0) { const long long t_ = clock64(); s_clk[slot] += t_ - t_last; t_last = t_; } } while (0)
#else
#define DPP_CLK(slot) do { } while (0)
#endif

// Final outputs of the fused request path for request b, by all threads of the cluster's leader CTA: the picks' rows
// and rank scores, or — status != 0, the reference returns the list unchanged (sort/dpp_sort.go:317-320) — the first
// top_n entries of the sorted list.  Same result as pipeline.cu's final_gather_kernel, without its launch.
__device__ __forceinline__ void dpp_write_final(const DppClArgs& a, int b, int tid, int st, int total, const int32_t* order,
                                                const int32_t* res) {
  if (!a.fin_row) return;
  const int T = a.p.top_n, n = a.n;
  const bool unchanged = st != 0;
  const int c = unchanged ? (n < T ? n : T) : total;
  for (int t = tid; t < T; t += kClThreads) {
    const size_t o = (size_t)b * T + t;
    if (t < c) {
      const size_t src = (size_t)b * n + (unchanged ? t : order[res[t]]);
      a.fin_row[o] = a.rows[src];
      a.fin_score[o] = a.score[src];
    } else {
      a.fin_row[o] = 0xFFFFFFFFu;
      a.fin_score[o] = 0.0;
    }
  }
  if (tid == 0) a.fin_n[b] = c;
}

template <int D>
__global__ void __launch_bounds__(kClThreads, 1) dpp_cluster_kernel(const DppClArgs a) {
  pdl_wait();                 // chained launch: the predecessor's writes are visible from here on
  pdl_launch_dependents();
  using CandRec = CandRecT<D>;
  using Cfg = ClCfg<D>;
  constexpr int LPC = Cfg::kLPC, R = Cfg::kR, CL = Cfg::kCL, TR = Cfg::kTR, FS = Cfg::kFS;
  cg::cluster_group cluster = cg::this_cluster();
  const unsigned rank = cluster.block_rank();
  const int b = blockIdx.x / kClCtas;
  extern __shared__ __align__(16) uint8_t csm[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int grp = tid / LPC, lam = tid % LPC;   // group of LPC lanes serves candidates grp*R .. grp*R+R-1 of this CTA
  const int q = lam & 3, bb = lam >> 2;         // this lane: DotUnitary chain q of k block bb, for each of the R candidates
  const int n = a.n, T_out = a.p.top_n;
  const int window = a.p.window_size > 0 ? a.p.window_size : 10;
  const int c_rows = T_out <= window ? T_out : window;
  const bool small_window = c_rows <= 10;  // then k <= 8: the <c_j, c_i> sum is fused into the Gram row code

  double* C = reinterpret_cast<double*>(csm);                                   // [c_rows][256]
  double* F = C + (size_t)c_rows * kClItems;                                    // [kSlots][512] features kept in smem
  CandRec* pub = reinterpret_cast<CandRec*>(F + Cfg::kFDoubles);                // [2][4]
  CandRec* srec = pub + 2 * kClCtas;                                            // [2] this CTA's own record, staged for the copy
  uint64_t* mbar = reinterpret_cast<uint64_t*>(srec + 2);                       // [2] one per record buffer
  double* inv_s = reinterpret_cast<double*>(mbar + 2);                          // [256]
  double* q_s = inv_s + kClItems;                                               // [256]
  uint64_t* red_k = reinterpret_cast<uint64_t*>(q_s + kClItems);                // [16]
  int32_t* red_i = reinterpret_cast<int32_t*>(red_k + 16);                      // [16]
  uint32_t* row_s = reinterpret_cast<uint32_t*>(red_i + 16);                    // [256]
  int32_t* order = reinterpret_cast<int32_t*>(row_s + kClItems);                // [1024]
  int32_t* res = order + kClMaxItems;                                           // [T_out]
  uint8_t* existed = reinterpret_cast<uint8_t*>(res + ((T_out + 3) & ~3));      // [1024]
  __shared__ int s_m, s_err, s_ny;
  __shared__ double s_p0, s_p1;

  const uint32_t* rows = a.rows + (size_t)b * n;
  const double* score = a.score + (size_t)b * n;
#ifdef PRG_DPP_PROFILE
  __shared__ long long s_clk[16];
  if (tid == 0)
    for (int i = 0; i < 16; ++i) s_clk[i] = 0;
  long long t_last = clock64();
#endif

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
    uint64_t* key = reinterpret_cast<uint64_t*>(csm);  // staging over C / F (not live yet); 12 B x P2 <= 48 KiB
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
    if (rank == 0) dpp_write_final(a, b, tid, (nv == 0) ? 0 : 2, 0, nullptr, nullptr);
    return;
  }
  __syncthreads();

  // ---- 1. abtest normalisation parameters (:382-405), redundantly per CTA
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
    if (rank == 0 && tid == 0) { a.out_n[b] = 0; a.status[b] = 1; }
    if (rank == 0) dpp_write_final(a, b, tid, 1, 0, nullptr, nullptr);
    return;
  }
  DPP_CLK(0);

  // ---- 2a. rows -> coalesced staging of the 256 embeddings in shared memory -> one thread per candidate: relevance,
  //          1/norm (table lookup), quality.
  //          * staging: thread-per-row reads of 512-B rows were 16 dependent HBM round trips; here every warp loads
  //            whole rows (16 independent 16-B loads per thread) into XS (over the F region, 16-B chunks XOR-swizzled
  //            by row so both the row-wise reads of 2a and the chain-wise reads of 2b are conflict-light).
  constexpr int CH = D / 4;  // 16-B chunks per embedding
  float* XS = reinterpret_cast<float*>(F);
  if (tid == 0) {  // (past the last early return: every CTA of the cluster gets here)
    mbar_init(&mbar[0], 1);
    mbar_init(&mbar[1], 1);
    mbar_fence_init();
  }
  cluster.barrier_arrive();  // peers may signal our mbarriers only after this; waited for before the first publication
  if (tid < kClItems) {
    const int c = (int)rank * kClItems + tid;
    const bool act = c < m;
    const uint32_t r = act ? rows[order[c]] : 0u;
    row_s[tid] = dpp_row_code(act, r, a.D_rows, act ? order[c] : 0);
  }
  for (int i = tid; i < kClMaxItems; i += kClThreads) existed[i] = 0;
  __syncthreads();
#pragma unroll
  for (int g = tid; g < kClItems * CH; g += kClThreads) {
    const int row = g / CH, c = g % CH;
    const uint32_t rw = row_s[row];
    const float4 v = (rw != 0xFFFFFFFFu) ? reinterpret_cast<const float4*>(dpp_row_ptr(a, rw, D))[c] : make_float4(0.f, 0.f, 0.f, 0.f);
    *reinterpret_cast<float4*>(XS + (size_t)row * D + 4 * (c ^ (row & (CH - 1)))) = v;
  }
  __syncthreads();
  if (tid < kClItems) {
    const int c = (int)rank * kClItems + tid;
    const bool act = c < m;
    double rel = act ? score[order[c]] : 0.0;
    if (a.p.norm_mode == 1) rel = __dsub_rn(rel, s_p0) / s_p1;
    else if (a.p.norm_mode == 2) rel = __dadd_rn(__dmul_rn(__dsub_rn(rel, s_p0) / s_p1, 1 - 1e-6), 1e-6);
    // 1 / ||e|| comes from the per-row table built when the matrix was set (dpp_inv_norm_kernel: gonum floats.Norm in
    // its exact operation order).  Computing it here cost 44k cycles per request — 128 dependent divide/accumulate steps
    // per candidate with only half the warps busy — for a value that is a pure function of the table row.  A candidate
    // without a table row takes a substitute direction (dpp_common.cuh) and its norm.
    const uint32_t rw = row_s[tid];
    const double inv = !a.p.normalize_emb ? 1.0 : (rw != 0xFFFFFFFFu ? dpp_row_inv(a, rw) : 1.0);
    inv_s[tid] = inv;
    q_s[tid] = act ? exp(__dmul_rn(a.p.alpha, rel)) : 0.0;
  }
  __syncthreads();
  DPP_CLK(1);

  // ---- 2b. the rounded fp64 features, built once.  Lane (bb, q) of a group holds, for each of the group's R
  //          candidates, chain q of block bb: features 64*bb + 4*t + q, t < CL; the first TR in registers, the rest
  //          in shared memory slots [r*(CL-TR) + t-TR][tid] (written after every thread has read its staged values:
  //          the slots overlay XS).
  double fr[R][TR];
  // shared-memory features: chain positions (t, t+1), t >= TR even, of candidate r form one 16-byte element [pair][tid]
  auto f_idx = [&](int r, int t) -> size_t {
    return ((size_t)(r * ((CL - TR) / 2) + ((t - TR) >> 1)) * kClThreads + tid) * 2 + ((t - TR) & 1);
  };
  {
    const bool do_norm = a.p.normalize_emb != 0;
    float xt[R][CL - TR > 0 ? CL - TR : 1];
    double invr[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int cand = grp * R + r;
      invr[r] = inv_s[cand];
      const float* src = XS + (size_t)cand * D + q;
#pragma unroll
      for (int t = 0; t < CL; ++t) {
        const float xf = src[4 * ((16 * bb + t) ^ (cand & (CH - 1)))];
        if (t < TR) {
          const double x = (double)xf;
          fr[r][t < TR ? t : 0] = do_norm ? __dmul_rn(__dmul_rn(x, invr[r]), kInvSqrt2c) : __dmul_rn(x, kInvSqrt2c);
        } else {
          xt[r][t >= TR ? t - TR : 0] = xf;
        }
      }
    }
    if (Cfg::kSlots > 0) {
      __syncthreads();
#pragma unroll
      for (int r = 0; r < R; ++r) {
#pragma unroll
        for (int t = TR; t < CL; ++t) {
          const double x = (double)xt[r][t - TR];
          F[f_idx(r, t)] = do_norm ? __dmul_rn(__dmul_rn(x, invr[r]), kInvSqrt2c) : __dmul_rn(x, kInvSqrt2c);
        }
      }
    }
  }
  const double cc = __dmul_rn(kInvSqrt2c, kInvSqrt2c);  // product of the constant feature with itself

  // S[j][i] in gonum Dgemm(NoTrans,Trans) order: per 64-wide k block DotUnitary = (s0+s2)+(s1+s3), block sums added to
  // C in block order from +0.  g = the other item's feature record (nullptr: the diagonal).  Every lane of the group
  // ends with the R dot products in S[0..R).
  //
  // WITH_SS (windows of at most 10, i.e. k <= 8): the k-term sequential sum <c_j, c_i> of the update step rides in the
  // same straight-line code, one row per loop step, so its latency hides under the Gram row instead of preceding it.
  double S[R];
  auto gram = [&](const double* g, auto with_ss, const double* wcj, int k, int c_it, double& ss) {
    constexpr bool WITH_SS = decltype(with_ss)::value;
    constexpr int LPI = 8 / (CL / 2);  // rows of C per loop step (1 for chains of 16, 2 for chains of 8)
    double acc[R];
#pragma unroll
    for (int r = 0; r < R; ++r) acc[r] = 0.0;
#pragma unroll
    for (int t = 0; t < CL; t += 2) {
      double2 g2 = make_double2(0.0, 0.0);
      if (g) g2 = *reinterpret_cast<const double2*>(g + lam * FS + t);
      if (WITH_SS) {
#pragma unroll
        for (int u = 0; u < LPI; ++u) {
          const int l = (t / 2) * LPI + u;
          // rows l >= k are read although the owners write row k in this very step; the value is discarded (tm == 0)
          // — racecheck reports this read/write pair as a warning; predicating the load costs 4 % of the kernel
          const double tm = (l < k) ? wcj[l] : 0.0;
          const double pr = __dmul_rn(tm, C[l * kClItems + c_it]);
          ss = (tm != 0) ? __dadd_rn(ss, pr) : ss;
        }
      }
#pragma unroll
      for (int r = 0; r < R; ++r) {
        double fa, fb;
        if (t < TR) {
          fa = fr[r][t < TR ? t : 0];
          fb = fr[r][t + 1 < TR ? t + 1 : 0];
        } else {
          const double2 f2 = *reinterpret_cast<const double2*>(F + f_idx(r, t));
          fa = f2.x;
          fb = f2.y;
        }
        acc[r] = __dadd_rn(acc[r], __dmul_rn(g ? g2.x : fa, fa));
        acc[r] = __dadd_rn(acc[r], __dmul_rn(g ? g2.y : fb, fb));
      }
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
      double u = acc[r];
      if (!Cfg::kConstOwnBlock && q == 0 && bb == Cfg::kBlocks - 1) u = __dadd_rn(u, cc);  // tail element -> s0
      double o = shfl_xor_f64(u, 2);
      const double pr = (q & 2) ? __dadd_rn(o, u) : __dadd_rn(u, o);      // s0+s2 | s1+s3
      o = shfl_xor_f64(pr, 1);
      const double bs = (q & 1) ? __dadd_rn(o, pr) : __dadd_rn(pr, o);    // (s0+s2)+(s1+s3)
      double tot;
      if (Cfg::kBlocks == 2) {
        o = shfl_xor_f64(bs, 4);
        tot = __dadd_rn(__dadd_rn(0.0, bb ? o : bs), bb ? bs : o);
      } else {
        tot = __dadd_rn(0.0, bs);
      }
      if (Cfg::kConstOwnBlock) tot = __dadd_rn(tot, cc);
      S[r] = tot;
    }
  };
  auto pick_own = [&]() -> double {  // the dot product of the candidate this lane owns (lam < R)
    double v = S[0];
#pragma unroll
    for (int r = 1; r < R; ++r) v = (lam == r) ? S[r] : v;
    return v;
  };

  const bool owner = lam < R;                      // lane r of a group owns candidate grp*R + r: d2, C, quality
  const int it = grp * R + (owner ? lam : 0);
  const int gi = (int)rank * kClItems + it;        // index in the truncated list
  const bool active = owner && gi < m;
  const double qi = q_s[it];
  DPP_CLK(2);
  {
    double unused = 0.0;
    gram(nullptr, std::false_type{}, nullptr, 0, 0, unused);
  }
  const double diag = active ? __dmul_rn(__dmul_rn(qi, pick_own()), qi) : CUDART_NAN;
  DPP_CLK(3);

  // cluster-wide first-maximum argmax over the owners; every CTA ends up with the winner's record (pub[par][w], or its own srec[par]).
  // Two CTA barriers (per-warp maxima -> every warp reduces them redundantly; record staged) + one mbarrier wait.
  int par = 0;
  uint32_t mb_phase = 0;      // bit p: parity to wait for on mbar[p]
  cluster.barrier_wait();     // every CTA's mbarriers are initialised (arrive was before phase 2a)
  auto cluster_argmax = [&](double v, int krows) -> const CandRec* {
    uint64_t wk;
    int wi;
    warp_first_max(owner ? d2_key(v) : 0ull, it, &wk, &wi);
    if (lane == 0) { red_k[warp] = wk; red_i[warp] = wi; }
    DPP_CLK(6);
    __syncthreads();  // also: row k of C (written by the owners just before) is complete
    DPP_CLK(7);
    uint64_t bk;
    int li;
    warp_first_max(red_k[lane & 15], red_i[lane & 15], &bk, &li);
    if (bk == 0) li = 0;
    // publish this CTA's candidate `li`: the record is assembled in srec[par] (features by the winner's group straight
    // from registers / shared memory, its column of C by warp 15, the scalars by one thread), then ONE thread sends it to
    // the four CTAs as bulk copies that complete on the receivers' mbarriers — no remote stores, no release fence, no
    // cluster barrier on the critical path (the st.shared::cluster + cluster.sync version spent 23 % of the kernel in
    // UCGABAR / membar stalls).
    CandRec* mine = &srec[par];
    const int wg = li / R, wr = li - wg * R;
    if (grp == wg) {
#pragma unroll
      for (int r = 0; r < R; ++r)
        if (r == wr) {
#pragma unroll
          for (int t = 0; t < TR; t += 2) *reinterpret_cast<double2*>(mine->f + lam * FS + t) = make_double2(fr[r][t], fr[r][t + 1]);
        }
#pragma unroll
      for (int t = TR; t < CL; t += 2)
        *reinterpret_cast<double2*>(mine->f + lam * FS + t) = *reinterpret_cast<const double2*>(F + f_idx(wr, t));
    }
    if (tid >= kClThreads - 32 && tid - (kClThreads - 32) < krows) {  // the candidate's column of C
      const int l = tid - (kClThreads - 32);
      mine->cj[l] = C[l * kClItems + li];
    }
    if (tid == kClThreads - 33) {
      const double dv = d2_from_key(bk);
      mine->v = dv;
      mine->inv_dj = 1.0 / sqrt(dv);
      mine->key = bk;
      mine->q = q_s[li];
      mine->idx = (int)rank * kClItems + li;
      mine->row = row_s[li];
    }
    DPP_CLK(8);
    __syncthreads();
    if (tid == 0) {
      fence_proxy_async_smem();
      // three PEERS receive the record; this CTA reads its own straight from srec[par] (a shared::cluster bulk copy
      // addressed to the issuing CTA itself is flagged by compute-sanitizer: "not located in remote CTA")
      mbar_arrive_expect_tx(&mbar[par], (uint32_t)((kClCtas - 1) * sizeof(CandRec)));
      const uint32_t src = smem_u32(mine), dst = smem_u32(&pub[par * kClCtas + rank]), bar = smem_u32(&mbar[par]);
#pragma unroll
      for (int r = 0; r < kClCtas; ++r)
        if (r != (int)rank) bulk_copy_to_peer(mapa_u32(dst, r), src, (uint32_t)sizeof(CandRec), mapa_u32(bar, r));
    }
    mbar_wait_bounded(&mbar[par], (mb_phase >> par) & 1);
    mb_phase ^= 1u << par;
    DPP_CLK(9);
    auto rec_of = [&](int r) -> const CandRec* { return r == (int)rank ? mine : &pub[par * kClCtas + r]; };
    uint64_t best_k = rec_of(0)->key;
    int best_r = 0;
#pragma unroll
    for (int r = 1; r < kClCtas; ++r) {
      const uint64_t rk = rec_of(r)->key;
      if (rk > best_k) { best_k = rk; best_r = r; }  // lower rank == lower index wins ties
    }
    const CandRec* used = rec_of(best_r);
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
    const CandRec* wrec = cluster_argmax(d2, 0);
    int j = isnan(wrec->v) ? 0 : wrec->idx;
    if (tid == 0) res[total] = j;
    int ny = 1;
    bool broke = false;
    while (ny < top) {
      const CandRec& W = *wrec;
      double dj = W.v;  // == d2[j]; NaN when every candidate is used up (the reference then repeats index 0)
      if (dj < 1e-10) { broke = true; break; }
      const int k = ny - 1;
      // <c_j, c_i> over the k rows so far: sequential adds, zero terms skipped as in the reference
      double ss = 0.0;
      const double inv_dj = W.inv_dj;
      const double q_j = W.q;
      DPP_CLK(10);
      if (small_window) {
        gram(W.f, std::true_type{}, W.cj, k, it, ss);
      } else {
        if (active) {
#pragma unroll 1
          for (int l0 = 0; l0 < k; l0 += 4) {
            double tm[4], cv[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const bool in = l0 + u < k;
              tm[u] = in ? W.cj[l0 + u] : 0.0;
              cv[u] = in ? C[(l0 + u) * kClItems + it] : 0.0;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const double pr = __dmul_rn(tm[u], cv[u]);
              ss = (tm[u] != 0) ? __dadd_rn(ss, pr) : ss;
            }
          }
        }
        gram(W.f, std::false_type{}, nullptr, 0, 0, ss);
      }
      DPP_CLK(4);
      if (owner) {
        if (active) {
          const double Lji = __dmul_rn(__dmul_rn(q_j, pick_own()), qi);
          const double e = (k == 0) ? __dmul_rn(inv_dj, Lji) : __dmul_rn(inv_dj, __dsub_rn(Lji, ss));
          C[k * kClItems + it] = e;
          d2 = __dsub_rn(d2, __dmul_rn(e, e));
        }
        if (gi == j) d2 = CUDART_NAN;
      }
      DPP_CLK(5);
      wrec = cluster_argmax(d2, ny);  // its CTA barrier also orders the C[k] writes before the column reads
      j = isnan(wrec->v) ? 0 : wrec->idx;
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
  DPP_CLK(11);
#ifdef PRG_DPP_PROFILE
  if (blockIdx.x == 0 && tid == 0)
    for (int i = 0; i < 16; ++i) g_dpp_clk[i] += s_clk[i];
#endif
  if (rank == 0) {
    for (int t = tid; t < total; t += kClThreads) a.out_idx[(size_t)b * T_out + t] = order[res[t]];
    if (tid == 0) { a.out_n[b] = total; a.status[b] = 0; }
    dpp_write_final(a, b, tid, 0, total, order, res);
  }
  cluster.sync();  // no CTA may exit while peers can still write into its shared memory
}

// 1 / ||row|| for every row of an f32 table, gonum floats.Norm (scaled sum of squares, true divisions) then
// 1 / (scale * sqrt(sumsq)) — the value DPPSort's normalisation divides by (sort/dpp_sort.go:372-381).  One thread per row;
// runs once per prg_set_diversity_matrix.
__global__ void dpp_inv_norm_kernel(const float* __restrict__ D, uint64_t rows, int dim, double* __restrict__ out) {
  const uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  const float* x = D + r * dim;
  double scale = 0.0, sumsq = 1.0;
  for (int d = 0; d < dim; ++d) {
    const double v = (double)x[d];
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
  out[r] = 1.0 / __dmul_rn(scale, sqrt(sumsq));
}

__global__ void dpp_substitute_kernel(float* __restrict__ out, int dim) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < (int)kDppSubRows * dim) out[i] = dpp_substitute_value((uint32_t)(i / dim), (uint32_t)(i % dim));
}

int dpp_cluster_prepare(prg_handle* h) {
  if (h->D_rows > (uint64_t)kDppMissBase) return fail(PRG_EUNSUPPORTED, "diversity table has more than 0xFFFFF000 rows");
  // substitute directions for candidates without a table row (any table dtype; the values are f32)
  PRG_TRY(h->D_sub.ensure((size_t)kDppSubRows * h->D_dim * 4));
  PRG_TRY(h->D_sub_inv.ensure((size_t)kDppSubRows * 8));
  dpp_substitute_kernel<<<(kDppSubRows * h->D_dim + 255) / 256, 256, 0, h->stream>>>((float*)h->D_sub.p, (int)h->D_dim);
  dpp_inv_norm_kernel<<<(kDppSubRows + 255) / 256, 256, 0, h->stream>>>((const float*)h->D_sub.p, kDppSubRows, (int)h->D_dim,
                                                                      (double*)h->D_sub_inv.p);
  PRG_CUDA(cudaGetLastError());
  if (h->D_dtype != PRG_F32) {  // f64 tables take the generic kernel, which normalises on the fly
    PRG_CUDA(cudaStreamSynchronize(h->stream));
    return PRG_OK;
  }
  PRG_TRY(h->D_inv.ensure((size_t)h->D_rows * 8));
  const unsigned blocks = (unsigned)((h->D_rows + 255) / 256);
  dpp_inv_norm_kernel<<<blocks, 256, 0, h->stream>>>((const float*)h->D, h->D_rows, (int)h->D_dim, (double*)h->D_inv.p);
  PRG_CUDA(cudaGetLastError());
  PRG_CUDA(cudaStreamSynchronize(h->stream));
  return PRG_OK;
}

template <int D>
static size_t dpp_cluster_smem(int top_n, int c_rows) {
  using Cfg = ClCfg<D>;
  size_t live = (size_t)c_rows * kClItems * 8 + (size_t)Cfg::kFDoubles * 8 + (2 * kClCtas + 2) * sizeof(CandRecT<D>) + 16 +
                2 * kClItems * 8 + 16 * 8 + 16 * 4 + kClItems * 4 + kClMaxItems * 4 +
                (size_t)((top_n + 3) & ~3) * 4 + kClMaxItems + 64;
  const size_t presort_staging = (size_t)kClMaxN * 12;
  return live > presort_staging ? live : presort_staging;
}

template <int D>
static int launch_cluster(prg_handle* h, const DppClArgs& a, int B) {
  const int window = a.p.window_size > 0 ? a.p.window_size : 10;
  const size_t smem = dpp_cluster_smem<D>(a.p.top_n, a.p.top_n <= window ? a.p.top_n : window);
  PRG_CUDA(cudaFuncSetAttribute(dpp_cluster_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  PRG_CUDA(launch_chained(h, dpp_cluster_kernel<D>, dim3((unsigned)(B * kClCtas)), dim3(kClThreads), smem, kClCtas, a));
  count_launch(h);
  return PRG_OK;
}

// returns PRG_OK and sets *handled when the request shape is served by the cluster kernel
int dpp_cluster_device(prg_handle* h, const uint32_t* rows_dev, const double* score_dev, int B, int n,
                       const prg_dpp_params& p, int32_t* out_idx, int32_t* out_n, int32_t* status, bool* handled,
                       DppFinal* fin) {
  *handled = false;
  if (h->D_dtype != PRG_F32) return PRG_OK;
  if (h->D_dim != 32 && h->D_dim != 64 && h->D_dim != 128) return PRG_OK;
  if ((reinterpret_cast<uintptr_t>(h->D) & 15) != 0) return PRG_OK;
  if (n > kClMaxN || p.top_n > 2048) return PRG_OK;
  const int window = p.window_size > 0 ? p.window_size : 10;
  const int c_rows = p.top_n <= window ? p.top_n : window;
  if (c_rows > kClCRows) return PRG_OK;
  DppClArgs a{};
  a.rows = rows_dev; a.score = score_dev; a.n = n; a.D = (const float*)h->D; a.D_inv = (const double*)h->D_inv.p; a.D_rows = h->D_rows; a.p = p;
  a.D_sub = (const float*)h->D_sub.p; a.D_sub_inv = (const double*)h->D_sub_inv.p;
  a.out_idx = out_idx; a.out_n = out_n; a.status = status;
  if (fin) { a.fin_row = fin->row; a.fin_score = fin->score; a.fin_n = fin->n; }
  StageScope span(h, ST_DPP);
  int rc = PRG_OK;
  if (h->D_dim == 32) rc = launch_cluster<32>(h, a, B);
  else if (h->D_dim == 64) rc = launch_cluster<64>(h, a, B);
  else rc = launch_cluster<128>(h, a, B);
  if (rc == PRG_OK) *handled = true;
  if (rc == PRG_OK && fin) fin->done = true;
  return rc;
}

#ifdef PRG_DPP_PROFILE
extern "C" void prg_debug_dpp_clocks(long long* out, int reset) {
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(out, g_dpp_clk, sizeof(long long) * 16);
  if (reset) { long long z[16] = {0}; cudaMemcpyToSymbol(g_dpp_clk, z, sizeof(z)); }
}
#endif

}  // namespace prg
