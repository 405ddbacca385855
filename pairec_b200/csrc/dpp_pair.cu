// dpp_pair.cu — the default DPP kernel for 128-d f32 diversity tables (config "dpp_pair": 0 or PRG_DPP_PAIR=0 selects the
// 4-CTA cluster kernel of dpp_cluster.cu; verified on the B200 in round 2: same sequences, 0.283 -> 0.192 ms per batch):
// the DPP cluster kernel (dpp_cluster.cu; sort/dpp_sort.go:271-351, :372-475, :477-551) re-cut so that a whole
// 64-request batch is ONE wave.
//
// Why: dpp_cluster.cu gives a request a cluster of 4 CTAs (256 candidates each) because one SM's registers + shared
// memory hold the fp64 features of 256 candidates and no more.  64 requests x 4 CTAs = 256 CTAs on 148 SMs: two waves
// of ≈ 0.14 ms, and the stage is a third of the step.  An SM has a third on-chip store the kernel does not use: 256 KB
// of TENSOR MEMORY.  Here a request gets a cluster of 2 CTAs with 512 candidates each (64 x 2 = 128 CTAs: one wave),
// and a candidate's 128 features live in three places:
//     chain positions 0..3   registers                    (32 f64 per thread, as before)
//     chain positions 4..7   shared memory                (128 KB per CTA)
//     chain positions 8..15  tensor memory                (all 512 columns: 64 f64 per thread, thread-private: a warp
//                                                          reads and writes only its own 32 lanes, tcgen05.st / .ld)
// Tensor memory is used as plain scratch: no MMA touches it.  Per selection step a CTA then reads 128 KB from shared
// memory (≈ 1 k cycles) and 256 KB from tensor memory (≈ 4 k cycles at the 64 B/clk of the microarchitecture notes)
// against 2 760 cycles of Gram row today — per step slower, per batch one wave instead of two (expected ≈ 0.17 ms
// instead of 0.28 ms; to be measured).
//
// Everything else — arithmetic order, argmax / publication protocol, window logic, outputs — is dpp_cluster.cu's, with
// R = 8 candidates per 8-lane group (every lane owns one candidate).  Shapes: f32 table, dim 128, windows of at most
// 10; anything else falls through to the cluster kernel.
#include "dpp_common.cuh"
#include <cooperative_groups.h>
#include <type_traits>

namespace cg = cooperative_groups;

namespace prg {

namespace {

constexpr int kPrCtas = 2;
constexpr int kPrItems = 512;                    // candidates per CTA
constexpr int kPrThreads = 512;
constexpr int kPrMaxItems = kPrCtas * kPrItems;  // 1024
constexpr int kPrD = 128;
constexpr int kPrLPC = 8;                        // lanes per group: (k block bb, DotUnitary chain q)
constexpr int kPrR = 8;                          // candidates per group — every lane owns one
constexpr int kPrCL = 16;                        // chain length: features per (candidate, lane)
constexpr int kPrTR = 4;                         // chain positions [0, 4) in registers
constexpr int kPrTS = 8;                         // [4, 8) in shared memory, [8, 16) in tensor memory
constexpr int kPrFS = kPrCL + 2;                 // lane stride (doubles) of a feature record
constexpr int kPrCRows = 10;                     // rows of C: windows of at most 10
constexpr int kPrFDoubles = kPrR * (kPrTS - kPrTR) * kPrThreads;  // 16384 doubles = 128 KB (also the f32 staging of 256 rows)
constexpr uint32_t kPrTmemCols = 512;

struct __align__(16) PairRec {
  double v;       // d2 of the candidate (NaN if the CTA has none)
  uint64_t key;   // its order key (0 = none)
  double q;       // exp(alpha * rel)
  double inv_dj;  // 1 / sqrt(d2)
  int32_t idx;    // index in the truncated list
  uint32_t row;   // diversity-table row (diagnostics)
  double pad_;
  double cj[kPrCRows + 2];
  double f[kPrLPC * kPrFS];  // features by (lane of the group, chain position); the constant feature is implied
};

__device__ __forceinline__ uint32_t pr_mapa(uint32_t cta_smem_addr, uint32_t cta_rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(cta_smem_addr), "r"(cta_rank));
  return r;
}
__device__ __forceinline__ void pr_bulk_copy_to_peer(uint32_t dst_cluster_addr, uint32_t src_cta_addr, uint32_t bytes,
                                                     uint32_t mbar_cluster_addr) {
  asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   dst_cluster_addr),
               "r"(src_cta_addr), "r"(bytes), "r"(mbar_cluster_addr)
               : "memory");
}
__device__ __forceinline__ void pr_fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// bounded wait: a protocol bug must trap, not hang the GPU
__device__ __forceinline__ void pr_mbar_wait_bounded(uint64_t* bar, uint32_t parity) {
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

// tensor memory as thread-private scratch: 16 consecutive 32-bit columns of the calling thread's lane
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
      "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr));
}
// wait for the outstanding tensor-memory loads; the registers are operands so that no use of them is scheduled above
__device__ __forceinline__ void tmem_ld_wait16(uint32_t (&v)[16]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]), "+r"(v[8]),
                 "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15])
               :
               : "memory");
}
__device__ __forceinline__ double pack_f64(uint32_t lo, uint32_t hi) { return __hiloint2double((int)hi, (int)lo); }

// Final outputs of the fused request path (dpp_cluster.cu dpp_write_final; pipeline.cu final_gather_kernel semantics)
__device__ __forceinline__ void pair_write_final(const DppClArgs& a, int b, int tid, int st, int total, const int32_t* order,
                                                 const int32_t* res) {
  if (!a.fin_row) return;
  const int T = a.p.top_n, n = a.n;
  const bool unchanged = st != 0;
  const int c = unchanged ? (n < T ? n : T) : total;
  for (int t = tid; t < T; t += kPrThreads) {
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

}  // namespace

__global__ void __launch_bounds__(kPrThreads, 1) dpp_pair_kernel(const DppClArgs a) {
  pdl_wait();                 // chained launch: the predecessor's writes are visible from here on
  pdl_launch_dependents();
  constexpr int D = kPrD, LPC = kPrLPC, R = kPrR, CL = kPrCL, TR = kPrTR, TS = kPrTS, FS = kPrFS;
  cg::cluster_group cluster = cg::this_cluster();
  const unsigned rank = cluster.block_rank();
  const int b = blockIdx.x / kPrCtas;
  extern __shared__ __align__(16) uint8_t psm[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int grp = tid / LPC, lam = tid % LPC;   // group of 8 lanes serves candidates grp*8 .. grp*8+7 of this CTA
  const int q = lam & 3, bb = lam >> 2;         // this lane: DotUnitary chain q of k block bb, for each of the R candidates
  const int n = a.n, T_out = a.p.top_n;
  const int window = a.p.window_size > 0 ? a.p.window_size : 10;
  const int c_rows = T_out <= window ? T_out : window;   // <= kPrCRows (checked by the launcher)

  double* C = reinterpret_cast<double*>(psm);                                   // [c_rows][512]
  double* F = C + (size_t)c_rows * kPrItems;                                    // [16 double2 slots][512 threads]
  PairRec* pub = reinterpret_cast<PairRec*>(F + kPrFDoubles);                   // [2][2]
  PairRec* srec = pub + 2 * kPrCtas;                                            // [2] this CTA's own record, staged for the copy
  uint64_t* mbar = reinterpret_cast<uint64_t*>(srec + 2);                       // [2] one per record buffer
  double* inv_s = reinterpret_cast<double*>(mbar + 2);                          // [512]
  double* q_s = inv_s + kPrItems;                                               // [512]
  uint64_t* red_k = reinterpret_cast<uint64_t*>(q_s + kPrItems);                // [16]
  int32_t* red_i = reinterpret_cast<int32_t*>(red_k + 16);                      // [16]
  uint32_t* row_s = reinterpret_cast<uint32_t*>(red_i + 16);                    // [512]
  int32_t* order = reinterpret_cast<int32_t*>(row_s + kPrItems);                // [1024]
  int32_t* res = order + kPrMaxItems;                                           // [T_out]
  uint8_t* existed = reinterpret_cast<uint8_t*>(res + ((T_out + 3) & ~3));      // [1024]
  __shared__ int s_m, s_err, s_ny;
  __shared__ double s_p0, s_p1;
  __shared__ uint32_t s_tmem;

  const uint32_t* rows = a.rows + (size_t)b * n;
  const double* score = a.score + (size_t)b * n;

  // ---- 0. valid count, optional presort + truncation (:280-300); done redundantly by every CTA of the cluster
  if (tid == 0) { s_m = 0; s_err = 0; }
  __syncthreads();
  {
    int cnt = 0;
    for (int i = tid; i < n; i += kPrThreads) cnt += (rows[i] != 0xFFFFFFFFu);
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
    uint64_t* key = reinterpret_cast<uint64_t*>(psm);  // staging over C / F (not live yet); 12 B x P2 <= 48 KiB
    int32_t* idx = reinterpret_cast<int32_t*>(key + P2);
    for (uint32_t i = tid; i < P2; i += kPrThreads) {
      key[i] = (i < (uint32_t)nv) ? f64_ord_c(score[i]) : 0ull;
      idx[i] = (i < (uint32_t)nv) ? (int32_t)i : 0x7FFFFFFF;
    }
    __syncthreads();
    for (uint32_t size = 2; size <= P2; size <<= 1) {
      for (uint32_t stride = size >> 1; stride > 0; stride >>= 1) {
        for (uint32_t i = tid; i < (P2 >> 1); i += kPrThreads) {
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
    if (m <= kPrMaxItems)
      for (int i = tid; i < m; i += kPrThreads) order[i] = idx[i];
    __syncthreads();
  } else if (m <= kPrMaxItems) {
    for (int i = tid; i < m; i += kPrThreads) order[i] = i;
  }
  if (nv == 0 || m > kPrMaxItems) {  // uniform across the cluster: every CTA sees the same request
    if (rank == 0 && tid == 0) { a.out_n[b] = 0; a.status[b] = (nv == 0) ? 0 : 2; }
    if (rank == 0) pair_write_final(a, b, tid, (nv == 0) ? 0 : 2, 0, nullptr, nullptr);
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
    if (rank == 0) pair_write_final(a, b, tid, 1, 0, nullptr, nullptr);
    return;
  }

  // ---- 2a. barriers, tensor memory (all 512 columns: one CTA per SM), rows, relevance / 1/norm / quality
  if (tid == 0) {  // (past the last early return: every CTA of the cluster gets here)
    mbar_init(&mbar[0], 1);
    mbar_init(&mbar[1], 1);
    mbar_fence_init();
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "n"(kPrTmemCols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  cluster.barrier_arrive();  // peers may signal our mbarriers only after this; waited for before the first publication
  for (int i = tid; i < kPrItems; i += kPrThreads) {
    const int c = (int)rank * kPrItems + i;
    const bool act = c < m;
    const uint32_t r = act ? rows[order[c]] : 0u;
    const uint32_t rw = dpp_row_code(act, r, a.D_rows, act ? order[c] : 0);
    row_s[i] = rw;
    double rel = act ? score[order[c]] : 0.0;
    if (a.p.norm_mode == 1) rel = __dsub_rn(rel, s_p0) / s_p1;
    else if (a.p.norm_mode == 2) rel = __dadd_rn(__dmul_rn(__dsub_rn(rel, s_p0) / s_p1, 1 - 1e-6), 1e-6);
    // 1 / ||e|| from the per-row table built when the matrix was set; a candidate without a table row takes a
    // substitute direction and its norm (dpp_common.cuh)
    inv_s[i] = !a.p.normalize_emb ? 1.0 : (rw != 0xFFFFFFFFu ? dpp_row_inv(a, rw) : 1.0);
    q_s[i] = act ? exp(__dmul_rn(a.p.alpha, rel)) : 0.0;
  }
  for (int i = tid; i < kPrMaxItems; i += kPrThreads) existed[i] = 0;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  // this thread's tensor-memory scratch: its own lane, 128 columns; warps w, w+4, w+8, w+12 share a lane quarter
  const uint32_t tbase = s_tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * 128);

  // ---- 2b. the rounded fp64 features, built once.  Lane (bb, q) of a group holds, for each of the group's 8
  //          candidates, chain q of block bb: features 64*bb + 4*t + q, t < 16.  The 512 embeddings are staged through
  //          shared memory in two halves of 256 rows (the staging area IS the F region: 128 KB); the threads of a half
  //          take their values out of it — register features at once, tensor-memory features stored at once,
  //          shared-memory features parked as f32 in registers until both halves are through.
  //          Tensor-memory layout of a thread: granule (rh, tp) = 16 columns = candidates rh*4 .. rh*4+3 x chain
  //          positions 8+2tp, 9+2tp; granule index rh*4 + tp.
  constexpr int CH = D / 4;  // 16-B chunks per embedding
  float* XS = reinterpret_cast<float*>(F);
  double fr[R][TR];
  float xt[R][TS - TR];
  double invr[R];
  const bool do_norm = a.p.normalize_emb != 0;
  auto feat = [&](float xf, double inv) -> double {
    const double x = (double)xf;
    return do_norm ? __dmul_rn(__dmul_rn(x, inv), kInvSqrt2c) : __dmul_rn(x, kInvSqrt2c);
  };
#pragma unroll 1
  for (int hh = 0; hh < 2; ++hh) {
    for (int g = tid; g < 256 * CH; g += kPrThreads) {
      const int row = g / CH, c = g % CH;
      const uint32_t rw = row_s[hh * 256 + row];
      const float4 v = (rw != 0xFFFFFFFFu) ? reinterpret_cast<const float4*>(dpp_row_ptr(a, rw, D))[c] : make_float4(0.f, 0.f, 0.f, 0.f);
      *reinterpret_cast<float4*>(XS + (size_t)row * D + 4 * (c ^ (row & (CH - 1)))) = v;
    }
    __syncthreads();
    if ((tid >> 8) == hh) {   // warp-uniform: warps 0-7 own the first 256 candidates, warps 8-15 the rest
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const int cand = grp * R + r;       // CTA-local candidate
        const int cl = cand - hh * 256;     // its row in the staging area
        invr[r] = inv_s[cand];
        const float* src = XS + (size_t)cl * D + q;
#pragma unroll
        for (int t = 0; t < TS; ++t) {
          const float xf = src[4 * ((16 * bb + t) ^ (cl & (CH - 1)))];
          if (t < TR) fr[r][t < TR ? t : 0] = feat(xf, invr[r]);
          else xt[r][t >= TR ? t - TR : 0] = xf;
        }
      }
#pragma unroll
      for (int rh = 0; rh < 2; ++rh) {
#pragma unroll
        for (int tp = 0; tp < 4; ++tp) {
          uint32_t v[16];
#pragma unroll
          for (int r4 = 0; r4 < 4; ++r4) {
            const int r = rh * 4 + r4;
            const int cl = grp * R + r - hh * 256;
            const float* src = XS + (size_t)cl * D + q;
#pragma unroll
            for (int u = 0; u < 2; ++u) {
              const int t = TS + 2 * tp + u;
              const double f = feat(src[4 * ((16 * bb + t) ^ (cl & (CH - 1)))], invr[r]);
              v[4 * r4 + 2 * u] = (uint32_t)__double2loint(f);
              v[4 * r4 + 2 * u + 1] = (uint32_t)__double2hiint(f);
            }
          }
          tmem_st16(tbase + (uint32_t)((rh * 4 + tp) * 16), v);
        }
      }
      tmem_st_wait();
    }
    __syncthreads();   // the staging area is free again (second half / the F slots below)
  }
  // shared-memory features: chain positions (t, t+1), t in {4, 6}, of candidate r form one 16-byte element [slot][tid]
  auto f_idx = [&](int r, int t) -> size_t {
    return ((size_t)(r * ((TS - TR) / 2) + ((t - TR) >> 1)) * kPrThreads + tid) * 2 + ((t - TR) & 1);
  };
#pragma unroll
  for (int r = 0; r < R; ++r) {
#pragma unroll
    for (int t = TR; t < TS; ++t) F[f_idx(r, t)] = feat(xt[r][t - TR], invr[r]);
  }
  const double cc = __dmul_rn(kInvSqrt2c, kInvSqrt2c);  // product of the constant feature with itself

  // S[j][i] in gonum Dgemm(NoTrans,Trans) order: per 64-wide k block DotUnitary = (s0+s2)+(s1+s3), block sums added to
  // C in block order from +0, then the constant feature's block.  g = the other item's feature record (nullptr: the
  // diagonal).  Returns the dot product of the candidate THIS lane owns (candidate lam of the group).
  // The k-term sequential sum <c_j, c_i> of the update step rides along in the first half's straight-line code, one
  // row of C per pair of chain positions (k <= 8 <= CL / 2).
  auto gram = [&](const double* g, auto with_ss, const double* wcj, int k, int c_it, double& ss) -> double {
    constexpr bool WITH_SS = decltype(with_ss)::value;
    double own = 0.0;
#pragma unroll
    for (int rh = 0; rh < 2; ++rh) {
      double acc[4] = {0.0, 0.0, 0.0, 0.0};
      uint32_t ta[16], tb[16];
      tmem_ld16_nowait(tbase + (uint32_t)((rh * 4 + 0) * 16), ta);   // first tensor-memory granule: needed at t = 8
      auto ss_row = [&](int t) {
        if (WITH_SS && rh == 0) {
          const int l = t / 2;
          // rows l >= k are read although the owners write row k in this very step; the value is discarded (tm == 0)
          const double tm = (l < k) ? wcj[l] : 0.0;
          const double pr = __dmul_rn(tm, C[l * kPrItems + c_it]);
          ss = (tm != 0) ? __dadd_rn(ss, pr) : ss;
        }
      };
#pragma unroll
      for (int t = 0; t < TS; t += 2) {   // registers, then shared memory
        double2 g2 = make_double2(0.0, 0.0);
        if (g) g2 = *reinterpret_cast<const double2*>(g + lam * FS + t);
        ss_row(t);
#pragma unroll
        for (int r4 = 0; r4 < 4; ++r4) {
          const int r = rh * 4 + r4;
          double fa, fb;
          if (t < TR) {
            fa = fr[r][t < TR ? t : 0];
            fb = fr[r][t + 1 < TR ? t + 1 : 0];
          } else {
            const double2 f2 = *reinterpret_cast<const double2*>(F + f_idx(r, t));
            fa = f2.x;
            fb = f2.y;
          }
          acc[r4] = __dadd_rn(acc[r4], __dmul_rn(g ? g2.x : fa, fa));
          acc[r4] = __dadd_rn(acc[r4], __dmul_rn(g ? g2.y : fb, fb));
        }
      }
#pragma unroll
      for (int tp = 0; tp < 4; ++tp) {    // tensor memory: granule tp is in flight; request tp + 1 before using it
        const int t = TS + 2 * tp;
        double2 g2 = make_double2(0.0, 0.0);
        if (g) g2 = *reinterpret_cast<const double2*>(g + lam * FS + t);
        ss_row(t);
        uint32_t (&cur)[16] = (tp & 1) ? tb : ta;
        uint32_t (&nxt)[16] = (tp & 1) ? ta : tb;
        tmem_ld_wait16(cur);
        double fa[4], fb[4];
#pragma unroll
        for (int r4 = 0; r4 < 4; ++r4) {
          fa[r4] = pack_f64(cur[4 * r4], cur[4 * r4 + 1]);
          fb[r4] = pack_f64(cur[4 * r4 + 2], cur[4 * r4 + 3]);
        }
        if (tp < 3) tmem_ld16_nowait(tbase + (uint32_t)((rh * 4 + tp + 1) * 16), nxt);
#pragma unroll
        for (int r4 = 0; r4 < 4; ++r4) {
          acc[r4] = __dadd_rn(acc[r4], __dmul_rn(g ? g2.x : fa[r4], fa[r4]));
          acc[r4] = __dadd_rn(acc[r4], __dmul_rn(g ? g2.y : fb[r4], fb[r4]));
        }
      }
#pragma unroll
      for (int r4 = 0; r4 < 4; ++r4) {
        const double u = acc[r4];
        double o = shfl_xor_f64(u, 2);
        const double pr = (q & 2) ? __dadd_rn(o, u) : __dadd_rn(u, o);      // s0+s2 | s1+s3
        o = shfl_xor_f64(pr, 1);
        const double bs = (q & 1) ? __dadd_rn(o, pr) : __dadd_rn(pr, o);    // (s0+s2)+(s1+s3)
        o = shfl_xor_f64(bs, 4);
        double tot = __dadd_rn(__dadd_rn(0.0, bb ? o : bs), bb ? bs : o);   // block 0, then block 1
        tot = __dadd_rn(tot, cc);                                           // the constant feature's own block
        own = (lam == rh * 4 + r4) ? tot : own;
      }
    }
    return own;
  };

  const int it = grp * R + lam;                    // every lane owns one candidate: d2, C, quality
  const int gi = (int)rank * kPrItems + it;        // index in the truncated list
  const bool active = gi < m;
  const double qi = q_s[it];
  double diag;
  {
    double unused = 0.0;
    const double sii = gram(nullptr, std::false_type{}, nullptr, 0, 0, unused);
    diag = active ? __dmul_rn(__dmul_rn(qi, sii), qi) : CUDART_NAN;
  }

  // cluster-wide first-maximum argmax; every CTA ends up with the winner's record (pub[par][w], or its own srec[par])
  int par = 0;
  uint32_t mb_phase = 0;      // bit p: parity to wait for on mbar[p]
  cluster.barrier_wait();     // every CTA's mbarriers are initialised (arrive was before phase 2a)
  auto cluster_argmax = [&](double v, int krows) -> const PairRec* {
    uint64_t wk;
    int wi;
    warp_first_max(d2_key(v), it, &wk, &wi);
    if (lane == 0) { red_k[warp] = wk; red_i[warp] = wi; }
    __syncthreads();  // also: row k of C (written by the owners just before) is complete
    uint64_t bk;
    int li;
    warp_first_max(red_k[lane & 15], red_i[lane & 15], &bk, &li);
    if (bk == 0) li = 0;
    // publish this CTA's candidate `li`: its features come from the three stores of its group's lanes
    PairRec* mine = &srec[par];
    const int wg = li / R, wr = li - wg * R;
    if (warp == wg / 4) {   // the winner's warp, all 32 lanes: the tensor-memory loads are warp-collective
      const int rh = wr >> 2, r4w = wr & 3;
      double tf[CL - TS];
#pragma unroll
      for (int tp = 0; tp < 4; ++tp) {
        uint32_t tv[16];
        tmem_ld16_nowait(tbase + (uint32_t)((rh * 4 + tp) * 16), tv);
        tmem_ld_wait16(tv);
#pragma unroll
        for (int r4 = 0; r4 < 4; ++r4)
          if (r4 == r4w) {
            tf[2 * tp] = pack_f64(tv[4 * r4], tv[4 * r4 + 1]);
            tf[2 * tp + 1] = pack_f64(tv[4 * r4 + 2], tv[4 * r4 + 3]);
          }
      }
      if (grp == wg) {
#pragma unroll
        for (int r = 0; r < R; ++r)
          if (r == wr) {
#pragma unroll
            for (int t = 0; t < TR; t += 2) *reinterpret_cast<double2*>(mine->f + lam * FS + t) = make_double2(fr[r][t], fr[r][t + 1]);
          }
#pragma unroll
        for (int t = TR; t < TS; t += 2)
          *reinterpret_cast<double2*>(mine->f + lam * FS + t) = *reinterpret_cast<const double2*>(F + f_idx(wr, t));
#pragma unroll
        for (int t = TS; t < CL; t += 2)
          *reinterpret_cast<double2*>(mine->f + lam * FS + t) = make_double2(tf[t - TS], tf[t - TS + 1]);
      }
    }
    if (tid >= kPrThreads - 32 && tid - (kPrThreads - 32) < krows) {  // the candidate's column of C
      const int l = tid - (kPrThreads - 32);
      mine->cj[l] = C[l * kPrItems + li];
    }
    if (tid == kPrThreads - 33) {
      const double dv = d2_from_key(bk);
      mine->v = dv;
      mine->inv_dj = 1.0 / sqrt(dv);
      mine->key = bk;
      mine->q = q_s[li];
      mine->idx = (int)rank * kPrItems + li;
      mine->row = row_s[li];
    }
    __syncthreads();
    if (tid == 0) {
      pr_fence_proxy_async_smem();
      // the PEER receives the record; this CTA reads its own straight from srec[par]
      mbar_arrive_expect_tx(&mbar[par], (uint32_t)((kPrCtas - 1) * sizeof(PairRec)));
      const uint32_t src = smem_u32(mine), dst = smem_u32(&pub[par * kPrCtas + rank]), bar = smem_u32(&mbar[par]);
#pragma unroll
      for (int r = 0; r < kPrCtas; ++r)
        if (r != (int)rank) pr_bulk_copy_to_peer(pr_mapa(dst, r), src, (uint32_t)sizeof(PairRec), pr_mapa(bar, r));
    }
    pr_mbar_wait_bounded(&mbar[par], (mb_phase >> par) & 1);
    mb_phase ^= 1u << par;
    auto rec_of = [&](int r) -> const PairRec* { return r == (int)rank ? mine : &pub[par * kPrCtas + r]; };
    const PairRec* r0 = rec_of(0);
    const PairRec* r1 = rec_of(1);
    const PairRec* used = (r1->key > r0->key) ? r1 : r0;  // lower rank == lower index wins ties
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
    const PairRec* wrec = cluster_argmax(d2, 0);
    int j = isnan(wrec->v) ? 0 : wrec->idx;
    if (tid == 0) res[total] = j;
    int ny = 1;
    bool broke = false;
    while (ny < top) {
      const PairRec& W = *wrec;
      const double dj = W.v;  // == d2[j]; NaN when every candidate is used up (the reference then repeats index 0)
      if (dj < 1e-10) { broke = true; break; }
      const int k = ny - 1;
      double ss = 0.0;        // <c_j, c_i> over the k rows so far: sequential adds, zero terms skipped as in the reference
      const double inv_dj = W.inv_dj;
      const double q_j = W.q;
      const double sji = gram(W.f, std::true_type{}, W.cj, k, it, ss);
      if (active) {
        const double Lji = __dmul_rn(__dmul_rn(q_j, sji), qi);
        const double e = (k == 0) ? __dmul_rn(inv_dj, Lji) : __dmul_rn(inv_dj, __dsub_rn(Lji, ss));
        C[k * kPrItems + it] = e;
        d2 = __dsub_rn(d2, __dmul_rn(e, e));
      }
      if (gi == j) d2 = CUDART_NAN;
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
  if (rank == 0) {
    for (int t = tid; t < total; t += kPrThreads) a.out_idx[(size_t)b * T_out + t] = order[res[t]];
    if (tid == 0) { a.out_n[b] = total; a.status[b] = 0; }
    pair_write_final(a, b, tid, 0, total, order, res);
  }
  tc_fence_before();
  cluster.sync();  // no CTA may exit while peers can still write into its shared memory
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(s_tmem), "n"(kPrTmemCols));
  }
}

static size_t dpp_pair_smem(int top_n, int c_rows) {
  size_t live = (size_t)c_rows * kPrItems * 8 + (size_t)kPrFDoubles * 8 + (2 * kPrCtas + 2) * sizeof(PairRec) + 16 +
                2 * kPrItems * 8 + 16 * 8 + 16 * 4 + kPrItems * 4 + kPrMaxItems * 4 + (size_t)((top_n + 3) & ~3) * 4 +
                kPrMaxItems + 64;
  const size_t presort_staging = (size_t)kClMaxN * 12;
  return live > presort_staging ? live : presort_staging;
}

// returns PRG_OK and sets *handled when the request shape is served by the pair kernel
int dpp_pair_device(prg_handle* h, const uint32_t* rows_dev, const double* score_dev, int B, int n,
                    const prg_dpp_params& p, int32_t* out_idx, int32_t* out_n, int32_t* status, bool* handled,
                    DppFinal* fin) {
  *handled = false;
  if (h->D_dtype != PRG_F32 || h->D_dim != kPrD || !h->D_inv.p) return PRG_OK;
  if ((reinterpret_cast<uintptr_t>(h->D) & 15) != 0) return PRG_OK;
  if (n > kClMaxN || p.top_n > 2048) return PRG_OK;
  const int window = p.window_size > 0 ? p.window_size : 10;
  const int c_rows = p.top_n <= window ? p.top_n : window;
  if (c_rows > kPrCRows) return PRG_OK;
  const size_t smem = dpp_pair_smem(p.top_n, c_rows);
  if (smem > 227 * 1024) return PRG_OK;
  DppClArgs a{};
  a.rows = rows_dev; a.score = score_dev; a.n = n; a.D = (const float*)h->D; a.D_inv = (const double*)h->D_inv.p;
  a.D_rows = h->D_rows; a.p = p;
  a.D_sub = (const float*)h->D_sub.p; a.D_sub_inv = (const double*)h->D_sub_inv.p;
  a.out_idx = out_idx; a.out_n = out_n; a.status = status;
  if (fin) { a.fin_row = fin->row; a.fin_score = fin->score; a.fin_n = fin->n; }
  StageScope span(h, ST_DPP);
  PRG_CUDA(cudaFuncSetAttribute(dpp_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  PRG_CUDA(launch_chained(h, dpp_pair_kernel, dim3((unsigned)(B * kPrCtas)), dim3(kPrThreads), smem, kPrCtas, a));
  count_launch(h);
  *handled = true;
  if (fin) fin->done = true;
  return PRG_OK;
}

}  // namespace prg
