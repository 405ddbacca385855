// group.cu — prg_group_*: the row-sharded request path over the G GPUs of one box INSIDE the library.
//
// The reference is one Go process (pairec.Run, pairec.go:61-86): it cannot call torch.distributed, and a cgo host that
// wants the item catalog sharded over the 8 GPUs of a box needs the whole exchange behind the C ABI.  One prg_group owns G
// prg_handles of ONE process — handle g holds row shard g of the item matrix (prg_set_item_matrix(..., row_base)) and
// replicas of the field / feature / diversity tables — and serves a batch of G*B requests per call:
//
//   every GPU       H2D of the whole batch (queries, user features)                       its own copy engine
//   phase 1  GPU g  sample keys of its shard for all G*B queries (prg_shard_sample)       kernels of recall.cu
//            GPU g  scatter_blocks_kernel: its block -> slot g of EVERY GPU's gather buffer   stores over NVLink (P2P)
//   phase 2  GPU g  tau from the G blocks; filter + exact re-score of its shard; lists of at most k keys per query
//            GPU g  scatter_lists_kernel: the VALID PREFIX of query q's list -> slot g of the buffer of q's owner
//                   (GPU q / B), plus one status word: an all-to-all of ~k/G keys per (shard, query) instead of the
//                   all-gather of k-padded lists the NCCL protocol moves (bench.py N>1: 4.1 MB per rank at G = 8, 59 % full)
//   phase 3  owner  check (overflow / fewer than k keys reach tau), merge of its B requests' G lists, gather + FM + MLP
//                   rank, score sort, DPP, D2H of its B results
// No NCCL, no host round trip inside a batch: GPU p orders itself behind GPU g's scatter with cudaStreamWaitEvent on an
// event g recorded after it; one host worker thread per GPU enqueues that GPU's work (a single thread issuing 8 x 15
// launches would be the bottleneck), the workers meet at a barrier between phases only so that an event is recorded
// before anybody waits on it.  A query that fails the check makes the host redo the batch with exact per-shard top-k
// lists (recall_topk_device) through the same scatter — the result is bit-identical to the unsharded path either way.
//
// Handles on the SAME device are allowed (no peer mapping needed): that is how the single-GPU test box exercises every
// line of this file (tests/test_group_gpu.py); real peer stores are covered by the 2- and 8-GPU runs.
#include "handle.h"
#include <condition_variable>
#include <cstring>
#include <thread>
#include <vector>

namespace prg {
int post_recall_device(prg_handle* h, int B, int k, int model, const prg_dpp_params& p, uint32_t* out_row, double* out_score,
                       int32_t* out_n, const prg_user_features& user);

// dst[p] + dst_off <- src (n16 16-byte words) for every peer p = blockIdx.y
__global__ void __launch_bounds__(256) scatter_blocks_kernel(const uint4* __restrict__ src, size_t n16, uint4* const* __restrict__ dst,
                                                             size_t dst_off16) {
  pdl_wait();
  pdl_launch_dependents();
  uint4* d = dst[blockIdx.y] + dst_off16;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) d[i] = src[i];
}

// One warp per query q of this shard's output (lists [Bg][k] sorted descending and 0-padded, then Bg status words): the
// valid prefix goes to slot `me` of the owner's buffer ([G][B*k + B]); the owner zeroed its buffer at the start of the
// batch, so nothing behind the prefix needs to travel.
__global__ void __launch_bounds__(128) scatter_lists_kernel(const uint64_t* __restrict__ lists, int Bg, int B, int k, int me,
                                                            uint64_t* const* __restrict__ owner_buf) {
  pdl_wait();
  pdl_launch_dependents();
  const int q = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  if (q >= Bg) return;
  const int o = q / B, ql = q - o * B;
  const size_t blk = (size_t)B * k + B;
  uint64_t* dst = owner_buf[o] + (size_t)me * blk;
  const uint64_t* src = lists + (size_t)q * k;
  for (int i0 = 0; i0 < k; i0 += 32) {
    const int i = i0 + lane;
    const uint64_t key = i < k ? src[i] : 0ull;
    if (key) dst[(size_t)ql * k + i] = key;
    if (__any_sync(0xffffffffu, key == 0ull)) break;   // sorted, 0-padded: the list ends here
  }
  if (lane == 0) dst[(size_t)B * k + ql] = lists[(size_t)Bg * k + q];
}

struct GroupDev {
  prg_handle* h = nullptr;
  DevBuf q, uid, udense, samp_local, samp_all, cand_local, cand_in, retry, peers1, peers2, out_row, out_score, out_n;
  cudaEvent_t e1 = nullptr, e2 = nullptr, done = nullptr;
  int32_t* retry_host = nullptr;   // pinned [2]
  uint32_t* row_host = nullptr;    // pinned results of this GPU's B requests
  double* score_host = nullptr;
  int32_t* n_host = nullptr;
  size_t res_cap = 0, n_cap = 0;   // capacities of row_host / score_host (results) and of n_host (requests)
  int rc = PRG_OK;
  std::string err;
};
}  // namespace prg

using namespace prg;

struct prg_group {
  int G = 0;
  std::vector<GroupDev> dev;
  std::mutex call_mu;     // one batch at a time
  // pinned input staging shared by all GPUs
  float* q_host = nullptr; uint32_t* uid_host = nullptr; float* udense_host = nullptr;
  size_t q_cap = 0, uid_cap = 0, udense_cap = 0;
  // job description (written under call_mu before the workers are released)
  int B = 0, Bg = 0, k = 0, model = 0, exact = 0;
  prg_dpp_params p{};
  bool has_uid = false, has_dense = false;
  // workers
  std::vector<std::thread> workers;
  std::mutex mu;
  std::condition_variable cv_start, cv_done, cv_bar;
  uint64_t job_seq = 0;
  int n_done = 0, bar_count = 0;
  uint64_t bar_gen = 0;
  bool quit = false;
  int rc1[16] = {0}, rc2[16] = {0};          // per-phase status of every member (read by all after the barrier)
  std::vector<void*> cur_p1, cur_p2;         // destination tables currently on the devices

  void barrier() {   // all G workers
    std::unique_lock<std::mutex> lk(mu);
    const uint64_t gen = bar_gen;
    if (++bar_count == G) { bar_count = 0; ++bar_gen; cv_bar.notify_all(); }
    else cv_bar.wait(lk, [&] { return bar_gen != gen; });
  }
  int run_dev(int g);
  void worker(int g);
};

static int ensure_pinned(void** p, size_t* cap, size_t need) {
  if (need <= *cap && *p) return PRG_OK;
  if (*p) cudaFreeHost(*p);
  *p = nullptr; *cap = 0;
  cudaError_t e = cudaHostAlloc(p, need ? need : 16, cudaHostAllocPortable);
  if (e != cudaSuccess) return fail(PRG_ENOMEM, std::string("cudaHostAlloc: ") + cudaGetErrorString(e));
  *cap = need;
  return PRG_OK;
}

// everything GPU g does for the current batch; called on worker g with the handle's lock held.  Every worker passes both
// barriers whatever happens to it: a phase's status goes to rc1[g] / rc2[g], and after the barrier all workers read the
// same array and take the same decision.
int prg_group::run_dev(int g) {
  GroupDev& d = dev[g];
  prg_handle* h = d.h;
  const uint32_t dim = h->E_dim, U = h->n_user_fields, nd = h->n_user_dense;
  const int r = shard_sample_len(k);
  const size_t blk = (size_t)B * k + B;
  cudaStream_t st = h->stream;
  const float* q_dev = (const float*)d.q.p;

  auto phase1 = [&]() -> int {
    PRG_CUDA(cudaSetDevice(h->device));
    // inputs + zeroed receive buffer (ordered before any peer's phase-2 stores through e1, see file header)
    PRG_CUDA(cudaMemcpyAsync(d.q.p, q_host, (size_t)Bg * dim * 4, cudaMemcpyHostToDevice, st));
    if (has_uid) PRG_CUDA(cudaMemcpyAsync(d.uid.p, uid_host, (size_t)Bg * U * 4, cudaMemcpyHostToDevice, st));
    if (has_dense) PRG_CUDA(cudaMemcpyAsync(d.udense.p, udense_host, (size_t)Bg * nd * 4, cudaMemcpyHostToDevice, st));
    PRG_CUDA(cudaMemsetAsync(d.cand_in.p, 0, (size_t)G * blk * 8, st));
    PRG_CUDA(cudaMemsetAsync(d.retry.p, 0, 8, st));
    // the user prefix of this GPU's own B requests runs on the side stream beside the whole recall (joined in phase 3)
    if (U + nd > 0 && (model == PRG_MODEL_FM || (h->mlp_layers > 0 && h->mlp_k_user == U * 16 + nd)))
      PRG_TRY(user_prefix_device(h, has_uid ? (const uint32_t*)d.uid.p + (size_t)g * B * U : nullptr,
                                 has_dense ? (const float*)d.udense.p + (size_t)g * B * nd : nullptr, B,
                                 model != PRG_MODEL_FM || (h->prerank_keep > 0 && h->prerank_model != PRG_MODEL_FM), /*ahead=*/true));
    if (!exact) {   // sample keys of this shard -> slot g of every GPU's gather buffer
      PRG_TRY(recall_shard_sample_device(h, q_dev, Bg, k, G, (uint64_t*)d.samp_local.p));
      const size_t n16 = (size_t)Bg * r * 8 / 16;
      PRG_CUDA(launch_chained(h, scatter_blocks_kernel, dim3(8, (unsigned)G), dim3(256), 0, 1, (const uint4*)d.samp_local.p, n16,
                              (uint4* const*)d.peers1.p, (size_t)g * n16));
      count_launch(h);
    }
    PRG_CUDA(cudaEventRecord(d.e1, st));
    return PRG_OK;
  };
  auto phase2 = [&]() -> int {
    for (int o = 0; o < G; ++o)
      if (o != g) PRG_CUDA(cudaStreamWaitEvent(st, dev[o].e1, 0));
    if (!exact) {
      PRG_TRY(recall_shard_candidates_device(h, q_dev, Bg, k, G, (const uint64_t*)d.samp_all.p, (uint64_t*)d.cand_local.p));
    } else {   // exact local top-k of this shard (resolves its own check at once), status words zero
      PRG_TRY(recall_topk_device(h, q_dev, Bg, k, (uint64_t*)d.cand_local.p, /*defer=*/false));
      PRG_CUDA(cudaMemsetAsync((uint64_t*)d.cand_local.p + (size_t)Bg * k, 0, (size_t)Bg * 8, st));
    }
    // the valid prefix of every list -> slot g of its owner's buffer
    PRG_CUDA(launch_chained(h, scatter_lists_kernel, dim3((unsigned)((Bg + 3) / 4)), dim3(128), 0, 1,
                            (const uint64_t*)d.cand_local.p, Bg, B, k, g, (uint64_t* const*)d.peers2.p));
    count_launch(h);
    PRG_CUDA(cudaEventRecord(d.e2, st));
    return PRG_OK;
  };
  auto phase3 = [&]() -> int {
    for (int o = 0; o < G; ++o)
      if (o != g) PRG_CUDA(cudaStreamWaitEvent(st, dev[o].e2, 0));
    if (!exact)   // shard_check restricted to the owner's queries: same layout with Bg := B, tau of query g*B + q
      PRG_TRY(shard_check_device(h, (const uint64_t*)d.cand_in.p, G, B, k, (int32_t*)d.retry.p, (const uint64_t*)h->tau.p + (size_t)g * B));
    PRG_CUDA(cudaMemcpyAsync(d.retry_host, d.retry.p, 8, cudaMemcpyDeviceToHost, st));
    PRG_TRY(h->topk_keys.ensure((size_t)B * k * 8));
    PRG_TRY(merge_keys_device(h, (const uint64_t*)d.cand_in.p, G, blk, B, k, (uint64_t*)h->topk_keys.p));
    prg_user_features u{has_uid ? (const uint32_t*)d.uid.p + (size_t)g * B * U : nullptr,
                        has_dense ? (const float*)d.udense.p + (size_t)g * B * nd : nullptr};
    PRG_TRY(post_recall_device(h, B, k, model, p, (uint32_t*)d.out_row.p, (double*)d.out_score.p, (int32_t*)d.out_n.p, u));
    const size_t TT = (size_t)B * p.top_n;
    PRG_CUDA(cudaMemcpyAsync(d.row_host, d.out_row.p, TT * 4, cudaMemcpyDeviceToHost, st));
    PRG_CUDA(cudaMemcpyAsync(d.score_host, d.out_score.p, TT * 8, cudaMemcpyDeviceToHost, st));
    PRG_CUDA(cudaMemcpyAsync(d.n_host, d.out_n.p, (size_t)B * 4, cudaMemcpyDeviceToHost, st));
    PRG_CUDA(cudaEventRecord(d.done, st));
    PRG_CUDA(cudaEventSynchronize(d.done));
    return PRG_OK;
  };

  int rc = phase1();
  if (rc != PRG_OK) d.err = prg_last_error();
  rc1[g] = rc;
  barrier();   // every e1 has been recorded (or its GPU failed)
  bool peer_failed = false;
  for (int o = 0; o < G; ++o) peer_failed |= rc1[o] != PRG_OK;
  if (peer_failed) {   // everybody sees the same rc1: everybody leaves here, nobody waits at the second barrier
    cudaStreamSynchronize(st);
    return rc != PRG_OK ? rc : fail(PRG_ESTATE, "a peer GPU of the group failed in phase 1");
  }
  rc = phase2();
  if (rc != PRG_OK) d.err = prg_last_error();
  rc2[g] = rc;
  barrier();   // every e2 has been recorded
  for (int o = 0; o < G; ++o) peer_failed |= rc2[o] != PRG_OK;
  if (peer_failed) {
    cudaStreamSynchronize(st);
    return rc != PRG_OK ? rc : fail(PRG_ESTATE, "a peer GPU of the group failed in phase 2");
  }
  return phase3();
}

void prg_group::worker(int g) {
  uint64_t seen = 0;
  for (;;) {
    {
      std::unique_lock<std::mutex> lk(mu);
      cv_start.wait(lk, [&] { return quit || job_seq != seen; });
      if (quit) return;
      seen = job_seq;
    }
    GroupDev& d = dev[g];
    int rc;
    {
      std::lock_guard<std::mutex> hl(d.h->mu);   // calls on the member handle from elsewhere wait for the batch
      d.h->pending.active = false;
      d.h->prefix_ahead = false;
      rc = run_dev(g);
      if (rc != PRG_OK) d.h->prefix_ahead = false;   // nobody will join the side stream for this batch
      if (rc != PRG_OK && d.err.empty()) d.err = prg_last_error();
    }
    d.rc = rc;
    {
      std::lock_guard<std::mutex> lk(mu);
      if (++n_done == G) cv_done.notify_all();
    }
  }
}

extern "C" {

int prg_group_create(prg_handle* const* handles, int G, prg_group** out) {
  if (!handles || !out || G < 1 || G > 16) return fail(PRG_EINVAL, "1 <= G <= 16 handles");
  *out = nullptr;
  for (int g = 0; g < G; ++g) {
    if (!handles[g]) return fail(PRG_EINVAL, "null handle in group");
    if (!handles[g]->E) return fail(PRG_ESTATE, "every member needs its item matrix shard (prg_set_item_matrix)");
    if (handles[g]->E_dim != handles[0]->E_dim || handles[g]->n_user_fields != handles[0]->n_user_fields ||
        handles[g]->n_user_dense != handles[0]->n_user_dense)
      return fail(PRG_EINVAL, "members of a group must agree on dim and user feature counts");
    for (int o = 0; o < g; ++o)
      if (handles[o] == handles[g]) return fail(PRG_EINVAL, "a handle appears twice in the group");
  }
  // peer mappings: every member's kernels store into every other member's buffers
  for (int a = 0; a < G; ++a) {
    for (int b = 0; b < G; ++b) {
      const int da = handles[a]->device, db = handles[b]->device;
      if (da == db) continue;
      int can = 0;
      PRG_CUDA(cudaDeviceCanAccessPeer(&can, da, db));
      if (!can) return fail(PRG_EUNSUPPORTED, "GPU " + std::to_string(da) + " cannot map GPU " + std::to_string(db) + " (no P2P)");
      PRG_CUDA(cudaSetDevice(da));
      cudaError_t e = cudaDeviceEnablePeerAccess(db, 0);
      if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
      else if (e != cudaSuccess) return fail(PRG_ECUDA, std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e));
    }
  }
  prg_group* grp = new prg_group();
  grp->G = G;
  grp->dev.resize((size_t)G);
  for (int g = 0; g < G; ++g) {
    GroupDev& d = grp->dev[g];
    d.h = handles[g];
    cudaSetDevice(d.h->device);
    cudaError_t e = cudaEventCreateWithFlags(&d.e1, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&d.e2, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&d.done, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaHostAlloc((void**)&d.retry_host, 8, cudaHostAllocPortable);
    if (e != cudaSuccess) { prg_group_destroy(grp); return fail(PRG_ECUDA, std::string("group setup: ") + cudaGetErrorString(e)); }
  }
  for (int g = 0; g < G; ++g) grp->workers.emplace_back([grp, g] { grp->worker(g); });
  *out = grp;
  return PRG_OK;
}

void prg_group_destroy(prg_group* grp) {
  if (!grp) return;
  {
    std::lock_guard<std::mutex> lk(grp->mu);
    grp->quit = true;
  }
  grp->cv_start.notify_all();
  for (auto& w : grp->workers)
    if (w.joinable()) w.join();
  for (GroupDev& d : grp->dev) {
    if (!d.h) continue;
    cudaSetDevice(d.h->device);
    cudaStreamSynchronize(d.h->stream);
    DevBuf* bufs[] = {&d.q, &d.uid, &d.udense, &d.samp_local, &d.samp_all, &d.cand_local, &d.cand_in, &d.retry, &d.peers1,
                      &d.peers2, &d.out_row, &d.out_score, &d.out_n};
    for (DevBuf* b : bufs) b->release();
    if (d.e1) cudaEventDestroy(d.e1);
    if (d.e2) cudaEventDestroy(d.e2);
    if (d.done) cudaEventDestroy(d.done);
    if (d.retry_host) cudaFreeHost(d.retry_host);
    if (d.row_host) cudaFreeHost(d.row_host);
    if (d.score_host) cudaFreeHost(d.score_host);
    if (d.n_host) cudaFreeHost(d.n_host);
  }
  if (grp->q_host) cudaFreeHost(grp->q_host);
  if (grp->uid_host) cudaFreeHost(grp->uid_host);
  if (grp->udense_host) cudaFreeHost(grp->udense_host);
  delete grp;
}

int prg_group_recommend(prg_group* grp, const float* q, int n_requests, int recall_k, int model, const prg_dpp_params* p,
                        const prg_user_features* user, uint32_t* out_row, double* out_score, int32_t* out_n,
                        int32_t* out_redone) {
  if (!grp || !q || !p || !out_row || !out_score || !out_n) return fail(PRG_EINVAL, "null argument");
  if (n_requests <= 0 || recall_k <= 0 || p->top_n <= 0) return fail(PRG_EINVAL, "n_requests, recall_k, top_n must be positive");
  std::lock_guard<std::mutex> call(grp->call_mu);
  const int G = grp->G;
  const int B = (n_requests + G - 1) / G, Bg = B * G;   // requests per GPU; the tail repeats the last request (dropped again)
  prg_handle* h0 = grp->dev[0].h;
  const uint32_t dim = h0->E_dim, U = h0->n_user_fields, nd = h0->n_user_dense;
  const int r = shard_sample_len(recall_k);
  const size_t blk = (size_t)B * recall_k + B, TT = (size_t)B * p->top_n;
  // ---- pinned input staging (zero-padded), shared by the G copy engines
  // (a zero query would tie every row of the catalog and fail the threshold check for the whole batch)
  PRG_TRY(ensure_pinned((void**)&grp->q_host, &grp->q_cap, (size_t)Bg * dim * 4));
  memcpy(grp->q_host, q, (size_t)n_requests * dim * 4);
  for (int i = n_requests; i < Bg; ++i) memcpy(grp->q_host + (size_t)i * dim, q + (size_t)(n_requests - 1) * dim, (size_t)dim * 4);
  grp->has_uid = user && user->ids && U;
  grp->has_dense = user && user->dense && nd;
  if (grp->has_uid) {
    PRG_TRY(ensure_pinned((void**)&grp->uid_host, &grp->uid_cap, (size_t)Bg * U * 4));
    memcpy(grp->uid_host, user->ids, (size_t)n_requests * U * 4);
    for (int i = n_requests; i < Bg; ++i) memcpy(grp->uid_host + (size_t)i * U, user->ids + (size_t)(n_requests - 1) * U, (size_t)U * 4);
  }
  if (grp->has_dense) {
    PRG_TRY(ensure_pinned((void**)&grp->udense_host, &grp->udense_cap, (size_t)Bg * nd * 4));
    memcpy(grp->udense_host, user->dense, (size_t)n_requests * nd * 4);
    for (int i = n_requests; i < Bg; ++i) memcpy(grp->udense_host + (size_t)i * nd, user->dense + (size_t)(n_requests - 1) * nd, (size_t)nd * 4);
  }
  // ---- per-GPU buffers and the tables of peer destinations
  std::vector<uint4*> p1((size_t)G);
  std::vector<uint64_t*> p2((size_t)G);
  for (int g = 0; g < G; ++g) {
    GroupDev& d = grp->dev[g];
    PRG_CUDA(cudaSetDevice(d.h->device));
    PRG_TRY(d.q.ensure((size_t)Bg * dim * 4));
    if (U) PRG_TRY(d.uid.ensure((size_t)Bg * U * 4));
    if (nd) PRG_TRY(d.udense.ensure((size_t)Bg * nd * 4));
    PRG_TRY(d.samp_local.ensure((size_t)Bg * r * 8));
    PRG_TRY(d.samp_all.ensure((size_t)G * Bg * r * 8));
    PRG_TRY(d.cand_local.ensure(((size_t)Bg * recall_k + Bg) * 8));
    PRG_TRY(d.cand_in.ensure((size_t)G * blk * 8));
    PRG_TRY(d.retry.ensure(8));
    PRG_TRY(d.out_row.ensure(TT * 4));
    PRG_TRY(d.out_score.ensure(TT * 8));
    PRG_TRY(d.out_n.ensure((size_t)B * 4));
    PRG_TRY(d.peers1.ensure((size_t)G * sizeof(void*)));
    PRG_TRY(d.peers2.ensure((size_t)G * sizeof(void*)));
    if (d.res_cap < TT || d.n_cap < (size_t)B) {   // (a later call may have fewer results per request but more requests)
      if (d.row_host) cudaFreeHost(d.row_host);
      if (d.score_host) cudaFreeHost(d.score_host);
      if (d.n_host) cudaFreeHost(d.n_host);
      d.row_host = nullptr; d.score_host = nullptr; d.n_host = nullptr; d.res_cap = 0; d.n_cap = 0;
      PRG_CUDA(cudaHostAlloc((void**)&d.row_host, TT * 4, cudaHostAllocPortable));
      PRG_CUDA(cudaHostAlloc((void**)&d.score_host, TT * 8, cudaHostAllocPortable));
      PRG_CUDA(cudaHostAlloc((void**)&d.n_host, (size_t)B * 4 + 16, cudaHostAllocPortable));
      d.res_cap = TT; d.n_cap = (size_t)B;
    }
    p1[(size_t)g] = (uint4*)d.samp_all.p;
    p2[(size_t)g] = (uint64_t*)d.cand_in.p;
  }
  {   // the destination tables change only when a buffer was re-allocated (a larger batch than ever before)
    std::vector<void*> v1(p1.begin(), p1.end()), v2(p2.begin(), p2.end());
    if (v1 != grp->cur_p1 || v2 != grp->cur_p2) {
      for (int g = 0; g < G; ++g) {
        GroupDev& d = grp->dev[g];
        PRG_CUDA(cudaSetDevice(d.h->device));
        PRG_CUDA(cudaMemcpyAsync(d.peers1.p, p1.data(), (size_t)G * sizeof(void*), cudaMemcpyHostToDevice, d.h->stream));
        PRG_CUDA(cudaMemcpyAsync(d.peers2.p, p2.data(), (size_t)G * sizeof(void*), cudaMemcpyHostToDevice, d.h->stream));
        PRG_CUDA(cudaStreamSynchronize(d.h->stream));   // p1 / p2 live on this stack frame
      }
      grp->cur_p1 = v1; grp->cur_p2 = v2;
    }
  }
  grp->B = B; grp->Bg = Bg; grp->k = recall_k; grp->model = model; grp->p = *p;
  int redone = 0;
  for (int attempt = 0; attempt < 2; ++attempt) {
    grp->exact = attempt;
    {
      std::lock_guard<std::mutex> lk(grp->mu);
      grp->n_done = 0;
      for (GroupDev& d : grp->dev) { d.rc = PRG_OK; d.err.clear(); }
      ++grp->job_seq;
    }
    grp->cv_start.notify_all();
    {
      std::unique_lock<std::mutex> lk(grp->mu);
      grp->cv_done.wait(lk, [&] { return grp->n_done == G; });
    }
    for (GroupDev& d : grp->dev)
      if (d.rc != PRG_OK) return fail(d.rc, "prg_group_recommend (GPU " + std::to_string(d.h->device) + "): " + d.err);
    bool retry = false;
    for (GroupDev& d : grp->dev) retry |= (attempt == 0 && d.retry_host[0] != 0);
    if (!retry) break;
    redone = 1;   // a query failed the global-threshold check somewhere: every GPU redoes the batch with exact local lists
  }
  if (out_redone) *out_redone = redone;
  const size_t T = (size_t)p->top_n;
  for (int i = 0; i < n_requests; ++i) {
    const GroupDev& d = grp->dev[(size_t)(i / B)];
    const int j = i % B;
    memcpy(out_row + (size_t)i * T, d.row_host + (size_t)j * T, T * 4);
    memcpy(out_score + (size_t)i * T, d.score_host + (size_t)j * T, T * 8);
    out_n[i] = d.n_host[j];
  }
  return PRG_OK;
}

int prg_group_size(prg_group* grp) { return grp ? grp->G : 0; }

}  // extern "C"
