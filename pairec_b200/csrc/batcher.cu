// batcher.cu — cross-call request batcher over prg_recommend (SURVEY §8b "internal cross-call batcher").
//
// The reference serves ONE request per goroutine: UserRecommendService.Recommend (service/user_recommend.go:46) runs
// recall -> rank -> sort for a single user, and RankService.Rank fans a request out into BatchCount-sized RPCs
// (service/rank/rank_service.go:163-166, :264-289).  The kernels behind prg_recommend want the opposite shape: one
// pass over the item matrix serves up to 256 queries for the price of one.  The batcher sits between the two: any
// number of host threads (cgo calls from request goroutines) block in prg_batcher_recommend with ONE query each; a
// worker thread coalesces whatever has arrived into one prg_recommend call and hands every caller its own slice of
// the result.
//
// Policy (continuous batching): the worker dispatches as soon as the GPU is free and at least one request waits, so
// an idle server adds no queueing delay; while a batch runs (≈ 1 ms) the next one fills.  `max_wait_us` > 0 lets a
// batch that is not yet full wait that long (measured from its first request) for more requests to join.
// Slots: two pinned host staging areas; a slot is FILLING (callers copy their query in), RUNNING (owned by the
// worker), DRAINING (callers copy their results out) or FREE.
#include "handle.h"
#include <chrono>
#include <condition_variable>
#include <cstring>
#include <thread>

using namespace prg;

namespace {
enum SlotState { SLOT_FREE = 0, SLOT_FILLING = 1, SLOT_RUNNING = 2, SLOT_DRAINING = 3 };

struct Slot {
  int state = SLOT_FREE;
  int count = 0;     // requests in the slot
  int readers = 0;   // callers that still have to copy their result out (DRAINING)
  int rc = PRG_OK;
  std::string err;
  uint64_t ticket = 0;  // generation: a caller waits for ITS batch, not for whatever occupies the slot later
  std::chrono::steady_clock::time_point first;
  float* q = nullptr;          // pinned [max_batch][dim]
  uint32_t* rows = nullptr;    // pinned [max_batch][top_n]
  double* scores = nullptr;    // pinned [max_batch][top_n]
  int32_t* n = nullptr;        // pinned [max_batch]
};
}  // namespace

struct prg_batcher {
  prg_handle* h = nullptr;
  prg_batcher_config cfg{};
  uint32_t dim = 0;
  std::mutex mu;
  std::condition_variable cv_worker;   // a request arrived / stop
  std::condition_variable cv_callers;  // a batch finished / a slot became free
  Slot slot[2];
  int filling = -1;  // index of the FILLING slot, -1 if none
  bool stop = false;
  int inflight = 0;  // callers inside prg_batcher_recommend (prg_batcher_stop waits for them before freeing)
  std::thread worker;
  uint64_t n_requests = 0, n_batches = 0, next_ticket = 1;
  uint64_t hist[9] = {0};  // batch sizes: 1, 2, 3-4, 5-8, 9-16, 17-32, 33-64, 65-128, 129+

  void run();
};

void prg_batcher::run() {
  cudaSetDevice(h->device);
  std::unique_lock<std::mutex> lk(mu);
  for (;;) {
    cv_worker.wait(lk, [&] { return stop || (filling >= 0 && slot[filling].count > 0); });
    if (stop && (filling < 0 || slot[filling].count == 0)) return;
    Slot& s = slot[filling];
    if (cfg.max_wait_us > 0 && s.count < cfg.max_batch && !stop) {
      const auto deadline = s.first + std::chrono::microseconds(cfg.max_wait_us);
      cv_worker.wait_until(lk, deadline, [&] { return stop || s.count >= cfg.max_batch; });
    }
    s.state = SLOT_RUNNING;
    filling = -1;
    const int B = s.count;
    cv_callers.notify_all();  // callers waiting for a slot to fill may now open the other one
    lk.unlock();
    const int rc = prg_recommend(h, s.q, B, cfg.recall_k, cfg.model, &cfg.dpp, s.rows, s.scores, s.n, PRG_MEM_HOST);
    std::string err = rc == PRG_OK ? std::string() : std::string(prg_last_error());
    lk.lock();
    s.rc = rc;
    s.err = std::move(err);
    s.readers = B;
    s.state = SLOT_DRAINING;
    n_batches += 1;
    n_requests += (uint64_t)B;
    int bin = 0;
    for (int v = B - 1; v > 0 && bin < 8; v >>= 1) ++bin;
    hist[bin] += 1;
    cv_callers.notify_all();
  }
}

extern "C" {

int prg_batcher_start(prg_handle* h, const prg_batcher_config* cfg, prg_batcher** out) {
  if (!h || !cfg || !out) return fail(PRG_EINVAL, "null argument");
  if (cfg->max_batch <= 0 || cfg->max_batch > 256) return fail(PRG_EINVAL, "max_batch must be in 1..256");
  if (cfg->recall_k <= 0 || cfg->dpp.top_n <= 0 || cfg->max_wait_us < 0) return fail(PRG_EINVAL, "bad batcher config");
  uint32_t dim;
  {
    std::lock_guard<std::mutex> g(h->mu);
    if (!h->E) return fail(PRG_ESTATE, "item matrix not set (prg_set_item_matrix)");
    dim = h->E_dim;
  }
  cudaSetDevice(h->device);
  prg_batcher* b = new prg_batcher();
  b->h = h;
  b->cfg = *cfg;
  b->dim = dim;
  const size_t T = (size_t)cfg->dpp.top_n, M = (size_t)cfg->max_batch;
  for (Slot& s : b->slot) {
    cudaError_t e = cudaHostAlloc((void**)&s.q, M * dim * 4, cudaHostAllocDefault);
    if (e == cudaSuccess) e = cudaHostAlloc((void**)&s.rows, M * T * 4, cudaHostAllocDefault);
    if (e == cudaSuccess) e = cudaHostAlloc((void**)&s.scores, M * T * 8, cudaHostAllocDefault);
    if (e == cudaSuccess) e = cudaHostAlloc((void**)&s.n, M * 4, cudaHostAllocDefault);
    if (e != cudaSuccess) {
      for (Slot& t : b->slot) { cudaFreeHost(t.q); cudaFreeHost(t.rows); cudaFreeHost(t.scores); cudaFreeHost(t.n); }
      delete b;
      return fail(PRG_ENOMEM, std::string("cudaHostAlloc: ") + cudaGetErrorString(e));
    }
  }
  b->worker = std::thread([b] { b->run(); });
  *out = b;
  return PRG_OK;
}

int prg_batcher_recommend(prg_batcher* b, const float* q, uint32_t* out_row, double* out_score, int32_t* out_n) {
  if (!b || !q || !out_row || !out_score || !out_n) return fail(PRG_EINVAL, "null argument");
  const size_t T = (size_t)b->cfg.dpp.top_n;
  std::unique_lock<std::mutex> lk(b->mu);
  struct InFlight {  // constructed and destroyed under the lock
    prg_batcher* b;
    explicit InFlight(prg_batcher* bb) : b(bb) { ++b->inflight; }
    ~InFlight() { if (--b->inflight == 0 && b->stop) b->cv_callers.notify_all(); }
  } inflight_guard(b);
  // 1. a place in the filling slot (open a FREE slot if there is none)
  int si = -1;
  for (;;) {
    if (b->stop) return fail(PRG_ESTATE, "batcher stopped");
    if (b->filling >= 0 && b->slot[b->filling].count < b->cfg.max_batch) { si = b->filling; break; }
    if (b->filling < 0) {
      for (int i = 0; i < 2; ++i)
        if (b->slot[i].state == SLOT_FREE) { si = i; break; }
      if (si >= 0) {
        Slot& s = b->slot[si];
        s.state = SLOT_FILLING;
        s.count = 0;
        s.ticket = b->next_ticket++;
        s.first = std::chrono::steady_clock::now();
        b->filling = si;
        break;
      }
    }
    b->cv_callers.wait(lk);
  }
  Slot& s = b->slot[si];
  const int me = s.count++;
  const uint64_t ticket = s.ticket;
  std::memcpy(s.q + (size_t)me * b->dim, q, (size_t)b->dim * 4);  // short copy under the lock: 256-512 bytes
  if (me == 0 || s.count >= b->cfg.max_batch) b->cv_worker.notify_one();
  // 2. wait for this batch
  b->cv_callers.wait(lk, [&] { return s.ticket == ticket && s.state == SLOT_DRAINING; });
  const int rc = s.rc;
  if (rc == PRG_OK) {
    std::memcpy(out_row, s.rows + (size_t)me * T, T * 4);
    std::memcpy(out_score, s.scores + (size_t)me * T, T * 8);
    *out_n = s.n[me];
  } else {
    set_error("batched prg_recommend: " + s.err);
  }
  if (--s.readers == 0) {
    s.state = SLOT_FREE;
    s.count = 0;
    b->cv_callers.notify_all();
  }
  return rc;
}

int prg_batcher_stats(prg_batcher* b, uint64_t* n_requests, uint64_t* n_batches, uint64_t* size_hist9) {
  if (!b) return fail(PRG_EINVAL, "null batcher");
  std::lock_guard<std::mutex> g(b->mu);
  if (n_requests) *n_requests = b->n_requests;
  if (n_batches) *n_batches = b->n_batches;
  if (size_hist9) std::memcpy(size_hist9, b->hist, sizeof(b->hist));
  return PRG_OK;
}

int prg_batcher_drive(prg_batcher* b, const float* q_pool, int n_pool, int n_threads, int per_thread, float* latency_us,
                      double* wall_s, uint32_t* rows_out, int32_t* n_out) {
  if (!b || !q_pool || n_pool <= 0 || n_threads <= 0 || per_thread <= 0) return fail(PRG_EINVAL, "bad arguments");
  const size_t T = (size_t)b->cfg.dpp.top_n;
  const uint32_t dim = b->dim;
  std::vector<std::thread> th;
  std::vector<int> rcs((size_t)n_threads, PRG_OK);
  std::vector<std::string> errs((size_t)n_threads);
  const auto t0 = std::chrono::steady_clock::now();
  for (int t = 0; t < n_threads; ++t) {
    th.emplace_back([=, &rcs, &errs] {
      std::vector<uint32_t> rows(T);
      std::vector<double> scores(T);
      int32_t n = 0;
      for (int j = 0; j < per_thread; ++j) {
        const size_t i = (size_t)t * per_thread + j;
        const int qi = (int)(i % (size_t)n_pool);
        const auto a = std::chrono::steady_clock::now();
        const int rc = prg_batcher_recommend(b, q_pool + (size_t)qi * dim, rows.data(), scores.data(), &n);
        const auto z = std::chrono::steady_clock::now();
        if (rc != PRG_OK) { rcs[t] = rc; errs[t] = prg_last_error(); return; }
        if (latency_us) latency_us[i] = std::chrono::duration<float, std::micro>(z - a).count();
        if (rows_out && i < (size_t)n_pool) {  // first pass over the pool: keep the answers for the caller to verify
          std::memcpy(rows_out + i * T, rows.data(), T * 4);
          if (n_out) n_out[i] = n;
        }
      }
    });
  }
  for (auto& x : th) x.join();
  if (wall_s) *wall_s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  for (int t = 0; t < n_threads; ++t)
    if (rcs[t] != PRG_OK) return fail(rcs[t], "client thread: " + errs[t]);
  return PRG_OK;
}

void prg_batcher_stop(prg_batcher* b) {
  if (!b) return;
  {
    std::lock_guard<std::mutex> g(b->mu);
    b->stop = true;
  }
  b->cv_worker.notify_all();
  b->cv_callers.notify_all();
  if (b->worker.joinable()) b->worker.join();
  {  // callers of the last batch may still be copying their results out; late arrivals leave with PRG_ESTATE
    std::unique_lock<std::mutex> lk(b->mu);
    b->cv_callers.wait(lk, [&] { return b->inflight == 0; });
  }
  cudaSetDevice(b->h->device);
  for (Slot& t : b->slot) { cudaFreeHost(t.q); cudaFreeHost(t.rows); cudaFreeHost(t.scores); cudaFreeHost(t.n); }
  delete b;
}

}  // extern "C"
