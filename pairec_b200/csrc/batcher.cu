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
// Pipelining (default; PRG_BATCHER_PIPELINE=0 restores the one-batch-at-a-time turn taking): a FULL batch does not
// wait for the GPU to free up — it is enqueued behind the running one (recommend_begin / recommend_end, pipeline.cu),
// so under load the device never idles between batches (≈ 60 µs per batch of host wake-up, H2D and launch latency
// otherwise); at most two batches are in flight, and a batch that is not full still forms at the last moment.
// Requests wait in ONE FIFO queue (arrival order == service order); each caller sleeps on its own condition variable,
// so finishing a batch wakes exactly its callers.
#include "handle.h"
#include <chrono>
#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <thread>

using namespace prg;

namespace {
// one blocked caller; lives on that caller's stack for the duration of prg_batcher_recommend
struct Request {
  const float* q;
  const uint32_t* user_ids = nullptr;   // [U] or null (every user feature absent)
  const float* user_dense = nullptr;    // [n_dense] or null (zeros)
  uint32_t* out_row;
  double* out_score;
  int32_t* out_n;
  int rc = PRG_OK;
  bool done = false;
  std::string err;
  std::condition_variable cv;
  std::chrono::steady_clock::time_point arrived;
};

struct Staging {  // pinned host buffers of one worker
  float* q = nullptr;          // [max_batch][dim]
  uint32_t* rows = nullptr;    // [max_batch][top_n]
  double* scores = nullptr;    // [max_batch][top_n]
  int32_t* n = nullptr;        // [max_batch]
  uint32_t* uid = nullptr;     // [max_batch][U] user field ids
  float* udense = nullptr;     // [max_batch][n_dense]
  void release() { cudaFreeHost(q); cudaFreeHost(rows); cudaFreeHost(scores); cudaFreeHost(n); cudaFreeHost(uid); cudaFreeHost(udense); }
};
}  // namespace

struct prg_batcher {
  prg_handle* h = nullptr;
  prg_batcher_config cfg{};
  uint32_t dim = 0;
  uint32_t n_user = 0, n_dense = 0;    // user / context features per request (prg_set_user_fields at start time)
  std::mutex mu;                       // queue, flags, statistics
  std::mutex gpu_turn;                 // held by the worker that forms and runs the next batch
  std::condition_variable cv_worker;   // a request arrived / stop
  std::condition_variable cv_idle;     // inflight dropped to 0 after stop
  std::deque<Request*> queue;          // FIFO: requests are served in arrival order
  bool stop = false;
  int inflight = 0;  // callers inside prg_batcher_recommend (prg_batcher_stop waits for them before freeing)
  std::thread worker[2];
  Staging stage[2];
  std::string last_err[2];
  uint64_t n_requests = 0, n_batches = 0;
  uint64_t hist[9] = {0};  // batch sizes: 1, 2, 3-4, 5-8, 9-16, 17-32, 33-64, 65-128, 129+
  bool pipeline = true;    // a full batch is enqueued behind the running one
  int gpu_busy = 0;        // batches enqueued and not yet finished (pipelined mode; guarded by mu)
  cudaEvent_t done[2] = {nullptr, nullptr};

  void run(int w);
  void run_pipelined(int w);
  // queries and user features of a batch -> worker w's pinned staging; returns the user block to pass on (or null)
  const prg_user_features* stage_inputs(int w, const std::vector<Request*>& batch, prg_user_features* uf) {
    Staging& st = stage[w];
    const int B = (int)batch.size();
    for (int i = 0; i < B; ++i) std::memcpy(st.q + (size_t)i * dim, batch[i]->q, (size_t)dim * 4);
    if (n_user + n_dense == 0) return nullptr;
    for (int i = 0; i < B; ++i) {
      if (n_user) {
        if (batch[i]->user_ids) std::memcpy(st.uid + (size_t)i * n_user, batch[i]->user_ids, (size_t)n_user * 4);
        else std::memset(st.uid + (size_t)i * n_user, 0xFF, (size_t)n_user * 4);
      }
      if (n_dense) {
        if (batch[i]->user_dense) std::memcpy(st.udense + (size_t)i * n_dense, batch[i]->user_dense, (size_t)n_dense * 4);
        else std::memset(st.udense + (size_t)i * n_dense, 0, (size_t)n_dense * 4);
      }
    }
    uf->ids = n_user ? st.uid : nullptr;
    uf->dense = n_dense ? st.udense : nullptr;
    return uf;
  }
  void deliver(int w, std::vector<Request*>& batch, int rc);
};

// Two workers take turns: the one holding `gpu_turn` forms its batch at the last moment (when the previous batch has
// left the GPU, so everything that arrived meanwhile joins), runs it, passes the turn on and only then hands out the
// results — the other worker's batch is already on the GPU while the callers of this one are being woken.
void prg_batcher::run(int w) {
  cudaSetDevice(h->device);
  Staging& st = stage[w];
  std::vector<Request*> batch;
  batch.reserve((size_t)cfg.max_batch);
  for (;;) {
    std::unique_lock<std::mutex> turn(gpu_turn);
    {
      std::unique_lock<std::mutex> lk(mu);
      cv_worker.wait(lk, [&] { return stop || !queue.empty(); });
      if (queue.empty()) return;  // stop, and nothing left to serve
      if (cfg.max_wait_us > 0 && (int)queue.size() < cfg.max_batch && !stop) {
        const auto deadline = queue.front()->arrived + std::chrono::microseconds(cfg.max_wait_us);
        cv_worker.wait_until(lk, deadline, [&] { return stop || (int)queue.size() >= cfg.max_batch; });
      }
      batch.clear();
      while (!queue.empty() && (int)batch.size() < cfg.max_batch) {
        batch.push_back(queue.front());
        queue.pop_front();
      }
    }
    const int B = (int)batch.size();
    prg_user_features uf{nullptr, nullptr};
    const prg_user_features* user = stage_inputs(w, batch, &uf);
    const int rc = prg_recommend_ex(h, st.q, B, cfg.recall_k, cfg.model, &cfg.dpp, user, st.rows, st.scores, st.n, PRG_MEM_HOST);
    if (rc != PRG_OK) last_err[w] = prg_last_error();
    turn.unlock();
    deliver(w, batch, rc);
  }
}

// results of worker w's staging buffers -> the callers' buffers; statistics; wake the callers
void prg_batcher::deliver(int w, std::vector<Request*>& batch, int rc) {
  Staging& st = stage[w];
  const size_t T = (size_t)cfg.dpp.top_n;
  const int B = (int)batch.size();
  if (rc == PRG_OK) {  // the callers are blocked: their output buffers are ours to fill
    for (int i = 0; i < B; ++i) {
      std::memcpy(batch[i]->out_row, st.rows + (size_t)i * T, T * 4);
      std::memcpy(batch[i]->out_score, st.scores + (size_t)i * T, T * 8);
      *batch[i]->out_n = st.n[i];
    }
  }
  std::lock_guard<std::mutex> lk(mu);
  n_batches += 1;
  n_requests += (uint64_t)B;
  int bin = 0;
  for (int v = B - 1; v > 0 && bin < 8; v >>= 1) ++bin;
  hist[bin] += 1;
  for (int i = 0; i < B; ++i) {  // notify under the lock: a Request may be destroyed as soon as its owner sees `done`
    batch[i]->rc = rc;
    if (rc != PRG_OK) batch[i]->err = last_err[w];
    batch[i]->done = true;
    batch[i]->cv.notify_one();
  }
}

// Pipelined turn taking: the worker holding `gpu_turn` forms a batch when the device is idle (gpu_busy == 0: exactly
// the policy of run()) OR when a full batch waits and at most one batch is in flight; it ENQUEUES the batch
// (recommend_begin returns once the copies and launches are in the stream), passes the turn on, and only then waits
// for its results.  The other worker's full batch therefore sits in the stream behind this one.
void prg_batcher::run_pipelined(int w) {
  cudaSetDevice(h->device);
  Staging& st = stage[w];
  std::vector<Request*> batch;
  batch.reserve((size_t)cfg.max_batch);
  for (;;) {
    std::unique_lock<std::mutex> turn(gpu_turn);
    {
      std::unique_lock<std::mutex> lk(mu);
      auto ready = [&] {
        if (queue.empty()) return stop;
        return gpu_busy == 0 || (gpu_busy < 2 && (int)queue.size() >= cfg.max_batch);
      };
      cv_worker.wait(lk, ready);
      if (queue.empty()) return;  // stop, and nothing left to serve
      if (cfg.max_wait_us > 0 && (int)queue.size() < cfg.max_batch && !stop) {
        const auto deadline = queue.front()->arrived + std::chrono::microseconds(cfg.max_wait_us);
        cv_worker.wait_until(lk, deadline, [&] { return stop || (int)queue.size() >= cfg.max_batch; });
      }
      batch.clear();
      while (!queue.empty() && (int)batch.size() < cfg.max_batch) {
        batch.push_back(queue.front());
        queue.pop_front();
      }
      ++gpu_busy;
    }
    const int B = (int)batch.size();
    prg_user_features uf{nullptr, nullptr};
    const prg_user_features* user = stage_inputs(w, batch, &uf);
    uint64_t seq = 0;
    int rc = recommend_begin(h, st.q, B, cfg.recall_k, cfg.model, cfg.dpp, user, st.rows, st.scores, st.n, done[w], &seq);
    if (rc != PRG_OK) last_err[w] = prg_last_error();
    turn.unlock();
    if (rc == PRG_OK) {
      rc = recommend_end(h, seq, done[w]);
      if (rc != PRG_OK) last_err[w] = prg_last_error();
    }
    {
      std::lock_guard<std::mutex> lk(mu);
      --gpu_busy;
    }
    cv_worker.notify_all();   // the device has one batch less: a waiting worker may form a partial batch now
    deliver(w, batch, rc);
  }
}

extern "C" {

int prg_batcher_start(prg_handle* h, const prg_batcher_config* cfg, prg_batcher** out) {
  if (!h || !cfg || !out) return fail(PRG_EINVAL, "null argument");
  if (cfg->max_batch <= 0 || cfg->max_batch > 256) return fail(PRG_EINVAL, "max_batch must be in 1..256");
  if (cfg->recall_k <= 0 || cfg->dpp.top_n <= 0 || cfg->max_wait_us < 0) return fail(PRG_EINVAL, "bad batcher config");
  uint32_t dim, n_user, n_dense;
  {
    std::lock_guard<std::mutex> g(h->mu);
    if (!h->E) return fail(PRG_ESTATE, "item matrix not set (prg_set_item_matrix)");
    dim = h->E_dim;
    n_user = h->n_user_fields;
    n_dense = h->n_user_dense;
  }
  cudaSetDevice(h->device);
  prg_batcher* b = new prg_batcher();
  b->h = h;
  b->cfg = *cfg;
  b->dim = dim;
  b->n_user = n_user;
  b->n_dense = n_dense;
  const size_t T = (size_t)cfg->dpp.top_n, M = (size_t)cfg->max_batch;
  for (Staging& s : b->stage) {
    cudaError_t e = cudaHostAlloc((void**)&s.q, M * dim * 4, cudaHostAllocDefault);
    if (e == cudaSuccess) e = cudaHostAlloc((void**)&s.rows, M * T * 4, cudaHostAllocDefault);
    if (e == cudaSuccess) e = cudaHostAlloc((void**)&s.scores, M * T * 8, cudaHostAllocDefault);
    if (e == cudaSuccess) e = cudaHostAlloc((void**)&s.n, M * 4, cudaHostAllocDefault);
    if (e == cudaSuccess && n_user) e = cudaHostAlloc((void**)&s.uid, M * n_user * 4, cudaHostAllocDefault);
    if (e == cudaSuccess && n_dense) e = cudaHostAlloc((void**)&s.udense, M * n_dense * 4, cudaHostAllocDefault);
    if (e != cudaSuccess) {
      for (Staging& t : b->stage) t.release();
      delete b;
      return fail(PRG_ENOMEM, std::string("cudaHostAlloc: ") + cudaGetErrorString(e));
    }
  }
  if (const char* ev = getenv("PRG_BATCHER_PIPELINE")) b->pipeline = atoi(ev) != 0;
  for (int w = 0; w < 2; ++w) {
    if (cudaEventCreateWithFlags(&b->done[w], cudaEventDisableTiming) != cudaSuccess) {
      for (Staging& t : b->stage) t.release();
      for (cudaEvent_t e : b->done) if (e) cudaEventDestroy(e);
      delete b;
      return fail(PRG_ECUDA, "cudaEventCreate failed");
    }
  }
  for (int w = 0; w < 2; ++w)
    b->worker[w] = std::thread([b, w] { if (b->pipeline) b->run_pipelined(w); else b->run(w); });
  *out = b;
  return PRG_OK;
}

int prg_batcher_recommend(prg_batcher* b, const float* q, uint32_t* out_row, double* out_score, int32_t* out_n) {
  return prg_batcher_recommend_ex(b, q, nullptr, nullptr, out_row, out_score, out_n);
}

int prg_batcher_recommend_ex(prg_batcher* b, const float* q, const uint32_t* user_ids, const float* user_dense,
                             uint32_t* out_row, double* out_score, int32_t* out_n) {
  if (!b || !q || !out_row || !out_score || !out_n) return fail(PRG_EINVAL, "null argument");
  Request r;
  r.q = q; r.user_ids = user_ids; r.user_dense = user_dense;
  r.out_row = out_row; r.out_score = out_score; r.out_n = out_n;
  std::unique_lock<std::mutex> lk(b->mu);
  if (b->stop) return fail(PRG_ESTATE, "batcher stopped");
  ++b->inflight;
  r.arrived = std::chrono::steady_clock::now();
  b->queue.push_back(&r);
  if (b->queue.size() == 1 || (int)b->queue.size() == b->cfg.max_batch) b->cv_worker.notify_all();
  r.cv.wait(lk, [&] { return r.done; });
  const int rc = r.rc;
  if (rc != PRG_OK) set_error("batched prg_recommend: " + r.err);
  if (--b->inflight == 0 && b->stop) b->cv_idle.notify_all();
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
    b->stop = true;   // late callers leave with PRG_ESTATE; what is queued is still served
  }
  b->cv_worker.notify_all();
  for (auto& w : b->worker)
    if (w.joinable()) w.join();
  {
    std::unique_lock<std::mutex> lk(b->mu);
    b->cv_idle.wait(lk, [&] { return b->inflight == 0; });
  }
  cudaSetDevice(b->h->device);
  for (Staging& t : b->stage) t.release();
  for (cudaEvent_t e : b->done) if (e) cudaEventDestroy(e);
  delete b;
}

}  // extern "C"
