#include "serving.h"

#include <algorithm>
#include <cstring>

namespace trbs {

static inline int64_t align16(int64_t x) { return (x + 15) & ~int64_t(15); }

Server::Server(const Config& cfg) : cfg_(cfg), ready_(cfg.num_gpus), outstanding_(cfg.num_gpus, 0) {
  for (int i = 0; i < std::max(1, cfg.batching_threads); ++i) threads_.emplace_back([this] { batching_loop(); });
}

Server::~Server() { shutdown(); }

void Server::shutdown() {
  if (stop_.exchange(true)) return;
  in_cv_.notify_all();
  buf_cv_.notify_all();
  rq_cv_.notify_all();
  for (auto& t : threads_) if (t.joinable()) t.join();
  // fail whatever is still pending
  std::lock_guard<std::mutex> g(in_mu_);
  for (auto& kv : live_) {
    std::lock_guard<std::mutex> rg(kv.second->mu);
    if (!kv.second->done) { kv.second->done = true; kv.second->status = -9; kv.second->cv.notify_all(); }
  }
}

void Server::add_buffer(void* ptr, int64_t bytes) {
  std::lock_guard<std::mutex> g(buf_mu_);
  buffers_.push_back(Buffer{(uint8_t*) ptr, bytes, false});
  buf_cv_.notify_all();
}

int64_t Server::submit(int32_t batch_size, int32_t num_float, const float* dense, int32_t idl_features, const int32_t* idl_lengths, const int64_t* idl_values,
                       int64_t idl_nvalues, int32_t ids_features, const int32_t* ids_lengths, const int64_t* ids_values, const float* ids_weights,
                       int64_t ids_nvalues) {
  auto r = std::make_shared<Request>();
  r->batch_size = batch_size;
  r->num_float = num_float;
  if (num_float > 0) r->dense.assign(dense, dense + (int64_t) batch_size * num_float);
  r->id_list.num_features = idl_features;
  if (idl_features > 0) {
    r->id_list.lengths.assign(idl_lengths, idl_lengths + (int64_t) idl_features * batch_size);
    r->id_list.values.assign(idl_values, idl_values + idl_nvalues);
  }
  r->id_score_list.num_features = ids_features;
  if (ids_features > 0) {
    r->id_score_list.lengths.assign(ids_lengths, ids_lengths + (int64_t) ids_features * batch_size);
    r->id_score_list.values.assign(ids_values, ids_values + ids_nvalues);
    r->id_score_list.weights.assign(ids_weights, ids_weights + ids_nvalues);
  }
  r->enqueued = Clock::now();
  {
    std::lock_guard<std::mutex> g(in_mu_);
    if ((int64_t) intake_.size() >= cfg_.max_queue_requests || stop_) { ++stats_.rejected; return -1; }
    r->id = next_request_++;
    intake_.push_back(r);
    live_[r->id] = r;
  }
  ++stats_.requests;
  in_cv_.notify_one();
  return r->id;
}

// Batching thread: wait for the first request, then keep collecting until the batch is full or the oldest request has
// waited `batching_interval_us`; requests that would overflow max_batch_size stay for the next batch.
void Server::batching_loop() {
  while (!stop_) {
    std::vector<std::shared_ptr<Request>> reqs;
    {
      std::unique_lock<std::mutex> g(in_mu_);
      in_cv_.wait(g, [this] { return stop_ || !intake_.empty(); });
      if (stop_) return;
      const auto deadline = intake_.front()->enqueued + std::chrono::microseconds(cfg_.batching_interval_us);
      int64_t have = 0;
      for (;;) {
        have = 0;
        for (auto& r : intake_) have += r->batch_size;
        if (have >= cfg_.max_batch_size || stop_) break;
        if (in_cv_.wait_until(g, deadline) == std::cv_status::timeout) break;
        if (intake_.empty()) break;  // another batching thread took them
      }
      if (intake_.empty()) continue;
      int64_t total = 0;
      const int32_t nf = intake_.front()->num_float, f1 = intake_.front()->id_list.num_features, f2 = intake_.front()->id_score_list.num_features;
      while (!intake_.empty()) {
        auto& r = intake_.front();
        // a batch only merges requests of the same schema; oversize single requests form their own batch
        if (r->num_float != nf || r->id_list.num_features != f1 || r->id_score_list.num_features != f2) break;
        if (!reqs.empty() && total + r->batch_size > cfg_.max_batch_size) break;
        total += r->batch_size;
        reqs.push_back(r);
        intake_.pop_front();
      }
    }
    if (!reqs.empty() && !form_batch(reqs)) {
      for (auto& r : reqs) {
        std::lock_guard<std::mutex> rg(r->mu);
        r->done = true; r->status = -2; r->cv.notify_all();
      }
    }
  }
}

int Server::acquire_buffer(int64_t bytes, uint8_t** ptr) {
  std::unique_lock<std::mutex> g(buf_mu_);
  for (;;) {
    bool any_big_enough = false;
    for (size_t i = 0; i < buffers_.size(); ++i) {
      if (buffers_[i].bytes >= bytes) {
        any_big_enough = true;
        if (!buffers_[i].busy) { buffers_[i].busy = true; *ptr = buffers_[i].ptr; return (int) i; }
      }
    }
    if (stop_ || (!any_big_enough && !buffers_.empty())) return -1;
    buf_cv_.wait_for(g, std::chrono::milliseconds(50));
  }
}

static void merge_sparse(const std::vector<std::shared_ptr<Request>>& reqs, bool score, int32_t F, int32_t B, int32_t* lengths, int64_t* values, float* weights) {
  // output layout: lengths [F][B_total] (requests concatenated along the batch), values key-major
  int64_t vpos = 0;
  std::vector<int64_t> cursor(reqs.size(), 0);  // per-request read offset into its values (key-major, so sequential)
  for (int32_t f = 0; f < F; ++f) {
    int32_t b0 = 0;
    for (size_t q = 0; q < reqs.size(); ++q) {
      const SparseInput& s = score ? reqs[q]->id_score_list : reqs[q]->id_list;
      const int32_t bq = reqs[q]->batch_size;
      const int32_t* len = s.lengths.data() + (int64_t) f * bq;
      std::memcpy(lengths + (int64_t) f * B + b0, len, sizeof(int32_t) * bq);
      int64_t n = 0;
      for (int32_t i = 0; i < bq; ++i) n += len[i];
      std::memcpy(values + vpos, s.values.data() + cursor[q], sizeof(int64_t) * n);
      if (score) std::memcpy(weights + vpos, s.weights.data() + cursor[q], sizeof(float) * n);
      cursor[q] += n;
      vpos += n;
      b0 += bq;
    }
  }
}

bool Server::form_batch(std::vector<std::shared_ptr<Request>>& reqs) {
  auto fb = std::make_shared<Formed>();
  BatchDesc& d = fb->desc;
  std::memset(&d, 0, sizeof(d));
  int32_t B = 0;
  int64_t nv1 = 0, nv2 = 0;
  for (auto& r : reqs) { B += r->batch_size; nv1 += (int64_t) r->id_list.values.size(); nv2 += (int64_t) r->id_score_list.values.size(); }
  const Request& r0 = *reqs[0];
  d.batch_size = B;
  d.num_requests = (int32_t) reqs.size();
  d.num_float = r0.num_float;
  d.id_list_features = r0.id_list.num_features;
  d.id_score_features = r0.id_score_list.num_features;
  int64_t off = 0;
  d.dense_off = d.num_float ? off : -1;             off = align16(off + (int64_t) B * d.num_float * 4);
  d.idl_lengths_off = d.id_list_features ? off : -1; off = align16(off + (int64_t) B * d.id_list_features * 4);
  d.idl_values_off = d.id_list_features ? off : -1;  off = align16(off + nv1 * 8);
  d.ids_lengths_off = d.id_score_features ? off : -1; off = align16(off + (int64_t) B * d.id_score_features * 4);
  d.ids_values_off = d.id_score_features ? off : -1;  off = align16(off + nv2 * 8);
  d.ids_weights_off = d.id_score_features ? off : -1; off = align16(off + nv2 * 4);
  d.idl_num_values = nv1;
  d.ids_num_values = nv2;
  d.total_bytes = off;
  uint8_t* base = nullptr;
  const int bi = acquire_buffer(off, &base);
  if (bi < 0) return false;
  d.buffer_index = bi;
  if (d.num_float) {
    float* dst = (float*) (base + d.dense_off);
    for (auto& r : reqs) { std::memcpy(dst, r->dense.data(), r->dense.size() * 4); dst += r->dense.size(); }
  }
  if (d.id_list_features) merge_sparse(reqs, false, d.id_list_features, B, (int32_t*) (base + d.idl_lengths_off), (int64_t*) (base + d.idl_values_off), nullptr);
  if (d.id_score_features) merge_sparse(reqs, true, d.id_score_features, B, (int32_t*) (base + d.ids_lengths_off), (int64_t*) (base + d.ids_values_off), (float*) (base + d.ids_weights_off));
  const auto now = Clock::now();
  int64_t oldest = 0;
  for (auto& r : reqs) {
    const int64_t us = std::chrono::duration_cast<std::chrono::microseconds>(now - r->enqueued).count();
    oldest = std::max(oldest, us);
    stats_.queue_us_sum += us;
  }
  int64_t prev = stats_.queue_us_max.load();
  while (oldest > prev && !stats_.queue_us_max.compare_exchange_weak(prev, oldest)) {}
  d.oldest_wait_us = oldest;
  d.batch_id = next_batch_++;
  fb->requests = std::move(reqs);
  fb->formed_at = now;
  {
    std::lock_guard<std::mutex> g(rq_mu_);
    // least-loaded GPU (ready + outstanding), round-robin on ties
    int best = 0;
    int64_t best_load = INT64_MAX;
    for (int k = 0; k < cfg_.num_gpus; ++k) {
      const int gidx = (rr_ + k) % cfg_.num_gpus;
      const int64_t load = (int64_t) ready_[gidx].size() + outstanding_[gidx];
      if (load < best_load) { best_load = load; best = gidx; }
    }
    rr_ = (best + 1) % cfg_.num_gpus;
    fb->gpu = best;
    ready_[best].push_back(fb);
  }
  ++stats_.batches;
  stats_.samples += B;
  rq_cv_.notify_all();
  return true;
}

int Server::pop_batch(int gpu, int64_t timeout_us, BatchDesc* out) {
  std::unique_lock<std::mutex> g(rq_mu_);
  const auto deadline = Clock::now() + std::chrono::microseconds(timeout_us);
  for (;;) {
    if (stop_) return 2;
    if (!ready_[gpu].empty() && outstanding_[gpu] < cfg_.max_outstanding_per_gpu) break;  // ResourceManager: bounded in-flight batches
    if (rq_cv_.wait_until(g, deadline) == std::cv_status::timeout) {
      if (!ready_[gpu].empty() && outstanding_[gpu] < cfg_.max_outstanding_per_gpu) break;
      return 1;
    }
  }
  auto fb = ready_[gpu].front();
  ready_[gpu].pop_front();
  ++outstanding_[gpu];
  inflight_[fb->desc.batch_id] = fb;
  *out = fb->desc;
  return 0;
}

int Server::complete(int64_t batch_id, const float* preds, int32_t per_sample, int status) {
  std::shared_ptr<Formed> fb;
  {
    std::lock_guard<std::mutex> g(rq_mu_);
    auto it = inflight_.find(batch_id);
    if (it == inflight_.end()) return -1;
    fb = it->second;
    inflight_.erase(it);
    --outstanding_[fb->gpu];
  }
  rq_cv_.notify_all();
  stats_.exec_us_sum += std::chrono::duration_cast<std::chrono::microseconds>(Clock::now() - fb->formed_at).count();
  // ResultSplit: request q owns rows [b0, b0 + batch_q)
  int64_t b0 = 0;
  for (auto& r : fb->requests) {
    {
      std::lock_guard<std::mutex> rg(r->mu);
      if (status == 0 && preds != nullptr) r->result.assign(preds + b0 * per_sample, preds + (b0 + r->batch_size) * per_sample);
      r->status = status;
      r->done = true;
    }
    r->cv.notify_all();
    b0 += r->batch_size;
  }
  {
    std::lock_guard<std::mutex> g(buf_mu_);
    buffers_[fb->desc.buffer_index].busy = false;
  }
  buf_cv_.notify_all();
  return 0;
}

int Server::wait(int64_t request_id, float* out, int64_t max_floats, int64_t timeout_us, int64_t* n_out) {
  std::shared_ptr<Request> r;
  {
    std::lock_guard<std::mutex> g(in_mu_);
    auto it = live_.find(request_id);
    if (it == live_.end()) return -3;
    r = it->second;
  }
  {
    std::unique_lock<std::mutex> rg(r->mu);
    if (!r->cv.wait_for(rg, std::chrono::microseconds(timeout_us), [&] { return r->done; })) { ++stats_.timeouts; return 1; }
    const int64_t n = std::min<int64_t>(max_floats, (int64_t) r->result.size());
    if (out && n > 0) std::memcpy(out, r->result.data(), n * sizeof(float));
    if (n_out) *n_out = (int64_t) r->result.size();
  }
  std::lock_guard<std::mutex> g(in_mu_);
  live_.erase(request_id);
  return r->status;
}

}  // namespace trbs

// ---- C ABI --------------------------------------------------------------------------------------------------------
#define TRB_API extern "C" __attribute__((visibility("default")))
using namespace trbs;

TRB_API void* trb_srv_create(int32_t max_batch_size, int64_t batching_interval_us, int32_t num_gpus, int32_t max_outstanding_per_gpu, int32_t batching_threads,
                             int64_t max_queue_requests) {
  Config c;
  c.max_batch_size = max_batch_size; c.batching_interval_us = batching_interval_us; c.num_gpus = num_gpus;
  c.max_outstanding_per_gpu = max_outstanding_per_gpu; c.batching_threads = batching_threads; c.max_queue_requests = max_queue_requests;
  return new Server(c);
}
TRB_API void trb_srv_destroy(void* h) { delete (Server*) h; }
TRB_API void trb_srv_shutdown(void* h) { ((Server*) h)->shutdown(); }
TRB_API void trb_srv_add_buffer(void* h, void* ptr, int64_t bytes) { ((Server*) h)->add_buffer(ptr, bytes); }
TRB_API int64_t trb_srv_submit(void* h, int32_t batch_size, int32_t num_float, const float* dense, int32_t idl_features, const int32_t* idl_lengths,
                               const int64_t* idl_values, int64_t idl_nvalues, int32_t ids_features, const int32_t* ids_lengths, const int64_t* ids_values,
                               const float* ids_weights, int64_t ids_nvalues) {
  return ((Server*) h)->submit(batch_size, num_float, dense, idl_features, idl_lengths, idl_values, idl_nvalues, ids_features, ids_lengths, ids_values, ids_weights, ids_nvalues);
}
TRB_API int trb_srv_pop_batch(void* h, int gpu, int64_t timeout_us, BatchDesc* out) { return ((Server*) h)->pop_batch(gpu, timeout_us, out); }
TRB_API int trb_srv_complete(void* h, int64_t batch_id, const float* preds, int32_t per_sample, int status) { return ((Server*) h)->complete(batch_id, preds, per_sample, status); }
TRB_API int trb_srv_wait(void* h, int64_t request_id, float* out, int64_t max_floats, int64_t timeout_us, int64_t* n_out) {
  return ((Server*) h)->wait(request_id, out, max_floats, timeout_us, n_out);
}
TRB_API void trb_srv_stats(void* h, int64_t* out8) {
  const Stats& s = ((Server*) h)->stats();
  out8[0] = s.requests; out8[1] = s.batches; out8[2] = s.samples; out8[3] = s.rejected; out8[4] = s.timeouts;
  out8[5] = s.queue_us_sum; out8[6] = s.queue_us_max; out8[7] = s.exec_us_sum;
}
