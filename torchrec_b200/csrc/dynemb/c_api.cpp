// C ABI of the dynamic-embedding runtime (loaded with ctypes; no torch headers, builds in seconds).
#include <condition_variable>
#include <deque>
#include <functional>
#include <map>
#include <future>
#include <thread>

#include "id_map.h"
#include "io.h"

#define TRB_API extern "C" __attribute__((visibility("default")))

using namespace trb;

// ---- id transformer ------------------------------------------------------------------------------------------
TRB_API void* trb_idt_create(int64_t num_embedding, int strategy, int min_used_freq_power, int partitions) {
  return new IdTransformer(num_embedding, (Strategy) strategy, min_used_freq_power, partitions);
}
TRB_API void trb_idt_destroy(void* h) { delete (IdTransformer*) h; }
TRB_API int64_t trb_idt_transform(void* h, const int64_t* gids, int64_t n, int64_t* slots, int64_t time, int threads) {
  return ((IdTransformer*) h)->transform(gids, n, slots, time, threads);
}
TRB_API int64_t trb_idt_evict(void* h, int64_t num, int64_t* out_pairs) { return ((IdTransformer*) h)->evict(num, out_pairs); }
TRB_API int64_t trb_idt_size(void* h) { return ((IdTransformer*) h)->size(); }
TRB_API int64_t trb_idt_pending_fetch(void* h) { return ((IdTransformer*) h)->pending_fetch(); }
TRB_API int64_t trb_idt_take_fetch(void* h, int64_t* out_pairs, int64_t max_pairs) { return ((IdTransformer*) h)->take_fetch(out_pairs, max_pairs); }
TRB_API int64_t trb_idt_save(void* h, int64_t* out_triples, int64_t max_entries) { return ((IdTransformer*) h)->save(out_triples, max_entries); }

// ---- parameter server client: async push/pull on a small worker pool ----------------------------------------------
namespace {
class Pool {
 public:
  explicit Pool(int n) {
    for (int i = 0; i < n; ++i) workers_.emplace_back([this] { run(); });
  }
  ~Pool() {
    { std::lock_guard<std::mutex> g(mu_); stop_ = true; }
    cv_.notify_all();
    for (auto& w : workers_) w.join();
  }
  std::shared_future<int> submit(std::function<int()> fn) {
    auto task = std::make_shared<std::packaged_task<int()>>(std::move(fn));
    std::shared_future<int> fut = task->get_future().share();
    { std::lock_guard<std::mutex> g(mu_); q_.push_back([task] { (*task)(); }); }
    cv_.notify_one();
    return fut;
  }

 private:
  void run() {
    for (;;) {
      std::function<void()> job;
      {
        std::unique_lock<std::mutex> g(mu_);
        cv_.wait(g, [this] { return stop_ || !q_.empty(); });
        if (stop_ && q_.empty()) return;
        job = std::move(q_.front());
        q_.pop_front();
      }
      job();
    }
  }
  std::mutex mu_;
  std::condition_variable cv_;
  std::deque<std::function<void()>> q_;
  std::vector<std::thread> workers_;
  bool stop_ = false;
};

struct PS {
  std::string table;
  std::shared_ptr<IOBackend> io;
  std::unique_ptr<Pool> pool;
  std::mutex mu;
  int64_t next_ticket = 1;
  std::map<int64_t, std::shared_future<int>> tickets;
};
}  // namespace


TRB_API void* trb_ps_create(const char* table, const char* url, int io_threads) {
  try {
    auto* ps = new PS;
    ps->table = table;
    ps->io = IORegistry::instance().open(url);
    ps->pool.reset(new Pool(io_threads > 0 ? io_threads : 2));
    return ps;
  } catch (const std::exception&) {
    return nullptr;
  }
}
TRB_API void trb_ps_destroy(void* h) { delete (PS*) h; }
TRB_API int trb_io_load_plugin(const char* scheme, const char* so_path) { return IORegistry::instance().load_plugin(scheme, so_path) ? 0 : -1; }

// The caller keeps gids / rows alive until trb_ps_wait(ticket) returns.
TRB_API int64_t trb_ps_push_async(void* h, const int64_t* gids, int64_t n, const uint8_t* rows, int64_t row_bytes) {
  auto* ps = (PS*) h;
  auto fut = ps->pool->submit([=] { try { ps->io->push(ps->table, gids, n, rows, row_bytes); return 0; } catch (...) { return -1; } });
  std::lock_guard<std::mutex> g(ps->mu);
  ps->tickets[ps->next_ticket] = fut;
  return ps->next_ticket++;
}
TRB_API int64_t trb_ps_pull_async(void* h, const int64_t* gids, int64_t n, uint8_t* rows, int64_t row_bytes, uint8_t* found) {
  auto* ps = (PS*) h;
  auto fut = ps->pool->submit([=] { try { ps->io->pull(ps->table, gids, n, rows, row_bytes, found); return 0; } catch (...) { return -1; } });
  std::lock_guard<std::mutex> g(ps->mu);
  ps->tickets[ps->next_ticket] = fut;
  return ps->next_ticket++;
}
TRB_API int trb_ps_wait(void* h, int64_t ticket) {
  auto* ps = (PS*) h;
  std::shared_future<int> fut;
  {
    std::lock_guard<std::mutex> g(ps->mu);
    auto it = ps->tickets.find(ticket);
    if (it == ps->tickets.end()) return -2;
    fut = it->second;
    ps->tickets.erase(it);
  }
  return fut.get();
}
TRB_API int64_t trb_ps_size(void* h) { auto* ps = (PS*) h; return ps->io->size(ps->table); }
