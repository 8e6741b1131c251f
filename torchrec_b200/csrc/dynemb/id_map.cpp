#include "id_map.h"

#include <algorithm>
#include <thread>

namespace trb {

IdTransformer::IdTransformer(int64_t num_embedding, Strategy strategy, int min_used_freq_power, int partitions)
    : num_embedding_(num_embedding), strategy_(strategy), min_power_(min_used_freq_power) {
  int P = std::max(1, partitions);
  while (P > 1 && num_embedding / P < 64) P >>= 1;
  const int64_t per = (num_embedding + P - 1) / P;
  for (int p = 0; p < P; ++p) {
    const int64_t base = (int64_t) p * per;
    const int64_t n = std::max<int64_t>(0, std::min(per, num_embedding - base));
    parts_.emplace_back(new Partition(base, n));
  }
}

// record layouts (uint32, smaller = evict first):
//   MIXED_LFU_LRU : [freq_power:5 | time:27]   (probabilistic power-of-two frequency counter)
//   LRU           : time
//   LFU           : saturating count
//   DISTANCE_LFU  : lxu = last access time, aux = hit count; key = count / (now - last + 1) in 16.16 fixed point at eviction
uint32_t IdTransformer::update_record(Entry* e, bool fresh, uint32_t time, std::mt19937_64& rng) const {
  switch (strategy_) {
    case Strategy::MIXED_LFU_LRU: {
      uint32_t power = fresh ? (uint32_t) min_power_ : (e->lxu >> 27);
      if (!fresh && power < 31) {
        const uint64_t bits = rng();
        if ((bits & ((1ULL << power) - 1)) == 0) ++power;  // carry with probability 2^-power
      }
      return (power << 27) | (time & 0x7ffffffu);
    }
    case Strategy::LRU:
      return time;
    case Strategy::LFU: {
      const uint32_t c = fresh ? 1u : e->lxu;
      return c == 0xffffffffu ? c : c + (fresh ? 0 : 1);
    }
    case Strategy::DISTANCE_LFU: {
      // lxu = last access time, aux = saturating hit count; the eviction key is formed at eviction time (see evict())
      e->aux = fresh ? 1u : (e->aux == 0xffffffffu ? e->aux : e->aux + 1);
      return time;
    }
  }
  return time;
}

int64_t IdTransformer::transform(const int64_t* gids, int64_t n, int64_t* slots, int64_t time, int threads) {
  std::lock_guard<std::mutex> g(mu_);
  const int P = (int) parts_.size();
  const uint32_t t32 = (uint32_t) time;
  now_ = t32;
  std::atomic<int64_t> resolved{0};
  std::vector<std::vector<int64_t>> fetch(P);
  auto work = [&](int p, const std::vector<int64_t>* idx_list) {
    std::mt19937_64 rng(0x9e3779b97f4a7c15ULL ^ (uint64_t) time * 1315423911u ^ (uint64_t) p);
    Partition& part = *parts_[p];
    int64_t ok = 0;
    auto one = [&](int64_t i) {
      const int64_t gid = gids[i];
      Entry* e = part.find(gid);
      bool fresh = false;
      if (e == nullptr) {
        e = part.insert(gid);
        if (e == nullptr) { slots[i] = -1; return; }
        fresh = true;
        fetch[p].push_back(gid);
        fetch[p].push_back(e->slot);
      }
      e->lxu = update_record(e, fresh, t32, rng);
      slots[i] = e->slot;
      ++ok;
    };
    if (idx_list) for (int64_t i : *idx_list) one(i);
    else for (int64_t i = 0; i < n; ++i) one(i);
    resolved += ok;
  };
  if (P == 1) {
    work(0, nullptr);
  } else {
    // route ids to partitions, then one worker per partition (no shared state on the hot path)
    std::vector<std::vector<int64_t>> idx(P);
    for (auto& v : idx) v.reserve(n / P + 16);
    for (int64_t i = 0; i < n; ++i) idx[mix64((uint64_t) gids[i]) % P].push_back(i);
    const int T = std::max(1, std::min(threads, P));
    if (T == 1 || n < 4096) {
      for (int p = 0; p < P; ++p) work(p, &idx[p]);
    } else {
      std::vector<std::thread> pool;
      std::atomic<int> next{0};
      for (int t = 0; t < T; ++t)
        pool.emplace_back([&] { for (int p = next++; p < P; p = next++) work(p, &idx[p]); });
      for (auto& th : pool) th.join();
    }
  }
  for (auto& f : fetch) fetch_.insert(fetch_.end(), f.begin(), f.end());
  return resolved.load();
}

int64_t IdTransformer::evict(int64_t num, int64_t* out_pairs) {
  std::lock_guard<std::mutex> g(mu_);
  if (num <= 0) return 0;
  std::priority_queue<EvictItem> heap;  // max-heap of the `num` smallest records
  const bool dist = strategy_ == Strategy::DISTANCE_LFU;
  const uint32_t now = now_;
  for (auto& part : parts_) {
    part->for_each([&](const Entry& e) {
      uint32_t key = e.lxu;
      if (dist) {
        const uint64_t age = (uint64_t) (now >= e.lxu ? now - e.lxu : 0) + 1;
        key = (uint32_t) std::min<uint64_t>(((uint64_t) std::min<uint32_t>(e.aux, 0xffffu) << 16) / age, 0xffffffffu);
      }
      if ((int64_t) heap.size() < num) heap.push(EvictItem{key, e.gid, e.slot});
      else if (key < heap.top().key) { heap.pop(); heap.push(EvictItem{key, e.gid, e.slot}); }
    });
  }
  const int P = (int) parts_.size();
  int64_t k = 0;
  while (!heap.empty()) {
    const EvictItem it = heap.top();
    heap.pop();
    out_pairs[2 * k] = it.gid;
    out_pairs[2 * k + 1] = it.slot;
    parts_[P == 1 ? 0 : mix64((uint64_t) it.gid) % P]->erase(it.gid);
    ++k;
  }
  return k;
}

int64_t IdTransformer::size() const {
  int64_t s = 0;
  for (auto& p : parts_) s += p->size();
  return s;
}

int64_t IdTransformer::take_fetch(int64_t* out_pairs, int64_t max_pairs) {
  std::lock_guard<std::mutex> g(mu_);
  const int64_t n = std::min<int64_t>(max_pairs, (int64_t) fetch_.size() / 2);
  std::memcpy(out_pairs, fetch_.data(), sizeof(int64_t) * 2 * n);
  fetch_.erase(fetch_.begin(), fetch_.begin() + 2 * n);
  return n;
}

int64_t IdTransformer::save(int64_t* out, int64_t max_entries) const {
  int64_t k = 0;
  for (auto& part : parts_) {
    part->for_each([&](const Entry& e) {
      if (k < max_entries) { out[3 * k] = e.gid; out[3 * k + 1] = e.slot; out[3 * k + 2] = e.lxu; ++k; }
    });
  }
  return k;
}

}  // namespace trb
