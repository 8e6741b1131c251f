// Native tests + micro-benchmark of the dynamic-embedding id transformer (reference: test/cpp/dynamic_embedding/*.cpp with gtest,
// benchmarks/cpp/dynamic_embedding/*.cpp with google-benchmark; neither library is in this image, so this is a plain executable:
// `dynemb_test` runs the checks (exit code != 0 on failure), `dynemb_test --bench` prints transform throughput).
// Built by torchrec_b200/csrc/build.py: build_native_test("dynemb").
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <set>
#include <vector>

#include "../dynemb/id_map.h"

using namespace trb;

static int g_failed = 0;
#define CHECK(cond)                                                              \
  do {                                                                           \
    if (!(cond)) {                                                               \
      std::fprintf(stderr, "CHECK failed %s:%d: %s\n", __FILE__, __LINE__, #cond); \
      ++g_failed;                                                                \
    }                                                                            \
  } while (0)

static void test_bitmap() {
  Bitmap b(130);
  std::set<int64_t> seen;
  for (int i = 0; i < 130; ++i) {
    const int64_t s = b.next_free();
    CHECK(s >= 0 && s < 130 && !seen.count(s));
    seen.insert(s);
  }
  CHECK(b.full() && b.next_free() == -1);
  b.release(77);
  CHECK(!b.full() && b.free_count() == 1 && b.next_free() == 77);
}

static void test_transform_is_stable_and_dense(Strategy st) {
  IdTransformer t(1000, st, 5, 4);
  std::vector<int64_t> ids(600), slots(600), again(600);
  for (int i = 0; i < 600; ++i) ids[i] = 1000003LL * i + 17;
  CHECK(t.transform(ids.data(), 600, slots.data(), 1, 4) == 600);
  std::set<int64_t> uniq(slots.begin(), slots.end());
  CHECK(uniq.size() == 600 && *uniq.begin() >= 0 && *uniq.rbegin() < 1000);
  CHECK(t.size() == 600 && t.pending_fetch() == 600);
  CHECK(t.transform(ids.data(), 600, again.data(), 2, 1) == 600);   // same ids -> same slots, nothing new to fetch
  CHECK(std::memcmp(slots.data(), again.data(), sizeof(int64_t) * 600) == 0);
  std::vector<int64_t> pairs(1200);
  CHECK(t.take_fetch(pairs.data(), 600) == 600 && t.pending_fetch() == 0);
  for (int i = 0; i < 600; ++i) CHECK(uniq.count(pairs[2 * i + 1]) == 1);
}

static void test_eviction_prefers_cold_ids() {
  IdTransformer t(64, Strategy::LFU, 5, 2);
  std::vector<int64_t> hot(16), cold(48), slots(64);
  for (int i = 0; i < 16; ++i) hot[i] = i;
  for (int i = 0; i < 48; ++i) cold[i] = 1000 + i;
  for (int rep = 0; rep < 20; ++rep) t.transform(hot.data(), 16, slots.data(), rep, 2);  // hot ids: frequency 20
  t.transform(cold.data(), 48, slots.data(), 21, 2);                                     // cold ids: frequency 1
  CHECK(t.size() == 64);
  std::vector<int64_t> ev(2 * 32);
  const int64_t n = t.evict(32, ev.data());
  CHECK(n == 32);
  for (int64_t i = 0; i < n; ++i) CHECK(ev[2 * i] >= 1000);                              // only cold ids were evicted
  CHECK(t.size() == 32);
  std::vector<int64_t> fresh(32), fs(32);
  for (int i = 0; i < 32; ++i) fresh[i] = 5000 + i;
  CHECK(t.transform(fresh.data(), 32, fs.data(), 30, 2) == 32);                          // freed slots are reused
  std::vector<int64_t> dump(3 * 64);
  CHECK(t.save(dump.data(), 64) == 64);
}

static void test_full_cache_reports_unresolved() {
  IdTransformer t(8, Strategy::LRU, 5, 1);
  std::vector<int64_t> ids(12), slots(12);
  for (int i = 0; i < 12; ++i) ids[i] = 100 + i;
  const int64_t ok = t.transform(ids.data(), 12, slots.data(), 1, 1);
  CHECK(ok == 8);
  int unresolved = 0;
  for (int i = 0; i < 12; ++i) unresolved += slots[i] < 0;
  CHECK(unresolved == 4);
}

static void bench() {
  const int64_t cap = 1 << 22, n = 1 << 20;
  std::vector<int64_t> ids(n), slots(n);
  uint64_t x = 88172645463325252ULL;
  for (int64_t i = 0; i < n; ++i) {
    x ^= x << 13; x ^= x >> 7; x ^= x << 17;
    ids[i] = (int64_t) (x % (uint64_t) (cap / 2));   // ~20 % duplicates inside a batch, fits the cache
  }
  for (int threads : {1, 4, 8}) {
    IdTransformer t(cap, Strategy::MIXED_LFU_LRU, 5, 8);
    t.transform(ids.data(), n, slots.data(), 0, threads);  // cold: inserts
    auto t0 = std::chrono::steady_clock::now();
    const int reps = 5;
    for (int r = 0; r < reps; ++r) t.transform(ids.data(), n, slots.data(), r + 1, threads);  // warm: lookups + record updates
    const double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    std::printf("transform warm: %d threads  %.1f M ids/s\n", threads, reps * (double) n / s / 1e6);
  }
}

int main(int argc, char** argv) {
  if (argc > 1 && std::strcmp(argv[1], "--bench") == 0) {
    bench();
    return 0;
  }
  test_bitmap();
  for (Strategy st : {Strategy::MIXED_LFU_LRU, Strategy::LRU, Strategy::LFU, Strategy::DISTANCE_LFU}) test_transform_is_stable_and_dense(st);
  test_eviction_prefers_cold_ids();
  test_full_cache_reports_unresolved();
  if (g_failed) {
    std::fprintf(stderr, "%d check(s) failed\n", g_failed);
    return 1;
  }
  std::printf("dynemb native tests: all passed\n");
  return 0;
}
