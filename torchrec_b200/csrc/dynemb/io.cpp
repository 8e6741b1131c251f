#include "io.h"

#include <dlfcn.h>
#include <sys/stat.h>

#include <cstdio>
#include <cstring>
#include <map>
#include <mutex>
#include <shared_mutex>
#include <stdexcept>
#include <unordered_map>

namespace trb {

// ---- memory:// ------------------------------------------------------------------------------------------------
class MemoryBackend : public IOBackend {
 public:
  void push(const std::string& table, const int64_t* gids, int64_t n, const uint8_t* rows, int64_t row_bytes) override {
    std::unique_lock<std::shared_mutex> g(mu_);
    auto& t = tables_[table];
    for (int64_t i = 0; i < n; ++i) t[gids[i]].assign(rows + i * row_bytes, rows + (i + 1) * row_bytes);
  }
  void pull(const std::string& table, const int64_t* gids, int64_t n, uint8_t* rows, int64_t row_bytes, uint8_t* found) override {
    std::shared_lock<std::shared_mutex> g(mu_);
    auto it = tables_.find(table);
    for (int64_t i = 0; i < n; ++i) {
      found[i] = 0;
      if (it == tables_.end()) continue;
      auto r = it->second.find(gids[i]);
      if (r == it->second.end() || (int64_t) r->second.size() != row_bytes) continue;
      std::memcpy(rows + i * row_bytes, r->second.data(), row_bytes);
      found[i] = 1;
    }
  }
  int64_t size(const std::string& table) override {
    std::shared_lock<std::shared_mutex> g(mu_);
    auto it = tables_.find(table);
    return it == tables_.end() ? 0 : (int64_t) it->second.size();
  }

 private:
  std::shared_mutex mu_;
  std::unordered_map<std::string, std::unordered_map<int64_t, std::vector<uint8_t>>> tables_;
};

// ---- file:// : append-only log per table ([gid:int64][len:int64][blob]) + in-memory index of the latest offset --------
class FileBackend : public IOBackend {
 public:
  explicit FileBackend(std::string dir) : dir_(std::move(dir)) { ::mkdir(dir_.c_str(), 0755); }
  ~FileBackend() override {
    for (auto& kv : tables_) if (kv.second.f) std::fclose(kv.second.f);
  }
  void push(const std::string& table, const int64_t* gids, int64_t n, const uint8_t* rows, int64_t row_bytes) override {
    std::lock_guard<std::mutex> g(mu_);
    Table& t = open(table);
    std::fseek(t.f, 0, SEEK_END);
    for (int64_t i = 0; i < n; ++i) {
      const int64_t off = std::ftell(t.f);
      std::fwrite(&gids[i], 8, 1, t.f);
      std::fwrite(&row_bytes, 8, 1, t.f);
      std::fwrite(rows + i * row_bytes, 1, row_bytes, t.f);
      t.index[gids[i]] = off;
    }
    std::fflush(t.f);
  }
  void pull(const std::string& table, const int64_t* gids, int64_t n, uint8_t* rows, int64_t row_bytes, uint8_t* found) override {
    std::lock_guard<std::mutex> g(mu_);
    Table& t = open(table);
    for (int64_t i = 0; i < n; ++i) {
      found[i] = 0;
      auto it = t.index.find(gids[i]);
      if (it == t.index.end()) continue;
      int64_t hdr[2];
      std::fseek(t.f, it->second, SEEK_SET);
      if (std::fread(hdr, 8, 2, t.f) != 2 || hdr[1] != row_bytes) continue;
      if ((int64_t) std::fread(rows + i * row_bytes, 1, row_bytes, t.f) == row_bytes) found[i] = 1;
    }
  }
  int64_t size(const std::string& table) override {
    std::lock_guard<std::mutex> g(mu_);
    return (int64_t) open(table).index.size();
  }

 private:
  struct Table {
    FILE* f = nullptr;
    std::unordered_map<int64_t, int64_t> index;
  };
  Table& open(const std::string& table) {
    auto it = tables_.find(table);
    if (it != tables_.end()) return it->second;
    Table& t = tables_[table];
    const std::string path = dir_ + "/" + table + ".log";
    t.f = std::fopen(path.c_str(), "a+b");
    if (!t.f) throw std::runtime_error("cannot open " + path);
    // rebuild the index from an existing log
    std::fseek(t.f, 0, SEEK_SET);
    for (;;) {
      const int64_t off = std::ftell(t.f);
      int64_t hdr[2];
      if (std::fread(hdr, 8, 2, t.f) != 2) break;
      if (std::fseek(t.f, hdr[1], SEEK_CUR) != 0) break;
      t.index[hdr[0]] = off;
    }
    return t;
  }
  std::string dir_;
  std::mutex mu_;
  std::map<std::string, Table> tables_;
};

// ---- plugin adapter ----------------------------------------------------------------------------------------------
class PluginBackend : public IOBackend {
 public:
  PluginBackend(const IOPlugin& p, const std::string& url) : p_(p), self_(p.create(url.c_str())) {}
  ~PluginBackend() override { p_.destroy(self_); }
  void push(const std::string& t, const int64_t* g, int64_t n, const uint8_t* r, int64_t rb) override { p_.push(self_, t.c_str(), g, n, r, rb); }
  void pull(const std::string& t, const int64_t* g, int64_t n, uint8_t* r, int64_t rb, uint8_t* f) override { p_.pull(self_, t.c_str(), g, n, r, rb, f); }
  int64_t size(const std::string& t) override { return p_.size(self_, t.c_str()); }

 private:
  IOPlugin p_;
  void* self_;
};

struct IORegistry::Impl {
  std::mutex mu;
  std::map<std::string, IOFactory> factories;
  std::map<std::string, std::shared_ptr<IOBackend>> opened;  // one backend per URL (memory:// stores are shared by name)
};

IORegistry::IORegistry() : impl_(new Impl) {
  impl_->factories["memory"] = [](const std::string&) { return std::make_shared<MemoryBackend>(); };
  impl_->factories["file"] = [](const std::string& url) { return std::make_shared<FileBackend>(url.substr(url.find("://") + 3)); };
}

IORegistry& IORegistry::instance() {
  static IORegistry r;
  return r;
}

void IORegistry::add(const std::string& scheme, IOFactory f) {
  std::lock_guard<std::mutex> g(impl_->mu);
  impl_->factories[scheme] = std::move(f);
}

bool IORegistry::load_plugin(const std::string& scheme, const std::string& so_path) {
  void* h = ::dlopen(so_path.c_str(), RTLD_NOW | RTLD_LOCAL);
  if (!h) return false;
  auto* p = reinterpret_cast<IOPlugin*>(::dlsym(h, "trb_io_plugin"));
  if (!p || !p->create || !p->destroy || !p->push || !p->pull || !p->size) return false;
  const IOPlugin plugin = *p;
  add(scheme, [plugin](const std::string& url) { return std::make_shared<PluginBackend>(plugin, url); });
  return true;
}

std::shared_ptr<IOBackend> IORegistry::open(const std::string& url) {
  std::lock_guard<std::mutex> g(impl_->mu);
  auto it = impl_->opened.find(url);
  if (it != impl_->opened.end()) return it->second;
  const auto pos = url.find("://");
  if (pos == std::string::npos) throw std::runtime_error("bad IO url: " + url);
  auto f = impl_->factories.find(url.substr(0, pos));
  if (f == impl_->factories.end()) throw std::runtime_error("no IO backend registered for " + url);
  auto b = f->second(url);
  impl_->opened[url] = b;
  return b;
}

}  // namespace trb
