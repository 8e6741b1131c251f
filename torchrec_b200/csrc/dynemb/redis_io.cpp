// redis:// IO backend of the dynamic-embedding parameter server: a self-contained RESP2 client over plain TCP sockets (the image has no
// hiredis; the reference's plugin links it: torchrec/csrc/dynamic_embedding/details/redis/redis_io.{h,cpp}).
//
//   redis://[:password@]host[:port][/db][?prefix=<str>&chunk=<rows per command>&timeout_ms=<n>&pool=<connections>]
//
// Data model: one Redis HASH per embedding table, `<prefix><table>`; field = the 8 raw bytes of the global id (little endian), value =
// the row blob (weight row + optimizer-state rows, see dynamic_embedding/ps.py). So
//   push  = HSET  key f1 v1 f2 v2 ...     (chunked, chunks pipelined: all commands of a call are written before the first reply is read)
//   pull  = HMGET key f1 f2 ...           (nil reply = id unknown to the PS; a blob of another size = treated as unknown)
//   size  = HLEN  key
// A backend owns a small pool of connections so the PS's IO threads do not serialise on one socket; every connection authenticates
// (AUTH) and selects the db on connect, and reconnects once when the server closed it between calls.
#include <arpa/inet.h>
#include <netdb.h>
#include <netinet/in.h>
#include <netinet/tcp.h>
#include <sys/socket.h>
#include <sys/time.h>
#include <unistd.h>

#include <cerrno>
#include <condition_variable>
#include <cstring>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>

#include "io.h"

namespace trb {
namespace {

struct RedisUrl {
  std::string host = "127.0.0.1", password, prefix = "trb:";
  int port = 6379, db = 0, chunk = 512, timeout_ms = 5000, pool = 2;
};

int to_int(const std::string& s, int dflt) {
  try {
    return s.empty() ? dflt : std::stoi(s);
  } catch (...) {
    return dflt;
  }
}

RedisUrl parse_url(const std::string& url) {
  RedisUrl u;
  std::string rest = url.substr(url.find("://") + 3);
  std::string query;
  auto q = rest.find('?');
  if (q != std::string::npos) {
    query = rest.substr(q + 1);
    rest = rest.substr(0, q);
  }
  auto at = rest.rfind('@');
  if (at != std::string::npos) {
    std::string cred = rest.substr(0, at);
    rest = rest.substr(at + 1);
    auto c = cred.find(':');
    u.password = c == std::string::npos ? cred : cred.substr(c + 1);
  }
  auto slash = rest.find('/');
  if (slash != std::string::npos) {
    u.db = to_int(rest.substr(slash + 1), 0);
    rest = rest.substr(0, slash);
  }
  auto colon = rest.rfind(':');
  if (colon != std::string::npos) {
    u.port = to_int(rest.substr(colon + 1), 6379);
    rest = rest.substr(0, colon);
  }
  if (!rest.empty()) u.host = rest;
  size_t p = 0;
  while (p < query.size()) {
    auto amp = query.find('&', p);
    std::string kv = query.substr(p, amp == std::string::npos ? std::string::npos : amp - p);
    auto eq = kv.find('=');
    if (eq != std::string::npos) {
      const std::string k = kv.substr(0, eq), v = kv.substr(eq + 1);
      if (k == "prefix") u.prefix = v;
      else if (k == "chunk") u.chunk = std::max(1, to_int(v, u.chunk));
      else if (k == "timeout_ms") u.timeout_ms = std::max(1, to_int(v, u.timeout_ms));
      else if (k == "pool") u.pool = std::max(1, to_int(v, u.pool));
    }
    if (amp == std::string::npos) break;
    p = amp + 1;
  }
  return u;
}

// One RESP reply, flattened: arrays are returned element by element through the callbacks of the caller.
struct Reply {
  char type = 0;  // '+', '-', ':', '$', '*'
  int64_t integer = 0;  // ':' value, '$' length (-1 = nil), '*' element count
  std::string text;     // '+' / '-' line, '$' payload
};

class Conn {
 public:
  explicit Conn(const RedisUrl& u) : u_(u) {}
  ~Conn() { close_fd(); }

  void ensure() {
    if (fd_ >= 0) return;
    addrinfo hints{}, *res = nullptr;
    hints.ai_family = AF_UNSPEC;
    hints.ai_socktype = SOCK_STREAM;
    const std::string port = std::to_string(u_.port);
    if (::getaddrinfo(u_.host.c_str(), port.c_str(), &hints, &res) != 0 || !res) throw std::runtime_error("redis: cannot resolve " + u_.host);
    int fd = -1;
    for (addrinfo* a = res; a; a = a->ai_next) {
      fd = ::socket(a->ai_family, a->ai_socktype, a->ai_protocol);
      if (fd < 0) continue;
      timeval tv{u_.timeout_ms / 1000, (u_.timeout_ms % 1000) * 1000};
      ::setsockopt(fd, SOL_SOCKET, SO_RCVTIMEO, &tv, sizeof(tv));
      ::setsockopt(fd, SOL_SOCKET, SO_SNDTIMEO, &tv, sizeof(tv));
      int one = 1;
      ::setsockopt(fd, IPPROTO_TCP, TCP_NODELAY, &one, sizeof(one));
      if (::connect(fd, a->ai_addr, a->ai_addrlen) == 0) break;
      ::close(fd);
      fd = -1;
    }
    ::freeaddrinfo(res);
    if (fd < 0) throw std::runtime_error("redis: cannot connect to " + u_.host + ":" + port);
    fd_ = fd;
    rpos_ = rlen_ = 0;
    if (!u_.password.empty()) expect_ok({"AUTH", u_.password});
    if (u_.db != 0) expect_ok({"SELECT", std::to_string(u_.db)});
  }

  void close_fd() {
    if (fd_ >= 0) ::close(fd_);
    fd_ = -1;
  }

  // ---- request building: commands are appended to one buffer and flushed together (pipelining) ----
  void begin(size_t argc) {
    out_ += '*';
    out_ += std::to_string(argc);
    out_ += "\r\n";
  }
  void arg(const void* p, size_t n) {
    out_ += '$';
    out_ += std::to_string(n);
    out_ += "\r\n";
    out_.append(reinterpret_cast<const char*>(p), n);
    out_ += "\r\n";
  }
  void arg(const std::string& s) { arg(s.data(), s.size()); }

  void flush() {
    size_t off = 0;
    while (off < out_.size()) {
      const ssize_t w = ::send(fd_, out_.data() + off, out_.size() - off, MSG_NOSIGNAL);
      if (w <= 0) {
        if (w < 0 && errno == EINTR) continue;
        out_.clear();
        throw std::runtime_error("redis: send failed");
      }
      off += (size_t) w;
    }
    out_.clear();
  }
  bool has_pending() const { return !out_.empty(); }
  void drop_pending() { out_.clear(); }

  // ---- reply parsing ----
  Reply read_reply() {
    Reply r;
    std::string line = read_line();
    if (line.empty()) throw std::runtime_error("redis: empty reply");
    r.type = line[0];
    const std::string body = line.substr(1);
    switch (r.type) {
      case '+':
      case '-':
        r.text = body;
        break;
      case ':':
      case '*':
        r.integer = std::stoll(body);
        break;
      case '$':
        r.integer = std::stoll(body);
        if (r.integer >= 0) {
          r.text.resize((size_t) r.integer);
          read_exact(&r.text[0], (size_t) r.integer);
          char crlf[2];
          read_exact(crlf, 2);
        }
        break;
      default:
        throw std::runtime_error("redis: protocol error");
    }
    return r;
  }

  void expect_ok(const std::vector<std::string>& cmd) {
    begin(cmd.size());
    for (auto& a : cmd) arg(a);
    flush();
    Reply r = read_reply();
    if (r.type == '-') throw std::runtime_error("redis: " + r.text);
  }

 private:
  void fill() {
    if (rpos_ < rlen_) return;
    for (;;) {
      const ssize_t n = ::recv(fd_, rbuf_, sizeof(rbuf_), 0);
      if (n > 0) {
        rpos_ = 0;
        rlen_ = (size_t) n;
        return;
      }
      if (n < 0 && errno == EINTR) continue;
      throw std::runtime_error(n == 0 ? "redis: connection closed" : "redis: recv failed / timed out");
    }
  }
  std::string read_line() {
    std::string s;
    for (;;) {
      fill();
      while (rpos_ < rlen_) {
        const char c = rbuf_[rpos_++];
        if (c == '\n') {
          if (!s.empty() && s.back() == '\r') s.pop_back();
          return s;
        }
        s += c;
      }
    }
  }
  void read_exact(char* dst, size_t n) {
    while (n) {
      fill();
      const size_t k = std::min(n, rlen_ - rpos_);
      std::memcpy(dst, rbuf_ + rpos_, k);
      rpos_ += k;
      dst += k;
      n -= k;
    }
  }

  RedisUrl u_;
  int fd_ = -1;
  std::string out_;
  char rbuf_[1 << 16];
  size_t rpos_ = 0, rlen_ = 0;
};

class RedisBackend : public IOBackend {
 public:
  explicit RedisBackend(const std::string& url) : u_(parse_url(url)) {
    for (int i = 0; i < u_.pool; ++i) idle_.push_back(new Conn(u_));
    Lease l(*this);  // fail at open time, not at the first push, when the server is unreachable
    l.c->ensure();
  }
  ~RedisBackend() override {
    for (Conn* c : idle_) delete c;
  }

  void push(const std::string& table, const int64_t* gids, int64_t n, const uint8_t* rows, int64_t row_bytes) override {
    const std::string key = u_.prefix + table;
    with_retry([&](Conn& c) {
      int64_t cmds = 0;
      for (int64_t i0 = 0; i0 < n; i0 += u_.chunk) {
        const int64_t k = std::min<int64_t>(u_.chunk, n - i0);
        c.begin((size_t) (2 + 2 * k));
        c.arg("HSET", 4);
        c.arg(key);
        for (int64_t i = i0; i < i0 + k; ++i) {
          c.arg(&gids[i], 8);
          c.arg(rows + i * row_bytes, (size_t) row_bytes);
        }
        ++cmds;
      }
      c.flush();
      for (int64_t j = 0; j < cmds; ++j) {
        Reply r = c.read_reply();
        if (r.type == '-') throw std::runtime_error("redis HSET: " + r.text);
      }
    });
  }

  void pull(const std::string& table, const int64_t* gids, int64_t n, uint8_t* rows, int64_t row_bytes, uint8_t* found) override {
    const std::string key = u_.prefix + table;
    with_retry([&](Conn& c) {
      std::vector<int64_t> counts;
      for (int64_t i0 = 0; i0 < n; i0 += u_.chunk) {
        const int64_t k = std::min<int64_t>(u_.chunk, n - i0);
        c.begin((size_t) (2 + k));
        c.arg("HMGET", 5);
        c.arg(key);
        for (int64_t i = i0; i < i0 + k; ++i) c.arg(&gids[i], 8);
        counts.push_back(k);
      }
      c.flush();
      int64_t i = 0;
      for (int64_t k : counts) {
        Reply head = c.read_reply();
        if (head.type == '-') throw std::runtime_error("redis HMGET: " + head.text);
        if (head.type != '*' || head.integer != k) throw std::runtime_error("redis HMGET: unexpected reply shape");
        for (int64_t j = 0; j < k; ++j, ++i) {
          Reply r = c.read_reply();
          found[i] = 0;
          if (r.type == '$' && r.integer == row_bytes) {
            std::memcpy(rows + i * row_bytes, r.text.data(), (size_t) row_bytes);
            found[i] = 1;
          }
        }
      }
    });
  }

  int64_t size(const std::string& table) override {
    int64_t res = 0;
    const std::string key = u_.prefix + table;
    with_retry([&](Conn& c) {
      c.begin(2);
      c.arg("HLEN", 4);
      c.arg(key);
      c.flush();
      Reply r = c.read_reply();
      if (r.type == '-') throw std::runtime_error("redis HLEN: " + r.text);
      res = r.integer;
    });
    return res;
  }

 private:
  struct Lease {
    explicit Lease(RedisBackend& b) : b(b) {
      std::unique_lock<std::mutex> g(b.mu_);
      b.cv_.wait(g, [&] { return !b.idle_.empty(); });
      c = b.idle_.back();
      b.idle_.pop_back();
    }
    ~Lease() {
      {
        std::lock_guard<std::mutex> g(b.mu_);
        b.idle_.push_back(c);
      }
      b.cv_.notify_one();
    }
    RedisBackend& b;
    Conn* c;
  };

  template <typename F>
  void with_retry(F&& fn) {
    Lease l(*this);
    for (int attempt = 0;; ++attempt) {
      try {
        l.c->ensure();
        fn(*l.c);
        return;
      } catch (const std::exception&) {
        // a half-read pipeline cannot be resumed: drop the socket; one reconnect covers servers that closed an idle connection
        l.c->drop_pending();
        l.c->close_fd();
        if (attempt >= 1) throw;
      }
    }
  }

  RedisUrl u_;
  std::mutex mu_;
  std::condition_variable cv_;
  std::vector<Conn*> idle_;
};

struct Registrar {
  Registrar() {
    IORegistry::instance().add("redis", [](const std::string& url) { return std::make_shared<RedisBackend>(url); });
  }
} g_registrar;

}  // namespace
}  // namespace trb
