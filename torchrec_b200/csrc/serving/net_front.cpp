// Native network front of the inference server: accepts TCP connections, parses `predictor.PredictionRequest` messages, feeds the
// batching core (serving.cpp) and writes `predictor.PredictionResponse` messages back - no Python between the socket and the batching
// queue (Python only runs the model on the executor threads).
//
// The reference's front is a gRPC C++ service (torchrec/inference/server.cpp:51-179). grpc++ / protoc are not in this image, so the
// transport here is a minimal framed protocol instead of HTTP/2: every message is `uint32 little-endian length` + the protobuf bytes
// of the SAME messages the gRPC front uses (inference/server.py builds them with the protobuf runtime; this file reads / writes the
// wire format by hand: varints, length-delimited fields, packed floats). One thread per connection (requests of a connection are
// answered in order; concurrency = connections, as with blocking unary gRPC stubs), bounded by `max_connections`.
//
//   PredictionRequest { int32 batch_size = 1; FloatFeatures float_features = 2; SparseFeatures id_list_features = 3;
//                       SparseFeatures id_score_list_features = 4; ... }
//   FloatFeatures  { int32 num_features = 1; bytes values = 2; }                                (fp32, [batch, num_features])
//   SparseFeatures { int32 num_features = 1; bytes lengths = 2; bytes values = 3; bytes weights = 4; }   (int32 / int64 / fp32)
//   PredictionResponse { map<string, FloatVec> predictions = 1; }    FloatVec { repeated float data = 1; }
// A response with an empty map and the extra field `status = 15` (varint, non-zero) reports a failed request.
#include <arpa/inet.h>
#include <netinet/in.h>
#include <netinet/tcp.h>
#include <sys/socket.h>
#include <unistd.h>

#include <atomic>
#include <cerrno>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "serving.h"

namespace trbs {
namespace {

// ---- protobuf wire helpers ----------------------------------------------------------------------------------------------------------
struct Span {
  const uint8_t* p = nullptr;
  size_t n = 0;
};

bool read_varint(const uint8_t*& p, const uint8_t* end, uint64_t* out) {
  uint64_t v = 0;
  for (int shift = 0; shift < 64 && p < end; shift += 7) {
    const uint8_t b = *p++;
    v |= (uint64_t) (b & 0x7f) << shift;
    if (!(b & 0x80)) {
      *out = v;
      return true;
    }
  }
  return false;
}

// Walks the fields of one message; calls varint(field, value) / bytes(field, span). Unknown wire types 1 / 5 are skipped.
template <typename FV, typename FB>
bool walk(Span m, FV on_varint, FB on_bytes) {
  const uint8_t* p = m.p;
  const uint8_t* end = m.p + m.n;
  while (p < end) {
    uint64_t tag;
    if (!read_varint(p, end, &tag)) return false;
    const int field = (int) (tag >> 3), wt = (int) (tag & 7);
    if (wt == 0) {
      uint64_t v;
      if (!read_varint(p, end, &v)) return false;
      on_varint(field, v);
    } else if (wt == 2) {
      uint64_t len;
      if (!read_varint(p, end, &len) || len > (uint64_t) (end - p)) return false;
      on_bytes(field, Span{p, (size_t) len});
      p += len;
    } else if (wt == 1) {
      if (end - p < 8) return false;
      p += 8;
    } else if (wt == 5) {
      if (end - p < 4) return false;
      p += 4;
    } else {
      return false;
    }
  }
  return true;
}

void put_varint(std::string& s, uint64_t v) {
  while (v >= 0x80) {
    s.push_back((char) (v | 0x80));
    v >>= 7;
  }
  s.push_back((char) v);
}

struct Sparse {
  int32_t num_features = 0;
  Span lengths, values, weights;
};

bool parse_sparse(Span m, Sparse* s) {
  return walk(m, [&](int f, uint64_t v) { if (f == 1) s->num_features = (int32_t) v; },
              [&](int f, Span b) { if (f == 2) s->lengths = b; else if (f == 3) s->values = b; else if (f == 4) s->weights = b; });
}

struct Parsed {
  int32_t batch_size = 0, num_float = 0;
  Span dense;
  Sparse idl, ids;
};

bool parse_request(Span m, Parsed* r) {
  bool ok = true;
  ok &= walk(m, [&](int f, uint64_t v) { if (f == 1) r->batch_size = (int32_t) v; },
             [&](int f, Span b) {
               if (f == 2) {
                 ok &= walk(b, [&](int ff, uint64_t v) { if (ff == 1) r->num_float = (int32_t) v; }, [&](int ff, Span bb) { if (ff == 2) r->dense = bb; });
               } else if (f == 3) {
                 ok &= parse_sparse(b, &r->idl);
               } else if (f == 4) {
                 ok &= parse_sparse(b, &r->ids);
               }
             });
  if (!ok || r->batch_size <= 0) return false;
  // sizes must agree with what the fields claim (the core trusts its arguments)
  if (r->num_float > 0 && r->dense.n != (size_t) r->batch_size * r->num_float * 4) return false;
  auto check = [&](const Sparse& s, bool weighted) {
    if (s.num_features <= 0) return s.lengths.n == 0 && s.values.n == 0;
    if (s.lengths.n != (size_t) s.num_features * r->batch_size * 4 || s.values.n % 8) return false;
    int64_t total = 0;
    for (size_t i = 0; i < s.lengths.n / 4; ++i) {
      int32_t l;
      std::memcpy(&l, s.lengths.p + 4 * i, 4);
      if (l < 0) return false;
      total += l;
    }
    if ((size_t) total * 8 != s.values.n) return false;
    return !weighted || s.weights.n == (size_t) total * 4;
  };
  return check(r->idl, false) && check(r->ids, true);
}

std::string encode_response(const std::string& task, const float* data, int64_t n, int status) {
  std::string out;
  if (status == 0) {
    std::string vec;  // FloatVec { repeated float data = 1 [packed] }
    vec.push_back((char) 0x0A);
    put_varint(vec, (uint64_t) n * 4);
    vec.append(reinterpret_cast<const char*>(data), (size_t) n * 4);
    std::string entry;  // map entry { string key = 1; FloatVec value = 2; }
    entry.push_back((char) 0x0A);
    put_varint(entry, task.size());
    entry += task;
    entry.push_back((char) 0x12);
    put_varint(entry, vec.size());
    entry += vec;
    out.push_back((char) 0x0A);  // predictions = 1
    put_varint(out, entry.size());
    out += entry;
  } else {
    out.push_back((char) 0x78);  // field 15, varint
    put_varint(out, (uint64_t) (uint32_t) status);
  }
  return out;
}

bool read_exact(int fd, void* dst, size_t n) {
  uint8_t* p = (uint8_t*) dst;
  while (n) {
    const ssize_t r = ::recv(fd, p, n, 0);
    if (r > 0) {
      p += r;
      n -= (size_t) r;
    } else if (r < 0 && errno == EINTR) {
      continue;
    } else {
      return false;
    }
  }
  return true;
}

bool write_all(int fd, const void* src, size_t n) {
  const uint8_t* p = (const uint8_t*) src;
  while (n) {
    const ssize_t w = ::send(fd, p, n, MSG_NOSIGNAL);
    if (w > 0) {
      p += w;
      n -= (size_t) w;
    } else if (w < 0 && errno == EINTR) {
      continue;
    } else {
      return false;
    }
  }
  return true;
}

// 8-byte aligned copies of the payload arrays (protobuf bytes fields sit at arbitrary offsets of the receive buffer)
template <typename T>
const T* aligned(Span s, std::vector<T>& store) {
  if (s.n == 0) return nullptr;
  store.resize(s.n / sizeof(T));
  std::memcpy(store.data(), s.p, store.size() * sizeof(T));
  return store.data();
}

}  // namespace

class NetFront {
 public:
  NetFront(Server* core, std::string task, int outputs_per_sample, int max_connections, int64_t request_timeout_us)
      : core_(core), task_(std::move(task)), per_sample_(outputs_per_sample), max_conn_(max_connections), timeout_us_(request_timeout_us) {}
  ~NetFront() { stop(); }

  int listen_on(int port) {
    fd_ = ::socket(AF_INET, SOCK_STREAM, 0);
    if (fd_ < 0) return -1;
    int one = 1;
    ::setsockopt(fd_, SOL_SOCKET, SO_REUSEADDR, &one, sizeof(one));
    sockaddr_in a{};
    a.sin_family = AF_INET;
    a.sin_addr.s_addr = htonl(INADDR_LOOPBACK);
    a.sin_port = htons((uint16_t) port);
    if (::bind(fd_, (sockaddr*) &a, sizeof(a)) != 0 || ::listen(fd_, 128) != 0) {
      ::close(fd_);
      fd_ = -1;
      return -1;
    }
    socklen_t len = sizeof(a);
    ::getsockname(fd_, (sockaddr*) &a, &len);
    port_ = ntohs(a.sin_port);
    acceptor_ = std::thread([this] { accept_loop(); });
    return port_;
  }

  void stop() {
    if (stop_.exchange(true)) return;
    if (fd_ >= 0) {
      ::shutdown(fd_, SHUT_RDWR);
      ::close(fd_);
    }
    if (acceptor_.joinable()) acceptor_.join();
    std::vector<std::thread> ws;
    {
      std::lock_guard<std::mutex> g(mu_);
      for (int c : conns_) ::shutdown(c, SHUT_RDWR);
      ws.swap(workers_);
    }
    for (auto& t : ws)
      if (t.joinable()) t.join();
  }

  std::atomic<int64_t> served{0}, malformed{0}, refused{0};

 private:
  void accept_loop() {
    while (!stop_) {
      const int c = ::accept(fd_, nullptr, nullptr);
      if (c < 0) {
        if (errno == EINTR) continue;
        return;
      }
      std::lock_guard<std::mutex> g(mu_);
      if ((int) conns_.size() >= max_conn_ || stop_) {
        ++refused;
        ::close(c);
        continue;
      }
      int one = 1;
      ::setsockopt(c, IPPROTO_TCP, TCP_NODELAY, &one, sizeof(one));
      conns_.push_back(c);
      workers_.emplace_back([this, c] { serve(c); });
    }
  }

  void serve(int c) {
    std::vector<uint8_t> buf;
    std::vector<float> dense, weights, result;
    std::vector<int32_t> l1, l2;
    std::vector<int64_t> v1, v2;
    while (!stop_) {
      uint32_t len = 0;
      if (!read_exact(c, &len, 4)) break;
      if (len == 0 || len > (256u << 20)) break;
      buf.resize(len);
      if (!read_exact(c, buf.data(), len)) break;
      Parsed r;
      std::string resp;
      if (!parse_request(Span{buf.data(), len}, &r)) {
        ++malformed;
        resp = encode_response(task_, nullptr, 0, 3 /* INVALID_ARGUMENT */);
      } else {
        const int64_t id = core_->submit(r.batch_size, r.num_float, aligned(r.dense, dense), r.idl.num_features, aligned(r.idl.lengths, l1), aligned(r.idl.values, v1),
                                         (int64_t) (r.idl.values.n / 8), r.ids.num_features, aligned(r.ids.lengths, l2), aligned(r.ids.values, v2),
                                         aligned(r.ids.weights, weights), (int64_t) (r.ids.values.n / 8));
        if (id < 0) {
          resp = encode_response(task_, nullptr, 0, 8 /* RESOURCE_EXHAUSTED: intake queue full */);
        } else {
          result.resize((size_t) r.batch_size * per_sample_);
          int64_t n = 0;
          const int rc = core_->wait(id, result.data(), (int64_t) result.size(), timeout_us_, &n);
          resp = rc == 0 ? encode_response(task_, result.data(), n, 0) : encode_response(task_, nullptr, 0, rc == 1 ? 4 /* DEADLINE_EXCEEDED */ : 13 /* INTERNAL */);
          ++served;
        }
      }
      const uint32_t rl = (uint32_t) resp.size();
      if (!write_all(c, &rl, 4) || !write_all(c, resp.data(), resp.size())) break;
    }
    ::close(c);
    std::lock_guard<std::mutex> g(mu_);
    for (auto it = conns_.begin(); it != conns_.end(); ++it)
      if (*it == c) {
        conns_.erase(it);
        break;
      }
  }

  Server* core_;
  std::string task_;
  int per_sample_, max_conn_;
  int64_t timeout_us_;
  int fd_ = -1, port_ = 0;
  std::atomic<bool> stop_{false};
  std::thread acceptor_;
  std::mutex mu_;
  std::vector<int> conns_;
  std::vector<std::thread> workers_;
};

}  // namespace trbs

#define TRB_API extern "C" __attribute__((visibility("default")))
using namespace trbs;

// Starts the front on 127.0.0.1:port (0 = any free port). Returns an opaque handle (nullptr on failure); *bound_port receives the port.
TRB_API void* trb_srv_listen(void* server, int port, const char* task_name, int outputs_per_sample, int max_connections, int64_t request_timeout_us, int* bound_port) {
  auto* f = new NetFront((Server*) server, task_name ? task_name : "default", outputs_per_sample > 0 ? outputs_per_sample : 1, max_connections > 0 ? max_connections : 64,
                         request_timeout_us > 0 ? request_timeout_us : 10'000'000);
  const int p = f->listen_on(port);
  if (p < 0) {
    delete f;
    return nullptr;
  }
  if (bound_port) *bound_port = p;
  return f;
}
TRB_API void trb_srv_listen_stop(void* front) { delete (NetFront*) front; }
TRB_API void trb_srv_listen_stats(void* front, int64_t* out3) {
  auto* f = (NetFront*) front;
  out3[0] = f->served;
  out3[1] = f->malformed;
  out3[2] = f->refused;
}
