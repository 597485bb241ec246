// transport.h — how a follower's pull reaches its upstream.  The reference uses fbthrift over TCP
// (ReplicatorAsyncClient::future_replicate, rocksdb_replicator/replicated_db.cpp:328; server side
// replicator_handler.cpp:24-41).  There is no thrift toolchain in this image, so the transport is an
// interface; LocalTransport delivers the request to a RocksDBReplicator registered under the upstream's
// port in the same process — the arrangement rocksdb_replicator/tests/rocksdb_replicator_test.cpp:137-144
// builds with several replicators on 127.0.0.1 ports.
#pragma once
#include <atomic>
#include <chrono>
#include <functional>
#include <memory>

#include "rocksdb_replicator/replicator_types.h"

namespace replicator {

// Where a pull round trip spends its time (diagnostics; off unless PullTrace::enabled): every stage adds its duration.
//   0 transport: replicate() called -> the callback delivers the response (upstream + wire)
//   1 hop to our executor                    2 response handling until HandleReplicateResponses returns (staging)
//   3 staged -> the engine's completion      4 completion -> continuation on our executor
//   5 continuation -> next replicate()
struct PullTrace {
  static constexpr int kStages = 6;
  std::atomic<bool> enabled{false};
  std::atomic<uint64_t> ns[kStages];
  std::atomic<uint64_t> n[kStages];
  PullTrace() { Reset(); }
  void Reset() { for (int i = 0; i < kStages; i++) { ns[i] = 0; n[i] = 0; } }
  static int64_t Now() { return std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
  void Add(int stage, int64_t from, int64_t to) {
    if (from && to >= from) { ns[stage].fetch_add((uint64_t)(to - from), std::memory_order_relaxed); n[stage].fetch_add(1, std::memory_order_relaxed); }
  }
  static PullTrace& Get() { static PullTrace t; return t; }
};

// result of one replicate() call: a response or an exception (folly::Try<ReplicateResponse>)
struct ReplicateResult {
  bool ok = false;
  ReplicateResponse response;
  bool is_replicate_exception = false;
  ReplicateException ex;
  std::string transport_error;  // std::exception path (connection errors)
  int64_t t_mark = 0;           // PullTrace: when the previous stage ended
};
using ReplicateCallback = std::function<void(ReplicateResult&&)>;

class Transport {
 public:
  virtual ~Transport() {}
  // Asynchronous: cb runs later on some thread (never inline in the caller's stack frame).
  virtual void replicate(const SocketAddress& upstream, const ReplicateRequest& request, uint32_t timeout_ms,
                         ReplicateCallback cb) = 0;
};

}  // namespace replicator
