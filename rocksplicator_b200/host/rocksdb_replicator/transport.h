// transport.h — how a follower's pull reaches its upstream.  The reference uses fbthrift over TCP
// (ReplicatorAsyncClient::future_replicate, rocksdb_replicator/replicated_db.cpp:328; server side
// replicator_handler.cpp:24-41).  There is no thrift toolchain in this image, so the transport is an
// interface; LocalTransport delivers the request to a RocksDBReplicator registered under the upstream's
// port in the same process — the arrangement rocksdb_replicator/tests/rocksdb_replicator_test.cpp:137-144
// builds with several replicators on 127.0.0.1 ports.
#pragma once
#include <functional>
#include <memory>

#include "rocksdb_replicator/replicator_types.h"

namespace replicator {

// result of one replicate() call: a response or an exception (folly::Try<ReplicateResponse>)
struct ReplicateResult {
  bool ok = false;
  ReplicateResponse response;
  bool is_replicate_exception = false;
  ReplicateException ex;
  std::string transport_error;  // std::exception path (connection errors)
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
