// replicator_types.h — the data shapes of rocksdb_replicator/thrift/replicator.thrift:21-92 as plain
// structs (no thrift toolchain here: the transport is pluggable, see transport.h), plus a stand-in for
// folly::SocketAddress.
#pragma once
#include <cstdint>
#include <exception>
#include <string>
#include <vector>

namespace replicator {

enum class ReplicaRole { NOOP = 0, FOLLOWER = 1, LEADER = 2, OBSERVER = 3 };  // replicator.thrift ReplicaRole
inline const char* ReplicaRoleString(ReplicaRole r) {
  switch (r) {
    case ReplicaRole::FOLLOWER: return "FOLLOWER";
    case ReplicaRole::LEADER: return "LEADER";
    case ReplicaRole::OBSERVER: return "OBSERVER";
    default: return "NOOP";
  }
}
// legacy spellings used by older callers of the reference
constexpr ReplicaRole DBRole_MASTER = ReplicaRole::LEADER;
constexpr ReplicaRole DBRole_SLAVE = ReplicaRole::FOLLOWER;

enum class ErrorCode { OTHER = 0, SOURCE_NOT_FOUND = 1, SOURCE_READ_ERROR = 2, SOURCE_REMOVED = 3 };

struct ReplicateException : public std::exception {
  ErrorCode code = ErrorCode::OTHER;
  std::string msg;
  const char* what() const noexcept override { return msg.c_str(); }
};

struct Update {
  std::string raw_data;   // the WriteBatch bytes (IOBuf in the reference)
  int64_t timestamp = 0;  // ms; the leader's LogData stamp
  uint64_t seq_no = 0;
  bool has_seq_no = false;
  void set_seq_no(uint64_t s) { seq_no = s; has_seq_no = true; }
};

struct ReplicateRequest {
  int64_t seq_no = 0;
  std::string db_name;
  int32_t max_wait_ms = 0;
  int32_t max_updates = 0;
  ReplicaRole role = ReplicaRole::FOLLOWER;
  bool has_role = false;
  void set_role(ReplicaRole r) { role = r; has_role = true; }
};

struct ReplicateResponse {
  std::vector<Update> updates;
  ReplicaRole role = ReplicaRole::NOOP;
  bool has_role = false;
  void set_role(ReplicaRole r) { role = r; has_role = true; }
};

// host:port of an upstream replicator (folly::SocketAddress in the reference)
struct SocketAddress {
  std::string host;
  uint16_t port = 0;
  SocketAddress() {}
  SocketAddress(const std::string& h, uint16_t p) : host(h), port(p) {}
  bool empty() const { return host.empty() && port == 0; }
  const std::string& getAddressStr() const { return host; }
  uint16_t getPort() const { return port; }
  void setFromIpPort(const std::string& h, uint16_t p) { host = h; port = p; }
  // common::getNetworkAddressStr (common/network_util.cpp:54-66): the address only
  std::string describe() const { return empty() ? "uninitialized_addr" : host; }
  bool operator==(const SocketAddress& o) const { return host == o.host && port == o.port; }
};

}  // namespace replicator
