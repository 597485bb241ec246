// message_ingestion.h — the second high-volume writer of the reference: AdminHandler::startMessageIngestion
// (rocksdb_admin/admin_handler.cpp:1855-2084) consumes a Kafka topic partition and writes every message straight into
// the shard's DB — Put / Delete / Merge by the payload's op code (rocksdb_admin.thrift:227-236), one DB call per
// message (:2019-2060) — recording the message timestamp every kafka_ts_update_interval messages (:2062-2074).
//
// librdkafka and the brokers are outside this image, so the consumer is an interface (MessageSource: what
// KafkaWatcher hands to its callback); the writer is the part on the hot path and is what is built here: messages are
// drained from the source in polls, every message becomes the same one-operation WriteBatch DB::Put / Delete / Merge
// would build, and a whole poll goes to the engine as ONE call (GpuDB::WriteMany -> rsp_apply_updates) that shares
// device ticks with the replication stream.  Per-message sequence numbers, statuses and counters are as in the
// reference (one sequence number per message, errors counted, the stream goes on).
#pragma once
#include <atomic>
#include <functional>
#include <memory>
#include <string>
#include <thread>
#include <vector>

#include "rocksdb_admin/application_db.h"

namespace admin {

enum class KafkaOperationCode : int32_t { PUT = 1, DELETE = 2, MERGE = 3 };  // rocksdb_admin.thrift:227-231

struct IngestMessage {
  std::string key;
  std::string value;  // the deserialised payload value (empty for DELETE)
  KafkaOperationCode op_code = KafkaOperationCode::PUT;
  int64_t timestamp_ms = 0;
  int32_t partition = 0;
  int64_t offset = 0;
};

// KafkaWatcher's role (admin_handler.cpp:1941-1957): hands over the messages from replay_timestamp_ms on
class MessageSource {
 public:
  virtual ~MessageSource() {}
  // up to max_messages into *out (appended); returns how many; 0 after timeout_ms without messages
  virtual size_t Poll(std::vector<IngestMessage>* out, size_t max_messages, int timeout_ms) = 0;
};

struct MessageIngestionOptions {
  size_t max_poll_messages = 4096;
  int poll_timeout_ms = 10;
  int64_t ts_update_interval = 1000;  // FLAGS_kafka_ts_update_interval
  // called with the newest message timestamp every ts_update_interval messages (the reference writes it to meta_db)
  std::function<void(int64_t timestamp_ms)> on_timestamp;
};

class MessageIngestor {
 public:
  MessageIngestor(std::shared_ptr<ApplicationDB> db, std::shared_ptr<MessageSource> source, MessageIngestionOptions opt = {});
  ~MessageIngestor();  // stopMessageIngestion (admin_handler.cpp:2086-2123)
  void Start();
  void Stop();
  // one poll on the calling thread; returns the number of messages written (tests / manual drive)
  size_t PollOnce();
  uint64_t messages() const { return messages_.load(); }
  uint64_t errors() const { return errors_.load(); }
  int64_t last_timestamp_ms() const { return last_ts_.load(); }

 private:
  std::shared_ptr<ApplicationDB> db_;
  std::shared_ptr<MessageSource> source_;
  MessageIngestionOptions opt_;
  std::atomic<bool> stop_{false};
  std::atomic<uint64_t> messages_{0}, errors_{0};
  std::atomic<int64_t> last_ts_{0};
  std::thread thread_;
};

}  // namespace admin
