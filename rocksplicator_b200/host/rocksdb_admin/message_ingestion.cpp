// message_ingestion.cpp — see message_ingestion.h (rocksdb_admin/admin_handler.cpp:1855-2084).
#include "rocksdb_admin/message_ingestion.h"

#include "common/stats.h"
#include "gpu_db.h"

namespace admin {

MessageIngestor::MessageIngestor(std::shared_ptr<ApplicationDB> db, std::shared_ptr<MessageSource> source,
                                 MessageIngestionOptions opt)
    : db_(std::move(db)), source_(std::move(source)), opt_(std::move(opt)) {}

MessageIngestor::~MessageIngestor() { Stop(); }

void MessageIngestor::Start() {
  if (thread_.joinable()) return;
  stop_ = false;
  thread_ = std::thread([this] {
    while (!stop_.load()) PollOnce();
  });
}

void MessageIngestor::Stop() {
  stop_ = true;
  if (thread_.joinable()) thread_.join();
}

size_t MessageIngestor::PollOnce() {
  std::vector<IngestMessage> msgs;
  const size_t n = source_->Poll(&msgs, opt_.max_poll_messages, opt_.poll_timeout_ms);
  if (!n) return 0;
  // one single-operation WriteBatch per message: exactly what DB::Put / Delete / Merge builds (admin_handler.cpp:2028-2051)
  std::vector<rocksdb::WriteBatch> batches(n);
  std::vector<rocksdb::WriteBatch*> ptrs(n);
  std::vector<bool> valid(n, true);
  auto* stats = common::Stats::get();
  for (size_t i = 0; i < n; i++) {
    const IngestMessage& m = msgs[i];
    switch (m.op_code) {
      case KafkaOperationCode::PUT:
        stats->IncrStatic("kafka_db_put_message");
        batches[i].Put(m.key, m.value);
        break;
      case KafkaOperationCode::DELETE:
        stats->IncrStatic("kafka_db_del_message");
        batches[i].Delete(m.key);
        break;
      case KafkaOperationCode::MERGE:
        stats->IncrStatic("kafka_db_merge_message");
        batches[i].Merge(m.key, m.value);
        break;
      default:
        stats->IncrStatic("kafka_invalid_opcode");  // logged and skipped in the reference
        valid[i] = false;
    }
    ptrs[i] = &batches[i];
  }
  std::vector<rocksdb::WriteBatch*> todo;
  todo.reserve(n);
  for (size_t i = 0; i < n; i++) if (valid[i]) todo.push_back(ptrs[i]);
  std::vector<rocksdb::Status> st;
  auto* gdb = dynamic_cast<b200::GpuDB*>(db_->rocksdb());
  if (gdb) {
    st = gdb->WriteMany(rocksdb::WriteOptions(), todo);  // the whole poll in one engine call
  } else {
    for (auto* b : todo) st.push_back(db_->rocksdb()->Write(rocksdb::WriteOptions(), b));
  }
  uint64_t bad = 0;
  for (auto& s : st) if (!s.ok()) bad++;
  if (bad) {
    errors_ += bad;
    stats->IncrStatic("kafka_db_write_errors", bad);  // the reference counts per operation kind and goes on
  }
  // the newest timestamp every ts_update_interval messages (admin_handler.cpp:2062-2074)
  const uint64_t before = messages_.fetch_add(n);
  last_ts_.store(msgs.back().timestamp_ms);
  if (opt_.on_timestamp && opt_.ts_update_interval > 0 &&
      (before + n) / (uint64_t)opt_.ts_update_interval != before / (uint64_t)opt_.ts_update_interval)
    opt_.on_timestamp(msgs.back().timestamp_ms);
  return n;
}

}  // namespace admin
