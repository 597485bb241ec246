// application_db.cpp — the read/write facade of one shard.
//
// Behavioural target: rocksdb_admin/application_db.cpp:52-234 in the reference — registration with the
// replicator on construction (:52-70), de-registration on destruction (:72-76), reads forwarded to the DB with a
// counter + latency metric each (:78-120, metric names :29-39), writes routed through ReplicatedDB::Write when the
// shard is replicated (:122-136), the two applicationdb.* properties (:171-225).
#include "rocksdb_admin/application_db.h"

#include <algorithm>

#include "common/stats.h"

namespace admin {

bool FLAGS_disable_rocksplicator_db_stats = false;

const std::string ApplicationDB::Properties::kNumLevels = "applicationdb.num-levels";
const std::string ApplicationDB::Properties::kHighestEmptyLevel = "applicationdb.highest-empty-level";

namespace {

// every facade operation reports "<name>" (count) and "<name>_ms" (latency); Write also "<name>_bytes"
enum class Op { kNewIterator, kGet, kMultiGet, kWrite, kCompactRange };
const char* OpName(Op op) {
  switch (op) {
    case Op::kNewIterator: return "rocksdb_new_iterator";
    case Op::kGet: return "rocksdb_get";
    case Op::kMultiGet: return "rocksdb_multi_get";
    case Op::kWrite: return "rocksdb_write";
    case Op::kCompactRange: return "rocksdb_compact_range";
  }
  return "rocksdb_unknown";
}
const char* OpMsName(Op op) {
  switch (op) {
    case Op::kNewIterator: return "rocksdb_new_iterator_ms";
    case Op::kGet: return "rocksdb_get_ms";
    case Op::kMultiGet: return "rocksdb_multi_get_ms";
    case Op::kWrite: return "rocksdb_write_ms";
    case Op::kCompactRange: return "rocksdb_compact_range_ms";
  }
  return "rocksdb_unknown_ms";
}

// counts the call now, records its duration when the scope ends (thread-local cells found by the static names)
class Metered {
 public:
  explicit Metered(Op op) : timer_(OpMsName(op)) { common::Stats::get()->IncrStatic(OpName(op)); }

 private:
  common::Timer timer_;
};

bool HasPrefix(const rocksdb::Slice& s, const char* prefix) { return s.starts_with(rocksdb::Slice(prefix)); }

}  // namespace

ApplicationDB::ApplicationDB(const std::string& db_name, std::shared_ptr<rocksdb::DB> db, replicator::ReplicaRole role,
                             std::unique_ptr<replicator::SocketAddress> upstream_addr,
                             replicator::RocksDBReplicator* repl)
    : db_name_(db_name),
      db_(std::move(db)),
      role_(role),
      upstream_addr_(std::move(upstream_addr)),
      replicator_(repl != nullptr ? repl : replicator::RocksDBReplicator::instance()),
      replicated_db_(nullptr) {
  // a follower without an upstream is a plain local DB; everything else joins the replication library
  const bool standalone = IsSlave() && upstream_addr_ == nullptr;
  if (standalone) return;
  const replicator::SocketAddress upstream = upstream_addr_ ? *upstream_addr_ : replicator::SocketAddress();
  const replicator::ReturnCode rc = replicator_->addDB(db_name_, db_, role_, upstream, &replicated_db_);
  if (rc != replicator::ReturnCode::OK) throw rc;
}

ApplicationDB::~ApplicationDB() {
  if (replicated_db_ != nullptr) replicator_->removeDB(db_name_);
}

rocksdb::Iterator* ApplicationDB::NewIterator(const rocksdb::ReadOptions& options) {
  Metered m(Op::kNewIterator);
  return db_->NewIterator(options);
}

// Get is the hottest call of the service ("nearly 10M times per second" in the reference's own comment), so its
// accounting can be switched off.
rocksdb::Status ApplicationDB::Get(const rocksdb::ReadOptions& options, const rocksdb::Slice& key, std::string* value) {
  if (!FLAGS_disable_rocksplicator_db_stats) {
    Metered m(Op::kGet);
    return db_->Get(options, key, value);
  }
  return db_->Get(options, key, value);
}

rocksdb::Status ApplicationDB::Get(const rocksdb::ReadOptions& options, const rocksdb::Slice& key,
                                   rocksdb::PinnableSlice* value) {
  rocksdb::ColumnFamilyHandle* cf = db_->DefaultColumnFamily();
  if (!FLAGS_disable_rocksplicator_db_stats) {
    Metered m(Op::kGet);
    return db_->Get(options, cf, key, value);
  }
  return db_->Get(options, cf, key, value);
}

std::vector<rocksdb::Status> ApplicationDB::MultiGet(const rocksdb::ReadOptions& options,
                                                     const std::vector<rocksdb::Slice>& keys,
                                                     std::vector<std::string>* values) {
  Metered m(Op::kMultiGet);
  return db_->MultiGet(options, keys, values);
}

rocksdb::Status ApplicationDB::Write(const rocksdb::WriteOptions& options, rocksdb::WriteBatch* write_batch) {
  Metered m(Op::kWrite);
  common::Stats::get()->IncrStatic("rocksdb_write_bytes", write_batch->GetDataSize());
  // replicated shards write through the replication library (leader check, timestamp, ACK modes); a shard that
  // the manager keeps without replication writes its DB directly
  return replicated_db_ != nullptr ? replicated_db_->Write(options, write_batch) : db_->Write(options, write_batch);
}

rocksdb::Status ApplicationDB::CompactRange(const rocksdb::CompactRangeOptions& options, const rocksdb::Slice* begin,
                                            const rocksdb::Slice* end) {
  Metered m(Op::kCompactRange);
  return db_->CompactRange(options, begin, end);
}

bool ApplicationDB::GetProperty(const rocksdb::Slice& property, std::string* value) {
  if (HasPrefix(property, "rocksdb.")) return db_->GetProperty(property, value);
  if (!HasPrefix(property, "applicationdb.")) return false;
  const std::string name = property.ToString();
  if (name == Properties::kNumLevels) {
    *value = std::to_string(db_->NumberLevels());
    return true;
  }
  if (name == Properties::kHighestEmptyLevel) {
    *value = std::to_string(getHighestEmptyLevel());
    return true;
  }
  return false;
}

bool ApplicationDB::DBLmaxEmpty() {
  std::string levels, highest_empty;
  if (!GetProperty(Properties::kNumLevels, &levels) || !GetProperty(Properties::kHighestEmptyLevel, &highest_empty)) return false;
  return std::stoi(highest_empty) == std::stoi(levels) - 1;
}

uint32_t ApplicationDB::getHighestEmptyLevel() {
  rocksdb::ColumnFamilyMetaData meta;
  db_->GetColumnFamilyMetaData(&meta);
  uint32_t highest = 0;
  for (const rocksdb::LevelMetaData& level : meta.levels)
    if (level.size == 0) highest = std::max(highest, static_cast<uint32_t>(level.level));
  return highest;
}

std::string ApplicationDB::Introspect() {
  return replicated_db_ != nullptr ? replicated_db_->Introspect() : std::string("__no_replicated_db__");
}

}  // namespace admin
