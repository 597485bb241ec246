// application_db.cpp — rocksdb_admin/application_db.cpp:52-234 restated over the shim headers.
#include "rocksdb_admin/application_db.h"

#include <set>

#include "common/stats.h"

namespace {
const std::string kRocksdbNewIterator = "rocksdb_new_iterator";
const std::string kRocksdbNewIteratorMs = "rocksdb_new_iterator_ms";
const std::string kRocksdbGet = "rocksdb_get";
const std::string kRocksdbGetMs = "rocksdb_get_ms";
const std::string kRocksdbMultiGet = "rocksdb_multi_get";
const std::string kRocksdbMultiGetMs = "rocksdb_multi_get_ms";
const std::string kRocksdbWrite = "rocksdb_write";
const std::string kRocksdbWriteBytes = "rocksdb_write_bytes";
const std::string kRocksdbWriteMs = "rocksdb_write_ms";
const std::string kRocksdbCompaction = "rocksdb_compact_range";
const std::string kRocksdbCompactionMs = "rocksdb_compact_range_ms";
}  // namespace

namespace admin {

bool FLAGS_disable_rocksplicator_db_stats = false;
static const std::string rocksdb_prefix = "rocksdb.";
static const std::string applicationdb_prefix = "applicationdb.";
const std::string ApplicationDB::Properties::kNumLevels = applicationdb_prefix + "num-levels";
const std::string ApplicationDB::Properties::kHighestEmptyLevel = applicationdb_prefix + "highest-empty-level";

ApplicationDB::ApplicationDB(const std::string& db_name, std::shared_ptr<rocksdb::DB> db, replicator::ReplicaRole role,
                             std::unique_ptr<replicator::SocketAddress> upstream_addr,
                             replicator::RocksDBReplicator* repl)
    : db_name_(db_name), db_(std::move(db)), role_(role), upstream_addr_(std::move(upstream_addr)),
      replicator_(repl ? repl : replicator::RocksDBReplicator::instance()), replicated_db_(nullptr) {
  if (!IsSlave() || upstream_addr_) {
    auto ret = replicator_->addDB(db_name_, db_, role_, upstream_addr_ ? *upstream_addr_ : replicator::SocketAddress(),
                                  &replicated_db_);
    if (ret != replicator::ReturnCode::OK) throw ret;
  }
}

ApplicationDB::~ApplicationDB() {
  if (replicated_db_) replicator_->removeDB(db_name_);
}

rocksdb::Iterator* ApplicationDB::NewIterator(const rocksdb::ReadOptions& options) {
  common::Stats::get()->Incr(kRocksdbNewIterator);
  common::Timer timer(kRocksdbNewIteratorMs);
  return db_->NewIterator(options);
}

rocksdb::Status ApplicationDB::Get(const rocksdb::ReadOptions& options, const rocksdb::Slice& slice, std::string* value) {
  // "We need to call Get() nearly 10M times per second" (application_db.cpp:89-91): stats are skippable
  if (FLAGS_disable_rocksplicator_db_stats) return db_->Get(options, slice, value);
  common::Stats::get()->Incr(kRocksdbGet);
  common::Timer timer(kRocksdbGetMs);
  return db_->Get(options, slice, value);
}

rocksdb::Status ApplicationDB::Get(const rocksdb::ReadOptions& options, const rocksdb::Slice& key,
                                   rocksdb::PinnableSlice* value) {
  if (FLAGS_disable_rocksplicator_db_stats) return db_->Get(options, db_->DefaultColumnFamily(), key, value);
  common::Stats::get()->Incr(kRocksdbGet);
  common::Timer timer(kRocksdbGetMs);
  return db_->Get(options, db_->DefaultColumnFamily(), key, value);
}

std::vector<rocksdb::Status> ApplicationDB::MultiGet(const rocksdb::ReadOptions& options,
                                                     const std::vector<rocksdb::Slice>& slice,
                                                     std::vector<std::string>* value) {
  common::Stats::get()->Incr(kRocksdbMultiGet);
  common::Timer timer(kRocksdbMultiGetMs);
  return db_->MultiGet(options, slice, value);
}

rocksdb::Status ApplicationDB::Write(const rocksdb::WriteOptions& options, rocksdb::WriteBatch* write_batch) {
  common::Stats::get()->Incr(kRocksdbWrite);
  common::Stats::get()->Incr(kRocksdbWriteBytes, write_batch->GetDataSize());
  common::Timer timer(kRocksdbWriteMs);
  if (replicated_db_) return replicated_db_->Write(options, write_batch);
  // un-replicated instance (FOLLOWER without upstream): write the local db (application_db.cpp:129-135)
  return db_->Write(options, write_batch);
}

rocksdb::Status ApplicationDB::CompactRange(const rocksdb::CompactRangeOptions& options, const rocksdb::Slice* begin,
                                            const rocksdb::Slice* end) {
  common::Stats::get()->Incr(kRocksdbCompaction);
  common::Timer timer(kRocksdbCompactionMs);
  return db_->CompactRange(options, begin, end);
}

bool ApplicationDB::GetProperty(const rocksdb::Slice& property, std::string* value) {
  if (property.starts_with(applicationdb_prefix)) {
    if (property == rocksdb::Slice(Properties::kHighestEmptyLevel)) { *value = std::to_string(getHighestEmptyLevel()); return true; }
    if (property == rocksdb::Slice(Properties::kNumLevels)) { *value = std::to_string(db_->NumberLevels()); return true; }
  } else if (property.starts_with(rocksdb_prefix)) {
    return db_->GetProperty(property, value);
  }
  return false;
}

bool ApplicationDB::DBLmaxEmpty() {
  std::string num_levels, highest_empty_level;
  return GetProperty(Properties::kNumLevels, &num_levels) && GetProperty(Properties::kHighestEmptyLevel, &highest_empty_level) &&
         std::stoi(num_levels) - 1 == std::stoi(highest_empty_level);
}

uint32_t ApplicationDB::getHighestEmptyLevel() {
  rocksdb::ColumnFamilyMetaData cf_metadata;
  db_->GetColumnFamilyMetaData(&cf_metadata);
  std::set<uint32_t> empty_levels;
  for (const auto& level_meta : cf_metadata.levels)
    if (level_meta.size == 0) empty_levels.insert((uint32_t)level_meta.level);
  return empty_levels.empty() ? 0 : *empty_levels.rbegin();
}

std::string ApplicationDB::Introspect() {
  if (replicated_db_) return replicated_db_->Introspect();
  return "__no_replicated_db__";
}

}  // namespace admin
