// application_db.h — admin::ApplicationDB with the reference's surface (rocksdb_admin/application_db.h:46-132):
// a facade over a rocksdb::DB that registers with the replicator in its constructor, forwards reads, and
// routes writes through ReplicatedDB::Write.  folly::SocketAddress -> replicator::SocketAddress.
#pragma once
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "rocksdb/db.h"
#include "rocksdb_replicator/rocksdb_replicator.h"

namespace admin {

extern bool FLAGS_disable_rocksplicator_db_stats;  // application_db.cpp:24-25

class ApplicationDB {
 public:
  struct Properties {
    static const std::string kNumLevels;          // "applicationdb.num-levels"
    static const std::string kHighestEmptyLevel;  // "applicationdb.highest-empty-level"
  };
  // throws replicator::ReturnCode when the replicator refuses the DB (application_db.cpp:62-68)
  ApplicationDB(const std::string& db_name, std::shared_ptr<rocksdb::DB> db, replicator::ReplicaRole role,
                std::unique_ptr<replicator::SocketAddress> upstream_addr,
                replicator::RocksDBReplicator* replicator = nullptr /* default: RocksDBReplicator::instance() */);
  ~ApplicationDB();

  rocksdb::Iterator* NewIterator(const rocksdb::ReadOptions& options);
  rocksdb::Status Get(const rocksdb::ReadOptions& options, const rocksdb::Slice& key, std::string* value);
  rocksdb::Status Get(const rocksdb::ReadOptions& options, const rocksdb::Slice& key, rocksdb::PinnableSlice* value);
  std::vector<rocksdb::Status> MultiGet(const rocksdb::ReadOptions& options, const std::vector<rocksdb::Slice>& keys,
                                        std::vector<std::string>* values);
  rocksdb::Status Write(const rocksdb::WriteOptions& options, rocksdb::WriteBatch* write_batch);
  rocksdb::Status CompactRange(const rocksdb::CompactRangeOptions& options, const rocksdb::Slice* begin,
                               const rocksdb::Slice* end);
  bool GetProperty(const rocksdb::Slice& property, std::string* value);
  bool DBLmaxEmpty();
  bool IsSlave() const { return role_ == replicator::ReplicaRole::FOLLOWER; }
  const std::string& db_name() const { return db_name_; }
  rocksdb::DB* rocksdb() const { return db_.get(); }
  replicator::SocketAddress* upstream_addr() const { return upstream_addr_.get(); }
  std::string Introspect();

 private:
  uint32_t getHighestEmptyLevel();
  const std::string db_name_;
  std::shared_ptr<rocksdb::DB> db_;
  const replicator::ReplicaRole role_;
  std::unique_ptr<replicator::SocketAddress> upstream_addr_;
  replicator::RocksDBReplicator* replicator_;
  replicator::RocksDBReplicator::ReplicatedDB* replicated_db_;
  friend class ApplicationDBManager;
};

}  // namespace admin
