// application_db_manager.h — name -> ApplicationDB registry (rocksdb_admin/application_db_manager.h:42-76):
// addDB / getDB / removeDB / getAllDBNames / Introspect, shared_mutex guarded; removeDB waits until it
// holds the only reference and hands the raw DB back (application_db_manager.cpp:76-100, 167-175).
#pragma once
#include <memory>
#include <shared_mutex>
#include <string>
#include <unordered_map>
#include <vector>

#include "rocksdb_admin/application_db.h"

namespace admin {
class ApplicationDBManager {
 public:
  explicit ApplicationDBManager(replicator::RocksDBReplicator* replicator = nullptr) : replicator_(replicator) {}
  ~ApplicationDBManager();
  bool addDB(const std::string& db_name, std::unique_ptr<rocksdb::DB> db, replicator::ReplicaRole role,
             std::unique_ptr<replicator::SocketAddress> upstream_addr, std::string* error_message);
  bool addDB(const std::string& db_name, std::unique_ptr<rocksdb::DB> db, replicator::ReplicaRole role,
             std::string* error_message) { return addDB(db_name, std::move(db), role, nullptr, error_message); }
  const std::shared_ptr<ApplicationDB> getDB(const std::string& db_name, std::string* error_message);
  std::unique_ptr<rocksdb::DB> removeDB(const std::string& db_name, std::string* error_message);
  std::vector<std::string> getAllDBNames();
  std::string Introspect() const;

 private:
  void waitOnApplicationDBRef(const std::shared_ptr<ApplicationDB>& db);
  replicator::RocksDBReplicator* replicator_;
  std::unordered_map<std::string, std::shared_ptr<ApplicationDB>> dbs_;
  mutable std::shared_mutex dbs_lock_;
};
}  // namespace admin
