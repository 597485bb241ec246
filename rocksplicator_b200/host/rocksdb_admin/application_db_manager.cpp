#include "rocksdb_admin/application_db_manager.h"

#include <chrono>
#include <sstream>
#include <thread>

namespace admin {

const int kRemoveDBRefWaitMilliSec = 200;

ApplicationDBManager::~ApplicationDBManager() {
  // first drop each ApplicationDB, then release the DB it referred to (application_db_manager.cpp:154-164)
  for (auto& n : getAllDBNames()) removeDB(n, nullptr);
}

bool ApplicationDBManager::addDB(const std::string& db_name, std::unique_ptr<rocksdb::DB> db, replicator::ReplicaRole role,
                                 std::unique_ptr<replicator::SocketAddress> upstream_addr, std::string* error_message) {
  std::unique_lock<std::shared_mutex> lock(dbs_lock_);
  if (dbs_.find(db_name) != dbs_.end()) {
    if (error_message) *error_message = db_name + " has already been added";
    return false;
  }
  // ApplicationDB gets a NON-owning shared_ptr: the manager hands the raw DB back on removeDB
  // (application_db_manager.cpp:53-54)
  auto rocksdb_ptr = std::shared_ptr<rocksdb::DB>(db.get(), [](rocksdb::DB*) {});
  try {
    auto application_db_ptr =
        std::make_shared<ApplicationDB>(db_name, std::move(rocksdb_ptr), role, std::move(upstream_addr), replicator_);
    dbs_.emplace(db_name, std::move(application_db_ptr));
  } catch (const replicator::ReturnCode rc) {
    if (error_message) *error_message = "replicator refused " + db_name + ": " + std::to_string((int)rc);
    return false;
  }
  db.release();
  return true;
}

const std::shared_ptr<ApplicationDB> ApplicationDBManager::getDB(const std::string& db_name, std::string* error_message) {
  std::shared_lock<std::shared_mutex> lock(dbs_lock_);
  auto itor = dbs_.find(db_name);
  if (itor == dbs_.end()) {
    if (error_message) *error_message = db_name + " does not exist";
    return nullptr;
  }
  return itor->second;
}

std::unique_ptr<rocksdb::DB> ApplicationDBManager::removeDB(const std::string& db_name, std::string* error_message) {
  std::shared_ptr<ApplicationDB> ret;
  {
    std::unique_lock<std::shared_mutex> lock(dbs_lock_);
    auto itor = dbs_.find(db_name);
    if (itor == dbs_.end()) {
      if (error_message) *error_message = db_name + " does not exist";
      return nullptr;
    }
    ret = std::move(itor->second);
    dbs_.erase(itor);
  }
  waitOnApplicationDBRef(ret);
  rocksdb::DB* raw = ret->rocksdb();
  ret.reset();  // ~ApplicationDB: RocksDBReplicator::removeDB
  return std::unique_ptr<rocksdb::DB>(raw);
}

std::vector<std::string> ApplicationDBManager::getAllDBNames() {
  std::vector<std::string> db_names;
  std::shared_lock<std::shared_mutex> lock(dbs_lock_);
  db_names.reserve(dbs_.size());
  for (const auto& db : dbs_) db_names.push_back(db.first);
  return db_names;
}

std::string ApplicationDBManager::Introspect() const {
  // exact text asserted by rocksdb_admin/tests/application_db_manager_test.cpp:54-85
  std::stringstream ss;
  ss << "ApplicationDBManager:" << std::endl;
  std::shared_lock<std::shared_mutex> lock(dbs_lock_);
  for (const auto& db : dbs_) {
    ss << db.first << ":" << std::endl;
    ss << " " << db.second->Introspect() << std::endl;
  }
  return ss.str();
}

void ApplicationDBManager::waitOnApplicationDBRef(const std::shared_ptr<ApplicationDB>& db) {
  while (db.use_count() > 1) std::this_thread::sleep_for(std::chrono::milliseconds(kRemoveDBRefWaitMilliSec));
}

}  // namespace admin
