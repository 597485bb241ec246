// application_db_manager.cpp — the name -> ApplicationDB registry of one process.
//
// Behavioural target: rocksdb_admin/application_db_manager.cpp in the reference — duplicate names are refused
// (:44-49), the manager (not ApplicationDB) owns the rocksdb::DB and hands it back from removeDB once it holds the
// last reference (:76-100, 167-175), Introspect() text as rocksdb_admin/tests/application_db_manager_test.cpp:54-85
// expects it.
#include "rocksdb_admin/application_db_manager.h"

#include <chrono>
#include <thread>

namespace admin {

namespace {
constexpr std::chrono::milliseconds kRefPollInterval{200};
void NonOwningDelete(rocksdb::DB*) {}
void Explain(std::string* error_message, const std::string& text) {
  if (error_message != nullptr) *error_message = text;
}
}  // namespace

ApplicationDBManager::~ApplicationDBManager() {
  // each ApplicationDB goes first (it de-registers from the replicator), then the DB it pointed at
  for (const std::string& name : getAllDBNames()) removeDB(name, nullptr);
}

bool ApplicationDBManager::addDB(const std::string& db_name, std::unique_ptr<rocksdb::DB> db, replicator::ReplicaRole role,
                                 std::unique_ptr<replicator::SocketAddress> upstream_addr, std::string* error_message) {
  std::unique_lock<std::shared_mutex> guard(dbs_lock_);
  if (dbs_.count(db_name) != 0) {
    Explain(error_message, db_name + " has already been added");
    return false;
  }
  // ownership stays here: ApplicationDB only borrows the DB
  std::shared_ptr<rocksdb::DB> borrowed(db.get(), NonOwningDelete);
  std::shared_ptr<ApplicationDB> app_db;
  try {
    app_db = std::make_shared<ApplicationDB>(db_name, std::move(borrowed), role, std::move(upstream_addr), replicator_);
  } catch (const replicator::ReturnCode rc) {
    Explain(error_message, "replicator refused " + db_name + ": " + std::to_string(static_cast<int>(rc)));
    return false;
  }
  dbs_[db_name] = std::move(app_db);
  db.release();
  return true;
}

const std::shared_ptr<ApplicationDB> ApplicationDBManager::getDB(const std::string& db_name, std::string* error_message) {
  std::shared_lock<std::shared_mutex> guard(dbs_lock_);
  const auto found = dbs_.find(db_name);
  if (found != dbs_.end()) return found->second;
  Explain(error_message, db_name + " does not exist");
  return nullptr;
}

std::unique_ptr<rocksdb::DB> ApplicationDBManager::removeDB(const std::string& db_name, std::string* error_message) {
  std::shared_ptr<ApplicationDB> victim;
  {
    std::unique_lock<std::shared_mutex> guard(dbs_lock_);
    const auto found = dbs_.find(db_name);
    if (found == dbs_.end()) {
      Explain(error_message, db_name + " does not exist");
      return nullptr;
    }
    victim.swap(found->second);
    dbs_.erase(found);
  }
  // requests in flight still hold the ApplicationDB: wait them out before the DB changes hands
  waitOnApplicationDBRef(victim);
  std::unique_ptr<rocksdb::DB> owned(victim->rocksdb());
  victim.reset();
  return owned;
}

std::vector<std::string> ApplicationDBManager::getAllDBNames() {
  std::shared_lock<std::shared_mutex> guard(dbs_lock_);
  std::vector<std::string> names;
  names.reserve(dbs_.size());
  for (const auto& kv : dbs_) names.push_back(kv.first);
  return names;
}

std::string ApplicationDBManager::Introspect() const {
  std::shared_lock<std::shared_mutex> guard(dbs_lock_);
  std::string out = "ApplicationDBManager:\n";
  for (const auto& kv : dbs_) {
    out += kv.first;
    out += ":\n ";
    out += kv.second->Introspect();
    out += "\n";
  }
  return out;
}

void ApplicationDBManager::waitOnApplicationDBRef(const std::shared_ptr<ApplicationDB>& db) {
  while (db.use_count() > 1) std::this_thread::sleep_for(kRefPollInterval);
}

}  // namespace admin
