// db_wrapper.h — THE plugin boundary of the replication library: rocksdb_replicator/db_wrapper.h:6-15.
// Same four virtuals, same argument meaning.  RocksDbWrapper (rocksdb_wrapper.cpp) is the reference's
// implementation over rocksdb::DB; GpuDbWrapper (gpu_db_wrapper.h) is ours over the B200 engine;
// cdc_admin/cdc_application_db.cpp:19-37 and test_db_proxy.cpp are further implementations in the reference.
#pragma once
#include <memory>

#include "rocksdb/db.h"
#include "rocksdb_replicator/replicator_types.h"

namespace replicator {
class DbWrapper {
 public:
  virtual ~DbWrapper() {}
  virtual rocksdb::Status WriteToLeader(const rocksdb::WriteOptions& options, rocksdb::WriteBatch* updates) = 0;
  virtual rocksdb::Status GetUpdatesFromLeader(rocksdb::SequenceNumber seq_number,
                                               std::unique_ptr<rocksdb::TransactionLogIterator>* iter) = 0;
  virtual uint64_t LatestSequenceNumber() = 0;
  virtual bool HandleReplicateResponse(Update* update) = 0;
};
}  // namespace replicator
