#include "rocksdb_replicator/gpu_db_wrapper.h"

#include <cstdio>

#include "gpu_db.h"

namespace replicator {

GpuDbWrapper::GpuDbWrapper(const std::string& db_name, std::shared_ptr<rocksdb::DB> db)
    : db_name_(db_name), db_(std::move(db)), write_options_() {}

uint64_t GpuDbWrapper::LatestSequenceNumber() { return db_->GetLatestSequenceNumber(); }

rocksdb::Status GpuDbWrapper::WriteToLeader(const rocksdb::WriteOptions& options, rocksdb::WriteBatch* updates) {
  return db_->Write(options, updates);
}

rocksdb::Status GpuDbWrapper::GetUpdatesFromLeader(rocksdb::SequenceNumber seq_number,
                                                   std::unique_ptr<rocksdb::TransactionLogIterator>* iter) {
  return db_->GetUpdatesSince(seq_number, iter);
}

bool GpuDbWrapper::HandleReplicateResponse(Update* update) {
  rocksdb::Status status;
  if (auto* gdb = dynamic_cast<b200::GpuDB*>(db_.get())) {
    // the bytes go to the device as they came off the wire; PutLogData(&timestamp, 8) and the record
    // walk happen in the apply kernels (no std::string / WriteBatch copies on the host)
    status = gdb->ApplyReplicated(rocksdb::Slice(update->raw_data), (uint64_t)update->timestamp);
  } else {
    // any other rocksdb::DB: the reference's own sequence (rocksdb_wrapper.cpp:17-22)
    rocksdb::WriteBatch write_batch(update->raw_data);
    write_batch.PutLogData(rocksdb::Slice(reinterpret_cast<const char*>(&update->timestamp), sizeof(update->timestamp)));
    status = db_->Write(write_options_, &write_batch);
  }
  if (!status.ok())
    fprintf(stderr, "Failed to apply updates to FOLLOWER %s %s\n", db_name_.c_str(), status.ToString().c_str());
  return status.ok();
}

void GpuDbWrapper::HandleReplicateResponses(std::vector<Update>* updates, AppliedCallback done) {
  auto* gdb = dynamic_cast<b200::GpuDB*>(db_.get());
  if (!gdb) return DbWrapper::HandleReplicateResponses(updates, std::move(done));
  const std::string name = db_name_;
  gdb->ApplyReplicatedBatch(*updates, [name, done](size_t n_applied, const rocksdb::Status& status) {
    if (!status.ok())
      fprintf(stderr, "Failed to apply updates to FOLLOWER %s %s\n", name.c_str(), status.ToString().c_str());
    done(n_applied);
  });
}

}  // namespace replicator
