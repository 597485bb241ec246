// rocksdb/transaction_log.h — TransactionLogIterator / BatchResult as the leader side consumes them
// (rocksdb_replicator/replicated_db.cpp:486-540, rocksdb_assumption_test.cpp:329-359).
#pragma once
#include <memory>

#include "rocksdb/status.h"
#include "rocksdb/write_batch.h"

namespace rocksdb {
struct BatchResult {
  SequenceNumber sequence = 0;
  std::unique_ptr<WriteBatch> writeBatchPtr;
};
class TransactionLogIterator {
 public:
  virtual ~TransactionLogIterator() {}
  virtual bool Valid() = 0;
  virtual void Next() = 0;
  virtual Status status() = 0;
  virtual BatchResult GetBatch() = 0;
};
}  // namespace rocksdb
