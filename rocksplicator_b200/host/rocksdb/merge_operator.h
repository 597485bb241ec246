// rocksdb/merge_operator.h — AssociativeMergeOperator as examples/counter_service/merge_operator.h and
// rocksdb_replicator/tests/rocksdb_assumption_test.cpp:58-77 subclass it.
#pragma once
#include <string>

#include "rocksdb/slice.h"

namespace rocksdb {
class Logger;
class MergeOperator {
 public:
  virtual ~MergeOperator() {}
  virtual const char* Name() const = 0;
  // associative form only (what the reference's operators implement)
  virtual bool Merge(const Slice& key, const Slice* existing_value, const Slice& value, std::string* new_value,
                     Logger* logger) const = 0;
};
class AssociativeMergeOperator : public MergeOperator {};
}  // namespace rocksdb
