// rocksdb/status.h — rocksdb::Status with RocksDB's code numbering and ToString() prefixes, as the
// reference observes them (rocksdb_replicator_test.cpp:586 compares against Status::TimedOut(...);
// rocksdb_wrapper.cpp:22-26 logs status.ToString()).
#pragma once
#include <string>

#include "rocksdb/slice.h"

namespace rocksdb {

class Status {
 public:
  enum Code { kOk = 0, kNotFound = 1, kCorruption = 2, kNotSupported = 3, kInvalidArgument = 4, kIOError = 5,
              kMergeInProgress = 6, kIncomplete = 7, kShutdownInProgress = 8, kTimedOut = 9, kAborted = 10,
              kBusy = 11, kExpired = 12, kTryAgain = 13 };
  Status() : code_(kOk) {}
  static Status OK() { return Status(); }
  static Status NotFound(const Slice& m = Slice()) { return Status(kNotFound, m); }
  static Status Corruption(const Slice& m = Slice()) { return Status(kCorruption, m); }
  static Status NotSupported(const Slice& m = Slice()) { return Status(kNotSupported, m); }
  static Status InvalidArgument(const Slice& m = Slice()) { return Status(kInvalidArgument, m); }
  static Status IOError(const Slice& m = Slice()) { return Status(kIOError, m); }
  static Status Incomplete(const Slice& m = Slice()) { return Status(kIncomplete, m); }
  static Status TimedOut(const Slice& m = Slice()) { return Status(kTimedOut, m); }
  static Status Aborted(const Slice& m = Slice()) { return Status(kAborted, m); }
  static Status Busy(const Slice& m = Slice()) { return Status(kBusy, m); }
  // from an engine code + "Prefix: message" text (rsp_last_error)
  static Status FromCode(int code, const std::string& text) {
    Status s;
    s.code_ = static_cast<Code>(code);
    const auto p = text.find(": ");
    s.msg_ = p == std::string::npos ? text : text.substr(p + 2);
    return s;
  }
  bool ok() const { return code_ == kOk; }
  bool IsNotFound() const { return code_ == kNotFound; }
  bool IsCorruption() const { return code_ == kCorruption; }
  bool IsNotSupported() const { return code_ == kNotSupported; }
  bool IsInvalidArgument() const { return code_ == kInvalidArgument; }
  bool IsIOError() const { return code_ == kIOError; }
  bool IsTimedOut() const { return code_ == kTimedOut; }
  bool IsIncomplete() const { return code_ == kIncomplete; }
  bool IsBusy() const { return code_ == kBusy; }
  Code code() const { return code_; }
  bool operator==(const Status& o) const { return code_ == o.code_; }
  bool operator!=(const Status& o) const { return code_ != o.code_; }
  std::string ToString() const {
    static const char* kPrefix[] = {"OK", "NotFound: ", "Corruption: ", "Not implemented: ", "Invalid argument: ",
                                    "IO error: ", "Merge in progress: ", "Result incomplete: ", "Shutdown in progress: ",
                                    "Operation timed out: ", "Operation aborted: ", "Resource busy: ", "Operation expired: ",
                                    "Operation failed. Try again.: "};
    if (code_ == kOk) return "OK";
    return std::string(kPrefix[code_]) + msg_;
  }

 private:
  Status(Code c, const Slice& m) : code_(c), msg_(m.data(), m.size()) {}
  Code code_;
  std::string msg_;
};

}  // namespace rocksdb
