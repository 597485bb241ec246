// rocksdb/write_batch.h — rocksdb::WriteBatch over its real wire format (the byte contract of
// rocksdb_replicator/rocksdb_wrapper.cpp:17-20 and replicated_db.cpp:115-117, 527-538).
#pragma once
#include <cstdint>
#include <string>

#include "rocksdb/slice.h"
#include "rocksdb/status.h"

namespace rocksdb {

typedef uint64_t SequenceNumber;

class WriteBatch {
 public:
  class Handler {
   public:
    virtual ~Handler() {}
    virtual Status PutCF(uint32_t cf, const Slice& key, const Slice& value) {
      if (cf == 0) { Put(key, value); return Status::OK(); }
      return Status::InvalidArgument("non-default column family and PutCF not implemented");
    }
    virtual void Put(const Slice& /*key*/, const Slice& /*value*/) {}
    virtual Status DeleteCF(uint32_t cf, const Slice& key) {
      if (cf == 0) { Delete(key); return Status::OK(); }
      return Status::InvalidArgument("non-default column family and DeleteCF not implemented");
    }
    virtual void Delete(const Slice& /*key*/) {}
    virtual Status SingleDeleteCF(uint32_t cf, const Slice& key) {
      if (cf == 0) { SingleDelete(key); return Status::OK(); }
      return Status::InvalidArgument("non-default column family and SingleDeleteCF not implemented");
    }
    virtual void SingleDelete(const Slice& /*key*/) {}
    virtual Status MergeCF(uint32_t cf, const Slice& key, const Slice& value) {
      if (cf == 0) { Merge(key, value); return Status::OK(); }
      return Status::InvalidArgument("non-default column family and MergeCF not implemented");
    }
    virtual void Merge(const Slice& /*key*/, const Slice& /*value*/) {}
    virtual void LogData(const Slice& /*blob*/) {}
    virtual bool Continue() { return true; }
  };

  WriteBatch() : rep_(kHeader, '\0') {}
  explicit WriteBatch(const std::string& rep) : rep_(rep) {}
  explicit WriteBatch(std::string&& rep) : rep_(std::move(rep)) {}

  Status Put(const Slice& key, const Slice& value) { Bump(); rep_.push_back(0x1); PutLP(key); PutLP(value); return Status::OK(); }
  Status Delete(const Slice& key) { Bump(); rep_.push_back(0x0); PutLP(key); return Status::OK(); }
  Status SingleDelete(const Slice& key) { Bump(); rep_.push_back(0x7); PutLP(key); return Status::OK(); }
  Status Merge(const Slice& key, const Slice& value) { Bump(); rep_.push_back(0x2); PutLP(key); PutLP(value); return Status::OK(); }
  Status PutLogData(const Slice& blob) { rep_.push_back(0x3); PutLP(blob); return Status::OK(); }
  void Clear() { rep_.assign(kHeader, '\0'); }
  int Count() const { uint32_t c; memcpy(&c, rep_.data() + 8, 4); return (int)c; }
  const std::string& Data() const { return rep_; }
  size_t GetDataSize() const { return rep_.size(); }
  SequenceNumber Sequence() const { uint64_t s; memcpy(&s, rep_.data(), 8); return s; }
  void SetSequence(SequenceNumber s) { memcpy(&rep_[0], &s, 8); }

  // WriteBatch::Iterate: walk the records, RocksDB's error classes
  Status Iterate(Handler* h) const {
    if (rep_.size() < kHeader) return Status::Corruption("malformed WriteBatch (too small)");
    const uint8_t* p = (const uint8_t*)rep_.data() + kHeader;
    const uint8_t* lim = (const uint8_t*)rep_.data() + rep_.size();
    int found = 0;
    while (p < lim && h->Continue()) {
      const uint8_t tag = *p++;
      uint32_t cf = 0;
      Slice k, v;
      Status s;
      switch (tag) {
        case 0x5: if (!GetVarint32(&p, lim, &cf)) return Status::Corruption("bad WriteBatch Put");  // fallthrough
        case 0x1: if (!GetLP(&p, lim, &k) || !GetLP(&p, lim, &v)) return Status::Corruption("bad WriteBatch Put");
                  s = h->PutCF(cf, k, v); found++; break;
        case 0x4: if (!GetVarint32(&p, lim, &cf)) return Status::Corruption("bad WriteBatch Delete");  // fallthrough
        case 0x0: if (!GetLP(&p, lim, &k)) return Status::Corruption("bad WriteBatch Delete");
                  s = h->DeleteCF(cf, k); found++; break;
        case 0x8: if (!GetVarint32(&p, lim, &cf)) return Status::Corruption("bad WriteBatch Delete");  // fallthrough
        case 0x7: if (!GetLP(&p, lim, &k)) return Status::Corruption("bad WriteBatch Delete");
                  s = h->SingleDeleteCF(cf, k); found++; break;
        case 0x6: if (!GetVarint32(&p, lim, &cf)) return Status::Corruption("bad WriteBatch Merge");  // fallthrough
        case 0x2: if (!GetLP(&p, lim, &k) || !GetLP(&p, lim, &v)) return Status::Corruption("bad WriteBatch Merge");
                  s = h->MergeCF(cf, k, v); found++; break;
        case 0x3: if (!GetLP(&p, lim, &k)) return Status::Corruption("bad WriteBatch Blob");
                  h->LogData(k); break;
        case 0xD: break;
        default: return Status::Corruption("unknown WriteBatch tag");
      }
      if (!s.ok()) return s;
    }
    if (found != Count()) return Status::Corruption("WriteBatch has wrong count");
    return Status::OK();
  }

  static const size_t kHeader = 12;

 private:
  void Bump() { uint32_t c; memcpy(&c, rep_.data() + 8, 4); c++; memcpy(&rep_[8], &c, 4); }
  void PutLP(const Slice& s) {
    uint32_t n = (uint32_t)s.size();
    while (n >= 0x80) { rep_.push_back((char)((n & 0x7f) | 0x80)); n >>= 7; }
    rep_.push_back((char)n);
    rep_.append(s.data(), s.size());
  }
  static bool GetVarint32(const uint8_t** p, const uint8_t* lim, uint32_t* v) {
    uint32_t r = 0;
    for (uint32_t shift = 0; shift <= 28 && *p < lim; shift += 7) {
      const uint32_t b = *(*p)++;
      if (b & 128) r |= (b & 127) << shift;
      else { *v = r | (b << shift); return true; }
    }
    return false;
  }
  static bool GetLP(const uint8_t** p, const uint8_t* lim, Slice* s) {
    uint32_t n;
    if (!GetVarint32(p, lim, &n) || (size_t)(lim - *p) < n) return false;
    *s = Slice((const char*)*p, n);
    *p += n;
    return true;
  }
  std::string rep_;
};

}  // namespace rocksdb
