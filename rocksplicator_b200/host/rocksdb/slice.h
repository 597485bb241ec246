// rocksdb/slice.h — the subset of rocksdb::Slice / PinnableSlice the reference's hot-path callers use
// (rocksdb_admin/application_db.cpp:85-120, examples/counter_service/counter_handler.cpp:88,152-158).
#pragma once
#include <cstring>
#include <string>

namespace rocksdb {

class Slice {
 public:
  Slice() : data_(""), size_(0) {}
  Slice(const char* d, size_t n) : data_(d), size_(n) {}
  Slice(const std::string& s) : data_(s.data()), size_(s.size()) {}  // NOLINT
  Slice(const char* s) : data_(s), size_(strlen(s)) {}               // NOLINT
  const char* data() const { return data_; }
  size_t size() const { return size_; }
  bool empty() const { return size_ == 0; }
  char operator[](size_t n) const { return data_[n]; }
  void clear() { data_ = ""; size_ = 0; }
  void remove_prefix(size_t n) { data_ += n; size_ -= n; }
  std::string ToString(bool hex = false) const {
    if (!hex) return std::string(data_, size_);
    static const char* d = "0123456789ABCDEF";
    std::string r;
    for (size_t i = 0; i < size_; i++) { r.push_back(d[(unsigned char)data_[i] >> 4]); r.push_back(d[data_[i] & 15]); }
    return r;
  }
  int compare(const Slice& b) const {
    const size_t m = size_ < b.size_ ? size_ : b.size_;
    int r = m ? memcmp(data_, b.data_, m) : 0;
    if (r == 0) r = size_ < b.size_ ? -1 : (size_ > b.size_ ? 1 : 0);
    return r;
  }
  bool starts_with(const Slice& x) const { return size_ >= x.size_ && memcmp(data_, x.data_, x.size_) == 0; }

 protected:
  const char* data_;
  size_t size_;
};
inline bool operator==(const Slice& a, const Slice& b) { return a.size() == b.size() && memcmp(a.data(), b.data(), a.size()) == 0; }
inline bool operator!=(const Slice& a, const Slice& b) { return !(a == b); }

class PinnableSlice : public Slice {
 public:
  PinnableSlice() = default;
  void PinSelf(const Slice& s) { buf_.assign(s.data(), s.size()); data_ = buf_.data(); size_ = buf_.size(); }
  void PinSelf() { data_ = buf_.data(); size_ = buf_.size(); }
  std::string* GetSelf() { return &buf_; }
  void Reset() { buf_.clear(); data_ = ""; size_ = 0; }

 private:
  std::string buf_;
};

}  // namespace rocksdb
