// max_number_box.h — keeps the largest number posted; wait(n, ms) blocks until max >= n or timeout
// (ms == 0: no timeout).  Contract of rocksdb_replicator/max_number_box.h:36-70; the leader's 2-ACK mode
// waits here for a follower's acknowledged sequence number (replicated_db.cpp:236-273, 452-456).
#pragma once
#include <chrono>
#include <condition_variable>
#include <cstdint>
#include <mutex>

namespace replicator {
namespace detail {

class MaxNumberBox {
 public:
  explicit MaxNumberBox(uint64_t init = 0) : max_(init) {}
  MaxNumberBox(const MaxNumberBox&) = delete;
  MaxNumberBox& operator=(const MaxNumberBox&) = delete;
  void post(uint64_t n) {
    {
      std::lock_guard<std::mutex> g(mu_);
      if (n <= max_) return;
      max_ = n;
    }
    cv_.notify_all();
  }
  bool wait(uint64_t n, uint64_t timeout_ms) {
    std::unique_lock<std::mutex> l(mu_);
    if (timeout_ms == 0) {
      cv_.wait(l, [&] { return max_ >= n; });
      return true;
    }
    return cv_.wait_for(l, std::chrono::milliseconds(timeout_ms), [&] { return max_ >= n; });
  }
  uint64_t value() {
    std::lock_guard<std::mutex> g(mu_);
    return max_;
  }

 private:
  std::mutex mu_;
  std::condition_variable cv_;
  uint64_t max_;
};

}  // namespace detail
}  // namespace replicator
