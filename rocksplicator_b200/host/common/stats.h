// common/stats.h — the two Stats calls the read/write facade makes (common/stats/stats.h:18-63 in the
// reference is a thread-local -> global counter library with an HTTP status server: out of scope).
// Counters are process-wide by name; enough for the facade and its tests.
#pragma once
#include <chrono>
#include <map>
#include <mutex>
#include <string>

namespace common {
class Stats {
 public:
  static Stats* get() { static Stats s; return &s; }
  void Incr(const std::string& name, uint64_t v = 1) { std::lock_guard<std::mutex> g(mu_); counters_[name] += v; }
  void AddMetric(const std::string& name, int64_t v) { std::lock_guard<std::mutex> g(mu_); metrics_sum_[name] += v; metrics_n_[name]++; }
  uint64_t GetCounter(const std::string& name) { std::lock_guard<std::mutex> g(mu_); return counters_[name]; }
 private:
  std::mutex mu_;
  std::map<std::string, uint64_t> counters_;
  std::map<std::string, int64_t> metrics_sum_;
  std::map<std::string, uint64_t> metrics_n_;
};
// common/timer.h:25-59: RAII elapsed-ms metric
class Timer {
 public:
  explicit Timer(const std::string& name) : name_(name), t0_(std::chrono::steady_clock::now()) {}
  ~Timer() {
    Stats::get()->AddMetric(name_, std::chrono::duration_cast<std::chrono::milliseconds>(std::chrono::steady_clock::now() - t0_).count());
  }
 private:
  std::string name_;
  std::chrono::steady_clock::time_point t0_;
};
}  // namespace common
