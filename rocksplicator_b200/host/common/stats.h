// common/stats.h — the two Stats calls the read/write facade and the replication library make.  Like the reference's
// common/stats/stats.h:18-63 the hot side is THREAD-LOCAL: a call finds its counter cell in a per-thread table (no
// lock, no shared cache line — the facade is called from hundreds of threads, Get "nearly 10M times per second"),
// and a reader sums the cells of all threads.  (The reference's HTTP status server is out of scope.)
#pragma once
#include <atomic>
#include <chrono>
#include <cstdint>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <unordered_map>
#include <vector>

namespace common {
class Stats {
 public:
  static Stats* get() { static Stats* s = new Stats(); return s; }  // never destroyed: threads may outlive main's statics
  void Incr(const std::string& name, uint64_t v = 1) { cell(name)->count.fetch_add(v, std::memory_order_relaxed); }
  void AddMetric(const std::string& name, int64_t v) {
    Cell* c = cell(name);
    c->sum.fetch_add(v, std::memory_order_relaxed);
    c->n.fetch_add(1, std::memory_order_relaxed);
  }
  // names that are string literals / static storage: found by pointer, no std::string on the hot path
  void IncrStatic(const char* name, uint64_t v = 1) { cell_static(name)->count.fetch_add(v, std::memory_order_relaxed); }
  void AddMetricStatic(const char* name, int64_t v) {
    Cell* c = cell_static(name);
    c->sum.fetch_add(v, std::memory_order_relaxed);
    c->n.fetch_add(1, std::memory_order_relaxed);
  }
  uint64_t GetCounter(const std::string& name) {
    std::lock_guard<std::mutex> g(mu_);
    uint64_t t = 0;
    auto r = cells_.equal_range(name);
    for (auto it = r.first; it != r.second; ++it) t += it->second->count.load(std::memory_order_relaxed);
    return t;
  }
  // (sum, samples) of a metric
  std::pair<int64_t, uint64_t> GetMetric(const std::string& name) {
    std::lock_guard<std::mutex> g(mu_);
    int64_t s = 0;
    uint64_t n = 0;
    auto r = cells_.equal_range(name);
    for (auto it = r.first; it != r.second; ++it) { s += it->second->sum.load(std::memory_order_relaxed); n += it->second->n.load(std::memory_order_relaxed); }
    return {s, n};
  }

 private:
  // written by one thread (relaxed atomics: readers may run concurrently), summed by readers
  struct alignas(64) Cell {
    std::atomic<uint64_t> count{0};
    std::atomic<int64_t> sum{0};
    std::atomic<uint64_t> n{0};
  };
  struct Local {
    std::unordered_map<std::string, Cell*> by_name;
    std::unordered_map<const char*, Cell*> by_ptr;
  };
  static Local& local() { static thread_local Local l; return l; }
  Cell* make(const std::string& name) {
    std::lock_guard<std::mutex> g(mu_);
    owned_.emplace_back(new Cell());
    cells_.emplace(name, owned_.back().get());
    return owned_.back().get();
  }
  Cell* cell(const std::string& name) {
    Local& l = local();
    auto it = l.by_name.find(name);
    if (it != l.by_name.end()) return it->second;
    Cell* c = make(name);
    l.by_name.emplace(name, c);
    return c;
  }
  Cell* cell_static(const char* name) {
    Local& l = local();
    auto it = l.by_ptr.find(name);
    if (it != l.by_ptr.end()) return it->second;
    Cell* c = cell(name);  // the same cell whichever way this thread names it
    l.by_ptr.emplace(name, c);
    return c;
  }
  std::mutex mu_;
  std::multimap<std::string, Cell*> cells_;
  std::vector<std::unique_ptr<Cell>> owned_;
};
// common/timer.h:25-59: RAII elapsed-ms metric (the name must outlive the timer: a literal or a static)
class Timer {
 public:
  explicit Timer(const char* static_name) : name_(static_name), t0_(std::chrono::steady_clock::now()) {}
  ~Timer() {
    Stats::get()->AddMetricStatic(name_, std::chrono::duration_cast<std::chrono::milliseconds>(std::chrono::steady_clock::now() - t0_).count());
  }
 private:
  const char* name_;
  std::chrono::steady_clock::time_point t0_;
};
}  // namespace common
