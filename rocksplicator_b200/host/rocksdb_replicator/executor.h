// executor.h — a fixed thread pool plus one timer thread (std only).  Stands in for the reference's
// folly CPUThreadPoolExecutor "rptor-worker-" (rocksdb_replicator/rocksdb_replicator.cpp:58-67) and for
// the EventBase::runAfterDelay / futures::sleep calls on the replication path (replicated_db.cpp:412-428,
// non_blocking_condition_variable.h:113-126).
#pragma once
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <deque>
#include <functional>
#include <map>
#include <mutex>
#include <thread>
#include <vector>

namespace replicator {

class Executor {
 public:
  explicit Executor(size_t n_threads) {
    for (size_t i = 0; i < n_threads; i++) workers_.emplace_back([this] { WorkLoop(); });
    timer_ = std::thread([this] { TimerLoop(); });
  }
  ~Executor() { Stop(); }
  void add(std::function<void()> f) {
    {
      std::lock_guard<std::mutex> g(mu_);
      if (stop_) return;
      q_.push_back(std::move(f));
    }
    cv_.notify_one();
  }
  // run f on the pool after delay_ms
  void addDelayed(std::function<void()> f, uint64_t delay_ms) {
    const auto when = std::chrono::steady_clock::now() + std::chrono::milliseconds(delay_ms);
    {
      std::lock_guard<std::mutex> g(tmu_);
      if (stop_) return;
      timed_.emplace(when, std::move(f));
    }
    tcv_.notify_one();
  }
  void Stop() {
    {
      std::lock_guard<std::mutex> g(mu_);
      std::lock_guard<std::mutex> g2(tmu_);
      if (stop_) return;
      stop_ = true;
    }
    cv_.notify_all();
    tcv_.notify_all();
    for (auto& t : workers_) t.join();
    timer_.join();
  }

 private:
  void WorkLoop() {
    for (;;) {
      std::function<void()> f;
      {
        std::unique_lock<std::mutex> l(mu_);
        cv_.wait(l, [this] { return stop_ || !q_.empty(); });
        if (stop_) return;
        f = std::move(q_.front());
        q_.pop_front();
      }
      f();
    }
  }
  void TimerLoop() {
    std::unique_lock<std::mutex> l(tmu_);
    for (;;) {
      if (stop_) return;
      if (timed_.empty()) {
        tcv_.wait(l);
        continue;
      }
      const auto when = timed_.begin()->first;
      if (std::chrono::steady_clock::now() < when) {
        tcv_.wait_until(l, when);
        continue;
      }
      auto f = std::move(timed_.begin()->second);
      timed_.erase(timed_.begin());
      l.unlock();
      add(std::move(f));
      l.lock();
    }
  }
  std::mutex mu_, tmu_;
  std::condition_variable cv_, tcv_;
  std::deque<std::function<void()>> q_;
  std::multimap<std::chrono::steady_clock::time_point, std::function<void()>> timed_;
  std::vector<std::thread> workers_;
  std::thread timer_;
  bool stop_ = false;
};

}  // namespace replicator
