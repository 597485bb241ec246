// non_blocking_condition_variable.h — park a task until a predicate holds, a notifyAll() arrives or a
// timeout expires, then run it ONCE on the executor: no thread is blocked per waiting task.
// Same contract as rocksdb_replicator/non_blocking_condition_variable.h:85-139 (asserted by its test:
// fires exactly once via predicate / notifyAll / timeout / destructor); the leader parks follower
// long-polls here (replicated_db.cpp:466-574).
#pragma once
#include <atomic>
#include <functional>
#include <memory>
#include <mutex>
#include <vector>

#include "rocksdb_replicator/executor.h"

namespace replicator {
namespace detail {

class NonBlockingConditionVariable {
 public:
  explicit NonBlockingConditionVariable(Executor* executor) : executor_(executor) {}
  NonBlockingConditionVariable(const NonBlockingConditionVariable&) = delete;
  NonBlockingConditionVariable& operator=(const NonBlockingConditionVariable&) = delete;
  ~NonBlockingConditionVariable() { notifyAll(); }

  template <typename Func, typename Predicate>
  void runIfConditionOrWaitForNotify(Func f, Predicate p, uint64_t timeout_ms) {
    if (p()) {
      executor_->add(std::move(f));
      return;
    }
    auto parked = std::make_shared<Parked>(std::move(f));
    {
      std::lock_guard<std::mutex> g(mu_);
      parked_.push_back(parked);
    }
    // the condition may have become true between the first check and parking
    if (p()) {
      Fire(parked);
      return;
    }
    if (timeout_ms > 0) {
      std::weak_ptr<Parked> weak = parked;
      Executor* ex = executor_;
      executor_->addDelayed([weak, ex] {
        if (auto t = weak.lock()) {
          if (!t->fired.exchange(true)) ex->add(std::move(t->fn));
        }
      }, timeout_ms);
    }
  }

  void notifyAll() {
    std::vector<std::shared_ptr<Parked>> local;
    {
      std::lock_guard<std::mutex> g(mu_);
      local.swap(parked_);
    }
    for (auto& t : local) Fire(t);
  }

 private:
  struct Parked {
    explicit Parked(std::function<void()> f) : fn(std::move(f)) {}
    std::function<void()> fn;
    std::atomic<bool> fired{false};
  };
  void Fire(const std::shared_ptr<Parked>& t) {
    if (!t->fired.exchange(true)) executor_->add(std::move(t->fn));
  }
  std::mutex mu_;
  std::vector<std::shared_ptr<Parked>> parked_;
  Executor* const executor_;
};

}  // namespace detail
}  // namespace replicator
