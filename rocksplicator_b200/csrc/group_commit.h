// group_commit.h — turns many concurrent single-update calls into few device ticks (std only, no CUDA).
//
// The reference's follower applies one update per DbWrapper::HandleReplicateResponse call, from >= 16 executor
// threads at once (rocksdb_replicator/rocksdb_replicator.cpp:58-67, replicated_db.cpp:369-383).  A device tick
// costs the same three launches and one synchronisation whether it carries 1 update or 50 000, so concurrent
// callers are combined: the first caller to arrive becomes the leader, drains everything queued so far
// (its own request included), runs ONE batched call, hands every waiter its status and wakes them; callers that
// arrive while a tick is running queue up for the next leader.  Requests of one caller thread stay ordered
// (a caller does not return before its request is done), which is all the per-shard FIFO rule needs.
#pragma once
#include <condition_variable>
#include <cstdint>
#include <functional>
#include <mutex>
#include <vector>

namespace rsp {

template <class Req>
class GroupCommit {
 public:
  // run(batch): executes all requests of the batch and fills each request's result fields
  using RunFn = std::function<void(std::vector<Req*>&)>;
  explicit GroupCommit(RunFn run, size_t max_batch = 65536) : run_(std::move(run)), max_batch_(max_batch) {}

  // blocks until `r` has been executed (by this thread as leader, or by another leader)
  void submit(Req* r) {
    std::unique_lock<std::mutex> l(mu_);
    Slot slot{r, false};
    queue_.push_back(&slot);
    for (;;) {
      if (slot.done) return;
      if (!leader_active_) break;
      cv_.wait(l);
    }
    // become the leader; keep leading until my own request is done (it is in the first batch I take)
    leader_active_ = true;
    while (!slot.done) {
      std::vector<Slot*> mine;
      const size_t take = queue_.size() < max_batch_ ? queue_.size() : max_batch_;
      mine.assign(queue_.begin(), queue_.begin() + take);
      queue_.erase(queue_.begin(), queue_.begin() + take);
      l.unlock();
      std::vector<Req*> batch;
      batch.reserve(mine.size());
      for (Slot* s : mine) batch.push_back(s->req);
      run_(batch);
      l.lock();
      for (Slot* s : mine) s->done = true;
      batches_++;
      requests_ += mine.size();
      cv_.notify_all();
    }
    leader_active_ = false;
    cv_.notify_all();  // someone queued meanwhile may take over
  }

  uint64_t batches() const { return batches_; }
  uint64_t requests() const { return requests_; }

 private:
  struct Slot {
    Req* req;
    bool done;
  };
  RunFn run_;
  size_t max_batch_;
  std::mutex mu_;
  std::condition_variable cv_;
  std::vector<Slot*> queue_;
  bool leader_active_ = false;
  uint64_t batches_ = 0, requests_ = 0;
};

}  // namespace rsp
