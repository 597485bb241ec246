// stager.h — multi-buffered staging combiner (std only, no CUDA): many caller threads, one device batch at a time.
//
// Why: the reference calls ApplicationDB::Get / MultiGet from up to 256 thrift worker threads
// (rocksdb_admin/application_db.cpp:85-120, examples/counter_service/counter.cpp:39,81) and applies replicated
// updates from >= 16 executor threads (rocksdb_replicator/rocksdb_replicator.cpp:58-67).  A device batch costs the same
// launch + synchronisation whether it carries 1 request or 100 000, so concurrent callers share batches:
//
//   caller:      begin()  -> a slice of the OPEN batch's pinned staging (items + bytes)     [short lock]
//                copy its own inputs into the slice, in parallel with every other caller     [no lock]
//                commit() -> wait() until the batch has run -> read its own results          [no lock]
//                release()
//   dispatcher:  one thread; as soon as it is idle and the open batch is not empty it CLOSES it (the other buffer
//                opens for new arrivals), waits for the callers still copying, runs the batch (H2D, kernels, D2H,
//                one synchronisation — the RunFn), fires the asynchronous completions and wakes the waiters.
//
// While the device works on batch k, callers fill batch k+1 and the callers of batch k-1 are still copying their results
// out (three buffers): batches grow with load on their own (no timer), a lone caller pays two thread hand-offs and
// nothing else.  Requests of one caller thread stay ordered (it does not return
// before its request has run); slices of one batch are ordered by begin() order, which is what per-shard FIFO needs.
#pragma once
#include <atomic>
#include <condition_variable>
#include <cstddef>
#include <cstdint>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

namespace rsp {

class Stager {
 public:
  struct Ticket {
    int buf = -1;
    size_t item0 = 0, byte0 = 0;
    uint64_t epoch = 0;
  };
  struct BatchInfo {
    int buf;
    size_t n_items, n_bytes;
    uint32_t klass;
    uint64_t epoch;
  };
  using RunFn = std::function<void(const BatchInfo&)>;
  using PostFn = std::function<void()>;  // dispatcher thread, after the asynchronous completions of a batch ran

  static constexpr int kBuffers = 3;  // filling | running | results being read

  Stager(size_t cap_items, size_t cap_bytes, RunFn run, PostFn post = nullptr)
      : cap_items_(cap_items), cap_bytes_(cap_bytes), run_(std::move(run)), post_(std::move(post)) {
    b_[0].state = OPEN;
    b_[0].epoch = ++epochs_;
    open_ = 0;
    thread_ = std::thread([this] { Loop(); });
  }
  ~Stager() { Stop(); }
  void Stop() {
    {
      std::lock_guard<std::mutex> g(mu_);
      if (stop_) return;
      stop_ = true;
    }
    cv_disp_.notify_all();
    cv_space_.notify_all();
    cv_done_.notify_all();
    if (thread_.joinable()) thread_.join();
  }

  size_t cap_items() const { return cap_items_; }
  size_t cap_bytes() const { return cap_bytes_; }

  // Reserve n_items / n_bytes in a batch of class `klass` (requests of different classes never share a batch);
  // max_items bounds the items of a batch of this class (<= cap_items).  Blocks while nothing can take the request.
  // false: the stager is stopping, or the request can never fit (the caller takes its direct path).
  bool begin(size_t n_items, size_t n_bytes, uint32_t klass, size_t max_items, Ticket* t) {
    if (max_items > cap_items_) max_items = cap_items_;
    if (n_items > max_items || n_bytes > cap_bytes_) return false;
    std::unique_lock<std::mutex> l(mu_);
    for (;;) {
      if (stop_) return false;
      if (open_ >= 0) {
        Batch& b = b_[open_];
        if (b.n_items == 0) b.klass = klass;
        if (b.klass == klass && b.n_items + n_items <= max_items && b.n_bytes + n_bytes <= cap_bytes_) {
          t->buf = open_;
          t->item0 = b.n_items;
          t->byte0 = b.n_bytes;
          t->epoch = b.epoch;
          b.n_items += n_items;
          b.n_bytes += n_bytes;
          b.copiers.fetch_add(1, std::memory_order_relaxed);
          b.users.fetch_add(1, std::memory_order_relaxed);
          if (b.n_items == n_items && disp_sleeping_) cv_disp_.notify_one();  // first request: the dispatcher may take the batch
          return true;
        }
        // full or of another class: the dispatcher closes it as soon as it can; wait for the next open batch
        if (disp_sleeping_) cv_disp_.notify_one();
      }
      cv_space_.wait(l);
    }
  }
  // the caller finished writing its slice.  Lock-free unless it is the last writer of a batch the dispatcher waits on.
  void commit(const Ticket& t) {
    Batch& b = b_[t.buf];
    if (b.copiers.fetch_sub(1, std::memory_order_acq_rel) == 1) {
      std::lock_guard<std::mutex> g(mu_);  // (the dispatcher checks the counter under mu_ before it sleeps)
      if (disp_sleeping_) cv_disp_.notify_one();
    }
  }
  // completion without a waiting thread: fn runs on the dispatcher thread once the batch has run; the slice is
  // released when fn returns (fn reads its results from the staging buffers itself)
  void commit_async(const Ticket& t, std::function<void()> fn) {
    Batch& b = b_[t.buf];
    std::lock_guard<std::mutex> g(mu_);
    b.async.push_back(std::move(fn));
    if (b.copiers.fetch_sub(1, std::memory_order_acq_rel) == 1 && disp_sleeping_) cv_disp_.notify_one();
  }
  // until the batch has run: a short spin on the batch's completion word (hundreds of callers would otherwise convoy
  // on one mutex just to learn that their batch is done), then a condition variable
  void wait(const Ticket& t) {
    Batch& b = b_[t.buf];
    for (int i = 0; i < 4000; i++) {
      if (b.epoch_done.load(std::memory_order_acquire) >= t.epoch) return;
      cpu_relax();
    }
    std::unique_lock<std::mutex> l(mu_);
    sleepers_++;
    while (b.epoch_done.load(std::memory_order_acquire) < t.epoch) cv_done_.wait(l);
    sleepers_--;
  }
  void release(const Ticket& t) {
    Batch& b = b_[t.buf];
    if (b.users.fetch_sub(1, std::memory_order_acq_rel) == 1) {
      std::lock_guard<std::mutex> g(mu_);
      MaybeFree(b);
    }
  }

  uint64_t batches() const { return batches_.load(std::memory_order_relaxed); }

 private:
  enum State { FREE, OPEN, CLOSED, DONE };
  struct Batch {
    State state = FREE;                 // mu_
    size_t n_items = 0, n_bytes = 0;    // mu_
    uint32_t klass = 0;
    uint64_t epoch = 0;
    std::atomic<uint32_t> copiers{0};   // callers still writing their slice
    std::atomic<uint32_t> users{0};     // callers (sync and async) that have not released their slice yet
    std::atomic<uint64_t> epoch_done{0};
    std::vector<std::function<void()>> async;  // mu_
  };
  static inline void cpu_relax() {
#if defined(__x86_64__) || defined(__i386__)
    __builtin_ia32_pause();
#else
    std::this_thread::yield();
#endif
  }

  void MaybeFree(Batch& b) {  // mu_ held: whoever sees "done and unused" first recycles the buffer
    if (b.state == DONE && b.users.load(std::memory_order_acquire) == 0) {
      b.state = FREE;
      b.n_items = b.n_bytes = 0;
      if (open_ < 0) OpenOne();
    }
  }
  void OpenOne() {  // mu_ held
    for (int i = 0; i < kBuffers; i++) {
      if (b_[i].state == FREE) {
        b_[i].state = OPEN;
        b_[i].epoch = ++epochs_;
        open_ = i;
        cv_space_.notify_all();
        return;
      }
    }
  }
  void Loop() {
    std::unique_lock<std::mutex> l(mu_);
    for (;;) {
      while (!(open_ >= 0 && b_[open_].n_items > 0)) {  // a stop request still lets queued work run: callers wait on it
        if (stop_) return;
        disp_sleeping_ = true;
        cv_disp_.wait(l);
        disp_sleeping_ = false;
      }
      const int bi = open_;
      Batch& b = b_[bi];
      b.state = CLOSED;
      open_ = -1;
      OpenOne();
      while (b.copiers.load(std::memory_order_acquire)) {
        disp_sleeping_ = true;
        cv_disp_.wait(l);
        disp_sleeping_ = false;
      }
      BatchInfo info{bi, b.n_items, b.n_bytes, b.klass, b.epoch};
      std::vector<std::function<void()>> async;
      async.swap(b.async);
      l.unlock();
      run_(info);
      for (auto& f : async) f();
      if (post_) post_();
      b.epoch_done.store(info.epoch, std::memory_order_release);  // spinning waiters go on at once
      l.lock();
      batches_++;
      b.state = DONE;
      if (sleepers_) cv_done_.notify_all();
      if (!async.empty()) b.users.fetch_sub((uint32_t)async.size(), std::memory_order_acq_rel);
      MaybeFree(b);
    }
  }

  const size_t cap_items_, cap_bytes_;
  RunFn run_;
  PostFn post_;
  std::mutex mu_;
  std::condition_variable cv_disp_, cv_space_, cv_done_;
  Batch b_[kBuffers];
  int open_ = -1;
  uint64_t epochs_ = 0;
  bool stop_ = false, disp_sleeping_ = false;
  uint32_t sleepers_ = 0;
  std::atomic<uint64_t> batches_{0};
  std::thread thread_;
};

}  // namespace rsp
