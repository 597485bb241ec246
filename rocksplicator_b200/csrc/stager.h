// stager.h — multi-buffered staging combiner (std only, no CUDA): many caller threads, one device batch at a time.
//
// Why: the reference calls ApplicationDB::Get / MultiGet from up to 256 thrift worker threads
// (rocksdb_admin/application_db.cpp:85-120, examples/counter_service/counter.cpp:39,81) and applies replicated
// updates from >= 16 executor threads (rocksdb_replicator/rocksdb_replicator.cpp:58-67).  A device batch costs the same
// launch + synchronisation whether it carries 1 request or 100 000, so concurrent callers share batches:
//
//   caller:      begin()  -> a slice of the OPEN batch's pinned staging (items + bytes)
//                copy its own inputs into the slice, in parallel with every other caller
//                commit() -> wait() until the batch has run -> read its own results -> release()
//   dispatcher:  one thread; as soon as it is idle and the open batch is not empty it CLOSES it (another buffer
//                opens for new arrivals), waits for the callers still copying, runs the batch (H2D, kernels, D2H,
//                one synchronisation — the RunFn), fires the asynchronous completions and publishes the batch's epoch.
//
// While the device works on batch k, callers fill batch k+1 and the callers of batch k-1 are still copying their results
// out (three buffers): batches grow with load on their own (no timer).  Requests of one caller thread stay ordered (it
// does not return before its request has run); slices of one batch are ordered by begin() order, which is what
// per-shard FIFO needs.
//
// Locking: every state change is a few loads and stores under a SPIN lock (nothing blocks while holding it).  A caller
// waits for its batch with a short spin and then sleeps on the batch's own futex word; the dispatcher wakes kWakeFan
// sleepers and every woken caller wakes kWakeFan more (a tree: nobody queues on a mutex, and the depth — every level costs
// the wake-up latency of an idle core, 50-100 us out of a deep C-state — stays at two for 256 callers; with a fan-out
// of two the eight levels were the whole 1.5 ms p50 of ApplicationDB::Get).  Both alternatives were measured with 256 ApplicationDB::Get threads on the
// 128-core host: one condition variable for everybody = 55 K Gets/s (the herd re-acquiring its mutex takes longer than
// the batch), spin-then-yield = p50 0.27 ms but p99 300 ms (spinners starve the dispatcher once threads outnumber
// cores).  Spins are SHORT (a few microseconds: RSP_WAIT_SPINS / RSP_DISPATCH_SPINS pauses): the GPU boxes cap the
// container's CPU time (16 CPUs of 128 visible), and spinning callers spend the budget the dispatcher needs — 64 Get
// callers: 151 K Gets/s with 1500-pause spins, 574 K without (profiles/r02_seams_trace.md).
#pragma once
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstddef>
#include <cstdint>
#include <cstdlib>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

#include <linux/futex.h>
#include <sys/syscall.h>
#include <time.h>
#include <unistd.h>

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

  static constexpr int kBuffers = 4;  // filling | running | results being read (callers still waking up: two)
  static constexpr int kWakeFan = 4;     // per futex word (kWakeWords of them per batch)
  static constexpr int kWakeWords = 16;

  Stager(size_t cap_items, size_t cap_bytes, RunFn run, PostFn post = nullptr)
      : cap_items_(cap_items), cap_bytes_(cap_bytes), run_(std::move(run)), post_(std::move(post)) {
    b_[0].state = OPEN;
    b_[0].epoch = ++epochs_;
    open_ = 0;
    thread_ = std::thread([this] { Loop(); });
  }
  ~Stager() { Stop(); }
  void Stop() {
    if (stop_.exchange(true)) return;
    WakeSleepers();
    if (thread_.joinable()) thread_.join();
  }

  size_t cap_items() const { return cap_items_; }
  size_t cap_bytes() const { return cap_bytes_; }

  // Reserve n_items / n_bytes in a batch of class `klass` (requests of different classes never share a batch);
  // max_items bounds the items of a batch of this class (<= cap_items).  Waits while nothing can take the request.
  // false: the stager is stopping, or the request can never fit (the caller takes its direct path).
  bool begin(size_t n_items, size_t n_bytes, uint32_t klass, size_t max_items, Ticket* t) {
    if (max_items > cap_items_) max_items = cap_items_;
    if (n_items > max_items || n_bytes > cap_bytes_) return false;
    for (int spins = 0;; spins++) {
      if (stop_.load(std::memory_order_acquire)) return false;
      const uint32_t seen_free = free32_.load(std::memory_order_acquire);
      {
        SpinGuard g(sl_);
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
            work32_.fetch_add(1, std::memory_order_seq_cst);
            break;
          }
          b.full = true;  // full, or of another class: the dispatcher closes it as soon as it can
        }
      }
      // no room right now: a buffer opens within a batch cycle.  Sleep on the word that counts openings (no polling:
      // the box's CPU time is capped, profiles/r02_seams_trace.md); bounded, in case the opening raced the read above
      if (spins < 20) cpu_relax();
      else {
        buf_waiters_.fetch_add(1, std::memory_order_seq_cst);
        FutexWait(&free32_, seen_free, 500000);
        buf_waiters_.fetch_sub(1, std::memory_order_seq_cst);
      }
    }
    if (disp_sleeping_.load(std::memory_order_seq_cst)) FutexWake(&work32_, 1);
    if (buf_waiters_.load(std::memory_order_relaxed)) FutexWake(&free32_, 2);  // (a few at a time: they all fit one batch)
    return true;
  }
  // the caller finished writing its slice
  void commit(const Ticket& t) { b_[t.buf].copiers.fetch_sub(1, std::memory_order_acq_rel); }
  // completion without a waiting thread: fn runs on the dispatcher thread once the batch has run; the slice is
  // released when fn returns (fn reads its results from the staging buffers itself)
  void commit_async(const Ticket& t, std::function<void()> fn) {
    Batch& b = b_[t.buf];
    {
      SpinGuard g(sl_);
      b.async.push_back(std::move(fn));
    }
    b.copiers.fetch_sub(1, std::memory_order_acq_rel);
  }
  // until the batch has run
  void wait(const Ticket& t) {
    Batch& b = b_[t.buf];
    static const int kSpins = [] { const char* v = getenv("RSP_WAIT_SPINS"); return v ? atoi(v) : 100; }();
    for (int spins = 0; spins < kSpins; spins++) {  // a batch cycle is often shorter than a sleep + wake-up
      if (b.epoch_done.load(std::memory_order_acquire) >= t.epoch) return;
      cpu_relax();
    }
    // sleepers spread over kWakeWords futex words (the kernel hashes a futex by its address: hundreds of callers waiting
    // on and waking ONE word queue on one hash-bucket lock inside the kernel)
    WakeWord& ww = b.wake[(uint32_t)((t.item0 * 2654435761u) >> 16) % kWakeWords];
    bool slept = false;
    for (;;) {
      const uint32_t s = ww.seq.load(std::memory_order_acquire);
      if (b.epoch_done.load(std::memory_order_acquire) >= t.epoch) break;
      ww.sleepers.fetch_add(1, std::memory_order_seq_cst);
      FutexWait(&ww.seq, s, 2000000);  // (bounded: 2 ms)
      ww.sleepers.fetch_sub(1, std::memory_order_seq_cst);
      slept = true;
    }
    if (slept && ww.sleepers.load(std::memory_order_seq_cst)) FutexWake(&ww.seq, kWakeFan);  // pass the wake-up on
  }
  void release(const Ticket& t) {
    Batch& b = b_[t.buf];
    if (b.users.fetch_sub(1, std::memory_order_acq_rel) == 1) {
      {
        SpinGuard g(sl_);
        MaybeFree(b);
      }
      WakeBufferWaiters();
    }
  }

  uint64_t batches() const { return batches_.load(std::memory_order_relaxed); }
  // diagnostics: batches run, items carried, ns inside the RunFn, ns waiting for callers still copying, ns idle
  void stats(uint64_t out[5]) const {
    out[0] = batches_.load(std::memory_order_relaxed); out[1] = st_items_.load(std::memory_order_relaxed);
    out[2] = st_run_ns_.load(std::memory_order_relaxed); out[3] = st_copy_ns_.load(std::memory_order_relaxed);
    out[4] = st_idle_ns_.load(std::memory_order_relaxed);
  }

 private:
  enum State { FREE, OPEN, CLOSED, DONE };
  struct alignas(64) WakeWord {
    std::atomic<uint32_t> seq{0};
    std::atomic<uint32_t> sleepers{0};  // callers inside FutexWait on seq (a wake-up is a system call: skipped when 0)
  };
  struct Batch {
    State state = FREE;                 // sl_
    size_t n_items = 0, n_bytes = 0;    // sl_
    uint32_t klass = 0;
    bool full = false;
    uint64_t epoch = 0;
    std::atomic<uint32_t> copiers{0};   // callers still writing their slice
    std::atomic<uint32_t> users{0};     // callers (sync and async) that have not released their slice yet
    std::atomic<uint64_t> epoch_done{0};
    WakeWord wake[kWakeWords];          // futex words: bumped when the batch has run
    std::vector<std::function<void()>> async;  // sl_
  };
  static inline void cpu_relax() {
#if defined(__x86_64__) || defined(__i386__)
    __builtin_ia32_pause();
#else
    std::this_thread::yield();
#endif
  }
  struct SpinLock {
    std::atomic<bool> f{false};
    void lock() {
      for (;;) {
        if (!f.exchange(true, std::memory_order_acquire)) return;
        for (int n = 0; f.load(std::memory_order_relaxed); n++) {
          if (n < 2000) cpu_relax();
          else {  // the holder was descheduled (threads may outnumber cores): get out of its way
            struct timespec ts = {0, 5000};
            nanosleep(&ts, nullptr);
          }
        }
      }
    }
    void unlock() { f.store(false, std::memory_order_release); }
  };
  struct SpinGuard {
    SpinLock& l;
    explicit SpinGuard(SpinLock& x) : l(x) { l.lock(); }
    ~SpinGuard() { l.unlock(); }
  };

  static void FutexWait(std::atomic<uint32_t>* w, uint32_t expected, long timeout_ns) {
    struct timespec ts;
    ts.tv_sec = timeout_ns / 1000000000L;
    ts.tv_nsec = timeout_ns % 1000000000L;
    syscall(SYS_futex, reinterpret_cast<uint32_t*>(w), FUTEX_WAIT_PRIVATE, expected, &ts, nullptr, 0);
  }
  static void FutexWake(std::atomic<uint32_t>* w, int n) {
    syscall(SYS_futex, reinterpret_cast<uint32_t*>(w), FUTEX_WAKE_PRIVATE, n, nullptr, nullptr, 0);
  }
  static int64_t NowNs() {
    return std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now().time_since_epoch()).count();
  }
  static void NapUs(long us) {
    struct timespec ts;
    ts.tv_sec = 0;
    ts.tv_nsec = us * 1000L;
    nanosleep(&ts, nullptr);
  }
  void WakeBufferWaiters() {  // after a buffer may have opened (sl_ NOT held: a system call)
    if (buf_waiters_.load(std::memory_order_seq_cst)) FutexWake(&free32_, kWakeFan);
  }
  void WakeSleepers() {  // Stop(): the dispatcher may be asleep, callers may wait for a buffer
    FutexWake(&work32_, 1);
    FutexWake(&free32_, 1 << 20);
  }

  bool MaybeFree(Batch& b) {  // sl_ held: whoever sees "done and unused" first recycles the buffer
    if (b.state == DONE && b.users.load(std::memory_order_acquire) == 0) {
      b.state = FREE;
      b.n_items = b.n_bytes = 0;
      if (open_ < 0) OpenOne();
      return true;
    }
    return false;
  }
  void OpenOne() {  // sl_ held
    for (int i = 0; i < kBuffers; i++) {
      if (b_[i].state == FREE) {
        b_[i].state = OPEN;
        b_[i].full = false;
        b_[i].epoch = ++epochs_;
        open_ = i;
        free32_.fetch_add(1, std::memory_order_seq_cst);  // (whoever called wakes the waiters after dropping sl_)
        return;
      }
    }
  }
  void Loop() {
    for (;;) {
      // ---- take the open batch once it holds work
      int bi = -1;
      const int64_t t_idle0 = NowNs();
      for (int spins = 0;; spins++) {
        {
          SpinGuard g(sl_);
          if (open_ >= 0 && b_[open_].n_items > 0) {
            bi = open_;
            b_[bi].state = CLOSED;
            open_ = -1;
            OpenOne();
            break;
          }
        }
        if (stop_.load(std::memory_order_acquire)) return;  // (queued work was taken above: callers wait on it)
        static const int kIdleSpins = [] { const char* v = getenv("RSP_DISPATCH_SPINS"); return v ? atoi(v) : 100; }();
        if (spins < kIdleSpins) cpu_relax();  // stay hot between batches under load
        else {
          disp_sleeping_.store(true, std::memory_order_seq_cst);
          const uint32_t w = work32_.load(std::memory_order_seq_cst);
          if (w == seen_work_ && !stop_.load()) FutexWait(&work32_, w, 2000000);
          disp_sleeping_.store(false, std::memory_order_seq_cst);
        }
      }
      seen_work_ = work32_.load(std::memory_order_seq_cst);
      WakeBufferWaiters();  // (the next buffer opened when this one was closed)
      Batch& b = b_[bi];
      const int64_t t_copy0 = NowNs();
      st_idle_ns_.fetch_add((uint64_t)(t_copy0 - t_idle0), std::memory_order_relaxed);
      for (int spins = 0; b.copiers.load(std::memory_order_acquire); spins++) {
        if (spins < 2000) cpu_relax(); else NapUs(10);  // (a copier may have been descheduled)
      }
      const int64_t t_run0 = NowNs();
      st_copy_ns_.fetch_add((uint64_t)(t_run0 - t_copy0), std::memory_order_relaxed);
      BatchInfo info;
      std::vector<std::function<void()>> async;
      {
        SpinGuard g(sl_);
        info = BatchInfo{bi, b.n_items, b.n_bytes, b.klass, b.epoch};
        async.swap(b.async);
      }
      run_(info);
      st_run_ns_.fetch_add((uint64_t)(NowNs() - t_run0), std::memory_order_relaxed);
      st_items_.fetch_add(info.n_items, std::memory_order_relaxed);
      for (auto& f : async) f();
      if (post_) post_();
      b.epoch_done.store(info.epoch, std::memory_order_release);  // spinning callers go on at once
      for (int w = 0; w < kWakeWords; w++) {  // sleeping ones: a few per word, who wake the others
        b.wake[w].seq.fetch_add(1, std::memory_order_seq_cst);
        if (b.wake[w].sleepers.load(std::memory_order_seq_cst)) FutexWake(&b.wake[w].seq, 2);
      }
      batches_.fetch_add(1, std::memory_order_relaxed);
      {
        SpinGuard g(sl_);
        b.state = DONE;
        if (!async.empty()) b.users.fetch_sub((uint32_t)async.size(), std::memory_order_acq_rel);
        MaybeFree(b);
      }
      WakeBufferWaiters();
    }
  }

  const size_t cap_items_, cap_bytes_;
  RunFn run_;
  PostFn post_;
  SpinLock sl_;
  Batch b_[kBuffers];
  int open_ = -1;
  uint64_t epochs_ = 0;
  std::atomic<bool> stop_{false}, disp_sleeping_{false};
  std::atomic<uint32_t> work32_{0};  // requests ever accepted (futex word): the idle dispatcher sleeps until it moves
  std::atomic<uint32_t> free32_{0};  // buffers ever opened (futex word): callers without a buffer sleep until it moves
  std::atomic<uint32_t> buf_waiters_{0};
  uint32_t seen_work_ = 0;
  std::atomic<uint64_t> batches_{0}, st_items_{0}, st_run_ns_{0}, st_copy_ns_{0}, st_idle_ns_{0};
  std::thread thread_;
};

}  // namespace rsp
