// TEST INFRASTRUCTURE ONLY — see tests/emul/include/cuda_runtime.h.  Fiber scheduler + the runtime calls csrc/ makes.
#include <cuda_runtime.h>
#include <stdio.h>
#include <ucontext.h>

#include <chrono>
#include <map>
#include <mutex>
#include <vector>

#if defined(__SANITIZE_ADDRESS__)
extern "C" void __sanitizer_start_switch_fiber(void** fake_stack_save, const void* bottom, size_t size);
extern "C" void __sanitizer_finish_switch_fiber(void* fake_stack_save, const void** bottom_old, size_t* size_old);
#define EMUL_ASAN 1
#else
#define EMUL_ASAN 0
#endif

emul_idx threadIdx, blockIdx;
dim3 blockDim, gridDim;
unsigned char* emul_dyn_smem = nullptr;

namespace {

enum State { RUNNABLE, WAIT_WARP, WAIT_BLOCK, YIELDED, DONE };
struct Fiber {
  ucontext_t ctx;
  State st;
  int op;
  unsigned mask, arg;
  uint64_t val, result;
  unsigned tid;
  char* stack;
  void* asan_fake = nullptr;
};
constexpr size_t kStack = 256 * 1024;

std::recursive_mutex g_mu;  // launches are serialised: one block's fibers at a time, process-wide
ucontext_t g_sched;
void* g_sched_fake = nullptr;
const void* g_sched_bottom = nullptr;
size_t g_sched_size = 0;
Fiber* g_cur = nullptr;
const std::function<void()>* g_body = nullptr;
std::vector<Fiber> g_fibers;
std::vector<char*> g_stacks;

void to_scheduler(Fiber* f, bool dying) {
#if EMUL_ASAN
  __sanitizer_start_switch_fiber(dying ? nullptr : &f->asan_fake, g_sched_bottom, g_sched_size);
#endif
  (void)dying;
  swapcontext(&f->ctx, &g_sched);
#if EMUL_ASAN
  __sanitizer_finish_switch_fiber(f->asan_fake, &g_sched_bottom, &g_sched_size);
#endif
}

void fiber_main() {
#if EMUL_ASAN
  __sanitizer_finish_switch_fiber(nullptr, &g_sched_bottom, &g_sched_size);
#endif
  (*g_body)();
  Fiber* f = g_cur;
  f->st = DONE;
  to_scheduler(f, true);
  abort();  // a finished fiber is never resumed
}

void resume(Fiber* f) {
  g_cur = f;
  threadIdx.x = f->tid % blockDim.x;
  threadIdx.y = (f->tid / blockDim.x) % blockDim.y;
  threadIdx.z = f->tid / (blockDim.x * blockDim.y);
  f->st = RUNNABLE;
#if EMUL_ASAN
  __sanitizer_start_switch_fiber(&g_sched_fake, f->stack, kStack);
#endif
  swapcontext(&g_sched, &f->ctx);
#if EMUL_ASAN
  __sanitizer_finish_switch_fiber(g_sched_fake, nullptr, nullptr);
#endif
  g_cur = nullptr;
}

// complete every collective of warp [w0, w0+n) whose participants have all arrived; true if anything was released
bool resolve_warp(unsigned w0, unsigned n) {
  bool any = false;
  bool handled[32] = {false};
  for (unsigned l = 0; l < n; l++) {
    Fiber& f = g_fibers[w0 + l];
    if (f.st != WAIT_WARP || handled[l]) continue;
    const unsigned M = f.mask;
    if (!((M >> l) & 1u)) {
      fprintf(stderr, "[cuda_emul] lane %u called a collective with mask %08x that does not name it\n", l, M);
      abort();
    }
    bool ready = true;
    for (unsigned j = 0; j < n && ready; j++) {
      if (!((M >> j) & 1u)) continue;
      const Fiber& g = g_fibers[w0 + j];
      if (g.st == DONE) continue;
      if (!(g.st == WAIT_WARP && g.mask == M && g.op == f.op)) ready = false;
    }
    if (!ready) continue;
    auto part = [&](unsigned j) { return j < n && ((M >> j) & 1u) && g_fibers[w0 + j].st == WAIT_WARP; };
    uint64_t res[32];
    unsigned ballot = 0;
    for (unsigned j = 0; j < n; j++)
      if (part(j) && g_fibers[w0 + j].val) ballot |= 1u << j;
    for (unsigned j = 0; j < n; j++) {
      if (!part(j)) continue;
      const Fiber& g = g_fibers[w0 + j];
      unsigned src = j;
      switch (f.op) {
        case EMUL_SHFL_IDX: src = g.arg & 31u; break;
        case EMUL_SHFL_UP: src = j >= g.arg ? j - g.arg : j; break;
        case EMUL_SHFL_XOR: src = j ^ (g.arg & 31u); break;
        default: break;
      }
      if (f.op == EMUL_BALLOT) res[j] = ballot;
      else if (f.op == EMUL_SYNCWARP) res[j] = 0;
      else res[j] = part(src) ? g_fibers[w0 + src].val : g.val;
    }
    for (unsigned j = 0; j < n; j++) {
      if (!part(j)) continue;
      g_fibers[w0 + j].result = res[j];
      g_fibers[w0 + j].st = RUNNABLE;
      handled[j] = true;
    }
    any = true;
  }
  return any;
}

void run_block(unsigned n_threads) {
  if (g_fibers.size() < n_threads) g_fibers.resize(n_threads);
  while (g_stacks.size() < n_threads) g_stacks.push_back((char*)malloc(kStack));
  for (unsigned t = 0; t < n_threads; t++) {
    Fiber& f = g_fibers[t];
    getcontext(&f.ctx);
    f.stack = g_stacks[t];
    f.ctx.uc_stack.ss_sp = f.stack;
    f.ctx.uc_stack.ss_size = kStack;
    f.ctx.uc_link = nullptr;
    makecontext(&f.ctx, fiber_main, 0);
    f.st = RUNNABLE;
    f.tid = t;
    f.asan_fake = nullptr;
  }
  unsigned alive = n_threads;
  while (alive) {
    bool progress = false;
    for (unsigned w0 = 0; w0 < n_threads; w0 += 32) {
      const unsigned n = std::min(32u, n_threads - w0);
      for (unsigned l = 0; l < n; l++) {
        Fiber& f = g_fibers[w0 + l];
        if (f.st != RUNNABLE && f.st != YIELDED) continue;
        resume(&f);
        progress = true;
        if (f.st == DONE) alive--;
      }
      if (resolve_warp(w0, n)) progress = true;
    }
    unsigned at_barrier = 0;
    for (unsigned t = 0; t < n_threads; t++) at_barrier += g_fibers[t].st == WAIT_BLOCK;
    if (alive && at_barrier == alive) {
      for (unsigned t = 0; t < n_threads; t++)
        if (g_fibers[t].st == WAIT_BLOCK) g_fibers[t].st = RUNNABLE;
      progress = true;
    }
    if (!progress) {
      fprintf(stderr, "[cuda_emul] deadlock in block (%u,%u,%u): ", blockIdx.x, blockIdx.y, blockIdx.z);
      for (unsigned t = 0; t < n_threads && t < 64; t++) fprintf(stderr, "%d", (int)g_fibers[t].st);
      fprintf(stderr, "\n");
      abort();
    }
  }
}

}  // namespace

uint64_t emul_collective(int op, unsigned mask, uint64_t value, unsigned arg) {
  Fiber* f = g_cur;
  f->op = op; f->mask = mask; f->val = value; f->arg = arg;
  f->st = WAIT_WARP;
  to_scheduler(f, false);
  return f->result;
}
void emul_syncthreads() {
  Fiber* f = g_cur;
  f->st = WAIT_BLOCK;
  to_scheduler(f, false);
}
void emul_yield() {
  Fiber* f = g_cur;
  f->st = YIELDED;
  to_scheduler(f, false);
}

void emul_launch(dim3 grid, dim3 block, size_t dyn_smem, const std::function<void()>& body) {
  std::lock_guard<std::recursive_mutex> g(g_mu);
  const unsigned n_threads = block.x * block.y * block.z;
  if (!n_threads || !grid.x || !grid.y || !grid.z) return;
  gridDim = grid;
  blockDim = block;
  std::vector<unsigned char> smem(dyn_smem + 64);
  emul_dyn_smem = (unsigned char*)(((uintptr_t)smem.data() + 63) & ~(uintptr_t)63);
  g_body = &body;
  for (unsigned z = 0; z < grid.z; z++)
    for (unsigned y = 0; y < grid.y; y++)
      for (unsigned x = 0; x < grid.x; x++) {
        blockIdx.x = x; blockIdx.y = y; blockIdx.z = z;
        run_block(n_threads);
      }
  g_body = nullptr;
  emul_dyn_smem = nullptr;
}

// ---- runtime ------------------------------------------------------------------------------------------------------
struct emul_stream { int id; };
struct emul_event { std::chrono::steady_clock::time_point t; };

// RSP_EMUL_DEVICES=<n>: pretend n devices (memory is malloc-backed, so they are interchangeable) — lets the
// several-engines-in-one-process paths (router) run on the CPU
int emul_sync_acc = 0;
static int emul_n_devices() { const char* e = getenv("RSP_EMUL_DEVICES"); const int n = e ? atoi(e) : 1; return n > 0 ? n : 1; }
static thread_local int g_cur_device = 0;
cudaError_t cudaGetDeviceCount(int* n) { *n = emul_n_devices(); return cudaSuccess; }
cudaError_t cudaSetDevice(int d) { if (d < 0 || d >= emul_n_devices()) return cudaErrorInvalidValue; g_cur_device = d; return cudaSuccess; }
cudaError_t cudaGetDevice(int* d) { *d = g_cur_device; return cudaSuccess; }
cudaError_t cudaDeviceSetLimit(cudaLimit, size_t) { return cudaSuccess; }
cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
// RSP_EMUL_DEVICE_BYTES: the emulated device's memory (tests of the out-of-memory paths); pinned host memory does not count
static std::mutex g_mem_mu;
static std::map<void*, size_t> g_dev_blocks;
static size_t g_dev_bytes = 0;
static size_t emul_device_limit() { static const size_t v = [] { const char* e = getenv("RSP_EMUL_DEVICE_BYTES"); return e ? (size_t)atoll(e) : (size_t)0; }(); return v; }
static cudaError_t emul_alloc(void** p, size_t n) {
  // exactly n bytes (ASan's red zone starts right behind them), 256-byte aligned like the real allocator
  if (posix_memalign(p, 256, n ? n : 1) != 0) { *p = nullptr; return cudaErrorMemoryAllocation; }
  if (n <= (16u << 20)) memset(*p, 0xCD, n);  // device memory is not zeroed: make a missing memset visible (big slabs stay lazy)
  return cudaSuccess;
}
cudaError_t cudaMalloc(void** p, size_t n) {
  if (emul_device_limit()) {
    std::lock_guard<std::mutex> g(g_mem_mu);
    if (g_dev_bytes + n > emul_device_limit()) { *p = nullptr; return cudaErrorMemoryAllocation; }
    const cudaError_t rc = emul_alloc(p, n);
    if (rc == cudaSuccess) { g_dev_blocks[*p] = n; g_dev_bytes += n; }
    return rc;
  }
  return emul_alloc(p, n);
}
cudaError_t cudaFree(void* p) {
  if (emul_device_limit() && p) {
    std::lock_guard<std::mutex> g(g_mem_mu);
    auto it = g_dev_blocks.find(p);
    if (it != g_dev_blocks.end()) { g_dev_bytes -= it->second; g_dev_blocks.erase(it); }
  }
  free(p);
  return cudaSuccess;
}
cudaError_t cudaHostAlloc(void** p, size_t n, unsigned) { return emul_alloc(p, n); }
cudaError_t cudaFreeHost(void* p) { free(p); return cudaSuccess; }
cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { if (n) memmove(d, s, n); return cudaSuccess; }
cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind k, cudaStream_t) {
  std::lock_guard<std::recursive_mutex> g(g_mu);  // ordered with launches, as a stream would order them
  return cudaMemcpy(d, s, n, k);
}
cudaError_t cudaMemset(void* d, int v, size_t n) { if (n) memset(d, v, n); return cudaSuccess; }
cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t) {
  std::lock_guard<std::recursive_mutex> g(g_mu);
  return cudaMemset(d, v, n);
}
cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = new emul_stream{0}; return cudaSuccess; }
cudaError_t cudaStreamDestroy(cudaStream_t s) { delete s; return cudaSuccess; }
cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }
cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = new emul_event{std::chrono::steady_clock::now()}; return cudaSuccess; }
cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { return cudaEventCreate(e); }
cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return cudaSuccess; }
cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t) { e->t = std::chrono::steady_clock::now(); return cudaSuccess; }
cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t a, cudaEvent_t b) {
  *ms = std::chrono::duration<float, std::milli>(b->t - a->t).count();
  return cudaSuccess;
}
cudaError_t cudaPeekAtLastError() { return cudaSuccess; }
cudaError_t cudaGetLastError() { return cudaSuccess; }
const char* cudaGetErrorName(cudaError_t e) { return e == cudaSuccess ? "cudaSuccess" : "cudaError(emulated)"; }
const char* cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : "emulated CUDA error"; }
