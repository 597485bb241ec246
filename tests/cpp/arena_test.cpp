// random alloc / release against csrc/arena.h over malloc'ed "slabs": no two live blocks overlap, contents survive,
// the allocator's invariants (tiling, coalescing, size index, counters) hold after every step, and memory that was
// released in small pieces is found again by a large request (what size-class free lists could not do)
#include <cstdio>
#include <cstring>
#include <random>
#include "../../rocksplicator_b200/csrc/arena.h"
static size_t g_slab_calls = 0;
int main() {
  rsp::Arena a;
  a.slab_bytes = 1 << 20;
  a.slab_alloc = [](size_t n) -> void* { g_slab_calls++; void* p = nullptr; if (posix_memalign(&p, 4096, n)) abort(); return p; };
  a.slab_free = [](void* p) { free(p); };
  std::mt19937_64 rng(12345);
  struct Live { unsigned char* p; size_t n; unsigned char tag; };
  std::vector<Live> live;
  long bad = 0;
  for (int step = 0; step < 60000; step++) {
    const bool do_alloc = live.empty() || (rng() % 100) < (live.size() < 400 ? 60u : 40u);
    if (do_alloc) {
      size_t n = 1 + rng() % ((rng() % 8 == 0) ? 300000 : 20000);
      if (rng() % 500 == 0) n = (1 << 20) + rng() % (1 << 20);  // beyond a slab: a slab of its own
      unsigned char* p = (unsigned char*)a.alloc(n);
      if (((uintptr_t)p & 255) != 0) bad++;
      const unsigned char tag = (unsigned char)(rng() | 1);
      memset(p, tag, n);
      live.push_back({p, n, tag});
    } else {
      const size_t k = rng() % live.size();
      const Live l = live[k];
      for (size_t i = 0; i < l.n; i += 97) if (l.p[i] != l.tag) { bad++; break; }  // nobody wrote into it
      if (l.p[l.n - 1] != l.tag) bad++;
      a.release(l.p, l.n);
      live[k] = live.back();
      live.pop_back();
    }
    if (step % 64 == 0 && !a.check()) { bad++; printf("invariants broken at step %d\n", step); break; }
  }
  for (auto& l : live) a.release(l.p, l.n);
  if (!a.check() || a.in_use != 0 || a.free_bytes != a.reserved) bad++;
  // everything coalesced: one free block per slab, and a request of a whole slab is served without a new one
  if (a.free_by_size.size() != a.slabs.size()) bad++;
  const size_t before = g_slab_calls;
  void* big = a.alloc(a.slab_bytes);
  if (g_slab_calls != before) bad++;
  a.release(big, a.slab_bytes);
  a.release(big, a.slab_bytes);  // released twice: ignored
  if (!a.check()) bad++;
  printf("arena: %zu slabs, %zu slab requests, reserved %zu bytes, bad %ld\n", a.slabs.size(), g_slab_calls, a.reserved, bad);
  a.destroy();
  return bad != 0;
}
