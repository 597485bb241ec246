// stress of csrc/stager.h: many callers, mixed classes and sizes, sync and async completions; every result checked
#include <cstdio>
#include <cstring>
#include <thread>
#include <vector>
#include "../../rocksplicator_b200/csrc/stager.h"
int main(int argc, char** argv) {
  const int T = argc > 1 ? atoi(argv[1]) : 64, N = argc > 2 ? atoi(argv[2]) : 2000;
  constexpr size_t CAP = 4096;
  static long in_buf[rsp::Stager::kBuffers][CAP], out_buf[rsp::Stager::kBuffers][CAP];
  std::atomic<long> executed{0}, async_done{0}, wrong{0};
  rsp::Stager st(CAP, CAP * 8, [&](const rsp::Stager::BatchInfo& b) {
    for (size_t i = 0; i < b.n_items; i++) {
      if ((uint32_t)(in_buf[b.buf][i] & 1) != (b.klass ? 1u : 0u)) wrong++;
      out_buf[b.buf][i] = in_buf[b.buf][i] * 2 + 1;
      executed++;
    }
    if (b.n_bytes != b.n_items * 8) wrong++;
  });
  std::vector<std::thread> th;
  for (int t = 0; t < T; t++)
    th.emplace_back([&, t] {
      for (int i = 0; i < N; i++) {
        rsp::Stager::Ticket k;
        const int n = 1 + (i * 7 + t) % 5;
        const uint32_t klass = (uint32_t)(t & 1) * 64u;  // classes 0 and 64
        if (!st.begin(n, n * 8, klass, CAP, &k)) { wrong++; continue; }
        for (int j = 0; j < n; j++) in_buf[k.buf][k.item0 + j] = (((long)t * 1000000 + i * 10 + j) << 1) | (klass ? 1 : 0);
        if (i % 4 == 3) {
          const int buf = k.buf; const size_t i0 = k.item0;
          st.commit_async(k, [&, buf, i0, n, t, i, klass] {
            for (int j = 0; j < n; j++) if (out_buf[buf][i0 + j] != (((((long)t * 1000000 + i * 10 + j) << 1) | (klass ? 1 : 0)) * 2 + 1)) wrong++;
            async_done++;
          });
        } else {
          st.commit(k);
          st.wait(k);
          for (int j = 0; j < n; j++) if (out_buf[k.buf][k.item0 + j] != in_buf[k.buf][k.item0 + j] * 2 + 1) wrong++;
          st.release(k);
        }
      }
    });
  for (auto& t : th) t.join();
  for (int i = 0; i < 2000 && async_done.load() != (long)T * (N / 4); i++) usleep(1000);
  printf("threads %d x %d: executed %ld items in %llu batches, async %ld / %ld, wrong %ld\n", T, N, executed.load(),
         (unsigned long long)st.batches(), async_done.load(), (long)T * (N / 4), wrong.load());
  return wrong.load() || async_done.load() != (long)T * (N / 4);
}
