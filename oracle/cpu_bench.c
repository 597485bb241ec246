/*
 * cpu_bench.c — times the CPU side of the hot path through the oracle interface (okv.h).
 * TEST / MEASUREMENT INFRASTRUCTURE: used only by bench.py's cpu_baseline and --impl reference legs.
 *
 * With oracle/_ref/libokv_ref.so it runs the reference's OWN RocksDB binary the way the reference
 * drives it: apply = WriteBatch(bytes) + PutLogData(ts) + DB::Write(default WriteOptions)
 * (rocksdb_replicator/rocksdb_wrapper.cpp:13-31); reads = DB::MultiGet per shard
 * (rocksdb_admin/application_db.cpp:113-120).  Shards are statically partitioned over the threads — the
 * shape of the replicator's executor fan-out (rocksdb_replicator/rocksdb_replicator.cpp:58-67).
 *
 * Workload = bench.py's: keys 16 B (BE index ‖ BE splitmix64(seed ^ index)), shard = index % shards,
 * values 64 B, single-Put replicated batches of 105 wire bytes applied in pull-sized groups of 50 per
 * shard, then (optionally) flush + full compaction, then MultiGet batches of `batch` uniform keys.
 *
 * usage: okv_cpu_bench <liboKV.so> <threads> <shards> <kv_total> <vlen> <wal 0|1> <apply_secs_cap>
 *                      <get_secs> <batch> <compact 0|1> <dir>
 * prints one JSON object.
 */
#define _GNU_SOURCE
#include <dlfcn.h>
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

typedef struct okv_db okv_db;
static okv_db* (*p_open)(const char*, int, int, char*, size_t);
static void (*p_close)(okv_db*);
static int (*p_apply)(okv_db*, const uint8_t*, size_t, uint64_t, char*, size_t);
static uint64_t (*p_latest_seq)(okv_db*);
static int (*p_multi_get)(okv_db*, size_t, const uint8_t*, const uint64_t*, int32_t*, uint8_t**, uint64_t*);
static int (*p_flush)(okv_db*);
static int (*p_compact)(okv_db*);
static void (*p_free)(void*);
static const char* (*p_kind)(void);

static double now_s(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec + ts.tv_nsec * 1e-9;
}
static uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  uint64_t z = x;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
static void put_be64(uint8_t* p, uint64_t v) {
  for (int i = 0; i < 8; i++) p[i] = (uint8_t)(v >> (56 - 8 * i));
}
#define SEED_DATA 0x5EED0001ull
#define SEED_QUERY 0x5EED0002ull
static void make_key(uint8_t* k, uint64_t idx) {
  put_be64(k, idx);
  put_be64(k + 8, splitmix64(SEED_DATA ^ idx));
}
static void make_value(uint8_t* v, uint32_t vlen, uint64_t shard, uint64_t idx, uint64_t version) {
  uint64_t s = splitmix64(SEED_DATA ^ (shard << 40) ^ (idx << 8) ^ version);
  for (uint32_t w = 0; w * 8 < vlen; w++) {
    s = splitmix64(s);
    uint32_t n = vlen - w * 8 < 8 ? vlen - w * 8 : 8;
    memcpy(v + w * 8, &s, n);
  }
}

typedef struct {
  int tid, threads, shards, wal, compact;
  uint64_t kv_total;
  uint32_t vlen, batch;
  double apply_cap, get_secs;
  okv_db** dbs;
  /* results */
  uint64_t applied, lookups, hits;
  double apply_s, compact_s, get_s;
  volatile int* stop;
} targ;

static pthread_barrier_t g_bar;

static void* worker(void* a_) {
  targ* a = (targ*)a_;
  const uint32_t vl = a->vlen;
  /* wire bytes of one replicated single-Put batch */
  size_t vv = vl < 128 ? 1 : 2;
  size_t L = 12 + 1 + 1 + 16 + vv + vl + 10;
  uint8_t* b = (uint8_t*)calloc(1, L);
  b[8] = 1;
  b[12] = 1;
  b[13] = 16;
  size_t vo = 14 + 16;
  if (vv == 1) b[vo] = (uint8_t)vl; else { b[vo] = (uint8_t)((vl & 0x7f) | 0x80); b[vo + 1] = (uint8_t)(vl >> 7); }
  size_t lo = vo + vv + vl;
  b[lo] = 3;
  b[lo + 1] = 8;
  /* my shards: s % threads == tid; keys of shard s: idx = s + j*shards */
  uint64_t per_shard = a->kv_total / a->shards;
  pthread_barrier_wait(&g_bar);
  double t0 = now_s();
  uint64_t applied = 0;
  int capped = 0;
  for (uint64_t j0 = 0; j0 < per_shard && !capped; j0 += 50) { /* pull-sized groups */
    for (int s = a->tid; s < a->shards; s += a->threads) {
      uint64_t j1 = j0 + 50 < per_shard ? j0 + 50 : per_shard;
      for (uint64_t j = j0; j < j1; j++) {
        uint64_t idx = (uint64_t)s + j * a->shards;
        make_key(b + 14, idx);
        make_value(b + vo + vv, vl, s, idx, 0);
        uint64_t ts = 1000 + idx;
        memcpy(b + lo + 2, &ts, 8);
        if (p_apply(a->dbs[s], b, L, ts, NULL, 0) != 0) { fprintf(stderr, "apply failed\n"); exit(2); }
        applied++;
      }
    }
    if (now_s() - t0 > a->apply_cap) capped = 1;
  }
  a->apply_s = now_s() - t0;
  a->applied = applied;
  uint64_t loaded_per_shard = applied / ((a->shards - a->tid + a->threads - 1) / a->threads);
  pthread_barrier_wait(&g_bar);
  t0 = now_s();
  if (a->compact)
    for (int s = a->tid; s < a->shards; s += a->threads) { p_flush(a->dbs[s]); p_compact(a->dbs[s]); }
  a->compact_s = now_s() - t0;
  pthread_barrier_wait(&g_bar);
  /* MultiGet: batches of `batch` keys uniform over my shards, split per shard (the router's job,
   * examples/counter_service/counter_router.cpp:36-66), one DB::MultiGet per shard */
  int my_n = (a->shards - a->tid + a->threads - 1) / a->threads;
  uint8_t* keys = (uint8_t*)malloc((size_t)a->batch * 16);
  uint64_t* koff = (uint64_t*)malloc(sizeof(uint64_t) * (a->batch + 1));
  int32_t* st = (int32_t*)malloc(sizeof(int32_t) * a->batch);
  uint64_t* voff = (uint64_t*)malloc(sizeof(uint64_t) * (a->batch + 1));
  uint32_t* cnt = (uint32_t*)calloc(my_n + 1, 4);
  uint32_t* pick_s = (uint32_t*)malloc(4 * a->batch);
  uint64_t* pick_j = (uint64_t*)malloc(8 * a->batch);
  uint64_t rng = splitmix64(SEED_QUERY ^ (uint64_t)a->tid);
  uint64_t lookups = 0, hits = 0;
  t0 = now_s();
  while (now_s() - t0 < a->get_secs && loaded_per_shard > 0) {
    memset(cnt, 0, 4 * (my_n + 1));
    for (uint32_t q = 0; q < a->batch; q++) {
      rng = splitmix64(rng);
      pick_s[q] = (uint32_t)(rng % my_n);
      pick_j[q] = (rng >> 32) % loaded_per_shard;
      cnt[pick_s[q] + 1]++;
    }
    for (int i = 0; i < my_n; i++) cnt[i + 1] += cnt[i];
    /* bucket by shard */
    uint32_t* fill = (uint32_t*)calloc(my_n, 4);
    for (uint32_t q = 0; q < a->batch; q++) {
      uint32_t pos = cnt[pick_s[q]] + fill[pick_s[q]]++;
      uint64_t s = (uint64_t)a->tid + (uint64_t)pick_s[q] * a->threads;
      make_key(keys + (size_t)pos * 16, s + pick_j[q] * a->shards);
    }
    free(fill);
    for (int i = 0; i < my_n; i++) {
      uint32_t n = cnt[i + 1] - cnt[i];
      if (!n) continue;
      for (uint32_t k = 0; k <= n; k++) koff[k] = (uint64_t)k * 16;
      uint8_t* vals = NULL;
      p_multi_get(a->dbs[a->tid + i * a->threads], n, keys + (size_t)cnt[i] * 16, koff, st, &vals, voff);
      for (uint32_t k = 0; k < n; k++) hits += st[k] == 0;
      p_free(vals);
      lookups += n;
    }
  }
  a->get_s = now_s() - t0;
  a->lookups = lookups;
  a->hits = hits;
  free(b); free(keys); free(koff); free(st); free(voff); free(cnt); free(pick_s); free(pick_j);
  return NULL;
}

int main(int argc, char** argv) {
  if (argc < 12) {
    fprintf(stderr, "usage: %s lib threads shards kv_total vlen wal apply_secs_cap get_secs batch compact dir\n", argv[0]);
    return 1;
  }
  void* h = dlopen(argv[1], RTLD_NOW | RTLD_GLOBAL);
  if (!h) { fprintf(stderr, "dlopen: %s\n", dlerror()); return 1; }
#define LD(v, n) do { *(void**)(&v) = dlsym(h, n); if (!v) { fprintf(stderr, "dlsym %s\n", n); return 1; } } while (0)
  LD(p_open, "okv_open"); LD(p_close, "okv_close"); LD(p_apply, "okv_apply"); LD(p_latest_seq, "okv_latest_seq");
  LD(p_multi_get, "okv_multi_get"); LD(p_flush, "okv_flush"); LD(p_compact, "okv_compact"); LD(p_free, "okv_free");
  LD(p_kind, "okv_kind");
  int threads = atoi(argv[2]), shards = atoi(argv[3]);
  uint64_t kv_total = strtoull(argv[4], NULL, 10);
  uint32_t vlen = (uint32_t)atoi(argv[5]);
  int wal = atoi(argv[6]);
  double apply_cap = atof(argv[7]), get_secs = atof(argv[8]);
  uint32_t batch = (uint32_t)atoi(argv[9]);
  int compact = atoi(argv[10]);
  const char* dir = argv[11];
  if (threads > shards) threads = shards;
  okv_db** dbs = (okv_db**)calloc(shards, sizeof(okv_db*));
  char path[4096], err[512];
  double t_open = now_s();
  for (int s = 0; s < shards; s++) {
    snprintf(path, sizeof(path), "%s/segment%05d", dir, s);
    dbs[s] = p_open(path, 0, wal, err, sizeof(err));
    if (!dbs[s]) { fprintf(stderr, "open %s: %s\n", path, err); return 1; }
  }
  t_open = now_s() - t_open;
  pthread_barrier_init(&g_bar, NULL, threads);
  pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * threads);
  targ* ta = (targ*)calloc(threads, sizeof(targ));
  for (int t = 0; t < threads; t++) {
    ta[t].tid = t; ta[t].threads = threads; ta[t].shards = shards; ta[t].wal = wal; ta[t].compact = compact;
    ta[t].kv_total = kv_total; ta[t].vlen = vlen; ta[t].batch = batch; ta[t].apply_cap = apply_cap;
    ta[t].get_secs = get_secs; ta[t].dbs = dbs;
    pthread_create(&th[t], NULL, worker, &ta[t]);
  }
  uint64_t applied = 0, lookups = 0, hits = 0;
  double apply_s = 0, get_s = 0, compact_s = 0;
  for (int t = 0; t < threads; t++) {
    pthread_join(th[t], NULL);
    applied += ta[t].applied; lookups += ta[t].lookups; hits += ta[t].hits;
    if (ta[t].apply_s > apply_s) apply_s = ta[t].apply_s;
    if (ta[t].get_s > get_s) get_s = ta[t].get_s;
    if (ta[t].compact_s > compact_s) compact_s = ta[t].compact_s;
  }
  uint64_t seq_sum = 0;
  for (int s = 0; s < shards; s++) seq_sum += p_latest_seq(dbs[s]);
  printf("{\"kind\": \"%s\", \"threads\": %d, \"shards\": %d, \"applied\": %llu, \"apply_s\": %.4f, "
         "\"applies_per_s\": %.1f, \"compact_s\": %.3f, \"lookups\": %llu, \"hits\": %llu, \"get_s\": %.4f, "
         "\"lookups_per_s\": %.1f, \"seq_sum\": %llu, \"open_s\": %.2f, \"wal\": %d, \"vlen\": %u, \"batch\": %u}\n",
         p_kind(), threads, shards, (unsigned long long)applied, apply_s, apply_s > 0 ? applied / apply_s : 0.0,
         compact_s, (unsigned long long)lookups, (unsigned long long)hits, get_s, get_s > 0 ? lookups / get_s : 0.0,
         (unsigned long long)seq_sum, t_open, wal, vlen, batch);
  for (int s = 0; s < shards; s++) p_close(dbs[s]);
  return 0;
}
