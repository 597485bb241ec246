/*
 * ref_driver.c — drives the reference's OWN RocksDB build through its exported C API.
 *
 * TEST INFRASTRUCTURE ONLY (see okv.h).  The reference's source tree cannot be compiled in this image
 * (folly / fbthrift / glog / gflags / RocksDB headers are absent, DESIGN.md §"oracle"), but it ships
 * the exact RocksDB binary its tests link: rocksdb_admin/tests/librocksdb.so.5.4.  oracle/build_ref.sh
 * places a stripped copy of that binary plus empty stub libraries for its missing sonames into
 * oracle/_ref/ next to this driver (outputs only; no reference SOURCE is copied).
 *
 * Every entry mirrors the reference call site named in okv.h, using the same RocksDB calls the
 * reference makes: apply = WriteBatch(bytes) + PutLogData(&ts, 8) + DB::Write(default WriteOptions)
 * (rocksdb_replicator/rocksdb_wrapper.cpp:13-31); reads = DB::Get / MultiGet / NewIterator
 * (rocksdb_admin/application_db.cpp:78-120).  DB options follow
 * examples/counter_service/rocksdb_options.cpp:61-103 (no compression, 4 KB blocks, 10-bit bloom,
 * LRU block cache, L0 trigger 4), with a smaller write buffer so that 1024 instances fit host RAM.
 */
#define _GNU_SOURCE
#include "okv.h"

#include <dlfcn.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>

const char* okv_kind(void) { return "reference"; }

typedef struct rocksdb_t rocksdb_t;
typedef struct rocksdb_options_t rocksdb_options_t;
typedef struct rocksdb_writeoptions_t rocksdb_writeoptions_t;
typedef struct rocksdb_readoptions_t rocksdb_readoptions_t;
typedef struct rocksdb_writebatch_t rocksdb_writebatch_t;
typedef struct rocksdb_iterator_t rocksdb_iterator_t;
typedef struct rocksdb_mergeoperator_t rocksdb_mergeoperator_t;
typedef struct rocksdb_flushoptions_t rocksdb_flushoptions_t;
typedef struct rocksdb_bbto_t rocksdb_bbto_t;
typedef struct rocksdb_cache_t rocksdb_cache_t;
typedef struct rocksdb_filterpolicy_t rocksdb_filterpolicy_t;

#define FN(ret, name, args) static ret(*p_##name) args
FN(rocksdb_options_t*, rocksdb_options_create, (void));
FN(void, rocksdb_options_destroy, (rocksdb_options_t*));
FN(void, rocksdb_options_set_create_if_missing, (rocksdb_options_t*, unsigned char));
FN(void, rocksdb_options_set_compression, (rocksdb_options_t*, int));
FN(void, rocksdb_options_set_write_buffer_size, (rocksdb_options_t*, size_t));
FN(void, rocksdb_options_set_max_write_buffer_number, (rocksdb_options_t*, int));
FN(void, rocksdb_options_set_min_write_buffer_number_to_merge, (rocksdb_options_t*, int));
FN(void, rocksdb_options_set_level0_file_num_compaction_trigger, (rocksdb_options_t*, int));
FN(void, rocksdb_options_set_max_bytes_for_level_base, (rocksdb_options_t*, uint64_t));
FN(void, rocksdb_options_set_max_open_files, (rocksdb_options_t*, int));
FN(void, rocksdb_options_set_keep_log_file_num, (rocksdb_options_t*, size_t));
FN(void, rocksdb_options_set_info_log_level, (rocksdb_options_t*, int));
FN(void, rocksdb_options_set_merge_operator, (rocksdb_options_t*, rocksdb_mergeoperator_t*));
FN(void, rocksdb_options_set_uint64add_merge_operator, (rocksdb_options_t*));
FN(void, rocksdb_options_set_block_based_table_factory, (rocksdb_options_t*, rocksdb_bbto_t*));
FN(rocksdb_bbto_t*, rocksdb_block_based_options_create, (void));
FN(void, rocksdb_block_based_options_destroy, (rocksdb_bbto_t*));
FN(void, rocksdb_block_based_options_set_block_size, (rocksdb_bbto_t*, size_t));
FN(void, rocksdb_block_based_options_set_filter_policy, (rocksdb_bbto_t*, rocksdb_filterpolicy_t*));
FN(void, rocksdb_block_based_options_set_block_cache, (rocksdb_bbto_t*, rocksdb_cache_t*));
FN(rocksdb_filterpolicy_t*, rocksdb_filterpolicy_create_bloom, (int));
FN(rocksdb_cache_t*, rocksdb_cache_create_lru, (size_t));
FN(void, rocksdb_cache_destroy, (rocksdb_cache_t*));
FN(rocksdb_mergeoperator_t*, rocksdb_mergeoperator_create,
   (void*, void (*)(void*),
    char* (*)(void*, const char*, size_t, const char*, size_t, const char* const*, const size_t*, int,
              unsigned char*, size_t*),
    char* (*)(void*, const char*, size_t, const char* const*, const size_t*, int, unsigned char*,
              size_t*),
    void (*)(void*, const char*, size_t), const char* (*)(void*)));
FN(rocksdb_t*, rocksdb_open, (const rocksdb_options_t*, const char*, char**));
FN(void, rocksdb_close, (rocksdb_t*));
FN(rocksdb_writeoptions_t*, rocksdb_writeoptions_create, (void));
FN(void, rocksdb_writeoptions_destroy, (rocksdb_writeoptions_t*));
FN(void, rocksdb_writeoptions_disable_WAL, (rocksdb_writeoptions_t*, int));
FN(rocksdb_readoptions_t*, rocksdb_readoptions_create, (void));
FN(void, rocksdb_readoptions_destroy, (rocksdb_readoptions_t*));
FN(rocksdb_writebatch_t*, rocksdb_writebatch_create_from, (const char*, size_t));
FN(void, rocksdb_writebatch_put_log_data, (rocksdb_writebatch_t*, const char*, size_t));
FN(void, rocksdb_writebatch_destroy, (rocksdb_writebatch_t*));
FN(void, rocksdb_write, (rocksdb_t*, const rocksdb_writeoptions_t*, rocksdb_writebatch_t*, char**));
FN(char*, rocksdb_get,
   (rocksdb_t*, const rocksdb_readoptions_t*, const char*, size_t, size_t*, char**));
FN(void, rocksdb_multi_get,
   (rocksdb_t*, const rocksdb_readoptions_t*, size_t, const char* const*, const size_t*, char**,
    size_t*, char**));
FN(void, rocksdb_free, (void*));
FN(rocksdb_iterator_t*, rocksdb_create_iterator, (rocksdb_t*, const rocksdb_readoptions_t*));
FN(void, rocksdb_iter_destroy, (rocksdb_iterator_t*));
FN(unsigned char, rocksdb_iter_valid, (const rocksdb_iterator_t*));
FN(void, rocksdb_iter_seek_to_first, (rocksdb_iterator_t*));
FN(void, rocksdb_iter_seek_to_last, (rocksdb_iterator_t*));
FN(void, rocksdb_iter_seek, (rocksdb_iterator_t*, const char*, size_t));
FN(void, rocksdb_iter_next, (rocksdb_iterator_t*));
FN(void, rocksdb_iter_prev, (rocksdb_iterator_t*));
FN(const char*, rocksdb_iter_key, (const rocksdb_iterator_t*, size_t*));
FN(const char*, rocksdb_iter_value, (const rocksdb_iterator_t*, size_t*));
FN(void, rocksdb_iter_get_error, (const rocksdb_iterator_t*, char**));
FN(rocksdb_flushoptions_t*, rocksdb_flushoptions_create, (void));
FN(void, rocksdb_flushoptions_destroy, (rocksdb_flushoptions_t*));
FN(void, rocksdb_flushoptions_set_wait, (rocksdb_flushoptions_t*, unsigned char));
FN(void, rocksdb_flush, (rocksdb_t*, const rocksdb_flushoptions_t*, char**));
FN(void, rocksdb_compact_range, (rocksdb_t*, const char*, size_t, const char*, size_t));
typedef struct rocksdb_envoptions_t rocksdb_envoptions_t;
typedef struct rocksdb_sstfilewriter_t rocksdb_sstfilewriter_t;
typedef struct rocksdb_ingestexternalfileoptions_t rocksdb_ingestexternalfileoptions_t;
FN(rocksdb_envoptions_t*, rocksdb_envoptions_create, (void));
FN(void, rocksdb_envoptions_destroy, (rocksdb_envoptions_t*));
FN(rocksdb_sstfilewriter_t*, rocksdb_sstfilewriter_create, (const rocksdb_envoptions_t*, const rocksdb_options_t*));
FN(void, rocksdb_sstfilewriter_open, (rocksdb_sstfilewriter_t*, const char*, char**));
FN(void, rocksdb_sstfilewriter_add, (rocksdb_sstfilewriter_t*, const char*, size_t, const char*, size_t, char**));
FN(void, rocksdb_sstfilewriter_finish, (rocksdb_sstfilewriter_t*, char**));
FN(void, rocksdb_sstfilewriter_destroy, (rocksdb_sstfilewriter_t*));
FN(rocksdb_ingestexternalfileoptions_t*, rocksdb_ingestexternalfileoptions_create, (void));
FN(void, rocksdb_ingestexternalfileoptions_destroy, (rocksdb_ingestexternalfileoptions_t*));
FN(void, rocksdb_ingestexternalfileoptions_set_move_files, (rocksdb_ingestexternalfileoptions_t*, unsigned char));
FN(void, rocksdb_ingestexternalfileoptions_set_allow_global_seqno, (rocksdb_ingestexternalfileoptions_t*, unsigned char));
FN(void, rocksdb_ingestexternalfileoptions_set_allow_blocking_flush, (rocksdb_ingestexternalfileoptions_t*, unsigned char));
FN(void, rocksdb_ingest_external_file, (rocksdb_t*, const char* const*, size_t, const rocksdb_ingestexternalfileoptions_t*, char**));
/* not in the C API: DBImpl::GetLatestSequenceNumber() const, called on *(DB**)rocksdb_t
 * (rocksdb_t is struct { DB* rep; }) — the call rocksdb_wrapper.cpp:4 makes */
static uint64_t (*p_latest_seq)(const void*);

static pthread_once_t g_once = PTHREAD_ONCE_INIT;
static int g_loaded = 0;
static char g_load_err[512];
static rocksdb_cache_t* g_cache; /* one shared block cache, as rocksdb_options.cpp:74-77 */

static void load_all(void) {
  Dl_info info;
  char dir[4096];
  if (!dladdr((void*)&okv_kind, &info) || !info.dli_fname) {
    snprintf(g_load_err, sizeof(g_load_err), "dladdr failed");
    return;
  }
  snprintf(dir, sizeof(dir), "%s", info.dli_fname);
  char* slash = strrchr(dir, '/');
  if (slash) *slash = 0; else snprintf(dir, sizeof(dir), ".");
  static const char* stubs[] = {"libsnappy.so.1", "libgflags.so.2", "libzstd.1.1.1024.so",
                                "libnuma.so.1", "libjemalloc.so.2", "libhdfs.so.0.0.0",
                                "libverify.so", "libjava.so", "libjvm.so"};
  char path[4400];
  for (size_t i = 0; i < sizeof(stubs) / sizeof(stubs[0]); i++) {
    snprintf(path, sizeof(path), "%s/%s", dir, stubs[i]);
    if (!dlopen(path, RTLD_NOW | RTLD_GLOBAL)) {
      snprintf(g_load_err, sizeof(g_load_err), "dlopen %s: %s", path, dlerror());
      return;
    }
  }
  snprintf(path, sizeof(path), "%s/librocksdb.so.5.4", dir);
  void* h = dlopen(path, RTLD_LAZY | RTLD_GLOBAL);
  if (!h) {
    snprintf(g_load_err, sizeof(g_load_err), "dlopen %s: %s", path, dlerror());
    return;
  }
#define LD(name)                                                                   \
  do {                                                                             \
    *(void**)(&p_##name) = dlsym(h, #name);                                        \
    if (!p_##name) {                                                               \
      snprintf(g_load_err, sizeof(g_load_err), "dlsym %s failed", #name);          \
      return;                                                                      \
    }                                                                              \
  } while (0)
  LD(rocksdb_options_create); LD(rocksdb_options_destroy); LD(rocksdb_options_set_create_if_missing);
  LD(rocksdb_options_set_compression); LD(rocksdb_options_set_write_buffer_size);
  LD(rocksdb_options_set_max_write_buffer_number);
  LD(rocksdb_options_set_min_write_buffer_number_to_merge);
  LD(rocksdb_options_set_level0_file_num_compaction_trigger);
  LD(rocksdb_options_set_max_bytes_for_level_base); LD(rocksdb_options_set_max_open_files);
  LD(rocksdb_options_set_keep_log_file_num); LD(rocksdb_options_set_info_log_level);
  LD(rocksdb_options_set_merge_operator); LD(rocksdb_options_set_uint64add_merge_operator);
  LD(rocksdb_options_set_block_based_table_factory); LD(rocksdb_block_based_options_create);
  LD(rocksdb_block_based_options_destroy); LD(rocksdb_block_based_options_set_block_size);
  LD(rocksdb_block_based_options_set_filter_policy); LD(rocksdb_block_based_options_set_block_cache);
  LD(rocksdb_filterpolicy_create_bloom); LD(rocksdb_cache_create_lru); LD(rocksdb_cache_destroy);
  LD(rocksdb_mergeoperator_create); LD(rocksdb_open); LD(rocksdb_close);
  LD(rocksdb_writeoptions_create); LD(rocksdb_writeoptions_destroy); LD(rocksdb_writeoptions_disable_WAL);
  LD(rocksdb_readoptions_create); LD(rocksdb_readoptions_destroy); LD(rocksdb_writebatch_create_from);
  LD(rocksdb_writebatch_put_log_data); LD(rocksdb_writebatch_destroy); LD(rocksdb_write);
  LD(rocksdb_get); LD(rocksdb_multi_get); LD(rocksdb_free); LD(rocksdb_create_iterator);
  LD(rocksdb_iter_destroy); LD(rocksdb_iter_valid); LD(rocksdb_iter_seek_to_first);
  LD(rocksdb_iter_seek_to_last); LD(rocksdb_iter_seek); LD(rocksdb_iter_next); LD(rocksdb_iter_prev);
  LD(rocksdb_iter_key); LD(rocksdb_iter_value); LD(rocksdb_iter_get_error);
  LD(rocksdb_flushoptions_create); LD(rocksdb_flushoptions_destroy); LD(rocksdb_flushoptions_set_wait);
  LD(rocksdb_flush); LD(rocksdb_compact_range);
  LD(rocksdb_envoptions_create); LD(rocksdb_envoptions_destroy); LD(rocksdb_sstfilewriter_create);
  LD(rocksdb_sstfilewriter_open); LD(rocksdb_sstfilewriter_add); LD(rocksdb_sstfilewriter_finish);
  LD(rocksdb_sstfilewriter_destroy); LD(rocksdb_ingestexternalfileoptions_create);
  LD(rocksdb_ingestexternalfileoptions_destroy); LD(rocksdb_ingestexternalfileoptions_set_move_files);
  LD(rocksdb_ingestexternalfileoptions_set_allow_global_seqno);
  LD(rocksdb_ingestexternalfileoptions_set_allow_blocking_flush); LD(rocksdb_ingest_external_file);
  *(void**)(&p_latest_seq) = dlsym(h, "_ZNK7rocksdb6DBImpl23GetLatestSequenceNumberEv");
  if (!p_latest_seq) {
    snprintf(g_load_err, sizeof(g_load_err), "dlsym DBImpl::GetLatestSequenceNumber failed");
    return;
  }
  g_cache = p_rocksdb_cache_create_lru((size_t)1 << 30);
  g_loaded = 1;
}

struct okv_db {
  rocksdb_t* db;
  rocksdb_options_t* opts;
  rocksdb_bbto_t* bbto;
  rocksdb_writeoptions_t* wo;
  rocksdb_readoptions_t* ro;
};

static int code_of(const char* e) {
  if (!e) return OKV_OK;
  if (!strncmp(e, "NotFound", 8)) return OKV_NOT_FOUND;
  if (!strncmp(e, "Corruption", 10)) return OKV_CORRUPTION;
  if (!strncmp(e, "Not implemented", 15)) return OKV_NOT_SUPPORTED;
  if (!strncmp(e, "Invalid argument", 16)) return OKV_INVALID_ARGUMENT;
  return OKV_IO_ERROR;
}
static int take_err(char* e, char* err, size_t cap) {
  int c = code_of(e);
  if (e) {
    if (err && cap) snprintf(err, cap, "%s", e);
    p_rocksdb_free(e);
  }
  return c;
}

/* ---- merge operators as C-API callbacks (AssociativeMergeOperator semantics: FullMerge folds the
 * operands oldest -> newest starting from the existing value; PartialMerge(l, r) = Merge(&l, r)) ---- */
static int counter_merge(int has, const char* ex, size_t exl, const char* v, size_t vl, char** out,
                         size_t* outl) {
  /* examples/counter_service/merge_operator.cpp:23-45 */
  if (!has) {
    *out = (char*)malloc(vl ? vl : 1);
    memcpy(*out, v, vl);
    *outl = vl;
    return 1;
  }
  if (exl != 8 || vl != 8) return 0;
  int64_t a, b;
  memcpy(&a, ex, 8);
  memcpy(&b, v, 8);
  b = (int64_t)((uint64_t)a + (uint64_t)b);
  *out = (char*)malloc(8);
  memcpy(*out, &b, 8);
  *outl = 8;
  return 1;
}
static int append_merge(int has, const char* ex, size_t exl, const char* v, size_t vl, char** out,
                        size_t* outl) {
  /* rocksdb_replicator/tests/rocksdb_assumption_test.cpp:58-77 */
  size_t n = (has ? exl : 0) + vl;
  *out = (char*)malloc(n ? n : 1);
  if (has) memcpy(*out, ex, exl);
  memcpy(*out + (has ? exl : 0), v, vl);
  *outl = n;
  return 1;
}
typedef int (*merge_fn)(int, const char*, size_t, const char*, size_t, char**, size_t*);

static char* fold(merge_fn fn, int has, const char* ex, size_t exl, const char* const* ops,
                  const size_t* opl, int n, int first, unsigned char* success, size_t* outl) {
  char* cur = NULL;
  size_t curl = 0;
  if (has) {
    cur = (char*)malloc(exl ? exl : 1);
    memcpy(cur, ex, exl);
    curl = exl;
  }
  for (int i = first; i < n; i++) {
    char* nv = NULL;
    size_t nl = 0;
    if (!fn(has, cur, curl, ops[i], opl[i], &nv, &nl)) {
      free(cur);
      *success = 0;
      *outl = 0;
      return NULL;
    }
    free(cur);
    cur = nv;
    curl = nl;
    has = 1;
  }
  *success = 1;
  *outl = curl;
  return cur;
}
static char* mo_full(void* st, const char* k, size_t kl, const char* ex, size_t exl,
                     const char* const* ops, const size_t* opl, int n, unsigned char* success,
                     size_t* outl) {
  (void)k; (void)kl;
  return fold((merge_fn)st, ex != NULL, ex, exl, ops, opl, n, 0, success, outl);
}
static char* mo_partial(void* st, const char* k, size_t kl, const char* const* ops,
                        const size_t* opl, int n, unsigned char* success, size_t* outl) {
  (void)k; (void)kl;
  if (n < 1) {
    *success = 0;
    return NULL;
  }
  return fold((merge_fn)st, 1, ops[0], opl[0], ops, opl, n, 1, success, outl);
}
static void mo_delete_value(void* st, const char* v, size_t vl) {
  (void)st; (void)vl;
  free((void*)v);
}
static void mo_destroy(void* st) { (void)st; }
static const char* mo_name_counter(void* st) { (void)st; return "CounterMergeOperator"; }
static const char* mo_name_append(void* st) { (void)st; return "SimpleMergeOperator"; }

okv_db* okv_open(const char* path, int merge_op, int wal, char* err, size_t errcap) {
  pthread_once(&g_once, load_all);
  if (!g_loaded) {
    if (err && errcap) snprintf(err, errcap, "%s", g_load_err);
    return NULL;
  }
  okv_db* d = (okv_db*)calloc(1, sizeof(okv_db));
  d->opts = p_rocksdb_options_create();
  p_rocksdb_options_set_create_if_missing(d->opts, 1);
  p_rocksdb_options_set_compression(d->opts, 0); /* kNoCompression (rocksdb_options.cpp:96) */
  d->bbto = p_rocksdb_block_based_options_create();
  p_rocksdb_block_based_options_set_block_size(d->bbto, 4096);
  p_rocksdb_block_based_options_set_filter_policy(d->bbto, p_rocksdb_filterpolicy_create_bloom(10));
  p_rocksdb_block_based_options_set_block_cache(d->bbto, g_cache);
  p_rocksdb_options_set_block_based_table_factory(d->opts, d->bbto);
  const char* wb = getenv("OKV_REF_WRITE_BUFFER_MB");
  size_t wbs = (size_t)(wb ? atoi(wb) : 8) << 20;
  p_rocksdb_options_set_write_buffer_size(d->opts, wbs);
  p_rocksdb_options_set_min_write_buffer_number_to_merge(d->opts, 1);
  p_rocksdb_options_set_level0_file_num_compaction_trigger(d->opts, 4);
  p_rocksdb_options_set_max_bytes_for_level_base(d->opts, (uint64_t)wbs * 4);
  p_rocksdb_options_set_max_open_files(d->opts, -1);
  p_rocksdb_options_set_keep_log_file_num(d->opts, 1);
  p_rocksdb_options_set_info_log_level(d->opts, 3 /* ERROR */);
  switch (merge_op) {
    case OKV_MERGE_COUNTER:
      p_rocksdb_options_set_merge_operator(
          d->opts, p_rocksdb_mergeoperator_create((void*)counter_merge, mo_destroy, mo_full, mo_partial,
                                                  mo_delete_value, mo_name_counter));
      break;
    case OKV_MERGE_UINT64ADD:
      p_rocksdb_options_set_uint64add_merge_operator(d->opts);
      break;
    case OKV_MERGE_APPEND:
      p_rocksdb_options_set_merge_operator(
          d->opts, p_rocksdb_mergeoperator_create((void*)append_merge, mo_destroy, mo_full, mo_partial,
                                                  mo_delete_value, mo_name_append));
      break;
    default:
      break;
  }
  mkdir(path, 0755);
  char* e = NULL;
  d->db = p_rocksdb_open(d->opts, path, &e);
  if (e || !d->db) {
    take_err(e, err, errcap);
    p_rocksdb_options_destroy(d->opts);
    free(d);
    return NULL;
  }
  d->wo = p_rocksdb_writeoptions_create(); /* defaults: WAL on, sync off (rocksdb_wrapper.cpp:31) */
  if (!wal) p_rocksdb_writeoptions_disable_WAL(d->wo, 1);
  d->ro = p_rocksdb_readoptions_create();
  return d;
}

void okv_close(okv_db* d) {
  if (!d) return;
  p_rocksdb_close(d->db);
  p_rocksdb_writeoptions_destroy(d->wo);
  p_rocksdb_readoptions_destroy(d->ro);
  p_rocksdb_options_destroy(d->opts);
  free(d);
}

void okv_free(void* p) { free(p); }

int okv_apply(okv_db* d, const uint8_t* batch, size_t len, uint64_t ts_ms, char* err, size_t errcap) {
  /* rocksdb_wrapper.cpp:13-28 */
  rocksdb_writebatch_t* wb = p_rocksdb_writebatch_create_from((const char*)batch, len);
  p_rocksdb_writebatch_put_log_data(wb, (const char*)&ts_ms, sizeof(ts_ms));
  char* e = NULL;
  p_rocksdb_write(d->db, d->wo, wb, &e);
  p_rocksdb_writebatch_destroy(wb);
  return take_err(e, err, errcap);
}

uint64_t okv_latest_seq(okv_db* d) { return p_latest_seq(*(void**)d->db); }

int okv_get(okv_db* d, const uint8_t* key, size_t klen, uint8_t** val, size_t* vlen, char* err,
            size_t errcap) {
  char* e = NULL;
  size_t n = 0;
  char* v = p_rocksdb_get(d->db, d->ro, (const char*)key, klen, &n, &e);
  *val = NULL;
  *vlen = 0;
  if (e) return take_err(e, err, errcap);
  if (!v) return OKV_NOT_FOUND;
  *val = (uint8_t*)malloc(n ? n : 1);
  memcpy(*val, v, n);
  *vlen = n;
  p_rocksdb_free(v);
  return OKV_OK;
}

int okv_multi_get(okv_db* d, size_t n, const uint8_t* keys, const uint64_t* koff, int32_t* st,
                  uint8_t** vals, uint64_t* voff) {
  const char** kp = (const char**)malloc(sizeof(char*) * (n ? n : 1));
  size_t* kl = (size_t*)malloc(sizeof(size_t) * (n ? n : 1));
  char** vp = (char**)calloc(n ? n : 1, sizeof(char*));
  size_t* vl = (size_t*)calloc(n ? n : 1, sizeof(size_t));
  char** ep = (char**)calloc(n ? n : 1, sizeof(char*));
  for (size_t i = 0; i < n; i++) {
    kp[i] = (const char*)keys + koff[i];
    kl[i] = (size_t)(koff[i + 1] - koff[i]);
  }
  p_rocksdb_multi_get(d->db, d->ro, n, kp, kl, vp, vl, ep);
  size_t tot = 0;
  for (size_t i = 0; i < n; i++) tot += vp[i] ? vl[i] : 0;
  uint8_t* out = (uint8_t*)malloc(tot ? tot : 1);
  size_t at = 0;
  for (size_t i = 0; i < n; i++) {
    voff[i] = at;
    if (ep[i]) {
      st[i] = code_of(ep[i]);
      p_rocksdb_free(ep[i]);
    } else if (!vp[i]) {
      st[i] = OKV_NOT_FOUND;
    } else {
      st[i] = OKV_OK;
      memcpy(out + at, vp[i], vl[i]);
      at += vl[i];
    }
    if (vp[i]) p_rocksdb_free(vp[i]);
  }
  voff[n] = at;
  *vals = out;
  free(kp); free(kl); free(vp); free(vl); free(ep);
  return OKV_OK;
}

struct okv_iter {
  rocksdb_iterator_t* it;
};
okv_iter* okv_iter_create(okv_db* d) {
  okv_iter* it = (okv_iter*)calloc(1, sizeof(okv_iter));
  it->it = p_rocksdb_create_iterator(d->db, d->ro);
  return it;
}
void okv_iter_destroy(okv_iter* it) {
  if (!it) return;
  p_rocksdb_iter_destroy(it->it);
  free(it);
}
void okv_iter_seek_to_first(okv_iter* it) { p_rocksdb_iter_seek_to_first(it->it); }
void okv_iter_seek_to_last(okv_iter* it) { p_rocksdb_iter_seek_to_last(it->it); }
void okv_iter_seek(okv_iter* it, const uint8_t* k, size_t kl) { p_rocksdb_iter_seek(it->it, (const char*)k, kl); }
void okv_iter_next(okv_iter* it) { p_rocksdb_iter_next(it->it); }
void okv_iter_prev(okv_iter* it) { p_rocksdb_iter_prev(it->it); }
int okv_iter_valid(okv_iter* it) { return p_rocksdb_iter_valid(it->it) ? 1 : 0; }
const uint8_t* okv_iter_key(okv_iter* it, size_t* kl) { return (const uint8_t*)p_rocksdb_iter_key(it->it, kl); }
const uint8_t* okv_iter_value(okv_iter* it, size_t* vl) { return (const uint8_t*)p_rocksdb_iter_value(it->it, vl); }
int okv_iter_status(okv_iter* it) {
  char* e = NULL;
  p_rocksdb_iter_get_error(it->it, &e);
  return take_err(e, NULL, 0);
}

int okv_flush(okv_db* d) {
  rocksdb_flushoptions_t* fo = p_rocksdb_flushoptions_create();
  p_rocksdb_flushoptions_set_wait(fo, 1);
  char* e = NULL;
  p_rocksdb_flush(d->db, fo, &e);
  p_rocksdb_flushoptions_destroy(fo);
  return take_err(e, NULL, 0);
}
int okv_compact(okv_db* d) {
  p_rocksdb_compact_range(d->db, NULL, 0, NULL, 0); /* CompactRange(nullptr, nullptr) */
  return OKV_OK;
}

/* ---- SST interchange (reference-only extras used by tests/test_sst_cpu.py) --------------------------------------
 * okv_write_sst: RocksDB's own SstFileWriter (rocksdb_assumption_test.cpp:209-243) -> an external SST file.
 * okv_ingest_sst: DB::IngestExternalFile (rocksdb_admin/tests/sst_binary.cpp:64-72, admin_handler.cpp:1820-1845). */
int okv_write_sst(const char* path, size_t n, const uint8_t* keys, const uint64_t* koff, const uint8_t* vals,
                  const uint64_t* voff, char* err, size_t errcap) {
  pthread_once(&g_once, load_all);
  if (!g_loaded) { if (err && errcap) snprintf(err, errcap, "%s", g_load_err); return OKV_IO_ERROR; }
  rocksdb_options_t* o = p_rocksdb_options_create();
  p_rocksdb_options_set_compression(o, 0);
  rocksdb_envoptions_t* eo = p_rocksdb_envoptions_create();
  rocksdb_sstfilewriter_t* w = p_rocksdb_sstfilewriter_create(eo, o);
  char* e = NULL;
  p_rocksdb_sstfilewriter_open(w, path, &e);
  for (size_t i = 0; i < n && !e; i++)
    p_rocksdb_sstfilewriter_add(w, (const char*)keys + koff[i], (size_t)(koff[i + 1] - koff[i]), (const char*)vals + voff[i],
                                (size_t)(voff[i + 1] - voff[i]), &e);
  if (!e) p_rocksdb_sstfilewriter_finish(w, &e);
  p_rocksdb_sstfilewriter_destroy(w);
  p_rocksdb_envoptions_destroy(eo);
  p_rocksdb_options_destroy(o);
  return take_err(e, err, errcap);
}

int okv_ingest_sst(okv_db* d, const char* path, int allow_global_seqno, char* err, size_t errcap) {
  rocksdb_ingestexternalfileoptions_t* io = p_rocksdb_ingestexternalfileoptions_create();
  p_rocksdb_ingestexternalfileoptions_set_move_files(io, 0);
  p_rocksdb_ingestexternalfileoptions_set_allow_global_seqno(io, allow_global_seqno ? 1 : 0);
  p_rocksdb_ingestexternalfileoptions_set_allow_blocking_flush(io, allow_global_seqno ? 1 : 0);
  const char* files[1] = {path};
  char* e = NULL;
  p_rocksdb_ingest_external_file(d->db, files, 1, io, &e);
  p_rocksdb_ingestexternalfileoptions_destroy(io);
  return take_err(e, err, errcap);
}
