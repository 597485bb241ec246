/*
 * okv.h — the ORACLE interface (TEST INFRASTRUCTURE, not product code).
 *
 * Two implementations export exactly these symbols:
 *   oracle/kv_oracle.c   -> oracle/libokv_port.so   CPU restatement ("port") of the algorithm the
 *                                                    reference's hot path runs inside RocksDB.
 *   oracle/ref_driver.c  -> oracle/_ref/libokv_ref.so  thin driver over the reference's OWN RocksDB
 *                                                    binary (rocksdb_admin/tests/librocksdb.so.5.4).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load
 * either library.  The product (librsp_b200.so) never links, loads or calls anything in oracle/.
 *
 * Each entry point names the reference call site it stands for.
 */
#ifndef OKV_H_
#define OKV_H_
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* rocksdb::Status::Code values (RocksDB include/rocksdb/status.h; the codes the reference surfaces
 * through rocksdb::Status at application_db.cpp:85-136). */
enum {
  OKV_OK = 0,
  OKV_NOT_FOUND = 1,
  OKV_CORRUPTION = 2,
  OKV_NOT_SUPPORTED = 3,
  OKV_INVALID_ARGUMENT = 4,
  OKV_IO_ERROR = 5
};

/* merge operators on the path */
enum {
  OKV_MERGE_NONE = 0,
  OKV_MERGE_COUNTER = 1,   /* examples/counter_service/merge_operator.cpp:23-45 */
  OKV_MERGE_UINT64ADD = 2, /* RocksDB built-in "uint64add"  */
  OKV_MERGE_APPEND = 3     /* rocksdb_replicator/tests/rocksdb_assumption_test.cpp:58-77 */
};

typedef struct okv_db okv_db;
typedef struct okv_iter okv_iter;

/* which implementation this is: "port" or "reference" */
const char* okv_kind(void);

/* DB::Open (admin_handler.cpp:640).  `path` is a scratch directory for the reference driver (it is
 * created; the port ignores it).  wal != 0 keeps RocksDB's default WriteOptions (WAL on, no fsync) as
 * rocksdb_wrapper.cpp:31 does; wal == 0 is the labelled WAL-off variant. */
okv_db* okv_open(const char* path, int merge_op, int wal, char* err, size_t errcap);
void okv_close(okv_db* db);

/* RocksDbWrapper::HandleReplicateResponse (rocksdb_wrapper.cpp:13-28): bytes -> WriteBatch ->
 * PutLogData(&timestamp, 8) -> DB::Write.  Returns the Status code; message in err. */
int okv_apply(okv_db* db, const uint8_t* batch, size_t len, uint64_t ts_ms, char* err, size_t errcap);

/* RocksDbWrapper::LatestSequenceNumber (rocksdb_wrapper.cpp:4). */
uint64_t okv_latest_seq(okv_db* db);

/* ApplicationDB::Get (application_db.cpp:85-111).  On OKV_OK *val is malloc'ed (free with okv_free),
 * possibly 0-length. */
int okv_get(okv_db* db, const uint8_t* key, size_t klen, uint8_t** val, size_t* vlen, char* err,
            size_t errcap);

/* ApplicationDB::MultiGet (application_db.cpp:113-120).  keys are concatenated, koff has n+1 offsets.
 * Fills st[n]; values are appended to a malloc'ed buffer *vals with voff[n+1] offsets. */
int okv_multi_get(okv_db* db, size_t n, const uint8_t* keys, const uint64_t* koff, int32_t* st,
                  uint8_t** vals, uint64_t* voff);

/* ApplicationDB::NewIterator (application_db.cpp:78-83) and rocksdb::Iterator. */
okv_iter* okv_iter_create(okv_db* db);
void okv_iter_destroy(okv_iter* it);
void okv_iter_seek_to_first(okv_iter* it);
void okv_iter_seek_to_last(okv_iter* it);
void okv_iter_seek(okv_iter* it, const uint8_t* key, size_t klen);
void okv_iter_next(okv_iter* it);
void okv_iter_prev(okv_iter* it);
int okv_iter_valid(okv_iter* it);
const uint8_t* okv_iter_key(okv_iter* it, size_t* klen);
const uint8_t* okv_iter_value(okv_iter* it, size_t* vlen);
int okv_iter_status(okv_iter* it);

/* DB::Flush / ApplicationDB::CompactRange(nullptr, nullptr) (application_db.cpp:138-144). */
int okv_flush(okv_db* db);
int okv_compact(okv_db* db);

void okv_free(void* p);

#ifdef __cplusplus
}
#endif
#endif /* OKV_H_ */
