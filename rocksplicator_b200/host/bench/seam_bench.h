// seam_bench.h — configuration / result records of rsp_seam_bench (host/bench/seam_bench.cpp): shared with bench.py
// (ctypes mirrors of these structs) and tests/cpp/host_tests.cpp.
#pragma once
#include <cstdint>

extern "C" {
typedef struct rsp_seam_cfg {
  int32_t device;
  uint32_t shards;
  uint64_t kv_total;           // loaded through the pull loop: kv_total / shards keys per shard
  uint32_t value_len;          // 64 (config 2) or 256 (config 5)
  uint32_t executor_threads;   // replicator executor threads (reference default 32, floor 16)
  uint32_t updates_per_response;  // replicator_max_updates_per_response (50)
  uint32_t update_rounds;      // mixed phase: this many more responses per shard while MultiGet runs
  uint32_t multiget_threads, multiget_batch;
  double multiget_secs;
  uint32_t get_threads;
  double get_secs;
  uint64_t seed;
  uint32_t first_shard_id;     // shard ids first .. first + shards - 1 (multi-rank runs)
  uint32_t steady_rounds;      // apply-only steady state: this many more responses per shard, nothing else running
} rsp_seam_cfg;

typedef struct rsp_seam_result {
  double load_s, load_applies_per_s;
  double resp_p50_ms, resp_p99_ms;        // response handed to the follower -> its next pull arrives (load phase)
  double compact_s;
  double mget_lookups_per_s, mget_p50_ms, mget_p99_ms;
  uint64_t mget_calls;
  double get_per_s, get_p50_us, get_p99_us;
  double mixed_applies_per_s, mixed_lookups_per_s, mixed_resp_p50_ms, mixed_resp_p99_ms;
  uint64_t applied_total, parity_errors, status_errors;
  uint64_t engine_launches;
  // steady state of the pull loops alone (shards already loaded, flushes running as memtables fill)
  double steady_applies_per_s, steady_resp_p50_ms, steady_resp_p99_ms;
  double trace_us[6];          // replicator::PullTrace: mean microseconds per stage of a pull round trip
  double apply_comb[9];        // apply combiner over the steady phase: batches, items, ms running, ms waiting for copiers, ms idle,
                               // ms (summed over responses) until the batch ran / until the callback started / inside callbacks, callbacks
  double read_comb[5];         // read combiner over the Get phase
  double cpu_s[4];             // CPU seconds of the process (getrusage: user + system) spent in the load, MultiGet, Get and
                               // steady phases — the GPU boxes cap the container's CPU time
} rsp_seam_result;

int rsp_seam_bench(const rsp_seam_cfg* cfg, rsp_seam_result* res);
}
