// seam_bench.cpp — the hot path driven THROUGH the reference's own seams, for bench.py's "seams" figures and for
// tests/cpp/host_tests.cpp:
//
//   applies : RocksDBReplicator::addDB(FOLLOWER) -> ReplicatedDB::pullFromUpstream -> DbWrapper -> engine
//             (rocksdb_replicator/replicated_db.cpp:314-433), one pull loop per shard on the replicator's executor
//             threads; the upstream is a synthetic leader behind the Transport interface that answers every pull with
//             the next <= max_updates single-Put WriteBatches of that shard (what a leader's handleReplicateRequest
//             returns, replicated_db.cpp:435-575), so the number measured is the FOLLOWER side alone.
//   reads   : admin::ApplicationDB::MultiGet(keys of ONE shard) and ApplicationDB::Get from many caller threads
//             (rocksdb_admin/application_db.cpp:85-120; the reference's callers are up to 256 thrift workers).
//
// Data are the synthetic shards of SURVEY.md §8(d) (the same generator as rocksplicator_b200/synth.py), so every value
// read back is checked against a pure function of (shard, key index, version).  No oracle is involved.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <memory>
#include <mutex>
#include <random>
#include <string>
#include <thread>
#include <vector>

#include <sys/resource.h>

#include "common/segment_utils.h"
#include "gpu_db.h"
#include "rocksdb_admin/application_db.h"
#include "rocksdb_replicator/rocksdb_replicator.h"

#include "bench/seam_bench.h"

namespace {

using Clock = std::chrono::steady_clock;
inline double secs_since(Clock::time_point t0) { return std::chrono::duration<double>(Clock::now() - t0).count(); }

// ---- rocksplicator_b200/synth.py, restated ----------------------------------------------------------
inline uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  uint64_t z = x;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
inline void put_be64(uint8_t* p, uint64_t v) { for (int i = 0; i < 8; i++) p[i] = (uint8_t)(v >> (56 - 8 * i)); }
inline void key16(uint64_t seed, uint64_t idx, uint8_t* out) {
  put_be64(out, idx);
  put_be64(out + 8, splitmix64(seed ^ idx));
}
inline void value_bytes(uint64_t seed, uint64_t shard, uint64_t idx, uint64_t version, uint32_t vlen, uint8_t* out) {
  uint64_t s = splitmix64(seed ^ (shard << 40) ^ (idx << 8) ^ version);
  for (uint32_t at = 0; at < vlen; at += 8) {
    s = splitmix64(s);
    memcpy(out + at, &s, std::min<uint32_t>(8, vlen - at));  // little-endian words
  }
}
inline size_t put_varint32(uint8_t* p, uint32_t v) {
  size_t n = 0;
  while (v >= 128) { p[n++] = (uint8_t)((v & 127) | 128); v >>= 7; }
  p[n++] = (uint8_t)v;
  return n;
}
// header(seq = 0, count = 1) Put(key, value) LogData(ts): what the leader serves (replicated_db.cpp:115-117, 527-530).
// One allocation, written in place: the synthetic leader shares the follower's CPU budget here (in production it is
// another machine), so it is kept as cheap as the generator allows.
inline void single_put_batch(const uint8_t* key, const uint8_t* val, uint32_t vlen, uint64_t ts, std::string* out) {
  uint8_t vl[5];
  const size_t nvl = put_varint32(vl, vlen);
  out->resize(12 + 2 + 16 + nvl + vlen + 10);
  uint8_t* p = reinterpret_cast<uint8_t*>(&(*out)[0]);
  memset(p, 0, 12);
  p[8] = 1;
  p[12] = 0x1;
  p[13] = 16;
  memcpy(p + 14, key, 16);
  memcpy(p + 30, vl, nvl);
  memcpy(p + 30 + nvl, val, vlen);
  uint8_t* q = p + 30 + nvl + vlen;
  q[0] = 0x3;
  q[1] = 8;
  memcpy(q + 2, &ts, 8);
}

double cpu_seconds() {
  struct rusage ru;
  getrusage(RUSAGE_SELF, &ru);
  return ru.ru_utime.tv_sec + ru.ru_stime.tv_sec + 1e-6 * (ru.ru_utime.tv_usec + ru.ru_stime.tv_usec);
}

uint64_t now_ms() {
  return (uint64_t)std::chrono::duration_cast<std::chrono::milliseconds>(std::chrono::system_clock::now().time_since_epoch()).count();
}

struct Percentiles {
  std::mutex mu;
  std::vector<float> v;
  void add(const std::vector<float>& more) { std::lock_guard<std::mutex> g(mu); v.insert(v.end(), more.begin(), more.end()); }
  void add1(float x) { std::lock_guard<std::mutex> g(mu); v.push_back(x); }
  void clear() { std::lock_guard<std::mutex> g(mu); v.clear(); }
  double pct(double p) {
    std::lock_guard<std::mutex> g(mu);
    if (v.empty()) return 0;
    std::sort(v.begin(), v.end());
    return v[std::min(v.size() - 1, (size_t)(p * v.size()))];
  }
};

// The upstream of every shard: answers a pull at sequence number q with updates q .. min(q + max_updates, target).
// Update j of a shard writes key ordinal (j % per_shard) with version (j / per_shard): the first per_shard updates
// load the shard, later ones rewrite its keys in order.  A pull that has reached the target is parked (the leader's
// long-poll, replicated_db.cpp:467-575) until the target moves.
class SyntheticLeader : public replicator::Transport {
 public:
  SyntheticLeader(const rsp_seam_cfg& cfg, size_t threads) : cfg_(cfg), per_shard_(cfg.kv_total / cfg.shards), pool_(threads), sh_(cfg.shards) {}
  void Stop() { stopped_ = true; pool_.Stop(); }
  void replicate(const replicator::SocketAddress&, const replicator::ReplicateRequest& req, uint32_t, replicator::ReplicateCallback cb) override {
    if (stopped_) return;
    auto r = std::make_shared<replicator::ReplicateRequest>(req);
    auto c = std::make_shared<replicator::ReplicateCallback>(std::move(cb));
    pool_.add([this, r, c] { Serve(r, c); });
  }
  void SetTargets(uint64_t target) {
    for (size_t i = 0; i < sh_.size(); i++) {
      std::shared_ptr<replicator::ReplicateRequest> r;
      std::shared_ptr<replicator::ReplicateCallback> c;
      {
        std::lock_guard<std::mutex> g(sh_[i].mu);
        sh_[i].target = target;
        r.swap(sh_[i].parked_req);
        c.swap(sh_[i].parked_cb);
        sh_[i].last_resp = Clock::time_point();
      }
      if (r) pool_.add([this, r, c] { Serve(r, c); });
    }
  }
  Percentiles lat_ms;

 private:
  struct Shard {
    std::mutex mu;
    uint64_t target = 0;
    std::shared_ptr<replicator::ReplicateRequest> parked_req;
    std::shared_ptr<replicator::ReplicateCallback> parked_cb;
    Clock::time_point last_resp;
  };
  void Serve(std::shared_ptr<replicator::ReplicateRequest> r, std::shared_ptr<replicator::ReplicateCallback> c) {
    const int id = common::ExtractShardId(r->db_name);
    const size_t si = (size_t)(id - (int)cfg_.first_shard_id);
    if (id < 0 || si >= sh_.size()) {
      replicator::ReplicateResult out;
      out.is_replicate_exception = true;
      out.ex.code = replicator::ErrorCode::SOURCE_NOT_FOUND;
      (*c)(std::move(out));
      return;
    }
    Shard& s = sh_[si];
    const uint64_t q = (uint64_t)r->seq_no;
    uint64_t target;
    {
      std::lock_guard<std::mutex> g(s.mu);
      if (s.last_resp != Clock::time_point()) {  // (sampled: every 4th response of a shard is stamped)
        lat_ms.add1((float)(1e3 * std::chrono::duration<double>(Clock::now() - s.last_resp).count()));
        s.last_resp = Clock::time_point();
      }
      target = s.target;
      if (q >= target) {  // nothing new: park the long-poll
        s.parked_req = r;
        s.parked_cb = c;
        return;
      }
    }
    replicator::ReplicateResult out;
    out.ok = true;
    out.response.set_role(replicator::ReplicaRole::LEADER);
    const uint64_t n = std::min<uint64_t>(target - q, (uint64_t)std::max(1, r->max_updates));
    out.response.updates.resize(n);
    uint8_t key[16];
    std::vector<uint8_t> val(cfg_.value_len);
    const uint64_t ts = now_ms();
    for (uint64_t k = 0; k < n; k++) {
      const uint64_t j = q + k, ordinal = j % per_shard_, version = j / per_shard_;
      const uint64_t idx = (uint64_t)id + ordinal * (uint64_t)total_shards();
      key16(cfg_.seed, idx, key);
      value_bytes(cfg_.seed, (uint64_t)id, idx, version, cfg_.value_len, val.data());
      replicator::Update& u = out.response.updates[k];
      single_put_batch(key, val.data(), cfg_.value_len, ts, &u.raw_data);
      u.timestamp = (int64_t)ts;
      u.set_seq_no(j + 1);
    }
    if ((q / std::max<uint64_t>(1, n)) % 4 == 0) {
      std::lock_guard<std::mutex> g(s.mu);
      s.last_resp = Clock::now();
    }
    (*c)(std::move(out));
  }
  uint32_t total_shards() const { return stride_ ? stride_ : cfg_.shards; }

 public:
  uint32_t stride_ = 0;  // key index = shard id + ordinal * stride (the global shard count of a multi-rank run)

 private:
  const rsp_seam_cfg cfg_;
  const uint64_t per_shard_;
  replicator::Executor pool_;
  std::vector<Shard> sh_;
  std::atomic<bool> stopped_{false};
};

}  // namespace

extern "C" int rsp_seam_bench(const rsp_seam_cfg* cfg_in, rsp_seam_result* res) {
  if (!cfg_in || !res || !cfg_in->shards || cfg_in->kv_total < cfg_in->shards) return 4;
  rsp_seam_cfg cfg = *cfg_in;
  memset(res, 0, sizeof(*res));
  if (!cfg.value_len) cfg.value_len = 64;
  if (!cfg.updates_per_response) cfg.updates_per_response = 50;
  if (!cfg.multiget_batch) cfg.multiget_batch = 4096;
  if (!cfg.seed) cfg.seed = 0x5EED0001;
  const uint32_t S = cfg.shards;
  const uint64_t per_shard = cfg.kv_total / S;
  auto& F = replicator::Flags();
  const auto saved_flags = F;
  F.rocksdb_replicator_executor_threads = (int32_t)std::max<uint32_t>(16, cfg.executor_threads);
  F.replicator_max_updates_per_response = (int32_t)cfg.updates_per_response;
  auto leader = std::make_shared<SyntheticLeader>(cfg, std::max<size_t>(16, cfg.executor_threads));  // the remote leaders' CPUs
  leader->stride_ = S;  // (single-rank key space: index = shard id + ordinal * S with ids offset by first_shard_id)
  std::atomic<uint64_t> parity_errors{0}, status_errors{0};
  {
    replicator::RocksDBReplicator repl((uint16_t)(19000 + (cfg.first_shard_id % 1000)), leader);
    std::vector<std::shared_ptr<rocksdb::DB>> dbs(S);
    std::vector<std::unique_ptr<admin::ApplicationDB>> adbs(S);
    rocksdb::Options opt;
    opt.write_buffer_size = 2u << 20;
    for (uint32_t i = 0; i < S; i++) {
      rocksdb::DB* raw = nullptr;
      const std::string name = common::SegmentToDbName("seam", (int)(cfg.first_shard_id + i));
      if (!b200::GpuDB::Open(opt, name, &raw, cfg.device).ok()) { F = saved_flags; leader->Stop(); return 5; }
      dbs[i].reset(raw);
    }
    auto* gdb0 = static_cast<b200::GpuDB*>(dbs[0].get());
    const uint64_t launches0 = rsp_kernel_launches(gdb0->engine());
    auto wait_seq = [&](uint64_t target, double timeout_s) {
      const auto t0 = Clock::now();
      for (;;) {
        bool all = true;
        for (uint32_t i = 0; i < S && all; i++) all = dbs[i]->GetLatestSequenceNumber() >= target;
        if (all) return true;
        if (secs_since(t0) > timeout_s) return false;
        std::this_thread::sleep_for(std::chrono::microseconds(200));
      }
    };
    // ---- load: every shard pulls its keys from the synthetic leader --------------------------------
    leader->SetTargets(per_shard);
    auto t0 = Clock::now();
    double cpu0 = cpu_seconds();
    for (uint32_t i = 0; i < S; i++)
      adbs[i].reset(new admin::ApplicationDB(common::SegmentToDbName("seam", (int)(cfg.first_shard_id + i)), dbs[i],
                                             replicator::ReplicaRole::FOLLOWER,
                                             std::make_unique<replicator::SocketAddress>("127.0.0.1", 1), &repl));
    if (!wait_seq(per_shard, 180)) status_errors++;
    res->load_s = secs_since(t0);
    res->cpu_s[0] = cpu_seconds() - cpu0;
    res->load_applies_per_s = (double)per_shard * S / res->load_s;
    res->resp_p50_ms = leader->lat_ms.pct(0.5);
    res->resp_p99_ms = leader->lat_ms.pct(0.99);
    leader->lat_ms.clear();
    // ---- ApplicationDB::CompactRange on every shard (admin_handler.cpp:1846) ------------------------
    t0 = Clock::now();
    {
      std::vector<std::thread> th;
      std::atomic<uint32_t> next{0};
      for (int t = 0; t < 8; t++) th.emplace_back([&] {
        for (uint32_t i; (i = next++) < S;)
          if (!adbs[i]->CompactRange(rocksdb::CompactRangeOptions(), nullptr, nullptr).ok()) status_errors++;
      });
      for (auto& t : th) t.join();
    }
    res->compact_s = secs_since(t0);

    // expected version of a key ordinal once `applied` updates per shard have run
    auto reader = [&](uint32_t threads, uint32_t batch, double secs, std::atomic<uint64_t>* applied_now, bool racing,
                      Percentiles* lat, double* per_s, uint64_t* calls_out, bool single_get) {
      std::atomic<uint64_t> total{0}, calls{0};
      std::vector<std::thread> th;
      const auto start = Clock::now();
      for (uint32_t t = 0; t < threads; t++) th.emplace_back([&, t] {
        std::mt19937_64 rng(cfg.seed * 7919 + t);
        const uint32_t POOL = single_get ? 1 : 4;
        // a small pool of pre-built key batches per thread (key generation is not what is measured)
        struct KB { uint32_t shard; std::vector<uint64_t> ord; std::vector<uint8_t> bytes; std::vector<rocksdb::Slice> keys; };
        std::vector<KB> pool(POOL);
        const uint32_t bn = single_get ? 4096 : batch;
        for (auto& kb : pool) {
          kb.shard = (uint32_t)(rng() % S);
          kb.ord.resize(bn);
          kb.bytes.resize((size_t)bn * 16);
          kb.keys.resize(bn);
          for (uint32_t k = 0; k < bn; k++) {
            kb.ord[k] = rng() % per_shard;
            key16(cfg.seed, (uint64_t)(cfg.first_shard_id + kb.shard) + kb.ord[k] * S, &kb.bytes[(size_t)k * 16]);
            kb.keys[k] = rocksdb::Slice((const char*)&kb.bytes[(size_t)k * 16], 16);
          }
        }
        std::vector<float> my_lat;
        std::vector<std::string> values;
        std::string one;
        std::vector<uint8_t> want(cfg.value_len);
        uint64_t n_calls = 0, n_keys = 0, bad = 0, bad_st = 0;
        auto check = [&](const KB& kb, uint32_t k, const std::string& got) {
          const uint64_t id = cfg.first_shard_id + kb.shard, idx = id + kb.ord[k] * S;
          // versions: ordinal o has been rewritten (applied - 1 - o) / per_shard + ... : version v is current once update
          // j = v * per_shard + o has run
          const uint64_t applied = applied_now->load(std::memory_order_acquire);
          const uint64_t vmax = applied > kb.ord[k] ? (applied - 1 - kb.ord[k]) / per_shard : 0;
          for (uint64_t v = vmax;; v--) {  // while updates race with reads either neighbour version is right
            value_bytes(cfg.seed, id, idx, v, cfg.value_len, want.data());
            if (got.size() == cfg.value_len && memcmp(got.data(), want.data(), cfg.value_len) == 0) return true;
            if (!racing || v == 0 || vmax - v >= 2) break;
          }
          if (racing) {  // the tick after the sampled sequence number may already be visible
            value_bytes(cfg.seed, id, idx, vmax + 1, cfg.value_len, want.data());
            if (got.size() == cfg.value_len && memcmp(got.data(), want.data(), cfg.value_len) == 0) return true;
          }
          return false;
        };
        while (secs_since(start) < secs) {
          const KB& kb = pool[n_calls % POOL];
          const auto c0 = Clock::now();
          if (single_get) {
            const uint32_t k = (uint32_t)(n_calls % bn);
            const rocksdb::Status s = adbs[kb.shard]->Get(rocksdb::ReadOptions(), kb.keys[k], &one);
            my_lat.push_back((float)(1e6 * std::chrono::duration<double>(Clock::now() - c0).count()));
            if (!s.ok()) bad_st++;
            else if (!check(kb, k, one)) bad++;
            n_keys++;
          } else {
            const auto st = adbs[kb.shard]->MultiGet(rocksdb::ReadOptions(), kb.keys, &values);
            my_lat.push_back((float)(1e3 * std::chrono::duration<double>(Clock::now() - c0).count()));
            for (uint32_t k = 0; k < bn; k++) if (!st[k].ok()) bad_st++;
            // every value of every 8th call, 1 in 64 of the others
            const uint32_t step = (n_calls % 8 == 0) ? 1 : 64;
            for (uint32_t k = (uint32_t)(n_calls % step); k < bn; k += step)
              if (st[k].ok() && !check(kb, k, values[k])) bad++;
            n_keys += bn;
          }
          n_calls++;
        }
        total += n_keys;
        calls += n_calls;
        parity_errors += bad;
        status_errors += bad_st;
        lat->add(my_lat);
      });
      for (auto& t : th) t.join();
      const double el = secs_since(start);
      *per_s = (double)total.load() / el;
      if (calls_out) *calls_out = calls.load();
    };

    std::atomic<uint64_t> applied_now{per_shard};
    // ---- ApplicationDB::MultiGet(batch) from many threads, compacted shards -------------------------
    if (cfg.multiget_threads && cfg.multiget_secs > 0) {
      Percentiles lat;
      cpu0 = cpu_seconds();
      reader(cfg.multiget_threads, cfg.multiget_batch, cfg.multiget_secs, &applied_now, false, &lat, &res->mget_lookups_per_s,
             &res->mget_calls, false);
      res->cpu_s[1] = cpu_seconds() - cpu0;
      res->mget_p50_ms = lat.pct(0.5);
      res->mget_p99_ms = lat.pct(0.99);
    }
    // ---- ApplicationDB::Get from many threads -------------------------------------------------------
    auto comb_delta = [&](int which, const uint64_t before[9], double* out) {
      uint64_t now[9];
      rsp_debug_combiner_stats(gdb0->engine(), which, now);
      out[0] = (double)(now[0] - before[0]);
      out[1] = (double)(now[1] - before[1]);
      for (int k = 2; k < 5; k++) out[k] = 1e-6 * (double)(now[k] - before[k]);
      if (which == 1) {
        for (int k = 5; k < 8; k++) out[k] = 1e-6 * (double)(now[k] - before[k]);
        out[8] = (double)(now[8] - before[8]);
      }
    };
    if (cfg.get_threads && cfg.get_secs > 0) {
      Percentiles lat;
      uint64_t before[9];
      rsp_debug_combiner_stats(gdb0->engine(), 0, before);
      cpu0 = cpu_seconds();
      reader(cfg.get_threads, 1, cfg.get_secs, &applied_now, false, &lat, &res->get_per_s, nullptr, true);
      res->cpu_s[2] = cpu_seconds() - cpu0;
      comb_delta(0, before, res->read_comb);
      res->get_p50_us = lat.pct(0.5);
      res->get_p99_us = lat.pct(0.99);
    }
    uint64_t cur_target = per_shard;
    // ---- steady state of the pull loops alone ---------------------------------------------------------
    if (cfg.steady_rounds) {
      const uint64_t target = cur_target + (uint64_t)cfg.steady_rounds * cfg.updates_per_response;
      leader->lat_ms.clear();
      uint64_t before[9];
      rsp_debug_combiner_stats(gdb0->engine(), 1, before);
      auto& tr = replicator::PullTrace::Get();
      tr.Reset();
      tr.enabled = true;
      t0 = Clock::now();
      cpu0 = cpu_seconds();
      leader->SetTargets(target);
      if (!wait_seq(target, 180)) status_errors++;
      const double el = secs_since(t0);
      res->cpu_s[3] = cpu_seconds() - cpu0;
      tr.enabled = false;
      res->steady_applies_per_s = (double)cfg.steady_rounds * cfg.updates_per_response * S / el;
      res->steady_resp_p50_ms = leader->lat_ms.pct(0.5);
      res->steady_resp_p99_ms = leader->lat_ms.pct(0.99);
      for (int k = 0; k < replicator::PullTrace::kStages; k++)
        res->trace_us[k] = tr.n[k].load() ? 1e-3 * (double)tr.ns[k].load() / (double)tr.n[k].load() : 0.0;
      comb_delta(1, before, res->apply_comb);
      leader->lat_ms.clear();
      cur_target = target;
      applied_now = target;
    }
    // ---- config 3: replicated updates flowing while MultiGet runs -----------------------------------
    if (cfg.update_rounds) {
      const uint64_t target = cur_target + (uint64_t)cfg.update_rounds * cfg.updates_per_response;
      std::atomic<bool> done{false};
      std::atomic<uint64_t> racing_seq{cur_target};
      std::thread watcher([&] {  // the smallest applied count over the shards, sampled: what a racing read may rely on
        while (!done.load()) {
          uint64_t mn = ~0ull;
          for (uint32_t i = 0; i < S; i++) mn = std::min<uint64_t>(mn, dbs[i]->GetLatestSequenceNumber());
          racing_seq.store(mn, std::memory_order_release);
          std::this_thread::sleep_for(std::chrono::microseconds(500));
        }
      });
      Percentiles lat;
      double reads_per_s = 0;
      std::thread readers;
      const bool with_reads = cfg.multiget_threads && cfg.multiget_secs > 0;
      t0 = Clock::now();
      leader->SetTargets(target);
      std::atomic<bool> reads_done{false};
      if (with_reads)
        readers = std::thread([&] {
          // read until the updates are through (bounded), verifying against the racing sequence numbers
          while (!done.load()) {
            double r = 0;
            reader(cfg.multiget_threads, cfg.multiget_batch, 0.25, &racing_seq, true, &lat, &r, nullptr, false);
            reads_per_s = reads_per_s == 0 ? r : 0.5 * (reads_per_s + r);
          }
          reads_done = true;
        });
      if (!wait_seq(target, 180)) status_errors++;
      const double el = secs_since(t0);
      done = true;
      if (with_reads) readers.join();
      watcher.join();
      res->mixed_applies_per_s = (double)cfg.update_rounds * cfg.updates_per_response * S / el;
      res->mixed_lookups_per_s = reads_per_s;
      res->mixed_resp_p50_ms = leader->lat_ms.pct(0.5);
      res->mixed_resp_p99_ms = leader->lat_ms.pct(0.99);
      applied_now = target;
      // after the updates: every key reads back at its final version
      Percentiles lat2;
      double r2 = 0;
      reader(std::min<uint32_t>(8, std::max<uint32_t>(1, cfg.multiget_threads)), cfg.multiget_batch, 0.2, &applied_now, false, &lat2, &r2, nullptr, false);
    }
    for (uint32_t i = 0; i < S; i++) res->applied_total += dbs[i]->GetLatestSequenceNumber();
    res->engine_launches = rsp_kernel_launches(gdb0->engine()) - launches0;
    leader->Stop();
    adbs.clear();  // removeDB
    dbs.clear();
  }
  res->parity_errors = parity_errors.load();
  res->status_errors = status_errors.load();
  F = saved_flags;
  return 0;
}
