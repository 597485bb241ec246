// replicator_stats.h — the counters and metrics the replication library reports, under the reference's names
// (rocksdb_replicator/replicator_stats.cpp:33-100), so dashboards keep working.  Reported through common::Stats;
// the reference's optional per-db / per-dataset tagging (" db=...", " dataset=...") is applied the same way.
#pragma once
#include <string>

#include "common/segment_utils.h"
#include "common/stats.h"

namespace replicator {

struct ReplicatorStatFlags {
  bool replicator_enable_per_db_stats = false;
  bool replicator_enable_per_dataset_stats = false;
};
inline ReplicatorStatFlags& StatFlags() { static ReplicatorStatFlags f; return f; }

#define RSP_STAT(id, text) static const char* const id = text
RSP_STAT(kReplicatorLatency, "replicator_latency_ms");
RSP_STAT(kReplicatorOutBytes, "replicator_out_bytes");
RSP_STAT(kReplicatorOutNumUpdates, "replicator_out_num_updates");
RSP_STAT(kReplicatorInBytes, "replicator_in_bytes");
RSP_STAT(kReplicatorWriteBytes, "replicator_write_bytes");
RSP_STAT(kReplicatorConnectionErrors, "replicator_connection_errors");
RSP_STAT(kReplicatorRemoteApplicationExceptions, "replicator_remote_app_exceptions");
RSP_STAT(kReplicatorGetUpdatesSinceErrors, "replicator_get_updates_since_errors");
RSP_STAT(kReplicatorWriteSuccess, "replicator_write_success");
RSP_STAT(kReplicatorWriteLeaderFailure, "replicator_write_leader_failure");
RSP_STAT(kReplicatorWriteWaitTimedOut, "replicator_write_wait_timed_out");
RSP_STAT(kReplicatorWriteToLeaderMs, "replicator_write_to_leader_ms");
RSP_STAT(kReplicatorWriteTwoAckDegraded, "replicator_write_two_ack_degraded");
RSP_STAT(kReplicatorWriteTwoAckRecovered, "replicator_write_two_ack_recovered");
RSP_STAT(kReplicatorLeaderSequenceNumbersBehind, "replicator_leader_sequence_numbers_behind");
RSP_STAT(kReplicatorPullRequests, "replicator_pull_requests");
RSP_STAT(kReplicatorPullRequestsSuccess, "replicator_pull_requests_success");
RSP_STAT(kReplicatorPullRequestsFailure, "replicator_pull_requests_failure");
RSP_STAT(kReplicatorPullRequestsNoUpdates, "replicator_pull_requests_no_updates");
RSP_STAT(kReplicatorPullFromNonLeader, "replicator_pull_from_non_leader");
RSP_STAT(kReplicatorHandleResponseFailure, "replicator_handle_response_failure");
RSP_STAT(kReplicatorResetUpstreamOnNoUpdates, "replicator_reset_upstream_on_no_updates_attempted");
RSP_STAT(kReplicatorHandleObserverRequests, "replicator_handle_observer_requests");
#undef RSP_STAT

inline std::string TaggedName(const char* name, const std::string& db_name) {
  if (!db_name.empty()) {
    if (StatFlags().replicator_enable_per_db_stats) return std::string(name) + " db=" + db_name;
    if (StatFlags().replicator_enable_per_dataset_stats) return std::string(name) + " dataset=" + common::DbNameToSegment(db_name);
  }
  return name;
}
inline bool Tagging() { return StatFlags().replicator_enable_per_db_stats || StatFlags().replicator_enable_per_dataset_stats; }
// the names above are static strings: without per-db tagging a report is a thread-local table hit by pointer
inline void incCounter(const char* name, uint64_t value, const std::string& db_name) {
  if (!Tagging()) common::Stats::get()->IncrStatic(name, value);
  else common::Stats::get()->Incr(TaggedName(name, db_name), value);
}
inline void logMetric(const char* name, int64_t value, const std::string& db_name) {
  if (!Tagging()) common::Stats::get()->AddMetricStatic(name, value);
  else common::Stats::get()->AddMetric(TaggedName(name, db_name), value);
}

}  // namespace replicator
