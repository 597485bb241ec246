// common/segment_utils.h — shard naming: segment + 5-digit shard id (common/segment_utils.cpp:26-52).
#pragma once
#include <cstdio>
#include <string>
namespace common {
inline std::string SegmentToDbName(const std::string& segment, int shard_id) {
  char buf[16];
  snprintf(buf, sizeof(buf), "%05d", shard_id);
  return segment + buf;
}
inline std::string DbNameToSegment(const std::string& db_name) { return db_name.size() < 5 ? db_name : db_name.substr(0, db_name.size() - 5); }
inline int ExtractShardId(const std::string& db_name) {
  if (db_name.size() < 5) return -1;
  try { return std::stoi(db_name.substr(db_name.size() - 5)); } catch (...) { return -1; }
}
}  // namespace common
