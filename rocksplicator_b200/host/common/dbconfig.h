// common/dbconfig.h — per-dataset replication settings, hot-swappable.
// Behavioural target: common/dbconfig.{h,cpp} in the reference: {"dataset": {"<segment>": {"ack_mode": N}}} is
// published atomically, getReplicationMode(db_name) resolves the db's segment (db name minus its 5-digit shard
// suffix) and the replication library uses max(flag, dataset setting) (replicated_db.cpp:131-136, 458-462).
#pragma once
#include <cctype>
#include <cstdint>
#include <map>
#include <memory>
#include <string>

#include "common/segment_utils.h"

namespace common {

class DBConfigManager {
 public:
  static DBConfigManager* get() { static DBConfigManager m; return &m; }

  uint32_t getReplicationMode(const std::string& db_name, uint32_t def_value = 0) const {
    const auto cfg = std::atomic_load(&cfg_);
    const auto it = cfg->find(DbNameToSegment(db_name));
    return it == cfg->end() ? def_value : it->second;
  }
  void setDatasetAckMode(const std::string& dataset, uint32_t mode) {
    auto next = std::make_shared<Map>(*std::atomic_load(&cfg_));
    (*next)[dataset] = mode;
    std::atomic_store(&cfg_, std::shared_ptr<const Map>(std::move(next)));
  }
  void clear() { std::atomic_store(&cfg_, std::shared_ptr<const Map>(std::make_shared<Map>())); }

  // accepts the reference's config document; anything else leaves the current config in place
  bool loadJsonText(const std::string& text) {
    Parser p{text, 0};
    Map out;
    if (!p.Object([&](const std::string& top) {
          if (top != "dataset") return p.Skip();
          return p.Object([&](const std::string& dataset) {
            return p.Object([&](const std::string& field) {
              if (field != "ack_mode") return p.Skip();
              uint64_t v;
              if (!p.Uint(&v)) return false;
              out[dataset] = (uint32_t)v;
              return true;
            });
          });
        }))
      return false;
    p.Ws();
    if (p.at != text.size()) return false;
    std::atomic_store(&cfg_, std::shared_ptr<const Map>(std::make_shared<Map>(std::move(out))));
    return true;
  }

 private:
  using Map = std::map<std::string, uint32_t>;
  DBConfigManager() : cfg_(std::make_shared<Map>()) {}

  // just enough JSON for the document above (objects, strings, numbers, literals, arrays are skipped)
  struct Parser {
    const std::string& s;
    size_t at;
    void Ws() { while (at < s.size() && isspace((unsigned char)s[at])) at++; }
    bool Eat(char c) { Ws(); if (at < s.size() && s[at] == c) { at++; return true; } return false; }
    bool String(std::string* out) {
      if (!Eat('"')) return false;
      out->clear();
      while (at < s.size() && s[at] != '"') {
        if (s[at] == '\\' && at + 1 < s.size()) at++;
        out->push_back(s[at++]);
      }
      return at < s.size() && s[at++] == '"';
    }
    bool Uint(uint64_t* v) {
      Ws();
      const size_t b = at;
      *v = 0;
      while (at < s.size() && isdigit((unsigned char)s[at])) *v = *v * 10 + (uint64_t)(s[at++] - '0');
      return at > b;
    }
    template <class F> bool Object(F on_member) {
      if (!Eat('{')) return false;
      if (Eat('}')) return true;
      do {
        std::string key;
        if (!String(&key) || !Eat(':') || !on_member(key)) return false;
      } while (Eat(','));
      return Eat('}');
    }
    bool Skip() {
      Ws();
      if (at >= s.size()) return false;
      if (s[at] == '{') return Object([&](const std::string&) { return Skip(); });
      if (s[at] == '"') { std::string t; return String(&t); }
      if (s[at] == '[') {
        at++;
        if (Eat(']')) return true;
        do { if (!Skip()) return false; } while (Eat(','));
        return Eat(']');
      }
      const size_t b = at;
      while (at < s.size() && (isalnum((unsigned char)s[at]) || s[at] == '-' || s[at] == '+' || s[at] == '.')) at++;
      return at > b;
    }
  };
  std::shared_ptr<const Map> cfg_;
};

}  // namespace common
