// fast_read_map.h — a map optimised for reads: readers take an immutable snapshot without blocking
// writers; writers copy, modify and publish.  Contract of rocksdb_replicator/fast_read_map.h:53-117
// (add / remove / get / clear; add of an existing key and remove of a missing key return false).
#pragma once
#include <memory>
#include <mutex>
#include <unordered_map>

namespace replicator {
namespace detail {

template <typename K, typename V>
class FastReadMap {
 public:
  using Map = std::unordered_map<K, V>;
  FastReadMap() : map_(std::make_shared<const Map>()) {}
  bool add(const K& k, const V& v) {
    std::lock_guard<std::mutex> g(write_mu_);
    auto cur = std::atomic_load(&map_);
    if (cur->count(k)) return false;
    auto next = std::make_shared<Map>(*cur);
    next->emplace(k, v);
    std::atomic_store(&map_, std::shared_ptr<const Map>(std::move(next)));
    return true;
  }
  bool remove(const K& k) {
    std::lock_guard<std::mutex> g(write_mu_);
    auto cur = std::atomic_load(&map_);
    if (!cur->count(k)) return false;
    auto next = std::make_shared<Map>(*cur);
    next->erase(k);
    std::atomic_store(&map_, std::shared_ptr<const Map>(std::move(next)));
    return true;
  }
  bool get(const K& k, V* v) const {
    auto cur = std::atomic_load(&map_);
    auto it = cur->find(k);
    if (it == cur->end()) return false;
    *v = it->second;
    return true;
  }
  void clear() {
    std::lock_guard<std::mutex> g(write_mu_);
    std::atomic_store(&map_, std::shared_ptr<const Map>(std::make_shared<const Map>()));
  }
  size_t size() const { return std::atomic_load(&map_)->size(); }

 private:
  std::shared_ptr<const Map> map_;
  std::mutex write_mu_;
};

}  // namespace detail
}  // namespace replicator
