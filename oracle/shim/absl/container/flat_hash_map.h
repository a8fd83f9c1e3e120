// TEST INFRASTRUCTURE ONLY (oracle/). std-only stand-in for absl::flat_hash_map.
// A map keyed by std::pair<int,int> (only KmerSetSet's `weights`,
// reference lib/core/kmer_set_set.h:187) is an ordered std::map so that the
// reference's "first strict maximum in iteration order" scan (:308-316)
// resolves ties to the smallest (j,k): a legal, reproducible refinement of an
// order the reference leaves unspecified.
#ifndef KMSC_ORACLE_SHIM_FLAT_HASH_MAP_H_
#define KMSC_ORACLE_SHIM_FLAT_HASH_MAP_H_
#include <map>
#include <unordered_map>
#include <utility>
#include "shim_hash.h"
namespace kmsc_shim {
template <typename K, typename V>
struct MapSelect { using type = std::unordered_map<K, V, Hash<K>>; };
template <typename V>
struct MapSelect<std::pair<int, int>, V> {
  struct type : std::map<std::pair<int, int>, V> {
    void reserve(std::size_t) {}
  };
};
}  // namespace kmsc_shim
namespace absl {
template <typename K, typename V>
using flat_hash_map = typename kmsc_shim::MapSelect<K, V>::type;
}
#endif
