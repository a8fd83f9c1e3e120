// TEST INFRASTRUCTURE ONLY (oracle/). std-only stand-in for absl::flat_hash_set.
#ifndef KMSC_ORACLE_SHIM_FLAT_HASH_SET_H_
#define KMSC_ORACLE_SHIM_FLAT_HASH_SET_H_
#include <unordered_set>
#include "shim_hash.h"
namespace absl {
template <typename T, typename H = kmsc_shim::Hash<T>>
using flat_hash_set = std::unordered_set<T, H>;
}
#endif
