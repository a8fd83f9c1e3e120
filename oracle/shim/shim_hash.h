// TEST INFRASTRUCTURE ONLY (oracle/). std-only stand-in for the tiny part of
// Abseil's hashing that the reference headers touch: types that declare a
// friend AbslHashValue(H, const T&) (reference lib/core/kmer.h:213-216) are
// hashed through it, everything else through std::hash.
#ifndef KMSC_ORACLE_SHIM_HASH_H_
#define KMSC_ORACLE_SHIM_HASH_H_
#include <cstddef>
#include <cstdint>
#include <functional>
#include <type_traits>
#include <utility>

namespace kmsc_shim {

struct HashState {
  std::uint64_t v = 0x9e3779b97f4a7c15ull;
  static std::uint64_t mix(std::uint64_t x) {
    x ^= x >> 33; x *= 0xff51afd7ed558ccdull;
    x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull;
    x ^= x >> 33; return x;
  }
  template <typename T>
  static HashState combine1(HashState h, const T& x) {
    h.v = mix(h.v ^ static_cast<std::uint64_t>(std::hash<T>()(x)));
    return h;
  }
  static HashState combine(HashState h) { return h; }
  template <typename T, typename... Ts>
  static HashState combine(HashState h, const T& x, const Ts&... xs) {
    return combine(combine1(std::move(h), x), xs...);
  }
};

template <typename T, typename = void>
struct HasAbslHashValue : std::false_type {};
template <typename T>
struct HasAbslHashValue<
    T, std::void_t<decltype(AbslHashValue(std::declval<HashState>(),
                                          std::declval<const T&>()))>>
    : std::true_type {};

template <typename T>
struct Hash {
  std::size_t operator()(const T& x) const {
    if constexpr (HasAbslHashValue<T>::value) {
      return AbslHashValue(HashState{}, x).v;
    } else {
      return HashState::mix(static_cast<std::uint64_t>(std::hash<T>()(x)));
    }
  }
};

template <typename A, typename B>
struct Hash<std::pair<A, B>> {
  std::size_t operator()(const std::pair<A, B>& p) const {
    return HashState::combine(HashState{}, p.first, p.second).v;
  }
};

}  // namespace kmsc_shim
#endif
