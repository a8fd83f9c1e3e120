// TEST INFRASTRUCTURE ONLY (oracle/). Seedable stand-in for absl::InsecureBitGen
// and absl::Uniform. KMSC_REF_SEED (env) makes every generator instance draw
// from one reproducible sequence (instance i is seeded seed+i); unset = random.
#ifndef KMSC_ORACLE_SHIM_RANDOM_H_
#define KMSC_ORACLE_SHIM_RANDOM_H_
#include <atomic>
#include <cstdint>
#include <cstdlib>
#include <random>
#include <type_traits>
namespace kmsc_shim {
inline std::atomic<std::uint64_t>& seed_counter() {
  static std::atomic<std::uint64_t> c{0};
  return c;
}
}  // namespace kmsc_shim
namespace absl {
struct IntervalClosedTag {};
constexpr IntervalClosedTag IntervalClosed{};
class InsecureBitGen {
 public:
  using result_type = std::uint64_t;
  InsecureBitGen() : eng_(NextSeed()) {}
  static constexpr result_type min() { return 0; }
  static constexpr result_type max() { return ~0ull; }
  result_type operator()() { return eng_(); }
 private:
  static std::uint64_t NextSeed() {
    std::atomic<std::uint64_t>& counter = kmsc_shim::seed_counter();
    const char* s = std::getenv("KMSC_REF_SEED");
    if (s) return std::strtoull(s, nullptr, 10) + counter.fetch_add(1);
    return std::random_device{}() * 0x9e3779b97f4a7c15ull + counter.fetch_add(1);
  }
  std::mt19937_64 eng_;
};
template <typename T = void, typename G, typename A, typename B>
auto Uniform(IntervalClosedTag, G&& g, A lo, B hi) {
  using R = std::conditional_t<std::is_void<T>::value, std::common_type_t<A, B>, T>;
  std::uniform_int_distribution<R> d(static_cast<R>(lo), static_cast<R>(hi));
  return d(g);
}
template <typename T = void, typename G, typename A, typename B>
auto Uniform(G&& g, A lo, B hi) {  // half-open [lo, hi)
  using R = std::conditional_t<std::is_void<T>::value, std::common_type_t<A, B>, T>;
  std::uniform_int_distribution<R> d(static_cast<R>(lo), static_cast<R>(hi) - 1);
  return d(g);
}
}  // namespace absl
#endif
