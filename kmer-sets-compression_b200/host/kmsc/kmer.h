// kmsc/kmer.h -- host value type Kmer<K> with the reference's interface
// (reference lib/core/kmer.h:17-241: same member names, same bit layout
// A=0 C=1 G=2 T=3, first base most significant) and the bucket/key split
// (lib/core/kmer_set.h:22-43). Written from the behaviour, not copied: the
// reverse complement uses the bit tricks the device code uses instead of a
// per-base loop.
#ifndef KMSC_HOST_KMER_H_
#define KMSC_HOST_KMER_H_

#include <array>
#include <cstddef>
#include <cstdint>
#include <functional>
#include <string>
#include <utility>

namespace kmsc {

template <int K>
class Kmer {
  static_assert(K >= 1 && K <= 32, "K must be in [1, 32]");

 public:
  Kmer() = default;
  explicit Kmer(std::uint64_t bits) : bits_(bits) {}
  explicit Kmer(const std::string& s) {
    std::uint64_t b = 0;
    for (int i = 0; i < K; i++) b = (b << 2) | Code(s[i]);
    bits_ = b;
  }

  std::string String() const {
    std::string s(K, 'A');
    std::uint64_t b = bits_;
    for (int i = K - 1; i >= 0; i--, b >>= 2) s[i] = "ACGT"[b & 3];
    return s;
  }
  char Last() const { return "ACGT"[bits_ & 3]; }

  // reverse the base order and complement every base (A<->T, C<->G)
  Kmer Complement() const {
    std::uint64_t x = ~bits_;
    x = ((x >> 2) & 0x3333333333333333ull) | ((x & 0x3333333333333333ull) << 2);
    x = ((x >> 4) & 0x0F0F0F0F0F0F0F0Full) | ((x & 0x0F0F0F0F0F0F0F0Full) << 4);
    x = ((x >> 8) & 0x00FF00FF00FF00FFull) | ((x & 0x00FF00FF00FF00FFull) << 8);
    x = ((x >> 16) & 0x0000FFFF0000FFFFull) | ((x & 0x0000FFFF0000FFFFull) << 16);
    x = (x >> 32) | (x << 32);
    return Kmer(x >> (64 - 2 * K));
  }
  Kmer Canonical() const {
    const Kmer c = Complement();
    return c.bits_ < bits_ ? c : *this;
  }
  Kmer Next(char c) const { return Kmer(((bits_ << 2) & Mask()) | Code(c)); }
  Kmer Prev(char c) const { return Kmer((bits_ >> 2) | (Code(c) << (2 * (K - 1)))); }
  std::array<Kmer, 4> Nexts() const { return {Next('A'), Next('C'), Next('G'), Next('T')}; }
  std::array<Kmer, 4> Prevs() const { return {Prev('A'), Prev('C'), Prev('G'), Prev('T')}; }

  std::uint64_t Bits() const { return bits_; }
  std::size_t Hash() const { return bits_; }

  static constexpr std::uint64_t Mask() { return K == 32 ? ~0ull : ((1ull << (2 * K)) - 1); }
  static std::uint64_t Code(char c) { return c == 'A' ? 0 : c == 'C' ? 1 : c == 'G' ? 2 : 3; }

 private:
  std::uint64_t bits_ = 0;
};

template <int K> bool operator==(const Kmer<K>& a, const Kmer<K>& b) { return a.Bits() == b.Bits(); }
template <int K> bool operator!=(const Kmer<K>& a, const Kmer<K>& b) { return a.Bits() != b.Bits(); }
template <int K> bool operator<(const Kmer<K>& a, const Kmer<K>& b) { return a.Bits() < b.Bits(); }
template <int K> bool operator>(const Kmer<K>& a, const Kmer<K>& b) { return a.Bits() > b.Bits(); }

template <int K, int N, typename KeyType>
std::pair<int, KeyType> GetBucketAndKeyFromKmer(const Kmer<K>& kmer) {
  constexpr int kb = 2 * K - N;
  static_assert(kb <= static_cast<int>(sizeof(KeyType) * 8), "key does not fit KeyType");
  const std::uint64_t bits = kmer.Bits();
  return {static_cast<int>(bits >> kb), static_cast<KeyType>(bits & ((1ull << kb) - 1))};
}

template <int K, int N, typename KeyType>
Kmer<K> GetKmerFromBucketAndKey(int bucket_id, KeyType key) {
  constexpr int kb = 2 * K - N;
  return Kmer<K>((static_cast<std::uint64_t>(bucket_id) << kb) | static_cast<std::uint64_t>(key));
}

}  // namespace kmsc

namespace std {
template <int K>
struct hash<kmsc::Kmer<K>> {
  size_t operator()(const kmsc::Kmer<K>& k) const { return std::hash<std::uint64_t>()(k.Bits()); }
};
}  // namespace std

#endif
