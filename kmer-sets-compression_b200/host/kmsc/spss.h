// kmsc/spss.h -- SPSS (spectrum-preserving string set) on the host side.
//
//  * GetKmerSetFromSPSS: the decode (reference lib/core/spss.h:1861-1941) runs on the
//    GPU through kmsc_set_from_spss (P2).
//  * GetSPSS / GetSPSSCanonical: construction (reference lib/core/spss.h:230-1858) is a
//    "next" row and stays on the host. The reference builds unitigs and then
//    a greedy path cover; any output that spells every k-mer of the set exactly once is
//    valid (test/spss.cc:57-68, 113-124). This builder walks greedy simplitigs over
//    the SORTED k-mer array (binary-search Contains, no hash tables): start at an
//    unvisited k-mer, extend right while an unvisited successor exists, then left.
//    Output is deterministic for a given set.
#ifndef KMSC_HOST_SPSS_H_
#define KMSC_HOST_SPSS_H_
#include <algorithm>
#include <cstdint>
#include <string>
#include <vector>

#include "kmsc/kmer.h"
#include "kmsc/kmer_set.h"

namespace kmsc {

namespace internal {
// index of value v in the sorted array, or -1
inline std::int64_t FindSorted(const std::vector<std::uint64_t>& a, std::uint64_t v) {
  auto it = std::lower_bound(a.begin(), a.end(), v);
  return (it != a.end() && *it == v) ? static_cast<std::int64_t>(it - a.begin()) : -1;
}
}  // namespace internal

// Complement of a string (reverse + A<->T, C<->G), reference spss.h:20-45.
inline std::string Complement(std::string s) {
  std::reverse(s.begin(), s.end());
  for (char& c : s) c = c == 'A' ? 'T' : c == 'C' ? 'G' : c == 'G' ? 'C' : 'A';
  return s;
}

template <int K, int N, typename KeyType>
std::vector<std::string> BuildSPSS(const KmerSet<K, N, KeyType>& kmer_set, bool canonical) {
  const std::vector<std::uint64_t>& a = kmer_set.SortedBits();
  std::vector<bool> visited(a.size(), false);
  std::vector<std::string> out;
  auto lookup = [&](const Kmer<K>& k) -> std::int64_t {
    return internal::FindSorted(a, canonical ? k.Canonical().Bits() : k.Bits());
  };
  std::string left;  // bases prepended while walking left (reversed)
  for (std::size_t start = 0; start < a.size(); start++) {
    if (visited[start]) continue;
    visited[start] = true;
    const Kmer<K> first(a[start]);
    std::string s = first.String();
    // extend to the right
    Kmer<K> cur = first;
    for (;;) {
      bool moved = false;
      for (char c : {'A', 'C', 'G', 'T'}) {
        const Kmer<K> nxt = cur.Next(c);
        const std::int64_t idx = lookup(nxt);
        if (idx >= 0 && !visited[static_cast<std::size_t>(idx)]) {
          visited[static_cast<std::size_t>(idx)] = true;
          s.push_back(c);
          cur = nxt;
          moved = true;
          break;
        }
      }
      if (!moved) break;
    }
    // extend to the left
    left.clear();
    cur = first;
    for (;;) {
      bool moved = false;
      for (char c : {'A', 'C', 'G', 'T'}) {
        const Kmer<K> prv = cur.Prev(c);
        const std::int64_t idx = lookup(prv);
        if (idx >= 0 && !visited[static_cast<std::size_t>(idx)]) {
          visited[static_cast<std::size_t>(idx)] = true;
          left.push_back(c);
          cur = prv;
          moved = true;
          break;
        }
      }
      if (!moved) break;
    }
    if (!left.empty()) {
      std::reverse(left.begin(), left.end());
      s = left + s;
    }
    out.push_back(std::move(s));
  }
  return out;
}

template <int K, int N, typename KeyType>
std::vector<std::string> GetSPSS(const KmerSet<K, N, KeyType>& kmer_set, int /*n_workers*/) {
  return BuildSPSS<K, N, KeyType>(kmer_set, false);
}
template <int K, int N, typename KeyType>
std::vector<std::string> GetSPSSCanonical(const KmerSet<K, N, KeyType>& kmer_set, bool /*fast*/, int /*n_workers*/) {
  return BuildSPSS<K, N, KeyType>(kmer_set, true);
}

// Reads SPSS and returns the corresponding k-mer set: device decode (P2, dedup).
template <int K, int N, typename KeyType>
KmerSet<K, N, KeyType> GetKmerSetFromSPSS(const std::vector<std::string>& spss, bool canonical, int /*n_workers*/) {
  std::string text;
  std::vector<std::int64_t> offs(spss.size() + 1, 0);
  std::size_t total = 0;
  for (const std::string& s : spss) total += s.size();
  text.reserve(total);
  for (std::size_t i = 0; i < spss.size(); i++) {
    text += spss[i];
    offs[i + 1] = static_cast<std::int64_t>(text.size());
  }
  kmsc_set* s = nullptr;
  {
    std::lock_guard<std::mutex> l(Device::Mu());
    Device::Check(kmsc_set_from_spss(Device::Ctx(), K, N, static_cast<int>(sizeof(KeyType)), text.data(), offs.data(),
                                     static_cast<std::int64_t>(spss.size()), canonical ? 1 : 0, /*dedup=*/1, 0, 1 << N, &s),
                  "kmsc_set_from_spss");
  }
  return KmerSet<K, N, KeyType>(MakeSetPtr(s));
}

}  // namespace kmsc
#endif
