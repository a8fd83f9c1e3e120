// kmsc/spss.h -- SPSS (spectrum-preserving string set) on the host side.
//
//  * GetKmerSetFromSPSS: the decode (reference lib/core/spss.h:1861-1941) runs on the
//    GPU through kmsc_set_from_spss (P2).
//  * GetSPSS / GetSPSSCanonical: construction (reference lib/core/spss.h:230-1858) is a
//    "next" row and stays on the host. The reference builds unitigs and then
//    a greedy path cover; any output that spells every k-mer of the set exactly once is
//    valid (test/spss.cc:57-68, 113-124). This builder asks the GPU for the de Bruijn
//    neighbours of every k-mer (kmsc_set_neighbors: one binary search per neighbour, the
//    Contains() calls that dominate the reference) and walks greedy simplitigs over that
//    table on the host. Output is deterministic for a given set.
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

// Greedy simplitigs over the device-computed neighbour table (kmsc_set_neighbors): start at an
// unvisited k-mer, extend right while an unvisited successor exists (bases tried in A, C, G, T
// order), then left. The walk keeps (k-mer index, orientation): orientation 1 means the string
// spells the reverse complement of the stored (canonical) k-mer, whose successors are the
// complemented predecessors of the stored one. Output is deterministic for a given set.
template <int K, int N, typename KeyType>
std::vector<std::string> BuildSPSS(const KmerSet<K, N, KeyType>& kmer_set, bool canonical) {
  const std::vector<std::uint64_t>& a = kmer_set.SortedBits();
  const std::size_t n = a.size();
  std::vector<std::string> out;
  if (n == 0) return out;
  std::vector<std::int32_t> nb(n * 8);
  {
    const SetPtr dev = kmer_set.Dev();
    std::lock_guard<std::mutex> l(Device::Mu());
    Device::Check(kmsc_set_neighbors(Device::Ctx(), dev->set, canonical ? 1 : 0, nb.data()), "kmsc_set_neighbors");
  }
  std::vector<bool> visited(n, false);
  // oriented neighbour of (i, o) when base c is appended (right = true) or prepended
  auto step = [&](std::size_t i, int o, int c, bool right, std::size_t* j, int* o2) -> bool {
    std::int32_t e;
    if (o == 0) e = nb[i * 8 + (right ? 0 : 4) + static_cast<std::size_t>(c)];
    else e = nb[i * 8 + (right ? 4 : 0) + static_cast<std::size_t>(3 - c)];
    if (e < 0) return false;
    *j = static_cast<std::size_t>(e >> 1);
    *o2 = o == 0 ? (e & 1) : !(e & 1);
    return true;
  };
  std::string left;  // bases prepended while walking left (reversed)
  for (std::size_t start = 0; start < n; start++) {
    if (visited[start]) continue;
    visited[start] = true;
    std::string s = Kmer<K>(a[start]).String();
    for (int dir = 0; dir < 2; dir++) {
      const bool right = dir == 0;
      left.clear();
      std::size_t cur = start;
      int o = 0;
      for (;;) {
        bool moved = false;
        for (int c = 0; c < 4; c++) {
          std::size_t j;
          int o2;
          if (step(cur, o, c, right, &j, &o2) && !visited[j]) {
            visited[j] = true;
            (right ? s : left).push_back("ACGT"[c]);
            cur = j;
            o = o2;
            moved = true;
            break;
          }
        }
        if (!moved) break;
      }
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
