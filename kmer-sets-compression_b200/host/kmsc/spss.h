// kmsc/spss.h -- SPSS (spectrum-preserving string set) on the host side.
//
//  * GetKmerSetFromSPSS: the decode (reference lib/core/spss.h:1861-1941) runs on the
//    GPU through kmsc_set_from_spss (P2).
//  * GetSPSS / GetSPSSCanonical: construction (reference lib/core/spss.h:230-1858) runs on the
//    GPU as well (kmsc_spss_build, csrc/spss.cu: port matching + pointer jumping). The reference
//    builds unitigs and then a greedy path cover; any output that spells every k-mer of the set
//    exactly once is valid (test/spss.cc:57-68, 113-124). Output is deterministic for a given set.
#ifndef KMSC_HOST_SPSS_H_
#define KMSC_HOST_SPSS_H_
#include <algorithm>
#include <cstdint>
#include <string>
#include <vector>

#include "kmsc/kmer.h"
#include "kmsc/kmer_set.h"

namespace kmsc {

// Complement of a string (reverse + A<->T, C<->G), reference spss.h:20-45.
inline std::string Complement(std::string s) {
  std::reverse(s.begin(), s.end());
  for (char& c : s) c = c == 'A' ? 'T' : c == 'C' ? 'G' : c == 'G' ? 'C' : 'A';
  return s;
}

// The strings of a device-built SPSS (kmsc_spss_build + kmsc_spss_fetch): the text comes back as one
// buffer and is cut at the string offsets.
template <int K, int N, typename KeyType>
std::vector<std::string> BuildSPSS(const KmerSet<K, N, KeyType>& kmer_set, bool canonical) {
  std::vector<std::string> out;
  if (kmer_set.Size() == 0) return out;
  std::int64_t n_strings = 0, n_chars = 0;
  std::string text;
  std::vector<std::int64_t> offs;
  {
    const SetPtr dev = kmer_set.Dev();
    std::lock_guard<std::mutex> l(Device::Mu());
    Device::Check(kmsc_spss_build(Device::Ctx(), dev->set, canonical ? 1 : 0, /*rounds=*/0, &n_strings, &n_chars), "kmsc_spss_build");
    text.resize(static_cast<std::size_t>(n_chars));
    offs.resize(static_cast<std::size_t>(n_strings) + 1);
    Device::Check(kmsc_spss_fetch(Device::Ctx(), text.data(), offs.data()), "kmsc_spss_fetch");
  }
  out.reserve(static_cast<std::size_t>(n_strings));
  for (std::int64_t i = 0; i < n_strings; i++)
    out.emplace_back(text, static_cast<std::size_t>(offs[i]), static_cast<std::size_t>(offs[i + 1] - offs[i]));
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
