// kmsc/kmer_set_compact.h -- KmerSetCompact<K,N,KeyType> with the reference's
// interface (lib/core/kmer_set_compact.h:31-347): an immutable SPSS store, two bits
// per base plus string lengths (minus K) in streamvbyte-0124. Same text file format
// (one SPSS string per line, optional external (de)compressor). The decode paths
// GetSampledKmerSet / ToKmerSet run on the GPU from the packed words
// (kmsc_set_from_packed): no ASCII is re-materialised.
#ifndef KMSC_HOST_KMER_SET_COMPACT_H_
#define KMSC_HOST_KMER_SET_COMPACT_H_
#include <cstdint>
#include <string>
#include <unordered_map>
#include <utility>
#include <vector>

#include "kmsc/io.h"
#include "kmsc/kmer_set.h"
#include "kmsc/spss.h"
#include "kmsc/streamvbyte0124.h"

namespace kmsc {

template <int K, int N, typename KeyType>
class KmerSetCompact {
 public:
  KmerSetCompact() = default;

  // SPSS of the set, built AND packed on the device (kmsc_spss_build + kmsc_spss_fetch_packed): what comes back is
  // this container's own 2-bit words and the string boundaries; no text crosses the bus (the reference:
  // GetSPSSCanonical / GetSPSS, then the constructor's per-base packing, lib/core/kmer_set_compact.h:89-118).
  static KmerSetCompact FromKmerSet(const KmerSet<K, N, KeyType>& kmer_set, bool canonical, bool /*fast*/, int /*n_workers*/) {
    KmerSetCompact c;
    if (kmer_set.Size() == 0) {
      c.words_.assign(2, 0);
      c.lengths_compressed_ = Svb0124Encode(std::vector<std::uint32_t>());
      return c;
    }
    std::int64_t n_strings = 0, n_chars = 0;
    std::vector<std::int64_t> offs;
    {
      const SetPtr dev = kmer_set.Dev();
      std::lock_guard<std::mutex> l(Device::Mu());
      Device::Check(kmsc_spss_build(Device::Ctx(), dev->set, canonical ? 1 : 0, /*rounds=*/0, &n_strings, &n_chars), "kmsc_spss_build");
      c.words_.assign(static_cast<std::size_t>((n_chars + 31) / 32 + 2), 0);
      offs.resize(static_cast<std::size_t>(n_strings) + 1);
      Device::Check(kmsc_spss_fetch_packed(Device::Ctx(), c.words_.data(), offs.data()), "kmsc_spss_fetch_packed");
    }
    c.n_ = n_strings;
    c.n_bases_ = n_chars;
    std::vector<std::uint32_t> lengths(static_cast<std::size_t>(n_strings));
    for (std::int64_t i = 0; i < n_strings; i++)
      lengths[static_cast<std::size_t>(i)] = static_cast<std::uint32_t>(offs[static_cast<std::size_t>(i) + 1] - offs[static_cast<std::size_t>(i)]) -
                                            static_cast<std::uint32_t>(K);
    c.lengths_compressed_ = Svb0124Encode(lengths);
    return c;
  }
  static KmerSetCompact FromStrings(const std::vector<std::string>& spss) { return KmerSetCompact(spss); }

  KmerSet<K, N, KeyType> ToKmerSet(bool canonical, int /*n_workers*/) const {
    return KmerSet<K, N, KeyType>(MakeSetPtr(Decode(canonical, /*dedup=*/true, 0, 1 << N)));
  }

  Status Dump(const std::string& file_name, const std::string& compressor, int n_workers) const {
    return WriteLines(file_name, compressor, ToStrings(n_workers));
  }
  static StatusOr<KmerSetCompact> Load(const std::string& file_name, const std::string& decompressor) {
    StatusOr<std::vector<std::string>> lines = ReadLines(file_name, decompressor);
    if (!lines.ok()) return lines.status();
    return KmerSetCompact(lines.value());
  }

  std::int64_t Size(int /*n_workers*/) const {
    std::int64_t s = 0;
    for (std::uint32_t l : Svb0124Decode(lengths_compressed_, static_cast<std::size_t>(n_))) s += static_cast<std::int64_t>(l) + 1;
    return s;
  }
  std::int64_t Weight() const { return n_bases_; }

  // sorted keys of the selected buckets, indexed by position in bucket_ids; duplicates
  // kept; a bucket id listed twice is filled only at its last position (reference :127-131)
  std::vector<std::vector<KeyType>> GetSampledKmerSet(const std::vector<int>& bucket_ids, bool canonical,
                                                      int /*n_workers*/) const {
    std::vector<std::vector<KeyType>> buckets(bucket_ids.size());
    if (bucket_ids.empty()) return buckets;
    int lo = bucket_ids[0], hi = bucket_ids[0];
    for (int b : bucket_ids) { lo = b < lo ? b : lo; hi = b > hi ? b : hi; }
    kmsc_set* s = Decode(canonical, /*dedup=*/false, lo, hi + 1);
    std::vector<std::int64_t> offs((std::size_t(1) << N) + 1);
    std::int64_t n = 0;
    std::vector<KeyType> keys;
    {
      std::lock_guard<std::mutex> l(Device::Mu());
      Device::Check(kmsc_set_size(Device::Ctx(), s, &n), "kmsc_set_size");
      keys.resize(static_cast<std::size_t>(n));
      Device::Check(kmsc_set_to_csr(Device::Ctx(), s, offs.data(), keys.data()), "kmsc_set_to_csr");
      kmsc_set_free(Device::Ctx(), s);
    }
    std::unordered_map<int, std::size_t> last;
    for (std::size_t i = 0; i < bucket_ids.size(); i++) last[bucket_ids[i]] = i;
    for (const auto& p : last) {
      const std::size_t b = static_cast<std::size_t>(p.first);
      buckets[p.second].assign(keys.begin() + offs[b], keys.begin() + offs[b + 1]);
    }
    return buckets;
  }

  // every set of a collection in one batched device call (kmsc_sets_from_packed_batch: the host-to-device
  // copies of later groups overlap the kernels of earlier ones, two synchronisations per group)
  static std::vector<KmerSet<K, N, KeyType>> ToKmerSetBatch(const std::vector<KmerSetCompact>& compact, bool canonical) {
    const std::size_t m = compact.size();
    std::vector<KmerSet<K, N, KeyType>> out;
    if (m == 0) return out;
    std::vector<const std::uint64_t*> words(m);
    std::vector<std::vector<std::int64_t>> offs(m);
    std::vector<const std::int64_t*> offp(m);
    std::vector<std::int64_t> nstr(m);
    for (std::size_t j = 0; j < m; j++) {
      words[j] = compact[j].words_.data();
      offs[j] = compact[j].StringOffsets();
      offp[j] = offs[j].data();
      nstr[j] = static_cast<std::int64_t>(offs[j].size()) - 1;
    }
    std::vector<kmsc_set*> sets(m, nullptr);
    {
      std::lock_guard<std::mutex> l(Device::Mu());
      Device::Check(kmsc_sets_from_packed_batch(Device::Ctx(), K, N, static_cast<int>(sizeof(KeyType)), static_cast<std::int32_t>(m),
                                                words.data(), offp.data(), nstr.data(), canonical ? 1 : 0, /*dedup=*/1, 0, 1 << N,
                                                sets.data()),
                    "kmsc_sets_from_packed_batch");
    }
    out.reserve(m);
    for (kmsc_set* s : sets) out.push_back(KmerSet<K, N, KeyType>(MakeSetPtr(s)));
    return out;
  }

  // the device set of the k-mers in [bucket_lo, bucket_hi) (a rank's prefix shard)
  KmerSet<K, N, KeyType> ToKmerSetShard(bool canonical, int bucket_lo, int bucket_hi) const {
    return KmerSet<K, N, KeyType>(MakeSetPtr(Decode(canonical, true, bucket_lo, bucket_hi)));
  }

  // the packed form the device decodes (2 bits per base, 32 bases per word, first base on top)
  // and the string boundaries in bases: what kmsc_set_from_packed / kmsc_sets_from_packed_batch take
  const std::vector<std::uint64_t>& PackedWords() const { return words_; }
  std::vector<std::int64_t> StringOffsets() const {
    const std::vector<std::uint32_t> lens = Svb0124Decode(lengths_compressed_, static_cast<std::size_t>(n_));
    std::vector<std::int64_t> offs(static_cast<std::size_t>(n_) + 1, 0);
    for (std::int64_t i = 0; i < n_; i++)
      offs[static_cast<std::size_t>(i) + 1] =
          offs[static_cast<std::size_t>(i)] + static_cast<std::uint32_t>(lens[static_cast<std::size_t>(i)] + static_cast<std::uint32_t>(K));
    return offs;
  }

  std::vector<std::string> ToStrings(int /*n_workers*/) const {
    const std::vector<std::uint32_t> lens = Svb0124Decode(lengths_compressed_, static_cast<std::size_t>(n_));
    std::vector<std::string> out(static_cast<std::size_t>(n_));
    std::int64_t pos = 0;
    for (std::int64_t i = 0; i < n_; i++) {
      // a string shorter than K is stored as (length - K) mod 2^32 and comes back by the same
      // 32-bit arithmetic (the reference's GetLengths, lib/core/kmer_set_compact.h:269-287)
      const std::int64_t len = static_cast<std::uint32_t>(lens[static_cast<std::size_t>(i)] + static_cast<std::uint32_t>(K));
      std::string& s = out[static_cast<std::size_t>(i)];
      s.resize(static_cast<std::size_t>(len));
      for (std::int64_t j = 0; j < len; j++, pos++)
        s[static_cast<std::size_t>(j)] = "ACGT"[(words_[static_cast<std::size_t>(pos >> 5)] >> (62 - 2 * (pos & 31))) & 3];
    }
    return out;
  }

 private:
  explicit KmerSetCompact(const std::vector<std::string>& spss) {
    n_ = static_cast<std::int64_t>(spss.size());
    std::vector<std::uint32_t> lengths(spss.size());
    std::int64_t total = 0;
    for (std::size_t i = 0; i < spss.size(); i++) {
      lengths[i] = static_cast<std::uint32_t>(spss[i].size()) - static_cast<std::uint32_t>(K);  // mod 2^32 for short lines, as the reference
      total += static_cast<std::int64_t>(spss[i].size());
    }
    n_bases_ = total;
    words_.assign(static_cast<std::size_t>((total + 31) / 32 + 2), 0);
    std::int64_t pos = 0;
    for (const std::string& s : spss)
      for (char c : s) {
        words_[static_cast<std::size_t>(pos >> 5)] |= Kmer<K>::Code(c) << (62 - 2 * (pos & 31));
        pos++;
      }
    lengths_compressed_ = Svb0124Encode(lengths);
  }

  kmsc_set* Decode(bool canonical, bool dedup, int bucket_lo, int bucket_hi) const {
    const std::vector<std::int64_t> offs = StringOffsets();
    kmsc_set* s = nullptr;
    std::lock_guard<std::mutex> l(Device::Mu());
    Device::Check(kmsc_set_from_packed(Device::Ctx(), K, N, static_cast<int>(sizeof(KeyType)), words_.data(), offs.data(), n_,
                                       canonical ? 1 : 0, dedup ? 1 : 0, bucket_lo, bucket_hi, &s), "kmsc_set_from_packed");
    return s;
  }

  std::int64_t n_ = 0;                            // number of strings
  std::int64_t n_bases_ = 0;                      // sum of string lengths = Weight()
  std::vector<std::uint8_t> lengths_compressed_;  // length - K per string, streamvbyte 0124
  std::vector<std::uint64_t> words_;              // 2 bits per base, 32 bases per word, first base on top
};

}  // namespace kmsc
#endif
