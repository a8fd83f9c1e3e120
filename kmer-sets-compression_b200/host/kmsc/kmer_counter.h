// kmsc/kmer_counter.h -- KmerCounter<K,N,KeyType,ValueType=uint8_t> with the
// reference's interface (lib/core/kmer_counter.h:48-299). FromReads / FromFASTA run
// on the GPU (kmsc_count_reads / kmsc_count_fasta: pack, mask windows touching 'N',
// partition, sort, run-length count with saturation at 255); the distinct k-mers
// stay on the device, the uint8 counts are mirrored to the host for Get().
#ifndef KMSC_HOST_KMER_COUNTER_H_
#define KMSC_HOST_KMER_COUNTER_H_
#include <algorithm>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <limits>
#include <string>
#include <type_traits>
#include <utility>
#include <vector>

#include "kmsc/io.h"
#include "kmsc/pinned_reader.h"
#include "kmsc/kmer_set.h"

namespace kmsc {

template <typename T>
T AddWithMax(T x, T y) {  // reference kmer_counter.h:28-38
  if (std::is_integral<T>::value) {
    const std::int64_t mx = static_cast<std::int64_t>(std::numeric_limits<T>::max());
    return static_cast<T>(std::min(mx, static_cast<std::int64_t>(x) + static_cast<std::int64_t>(y)));
  }
  return x + y;
}

template <int K, int N, typename KeyType, typename ValueType = std::uint8_t>
class KmerCounter {
  static_assert(std::is_same<ValueType, std::uint8_t>::value, "the device path counts in uint8 like the reference default");

 public:
  KmerCounter() = default;

  std::int64_t Size() const { return static_cast<std::int64_t>(kmers_.size()); }

  static KmerCounter FromReads(std::vector<std::string> reads, bool canonical, int /*n_workers*/) {
    std::string text;
    std::size_t total = 0;
    for (const std::string& r : reads) total += r.size() + 1;
    text.reserve(total);
    for (const std::string& r : reads) { text += r; text += '\n'; }
    KmerCounter c;
    Status st = c.Count(text, canonical, /*fasta=*/false);
    if (!st.ok()) { std::fprintf(stderr, "KmerCounter::FromReads: %s\n", st.ToString().c_str()); std::abort(); }
    return c;
  }
  static StatusOr<KmerCounter> FromFASTA(const std::string& file_name, const std::string& decompressor,
                                         bool canonical, int /*n_workers*/) {
    StatusOr<std::string> text = internal::ReadAll(file_name, decompressor);
    if (!text.ok()) return text.status();
    KmerCounter c;
    Status st = c.Count(text.value(), canonical, /*fasta=*/true);
    if (!st.ok()) return st;
    return c;
  }
  static StatusOr<KmerCounter> FromFASTA(std::vector<std::string> lines, bool canonical, int /*n_workers*/) {
    std::string text;
    for (const std::string& l : lines) { text += l; text += '\n'; }
    KmerCounter c;
    Status st = c.Count(text, canonical, /*fasta=*/true);
    if (!st.ok()) return st;
    return c;
  }

  // keeps k-mers with count >= cutoff; returns the set and the number dropped
  std::pair<KmerSet<K, N, KeyType>, std::int64_t> ToKmerSet(ValueType cutoff, int /*n_workers*/) const {
    std::vector<std::uint64_t> kept;
    kept.reserve(kmers_.size());
    std::int64_t cut = 0;
    for (std::size_t i = 0; i < kmers_.size(); i++) {
      if (counts_[i] < cutoff) { cut++; continue; }
      kept.push_back(kmers_[i]);
    }
    return {KmerSet<K, N, KeyType>::FromSortedBits(std::move(kept)), cut};
  }

  ValueType Get(const Kmer<K>& kmer) const {
    auto it = std::lower_bound(kmers_.begin(), kmers_.end(), kmer.Bits());
    if (it == kmers_.end() || *it != kmer.Bits()) return 0;
    return counts_[static_cast<std::size_t>(it - kmers_.begin())];
  }
  KmerCounter& Add(const Kmer<K>& kmer, ValueType v) {
    auto it = std::lower_bound(kmers_.begin(), kmers_.end(), kmer.Bits());
    const std::size_t pos = static_cast<std::size_t>(it - kmers_.begin());
    if (it == kmers_.end() || *it != kmer.Bits()) {
      kmers_.insert(it, kmer.Bits());
      counts_.insert(counts_.begin() + static_cast<std::ptrdiff_t>(pos), AddWithMax<ValueType>(0, v));
    } else {
      counts_[pos] = AddWithMax<ValueType>(counts_[pos], v);
    }
    return *this;
  }

  // one-shot device path used by kmerset-build: count + cutoff without mirroring counts
  static StatusOr<std::pair<KmerSet<K, N, KeyType>, std::int64_t>> CountToKmerSet(
      const std::string& fasta_text, bool canonical, int cutoff, std::int64_t* n_distinct) {
    kmsc_set* s = nullptr;
    std::int64_t cut = 0, nd = 0;
    std::lock_guard<std::mutex> l(Device::Mu());
    const int rc = kmsc_count_fasta(Device::Ctx(), K, N, static_cast<int>(sizeof(KeyType)), fasta_text.data(),
                                    static_cast<std::int64_t>(fasta_text.size()), canonical ? 1 : 0, cutoff, &s, &cut, &nd);
    if (rc == KMSC_E_FORMAT) return FailedPreconditionError(kmsc_last_error());
    if (rc != KMSC_OK) return InternalError(kmsc_last_error());
    if (n_distinct) *n_distinct = nd;
    return std::make_pair(KmerSet<K, N, KeyType>(MakeSetPtr(s)), cut);
  }

  // streaming device path for files that do not fit one call (config 4): the file is read in
  // chunks of whole records, every chunk is counted on the GPU and merged into a device counter
  static StatusOr<std::pair<KmerSet<K, N, KeyType>, std::int64_t>> CountFileToKmerSet(
      const std::string& file_name, const std::string& decompressor, bool canonical, int cutoff,
      std::size_t chunk_bytes, std::int64_t* n_distinct) {
    kmsc_counter* c = nullptr;
    {
      std::lock_guard<std::mutex> l(Device::Mu());
      Device::Check(kmsc_counter_create(Device::Ctx(), K, N, static_cast<int>(sizeof(KeyType)), canonical ? 1 : 0, &c),
                    "kmsc_counter_create");
    }
    // default: a reader thread freads into page-locked buffers while the device counts the previous chunk;
    // KMSC_IO_OVERLAP=1: pageable chunks, overlapped; =0: read and count in turn (diagnostics)
    const char* ov = std::getenv("KMSC_IO_OVERLAP");
    auto run = [&](auto&& sink) {
      if (ov && std::atoi(ov) == 0) return ReadRecordChunks(file_name, decompressor, chunk_bytes, sink);
      if (ov && std::atoi(ov) == 1) return ReadRecordChunksOverlapped(file_name, decompressor, chunk_bytes, sink);
      return ReadRecordChunksPinned(file_name, decompressor, chunk_bytes, sink, [&](const char* data, std::size_t n) {
        // the chunk after this one is already in its page-locked buffer: its copy overlaps the counting of this one
        std::lock_guard<std::mutex> l(Device::Mu());
        kmsc_counter_prefetch(Device::Ctx(), c, data, static_cast<std::int64_t>(n));
      });
    };
    const bool timing = std::getenv("KMSC_TIMING") != nullptr;
    if (timing) {   // the process's first CUDA call (context start-up) is kept out of the phase times
      const auto t_init = std::chrono::steady_clock::now();
      std::lock_guard<std::mutex> l(Device::Mu());
      Device::Check(kmsc_ctx_sync(Device::Ctx()), "kmsc_ctx_sync");
      std::fprintf(stderr, "[kmsc timing] device context ready: %.3f s\n",
                   std::chrono::duration<double>(std::chrono::steady_clock::now() - t_init).count());
    }
    const auto t_all = std::chrono::steady_clock::now();
    double t_dev = 0, t_first = 0;
    int n_calls = 0;
    Status st = run([&](const char* data, std::size_t n) -> Status {
      std::lock_guard<std::mutex> l(Device::Mu());
      const auto t0 = std::chrono::steady_clock::now();
      const int rc = kmsc_counter_add_fasta(Device::Ctx(), c, data, static_cast<std::int64_t>(n));
      const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
      t_dev += dt;
      if (n_calls++ == 0) t_first = dt;
      if (rc == KMSC_E_FORMAT) return FailedPreconditionError(kmsc_last_error());
      if (rc != KMSC_OK) return InternalError(kmsc_last_error());
      return OkStatus();
    });
    if (timing)
      std::fprintf(stderr, "[kmsc timing] file -> counter: %.3f s in all, %.3f s inside %d kmsc_counter_add_fasta calls (the first, with "
                   "module loading and allocations: %.3f s)\n",
                   std::chrono::duration<double>(std::chrono::steady_clock::now() - t_all).count(), t_dev, n_calls, t_first);
    std::lock_guard<std::mutex> l(Device::Mu());
    if (!st.ok()) { kmsc_counter_free(Device::Ctx(), c); return st; }
    kmsc_set* s = nullptr;
    std::int64_t cut = 0, nd = 0;
    const int rc = kmsc_counter_finish(Device::Ctx(), c, cutoff, &s, &cut, &nd);
    kmsc_counter_free(Device::Ctx(), c);
    if (rc != KMSC_OK) return InternalError(kmsc_last_error());
    if (n_distinct) *n_distinct = nd;
    return std::make_pair(KmerSet<K, N, KeyType>(MakeSetPtr(s)), cut);
  }

 private:
  Status Count(const std::string& text, bool canonical, bool fasta) {
    kmsc_set* s = nullptr;
    std::int64_t cut = 0, nd = 0;
    std::lock_guard<std::mutex> l(Device::Mu());
    const int rc = (fasta ? kmsc_count_fasta : kmsc_count_reads)(
        Device::Ctx(), K, N, static_cast<int>(sizeof(KeyType)), text.data(), static_cast<std::int64_t>(text.size()),
        canonical ? 1 : 0, /*cutoff=*/1, &s, &cut, &nd);
    if (rc == KMSC_E_FORMAT) return FailedPreconditionError(kmsc_last_error());
    if (rc != KMSC_OK) return InternalError(kmsc_last_error());
    // mirror: distinct k-mers (ascending) + counts
    std::vector<std::int64_t> offs((std::size_t(1) << N) + 1);
    std::vector<KeyType> keys(static_cast<std::size_t>(nd));
    Device::Check(kmsc_set_to_csr(Device::Ctx(), s, offs.data(), keys.data()), "kmsc_set_to_csr");
    counts_.resize(static_cast<std::size_t>(nd));
    Device::Check(kmsc_count_last_counts(Device::Ctx(), counts_.data(), nd), "kmsc_count_last_counts");
    kmers_.resize(static_cast<std::size_t>(nd));
    constexpr int kb = 2 * K - N;
    for (std::size_t b = 0; b + 1 < offs.size(); b++)
      for (std::int64_t i = offs[b]; i < offs[b + 1]; i++)
        kmers_[static_cast<std::size_t>(i)] = (static_cast<std::uint64_t>(b) << kb) | static_cast<std::uint64_t>(keys[static_cast<std::size_t>(i)]);
    kmsc_set_free(Device::Ctx(), s);
    return OkStatus();
  }

  std::vector<std::uint64_t> kmers_;   // ascending distinct k-mers
  std::vector<ValueType> counts_;      // aligned counts, saturated at 255
};

}  // namespace kmsc
#endif
