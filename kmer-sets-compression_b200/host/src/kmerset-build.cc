// kmerset-build -- drop-in for the reference executable (src/kmerset-build.cc:33-144):
// FASTA -> canonical k-mer counts -> cutoff -> KmerSet -> SPSS text file. Same flags
// (--k --workers --canonical --decompressor --compressor --cutoff --check --out).
#include <string>

#include "flags.h"
#include "kmsc/kmer_counter.h"
#include "kmsc/kmer_set_compact.h"

using namespace kmsc;
using namespace kmsc_cli;

template <int K, int N, typename KeyType>
int Main(const Flags& flags) {
  if (flags.positional.size() != 1) { Error("usage: kmerset-build [flags] <fasta>"); return 1; }
  const int n_workers = flags.Int("workers", 1);
  const bool canonical = flags.Bool("canonical", true);
  // the file is streamed in chunks of whole records (--chunk_mb, default 512) and counted
  // chunk by chunk on the device; the reference holds it in memory twice (kmer_counter.h:141-157)
  std::int64_t n_distinct = 0;
  const std::size_t chunk = flags.Int("chunk_bytes", 0) > 0 ? static_cast<std::size_t>(flags.Int("chunk_bytes", 0))
                                                            : static_cast<std::size_t>(flags.Int("chunk_mb", 512)) << 20;
  auto counted = KmerCounter<K, N, KeyType>::CountFileToKmerSet(flags.positional[0], flags.Str("decompressor", ""), canonical,
                                                                flags.Int("cutoff", 1), chunk, &n_distinct);
  if (!counted.ok()) { Error("failed to parse FASTA file: " + counted.status().ToString()); return 1; }
  KmerSet<K, N, KeyType>& kmer_set = counted.value().first;
  Info("constructed kmer_counter, size = " + std::to_string(n_distinct));
  Info("cutoff_count = " + std::to_string(counted.value().second));
  Info("kmer_set.Size() = " + std::to_string(kmer_set.Size()));
  Info("kmer_set.Hash() = " + std::to_string(kmer_set.Hash(n_workers)));
  Info("constructing kmer_set_compact");
  const KmerSetCompact<K, N, KeyType> compact = KmerSetCompact<K, N, KeyType>::FromKmerSet(kmer_set, canonical, true, n_workers);
  Info("kmer_set_compact.Size() = " + std::to_string(compact.Size(n_workers)));
  if (flags.Bool("check", false)) {
    if (kmer_set.Equals(compact.ToKmerSet(canonical, n_workers), n_workers)) Info("kmer_set_compact -> KmerSet: ok");
    else { Error("kmer_set_compact -> KmerSet: failed"); return 1; }
  }
  const std::string out = flags.Str("out", "");
  if (!out.empty()) {
    Status st = compact.Dump(out, flags.Str("compressor", ""), n_workers);
    if (!st.ok()) { Error("failed to dump kmer_set_compact: " + st.ToString()); return 1; }
  }
  return 0;
}

int main(int argc, char** argv) {
  const Flags flags = ParseFlags(argc, argv, {"debug", "canonical", "check"});
  switch (flags.Int("k", 15)) {
    case 15: return Main<15, 14, std::uint16_t>(flags);
    case 19: return Main<19, 10, std::uint32_t>(flags);
    case 23: return Main<23, 14, std::uint32_t>(flags);
    case 31: return Main<31, 14, std::uint64_t>(flags);
    default: Error("unsupported k"); return 1;
  }
}
