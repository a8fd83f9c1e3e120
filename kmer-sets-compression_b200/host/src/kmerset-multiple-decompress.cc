// kmerset-multiple-decompress -- drop-in for the reference executable
// (src/kmerset-multiple-decompress.cc:27-117): reconstructs every original set from a
// dumped directory and prints its XOR hash and size, the end-to-end parity observable.
#include <string>

#include "flags.h"
#include "kmsc/kmer_set_set.h"

using namespace kmsc;
using namespace kmsc_cli;

template <int K, int N, typename KeyType>
int Main(const Flags& flags) {
  if (flags.positional.size() != 1) { Error("usage: kmerset-multiple-decompress [flags] <directory>"); return 1; }
  const int n_workers = flags.Int("workers", 1);
  if (MstReader<K, N, KeyType>::IsMstDirectory(flags.positional[0], flags.Str("extension", "txt"), flags.Str("decompressor", ""))) {
    // a directory written by `kmerset-multiple-compress --driver=mst`
    auto m = MstReader<K, N, KeyType>::FromDirectory(flags.positional[0], flags.Str("extension", "txt"),
                                                     flags.Str("decompressor", ""), flags.Bool("canonical", true));
    if (!m.ok()) { Error("failed to load the spanning tree: " + m.status().ToString()); return 1; }
    Info("kmer_set_set_reader.Size() = " + std::to_string(m.value().Size()));
    const int n = flags.Int("n", m.value().Size());
    for (int i = 0; i < n; i++) {
      auto s = m.value().Get(i, n_workers);
      if (!s.ok()) { Error("failed to construct kmer_set: " + s.status().ToString()); return 1; }
      Info("constructed kmer_set: i = " + std::to_string(i));
      Info("kmer_set.Hash() = " + std::to_string(s.value().Hash(n_workers)));
      Info("kmer_set.Size() = " + std::to_string(s.value().Size()));
    }
    return 0;
  }
  auto r = KmerSetSetReader<K, N, KeyType>::FromDirectory(flags.positional[0], flags.Str("extension", "txt"),
                                                          flags.Str("decompressor", ""), flags.Bool("canonical", true));
  if (!r.ok()) { Error("failed to load kmer_set_set_reader: " + r.status().ToString()); return 1; }
  Info("kmer_set_set_reader.Size() = " + std::to_string(r.value().Size()));
  // the reference reconstructs the sets it was given: the first ids (nodes added by merges follow)
  const int n = flags.Int("n", r.value().Size());
  for (int i = 0; i < n; i++) {
    auto s = r.value().Get(i, n_workers);
    if (!s.ok()) { Error("failed to construct kmer_set: " + s.status().ToString()); return 1; }
    Info("constructed kmer_set: i = " + std::to_string(i));
    Info("kmer_set.Hash() = " + std::to_string(s.value().Hash(n_workers)));
    Info("kmer_set.Size() = " + std::to_string(s.value().Size()));
  }
  return 0;
}

int main(int argc, char** argv) {
  const Flags flags = ParseFlags(argc, argv, {"debug", "canonical"});
  switch (flags.Int("k", 15)) {
    case 15: return Main<15, 14, std::uint16_t>(flags);
    case 19: return Main<19, 10, std::uint32_t>(flags);
    case 23: return Main<23, 14, std::uint32_t>(flags);
    case 31: return Main<31, 14, std::uint64_t>(flags);
    default: Error("unsupported k"); return 1;
  }
}
