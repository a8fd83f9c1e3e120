// kmerset-multiple-compress -- drop-in for the reference executable
// (src/kmerset-multiple-compress.cc:33-163): same flags (--k --workers --canonical
// --decompressor --compressor --out --extension --out_graph), same input (one SPSS text
// file per set) and the same output directory format (meta.<ext> + <i>.<ext>, DOT graph).
// Extra flags: --exact (all-bucket weights instead of the 2% sample), --seed (bucket sample),
// --bucket_ids_file (explicit sample, one id per line: replays a seeded reference run),
// --max_iterations, --driver=greedy|mst (mst: Kruskal over the exact symmetric-difference
// matrix + one difference-set pair per tree edge; its own directory layout, see DumpMst),
// --trace=<file> (weight matrix, merges / tree edges with the sizes and XOR hashes of the
// difference sets: what the parity tests compare with the oracle), --gpus=N (mst driver: the
// k-mer prefix space is sharded over N GPUs of this box, one host thread and one NCCL rank per
// GPU; the number of sets must be a multiple of N).
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <fstream>
#include <string>
#include <thread>
#include <vector>

#include "flags.h"
#include "kmsc/kmer_set_compact.h"
#include "kmsc/kmer_set_set.h"
#include "kmsc/multi_gpu.h"

using namespace kmsc;
using namespace kmsc_cli;

template <int K, int N, typename KeyType>
int Main(const Flags& flags) {
  const int n_workers = flags.Int("workers", 1);
  const bool canonical = flags.Bool("canonical", true);
  const std::string decompressor = flags.Str("decompressor", "");
  const std::vector<std::string>& files = flags.positional;
  std::vector<KmerSetCompact<K, N, KeyType>> sets(files.size());
  const auto t_load = std::chrono::steady_clock::now();
  {
    // the files are independent: --workers threads read and pack them (the reference loads them in turn)
    std::vector<std::string> errors(files.size());
    std::atomic<std::size_t> next{0};
    auto work = [&] {
      for (std::size_t i = next++; i < files.size(); i = next++) {
        auto r = KmerSetCompact<K, N, KeyType>::Load(files[i], decompressor);
        if (!r.ok()) errors[i] = r.status().ToString();
        else sets[i] = std::move(r).value();
      }
    };
    std::vector<std::thread> pool;
    const int n_threads = std::max(1, std::min<int>(n_workers, static_cast<int>(files.size())));
    for (int t = 1; t < n_threads; t++) pool.emplace_back(work);
    work();
    for (std::thread& t : pool) t.join();
    for (std::size_t i = 0; i < files.size(); i++) {
      Info("loading file: " + files[i]);
      if (!errors[i].empty()) { Error("failed to load file: " + errors[i]); return 1; }
    }
  }
  const double s_load = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_load).count();
  for (std::size_t i = 0; i < sets.size(); i++)
    Info("i = " + std::to_string(i) + ", size = " + std::to_string(sets[i].Size(n_workers)));
  const std::string trace = flags.Str("trace", "");
  const std::string out = flags.Str("out", "");
  if (flags.Str("driver", "greedy") == "mst") {
    Info("constructing the spanning tree");
    const int n_gpus = flags.Int("gpus", 1);
    MstResult<K, N, KeyType> r;
    if (n_gpus > 1) {
      MultiGpuMst<K, N, KeyType> mg = BuildMstMultiGpu<K, N, KeyType>(sets, canonical, n_gpus);
      if (!mg.error.empty()) { Error("multi-GPU spanning tree failed: " + mg.error); return 1; }
      r.W = std::move(mg.W);
      r.edges = mg.edges;
      for (std::size_t i = 0; i < r.edges.size(); i++) {
        r.add.push_back(KmerSet<K, N, KeyType>::FromSortedBits(std::move(mg.add[i])));
        r.del.push_back(KmerSet<K, N, KeyType>::FromSortedBits(std::move(mg.del[i])));
      }
    } else {
      {
        const auto t_init = std::chrono::steady_clock::now();
        Device::Check(kmsc_ctx_sync(Device::Ctx()), "kmsc_ctx_sync");
        Info("device context ready, seconds = " + std::to_string(std::chrono::duration<double>(std::chrono::steady_clock::now() - t_init).count()));
      }
      const auto t_dec = std::chrono::steady_clock::now();
      std::vector<KmerSet<K, N, KeyType>> ksets = KmerSetCompact<K, N, KeyType>::ToKmerSetBatch(sets, canonical);
      const auto t_mst = std::chrono::steady_clock::now();
      r = BuildMst<K, N, KeyType>(ksets);
      Device::Check(kmsc_ctx_sync(Device::Ctx()), "kmsc_ctx_sync");
      const auto t_end = std::chrono::steady_clock::now();
      Info("phases: load files = " + std::to_string(s_load) + " s, decode = " +
           std::to_string(std::chrono::duration<double>(t_mst - t_dec).count()) + " s, matrix + tree + splits = " +
           std::to_string(std::chrono::duration<double>(t_end - t_mst).count()) + " s");
    }
    Info("constructed the spanning tree, edges = " + std::to_string(r.edges.size()));
    if (!trace.empty()) {
      std::ofstream t(trace);
      const std::size_t n = sets.size();
      t << "driver mst\nn " << n << "\nW";
      for (std::int64_t w : r.W) t << ' ' << w;
      t << "\n";
      for (std::size_t i = 0; i < r.edges.size(); i++)
        t << "edge " << r.edges[i].parent << ' ' << r.edges[i].child << ' ' << r.edges[i].distance << ' '
          << r.add[i].Size() << ' ' << r.add[i].Hash(n_workers) << ' ' << r.del[i].Size() << ' ' << r.del[i].Hash(n_workers) << "\n";
    }
    if (!out.empty() && !sets.empty()) {
      const auto t_dump = std::chrono::steady_clock::now();
      Status st = DumpMst<K, N, KeyType>(r, sets[0], static_cast<int>(sets.size()), out, flags.Str("compressor", ""),
                                         flags.Str("extension", "txt"), canonical, n_workers);
      if (!st.ok()) { Error("failed to dump the spanning tree: " + st.ToString()); return 1; }
      Info("dumped the spanning tree (SPSS of the root and of both difference sets of every edge), seconds = " +
           std::to_string(std::chrono::duration<double>(std::chrono::steady_clock::now() - t_dump).count()));
    }
    return 0;
  }
  Info("constructing kmer_set_set");
  KmerSetSetOptions opt;
  opt.exact = flags.Bool("exact", false);
  opt.seed = static_cast<std::uint64_t>(flags.Int("seed", 0));
  opt.max_iterations = flags.Int("max_iterations", -1);
  const std::string ids_file = flags.Str("bucket_ids_file", "");
  if (!ids_file.empty()) {
    std::ifstream f(ids_file);
    int id;
    while (f >> id) opt.bucket_ids.push_back(id);
    if (opt.bucket_ids.empty()) { Error("no bucket ids in " + ids_file); return 1; }
  }
  const std::size_t n0 = sets.size();
  {
    // the process's first CUDA call (driver + context start-up, 0.5-3 s on an 8-GPU box) is not part of the
    // constructor: it is logged on its own line
    const auto t_init = std::chrono::steady_clock::now();
    Device::Check(kmsc_ctx_sync(Device::Ctx()), "kmsc_ctx_sync");
    Info("device context ready, seconds = " + std::to_string(std::chrono::duration<double>(std::chrono::steady_clock::now() - t_init).count()));
  }
  const auto t_ctor = std::chrono::steady_clock::now();
  KmerSetSet<K, N, KeyType> kss(std::move(sets), canonical, n_workers, opt);
  Info("constructed kmer_set_set, size = " + std::to_string(kss.Size()) + ", merges = " + std::to_string(kss.Merges().size()) +
       ", seconds = " + std::to_string(std::chrono::duration<double>(std::chrono::steady_clock::now() - t_ctor).count()));
  if (!trace.empty()) {
    std::ofstream t(trace);
    t << "driver greedy\nn " << n0 << "\nW";
    for (std::int64_t w : kss.InitialWeights()) t << ' ' << w;
    t << "\n";
    for (const auto& m : kss.Merges()) t << "merge " << std::get<0>(m) << ' ' << std::get<1>(m) << ' ' << std::get<2>(m) << "\n";
  }
  const std::string out_graph = flags.Str("out_graph", "");
  if (!out_graph.empty()) {
    Status st = kss.DumpGraph(out_graph);
    if (!st.ok()) { Error("failed to dump graph: " + st.ToString()); return 1; }
  }
  if (!out.empty()) {
    Status st = kss.Dump(out, flags.Str("compressor", ""), flags.Str("extension", "txt"), n_workers);
    if (!st.ok()) { Error("failed to dump kmer_set_set: " + st.ToString()); return 1; }
  }
  return 0;
}

int main(int argc, char** argv) {
  const Flags flags = ParseFlags(argc, argv, {"debug", "canonical", "exact"});
  switch (flags.Int("k", 15)) {  // the reference's instantiations (:149-157)
    case 15: return Main<15, 14, std::uint16_t>(flags);
    case 19: return Main<19, 10, std::uint32_t>(flags);
    case 23: return Main<23, 14, std::uint32_t>(flags);
    case 31: return Main<31, 14, std::uint64_t>(flags);
    default: Error("unsupported k"); return 1;
  }
}
