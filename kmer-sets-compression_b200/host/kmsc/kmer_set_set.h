// kmsc/kmer_set_set.h -- KmerSetSet / KmerSetSetReader with the reference's interface
// (lib/core/kmer_set_set.h:100-775) and on-disk formats, driving the GPU primitives.
//
// greedy driver (the reference's constructor, :109-427):
//   weights  W[i][j] = sum over the bucket sample of |S_i[b] & S_j[b]|   -> kmsc_pair_counts
//   loop     (j,k) = argmax W (ties: smallest (j,k), a legal refinement of the reference's
//            hash-map order); n = S_j & S_k, S_j -= n, S_k -= n                -> kmsc_pair_split
//            children[j] += n, children[k] += n; re-weight 3n-2 pairs            -> kmsc_pair_counts
//   stop     every `interval` iterations if the relative drop of the total SPSS weight
//            is <= 0.1 * interval / N0 (float arithmetic as :267-302), or max weight 0.
// Sets stay resident on the device between iterations (the reference decodes SPSS text
// to hash sets and re-encodes three SPSS per iteration); SPSS is built on the host only
// for the nodes that changed (needed by the stop rule) and at Dump time.
//
// mst driver (north-star variant, no counterpart in the reference): exact all-bucket
// matrix, Kruskal over d(i,j) = |S_i| + |S_j| - 2 W[i][j] with ParallelDisjointSet, then
// per tree edge the two difference sets.
#ifndef KMSC_HOST_KMER_SET_SET_H_
#define KMSC_HOST_KMER_SET_SET_H_
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <filesystem>
#include <map>
#include <queue>
#include <random>
#include <sstream>
#include <string>
#include <tuple>
#include <utility>
#include <vector>

#include "kmsc/io.h"
#include "kmsc/kmer_set.h"
#include "kmsc/kmer_set_compact.h"
#include "kmsc/parallel_disjoint_set.h"

namespace kmsc {

namespace internal {

using AdjacencyList = std::map<int, std::vector<int>>;  // ordered: deterministic meta files

// "<#keys> {<key> <#children> <child>...}" with single spaces (reference :45-58)
inline std::string SerializeAdjacencyList(const AdjacencyList& a) {
  std::stringstream ss;
  ss << a.size();
  for (const auto& p : a) {
    ss << ' ' << p.first << ' ' << p.second.size();
    for (int c : p.second) ss << ' ' << c;
  }
  return ss.str();
}

inline AdjacencyList DeserializeAdjacencyList(const std::string& s) {
  std::stringstream ss(s);
  AdjacencyList a;
  std::size_t size = 0;
  ss >> size;
  for (std::size_t i = 0; i < size; i++) {
    int key = 0;
    std::size_t n = 0;
    ss >> key >> n;
    std::vector<int> v(n);
    for (std::size_t j = 0; j < n; j++) ss >> v[j];
    a[key] = std::move(v);
  }
  return a;
}

}  // namespace internal

// n unique sorted ints in [min, max] (reference lib/core/random.h:13-41); seed = 0 draws
// from std::random_device like the reference's InsecureBitGen.
inline std::vector<int> GetRandomInts(int n, int min, int max, std::uint64_t seed = 0) {
  std::mt19937_64 gen(seed ? seed : std::random_device{}());
  std::uniform_int_distribution<int> d(min, max);
  std::vector<int> v;
  std::vector<bool> seen(static_cast<std::size_t>(max - min + 1), false);
  while (static_cast<int>(v.size()) < n && static_cast<int>(v.size()) < max - min + 1) {
    const int x = d(gen);
    if (!seen[static_cast<std::size_t>(x - min)]) { seen[static_cast<std::size_t>(x - min)] = true; v.push_back(x); }
  }
  std::sort(v.begin(), v.end());
  return v;
}

struct KmerSetSetOptions {
  std::vector<int> bucket_ids;   // empty = draw (1 << N) / 50 random buckets like the reference
  bool exact = false;            // true = all buckets (exact weights)
  std::uint64_t seed = 0;        // bucket sample seed (0 = nondeterministic, like the reference)
  int max_iterations = -1;       // -1 = until the reference's stop rule fires
};

template <int K, int N, typename KeyType>
class KmerSetSet {
 public:
  using Set = KmerSet<K, N, KeyType>;
  using Compact = KmerSetCompact<K, N, KeyType>;

  KmerSetSet() = default;

  KmerSetSet(std::vector<Compact> kmer_sets_compact, bool canonical, int n_workers,
             const KmerSetSetOptions& opt = KmerSetSetOptions())
      : kmer_sets_compact_(std::move(kmer_sets_compact)) {
    const int n0 = static_cast<int>(kmer_sets_compact_.size());
    std::vector<int> bucket_ids = opt.bucket_ids;
    if (!opt.exact && bucket_ids.empty()) bucket_ids = GetRandomInts((1 << N) / 50, 0, (1 << N) - 1, opt.seed);
    const std::vector<std::int32_t> ids(bucket_ids.begin(), bucket_ids.end());

    // KMSC_TIMING=1: seconds per phase on stderr when the constructor returns
    const bool timing = std::getenv("KMSC_TIMING") != nullptr;
    double t_phase[5] = {0, 0, 0, 0, 0};  // decode, weights, split, SPSS re-encode, row re-weights
    auto now = [] { return std::chrono::steady_clock::now(); };
    auto since = [&](std::chrono::steady_clock::time_point t0) { return std::chrono::duration<double>(now() - t0).count(); };
    auto t0 = now();

    // resident device sets (the reference re-decodes SPSS on every use)
    std::vector<Set> sets = Compact::ToKmerSetBatch(kmer_sets_compact_, canonical);   // one batched decode
    t_phase[0] = since(t0); t0 = now();

    std::vector<std::int64_t> W = PairCounts(sets, opt.exact ? nullptr : &ids);  // dense n x n
    t_phase[1] = since(t0);
    initial_weights_ = W;
    bucket_ids_ = bucket_ids;
    int n = n0;
    merges_.clear();

    std::int64_t total_spss_weight = TotalWeight();
    const int interval = n0 / 8 + 1;
    const float improvement_threshold = 0.1 * interval / static_cast<std::size_t>(n0);

    for (int it = 0; opt.max_iterations < 0 || it < opt.max_iterations; it++) {
      if (it > 0 && it % interval == 0) {
        const std::int64_t updated = TotalWeight();
        const float improvement = static_cast<float>(total_spss_weight - updated) / total_spss_weight;
        if (improvement <= improvement_threshold) break;
        total_spss_weight = updated;
      }
      // argmax, strict '>' in ascending (j, k) order
      std::int64_t weight = 0;
      int j = -1, k = -1;
      for (int a = 0; a < n; a++)
        for (int b = a + 1; b < n; b++)
          if (W[static_cast<std::size_t>(a) * n + b] > weight) { weight = W[static_cast<std::size_t>(a) * n + b]; j = a; k = b; }
      if (weight == 0) break;
      merges_.push_back({j, k, weight});

      t0 = now();
      Set inter, jm, km;
      Set::Split(sets[j], sets[k], &inter, &jm, &km, opt.exact ? weight : -1);  // exact weight = |S_j & S_k|
      sets[j] = jm;
      sets[k] = km;
      sets.push_back(inter);
      t_phase[2] += since(t0); t0 = now();
      kmer_sets_compact_[j] = Compact::FromKmerSet(sets[j], canonical, true, n_workers);
      kmer_sets_compact_[k] = Compact::FromKmerSet(sets[k], canonical, true, n_workers);
      kmer_sets_compact_.push_back(Compact::FromKmerSet(inter, canonical, true, n_workers));
      t_phase[3] += since(t0); t0 = now();
      children_[j].push_back(n);
      children_[k].push_back(n);
      n += 1;

      // re-weight: rows j, k and the new node n-1 against everything (3n-2 pairs in the
      // reference, :385-425); the device recomputes the three rows in one call
      const std::vector<std::int64_t> rows = PairCountRows(sets, {j, k, n - 1}, opt.exact ? nullptr : &ids);
      std::vector<std::int64_t> W2(static_cast<std::size_t>(n) * n, 0);
      for (int a = 0; a < n - 1; a++)
        for (int b = 0; b < n - 1; b++) W2[static_cast<std::size_t>(a) * n + b] = W[static_cast<std::size_t>(a) * (n - 1) + b];
      const int changed[3] = {j, k, n - 1};
      for (int r = 0; r < 3; r++)
        for (int l = 0; l < n; l++) {
          const std::int64_t v = rows[static_cast<std::size_t>(r) * n + l];
          W2[static_cast<std::size_t>(changed[r]) * n + l] = v;
          W2[static_cast<std::size_t>(l) * n + changed[r]] = v;
        }
      W.swap(W2);
      t_phase[4] += since(t0);
    }
    if (timing)
      std::fprintf(stderr, "[kmsc timing] decode %.4f s, initial weights %.4f s, %zu merges: split %.4f s, SPSS re-encode %.4f s, "
                   "row re-weights %.4f s\n", t_phase[0], t_phase[1], merges_.size(), t_phase[2], t_phase[3], t_phase[4]);
  }

  int Size() const { return static_cast<int>(kmer_sets_compact_.size()); }

  // union of every node reachable from i (reference :433-454), merged on the device
  Set Get(int i, bool canonical, int n_workers) const {
    std::vector<Set> parts;
    std::queue<int> q;
    q.push(i);
    while (!q.empty()) {
      const int cur = q.front();
      q.pop();
      parts.push_back(kmer_sets_compact_[static_cast<std::size_t>(cur)].ToKmerSet(canonical, n_workers));
      auto itc = children_.find(cur);
      if (itc != children_.end())
        for (int c : itc->second) q.push(c);
    }
    return Union(parts);
  }

  Status Dump(const std::string& directory_name, const std::string& compressor, const std::string& extension,
              int n_workers) {
    try {
      std::filesystem::create_directories(directory_name);
    } catch (...) {
      return InternalError("failed to create a directory");
    }
    {
      std::vector<std::string> v;
      v.push_back(internal::SerializeAdjacencyList(children_));
      v.push_back(std::to_string(kmer_sets_compact_.size()));
      const std::string f = (std::filesystem::path(directory_name) / ("meta." + extension)).string();
      Status st = WriteLines(f, compressor, v);
      if (!st.ok()) return st;
    }
    int fail_count = 0;
    for (std::size_t i = 0; i < kmer_sets_compact_.size(); i++) {
      const std::string f = (std::filesystem::path(directory_name) / (std::to_string(i) + "." + extension)).string();
      if (!kmer_sets_compact_[i].Dump(f, compressor, n_workers).ok()) fail_count++;
    }
    if (fail_count > 0) return InternalError("failed to write " + std::to_string(fail_count) + " files");
    return OkStatus();
  }

  Status DumpGraph(const std::string& file_name) const {
    std::vector<std::string> lines;
    lines.emplace_back("digraph G {");
    for (const auto& p : children_)
      for (int c : p.second) lines.push_back("v" + std::to_string(p.first) + " -> v" + std::to_string(c));
    lines.emplace_back("}");
    return WriteLines(file_name, "", lines);
  }

  static StatusOr<KmerSetSet> Load(const std::string& directory_name, const std::string& decompressor,
                                   const std::string& extension, int /*n_workers*/) {
    const std::string meta = (std::filesystem::path(directory_name) / ("meta." + extension)).string();
    StatusOr<std::vector<std::string>> lines = ReadLines(meta, decompressor);
    if (!lines.ok()) return lines.status();
    if (lines.value().size() < 2) return InternalError("malformed meta file");
    KmerSetSet out;
    out.children_ = internal::DeserializeAdjacencyList(lines.value()[0]);
    const int n = std::atoi(lines.value()[1].c_str());
    out.kmer_sets_compact_.resize(static_cast<std::size_t>(n));
    int n_fail = 0;
    for (int i = 0; i < n; i++) {
      const std::string f = (std::filesystem::path(directory_name) / (std::to_string(i) + "." + extension)).string();
      StatusOr<Compact> c = Compact::Load(f, decompressor);
      if (!c.ok()) { n_fail++; continue; }
      out.kmer_sets_compact_[static_cast<std::size_t>(i)] = std::move(c).value();
    }
    if (n_fail > 0) return InternalError("failed to dump " + std::to_string(n_fail) + " files");
    return out;
  }

  // (j, k, weight) of every merge, in order: the observable the reference only logs (:324)
  const std::vector<std::tuple<int, int, std::int64_t>>& Merges() const { return merges_; }
  const internal::AdjacencyList& Children() const { return children_; }
  // the initial n0 x n0 weight matrix (reference :187-219) and the bucket sample it was taken over
  const std::vector<std::int64_t>& InitialWeights() const { return initial_weights_; }
  const std::vector<int>& BucketIds() const { return bucket_ids_; }

  // ---- device helpers shared with the mst driver -------------------------------------------
  static std::vector<std::int64_t> PairCounts(const std::vector<Set>& sets, const std::vector<std::int32_t>* ids) {
    const int n = static_cast<int>(sets.size());
    std::vector<const kmsc_set*> h(static_cast<std::size_t>(n));
    for (int i = 0; i < n; i++) h[static_cast<std::size_t>(i)] = sets[static_cast<std::size_t>(i)].Dev()->set;
    std::vector<std::int64_t> W(static_cast<std::size_t>(n) * n, 0);
    std::lock_guard<std::mutex> l(Device::Mu());
    Device::Check(kmsc_pair_counts(Device::Ctx(), h.data(), n, ids ? ids->data() : nullptr,
                                   ids ? static_cast<std::int32_t>(ids->size()) : 0, W.data(), nullptr), "kmsc_pair_counts");
    return W;
  }
  static std::vector<std::int64_t> PairCountRows(const std::vector<Set>& sets, const std::vector<std::int32_t>& rows,
                                                 const std::vector<std::int32_t>* ids) {
    const int n = static_cast<int>(sets.size());
    std::vector<const kmsc_set*> h(static_cast<std::size_t>(n));
    for (int i = 0; i < n; i++) h[static_cast<std::size_t>(i)] = sets[static_cast<std::size_t>(i)].Dev()->set;
    std::vector<std::int64_t> out(rows.size() * static_cast<std::size_t>(n), 0);
    std::lock_guard<std::mutex> l(Device::Mu());
    Device::Check(kmsc_pair_counts_rows(Device::Ctx(), h.data(), n, rows.data(), static_cast<std::int32_t>(rows.size()),
                                        ids ? ids->data() : nullptr, ids ? static_cast<std::int32_t>(ids->size()) : 0,
                                        out.data()), "kmsc_pair_counts_rows");
    return out;
  }
  static Set Union(const std::vector<Set>& parts) {
    std::vector<const kmsc_set*> h(parts.size());
    for (std::size_t i = 0; i < parts.size(); i++) h[i] = parts[i].Dev()->set;
    kmsc_set* u = nullptr;
    std::lock_guard<std::mutex> l(Device::Mu());
    Device::Check(kmsc_set_union(Device::Ctx(), h.data(), static_cast<std::int32_t>(h.size()), &u), "kmsc_set_union");
    return Set(MakeSetPtr(u));
  }

 private:
  std::int64_t TotalWeight() const {
    std::int64_t t = 0;
    for (const Compact& c : kmer_sets_compact_) t += c.Weight();
    return t;
  }

  internal::AdjacencyList children_;
  std::vector<Compact> kmer_sets_compact_;
  std::vector<std::tuple<int, int, std::int64_t>> merges_;
  std::vector<std::int64_t> initial_weights_;
  std::vector<int> bucket_ids_;
};

// Lazy reader over a dumped directory (reference :622-775)
template <int K, int N, typename KeyType>
class KmerSetSetReader {
 public:
  KmerSetSetReader() = default;

  static StatusOr<KmerSetSetReader> FromDirectory(std::string directory_name, std::string extension,
                                                  std::string decompressor, bool canonical) {
    const std::string meta = (std::filesystem::path(directory_name) / ("meta." + extension)).string();
    StatusOr<std::vector<std::string>> v = ReadLines(meta, decompressor);
    if (!v.ok()) return v.status();
    if (v.value().size() < 2) return InternalError("malformed meta file");
    KmerSetSetReader r;
    r.directory_name_ = std::move(directory_name);
    r.extension_ = std::move(extension);
    r.decompressor_ = std::move(decompressor);
    r.canonical_ = canonical;
    r.children_ = internal::DeserializeAdjacencyList(v.value()[0]);
    r.size_ = std::atoi(v.value()[1].c_str());
    return r;
  }

  int Size() const { return size_; }

  StatusOr<KmerSet<K, N, KeyType>> Get(int i, int n_workers) const {
    std::vector<int> ids;
    std::queue<int> q;
    q.push(i);
    while (!q.empty()) {
      const int cur = q.front();
      q.pop();
      ids.push_back(cur);
      auto it = children_.find(cur);
      if (it == children_.end()) continue;
      for (int c : it->second) q.push(c);
    }
    // a node is read, packed and decoded once: its device set stays with the reader (merge nodes are shared by
    // many of the original sets; the reference re-reads them for every Get)
    std::vector<KmerSet<K, N, KeyType>> parts;
    int n_fail = 0;
    for (int id : ids) {
      auto hit = cache_.find(id);
      if (hit == cache_.end()) {
        const std::string f = (std::filesystem::path(directory_name_) / (std::to_string(id) + "." + extension_)).string();
        StatusOr<KmerSetCompact<K, N, KeyType>> c = KmerSetCompact<K, N, KeyType>::Load(f, decompressor_);
        if (!c.ok()) { n_fail++; continue; }
        hit = cache_.emplace(id, c.value().ToKmerSet(canonical_, n_workers)).first;
      }
      parts.push_back(hit->second);
    }
    if (n_fail > 0) return InternalError("failed to load data from " + std::to_string(n_fail) + " files");
    return KmerSetSet<K, N, KeyType>::Union(parts);
  }

 private:
  std::string directory_name_, extension_, decompressor_;
  bool canonical_ = false;
  internal::AdjacencyList children_;
  int size_ = 0;
  mutable std::map<int, KmerSet<K, N, KeyType>> cache_;   // node id -> device set
};

// ---- mst driver ---------------------------------------------------------------------------
struct MstEdge;
inline std::vector<MstEdge> MstTree(const std::vector<std::int64_t>& W, int n);
struct MstEdge {
  int parent, child;
  std::int64_t distance;  // |S_p ^ S_c| (symmetric difference)
};

template <int K, int N, typename KeyType>
struct MstResult {
  std::vector<std::int64_t> W;            // exact n x n intersection matrix
  std::vector<MstEdge> edges;             // n - 1 tree edges, oriented away from root 0
  std::vector<KmerSet<K, N, KeyType>> add;  // per edge: S_c \ S_p
  std::vector<KmerSet<K, N, KeyType>> del;  // per edge: S_p \ S_c
};

// Minimum spanning tree of the symmetric-difference graph: edges sorted by
// (d ascending, i ascending, j ascending), Kruskal with ParallelDisjointSet, tree
// oriented by BFS from set 0, the difference sets of all edges from one batched device split.
// The tree itself, from an exact n x n intersection matrix (diagonal = set sizes): candidate edges by
// (d ascending, i ascending, j ascending), Kruskal with ParallelDisjointSet, oriented breadth-first from
// set 0 with neighbours in ascending order. Shared by the single-GPU and the multi-GPU driver.
inline std::vector<MstEdge> MstTree(const std::vector<std::int64_t>& W, int n) {
  std::vector<MstEdge> edges;
  std::vector<std::tuple<std::int64_t, int, int>> cand;
  for (int i = 0; i < n; i++)
    for (int j = i + 1; j < n; j++) {
      const std::int64_t d = W[static_cast<std::size_t>(i) * n + i] + W[static_cast<std::size_t>(j) * n + j] -
                             2 * W[static_cast<std::size_t>(i) * n + j];
      cand.emplace_back(d, i, j);
    }
  std::sort(cand.begin(), cand.end());
  ParallelDisjointSet dsu(n);
  std::vector<std::vector<std::pair<int, std::int64_t>>> adj(static_cast<std::size_t>(n));
  for (const auto& e : cand) {
    const int i = std::get<1>(e), j = std::get<2>(e);
    if (dsu.IsSame(i, j)) continue;
    dsu.Unite(i, j);
    adj[static_cast<std::size_t>(i)].emplace_back(j, std::get<0>(e));
    adj[static_cast<std::size_t>(j)].emplace_back(i, std::get<0>(e));
  }
  std::vector<bool> seen(static_cast<std::size_t>(n), false);
  std::queue<int> q;
  if (n > 0) { q.push(0); seen[0] = true; }
  while (!q.empty()) {
    const int p = q.front();
    q.pop();
    std::sort(adj[static_cast<std::size_t>(p)].begin(), adj[static_cast<std::size_t>(p)].end());
    for (const auto& e : adj[static_cast<std::size_t>(p)]) {
      if (seen[static_cast<std::size_t>(e.first)]) continue;
      seen[static_cast<std::size_t>(e.first)] = true;
      edges.push_back({p, e.first, e.second});
      q.push(e.first);
    }
  }
  return edges;
}

template <int K, int N, typename KeyType>
MstResult<K, N, KeyType> BuildMst(const std::vector<KmerSet<K, N, KeyType>>& sets) {
  using Set = KmerSet<K, N, KeyType>;
  MstResult<K, N, KeyType> r;
  const int n = static_cast<int>(sets.size());
  r.W = KmerSetSet<K, N, KeyType>::PairCounts(sets, nullptr);
  r.edges = MstTree(r.W, n);
  // the difference sets of all n - 1 tree edges in one streaming device pass; the exact matrix
  // gives every |S_p & S_c|, so the outputs are allocated exactly and written directly
  std::vector<const Set*> js, ks;
  std::vector<std::int64_t> inter_sizes;
  for (const MstEdge& e : r.edges) {
    js.push_back(&sets[static_cast<std::size_t>(e.parent)]);
    ks.push_back(&sets[static_cast<std::size_t>(e.child)]);
    inter_sizes.push_back(r.W[static_cast<std::size_t>(e.parent) * n + e.child]);
  }
  if (!r.edges.empty()) Set::SplitBatch(js, ks, inter_sizes, nullptr, &r.del, &r.add);
  return r;
}

// ---- on-disk form of an mst result (no counterpart in the reference; text like its formats) ----
// meta.<ext>: line 1 "mst <n>", then one line "<parent> <child> <distance>" per tree edge in
// BuildMst order; 0.<ext> = SPSS of the root set; <child>.add.<ext> / <child>.del.<ext> = SPSS of
// S_child \ S_parent and S_parent \ S_child.
template <int K, int N, typename KeyType>
Status DumpMst(const MstResult<K, N, KeyType>& r, const KmerSetCompact<K, N, KeyType>& root, int n_sets,
               const std::string& directory_name, const std::string& compressor, const std::string& extension,
               bool canonical, int n_workers) {
  using Compact = KmerSetCompact<K, N, KeyType>;
  try {
    std::filesystem::create_directories(directory_name);
  } catch (...) {
    return InternalError("failed to create a directory");
  }
  const std::filesystem::path dir(directory_name);
  std::vector<std::string> meta;
  meta.push_back("mst " + std::to_string(n_sets));
  for (const MstEdge& e : r.edges)
    meta.push_back(std::to_string(e.parent) + " " + std::to_string(e.child) + " " + std::to_string(e.distance));
  Status st = WriteLines((dir / ("meta." + extension)).string(), compressor, meta);
  if (!st.ok()) return st;
  st = root.Dump((dir / ("0." + extension)).string(), compressor, n_workers);
  if (!st.ok()) return st;
  int fail_count = 0;
  for (std::size_t i = 0; i < r.edges.size(); i++) {
    const std::string c = std::to_string(r.edges[i].child);
    if (!Compact::FromKmerSet(r.add[i], canonical, true, n_workers).Dump((dir / (c + ".add." + extension)).string(), compressor, n_workers).ok()) fail_count++;
    if (!Compact::FromKmerSet(r.del[i], canonical, true, n_workers).Dump((dir / (c + ".del." + extension)).string(), compressor, n_workers).ok()) fail_count++;
  }
  if (fail_count > 0) return InternalError("failed to write " + std::to_string(fail_count) + " files");
  return OkStatus();
}

template <int K, int N, typename KeyType>
class MstReader {
 public:
  static bool IsMstDirectory(const std::string& directory_name, const std::string& extension, const std::string& decompressor) {
    StatusOr<std::vector<std::string>> v = ReadLines((std::filesystem::path(directory_name) / ("meta." + extension)).string(), decompressor);
    return v.ok() && !v.value().empty() && v.value()[0].rfind("mst ", 0) == 0;
  }
  static StatusOr<MstReader> FromDirectory(std::string directory_name, std::string extension, std::string decompressor,
                                           bool canonical) {
    StatusOr<std::vector<std::string>> v = ReadLines((std::filesystem::path(directory_name) / ("meta." + extension)).string(), decompressor);
    if (!v.ok()) return v.status();
    if (v.value().empty() || v.value()[0].rfind("mst ", 0) != 0) return InternalError("malformed meta file");
    MstReader r;
    r.size_ = std::atoi(v.value()[0].c_str() + 4);
    r.parent_.assign(static_cast<std::size_t>(r.size_), -1);
    for (std::size_t i = 1; i < v.value().size(); i++) {
      std::stringstream ss(v.value()[i]);
      int p = -1, c = -1;
      std::int64_t d = 0;
      if (!(ss >> p >> c >> d)) continue;
      if (c < 0 || c >= r.size_ || p < 0 || p >= r.size_) return InternalError("malformed meta file");
      r.parent_[static_cast<std::size_t>(c)] = p;
    }
    r.directory_name_ = std::move(directory_name);
    r.extension_ = std::move(extension);
    r.decompressor_ = std::move(decompressor);
    r.canonical_ = canonical;
    return r;
  }
  int Size() const { return size_; }
  // S_i = ((S_root \ del) | add) applied along the tree path from the root to i
  StatusOr<KmerSet<K, N, KeyType>> Get(int i, int n_workers) const {
    using Compact = KmerSetCompact<K, N, KeyType>;
    // walk up to the nearest set this reader has already reconstructed (the root at the latest), then apply the
    // edges of the path downwards; every set on the way stays with the reader, so reconstructing all n sets
    // reads every file once
    std::vector<int> path;
    int c = i;
    while (c != 0 && cache_.find(c) == cache_.end()) {
      if (c < 0 || c >= size_ || path.size() > parent_.size()) return InternalError("malformed tree");
      path.push_back(c);
      c = parent_[static_cast<std::size_t>(c)];
    }
    const std::filesystem::path dir(directory_name_);
    if (c == 0 && cache_.find(0) == cache_.end()) {
      StatusOr<Compact> root = Compact::Load((dir / ("0." + extension_)).string(), decompressor_);
      if (!root.ok()) return root.status();
      cache_.emplace(0, root.value().ToKmerSet(canonical_, n_workers));
    }
    KmerSet<K, N, KeyType> s = cache_.at(c);
    for (auto it = path.rbegin(); it != path.rend(); ++it) {
      const std::string name = std::to_string(*it);
      StatusOr<Compact> add = Compact::Load((dir / (name + ".add." + extension_)).string(), decompressor_);
      StatusOr<Compact> del = Compact::Load((dir / (name + ".del." + extension_)).string(), decompressor_);
      if (!add.ok() || !del.ok()) return InternalError("failed to load data from 1 files");
      s.Sub(del.value().ToKmerSet(canonical_, n_workers), n_workers);
      s.Add(add.value().ToKmerSet(canonical_, n_workers), n_workers);
      cache_.emplace(*it, s);
    }
    return s;
  }

 private:
  std::string directory_name_, extension_, decompressor_;
  bool canonical_ = false;
  int size_ = 0;
  std::vector<int> parent_;
  mutable std::map<int, KmerSet<K, N, KeyType>> cache_;   // set id -> reconstructed device set
};

}  // namespace kmsc
#endif
