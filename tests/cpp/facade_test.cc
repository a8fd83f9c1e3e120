// facade_test.cc -- the reference's own gtest cases restated against the kmsc C++17
// facade (same class and method names), run on a GPU by tests/test_gpu_facade.py.
// Mirrors test/kmer.cc, test/kmer_counter.cc, test/kmer_set.cc, test/kmer_set_compact.cc,
// test/kmer_set_set.cc, test/parallel_disjoint_set.cc, test/spss.cc of the reference.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <filesystem>
#include <random>
#include <string>
#include <thread>
#include <vector>

#include "kmsc/kmer.h"
#include "kmsc/kmer_counter.h"
#include "kmsc/kmer_set.h"
#include "kmsc/kmer_set_compact.h"
#include "kmsc/kmer_set_set.h"
#include "kmsc/parallel_disjoint_set.h"
#include "kmsc/spss.h"

using namespace kmsc;

static int g_fail = 0;
#define CHECK(cond)                                                              \
  do {                                                                           \
    if (!(cond)) { std::printf("FAIL %s:%d: %s\n", __FILE__, __LINE__, #cond); g_fail++; } \
  } while (0)

static std::mt19937_64 rng(12345);

template <int K>
static std::string RandomRead() {  // like lib/random.h:35-50: 1..100 random k-mers, half of them doubled
  std::string s;
  const int n = 1 + static_cast<int>(rng() % 100);
  for (int j = 0; j < n; j++)
    for (int i = 0; i < K; i++) s += "ACGT"[rng() & 3];
  if (rng() & 1) s += s;
  return s;
}

template <int K, int N, typename KeyType>
static KmerSet<K, N, KeyType> RandomKmerSet(int n, bool canonical) {
  std::vector<std::uint64_t> v;
  while (static_cast<int>(v.size()) < n) {
    const std::string s = RandomRead<K>();
    for (std::size_t j = 0; j + K <= s.size() && static_cast<int>(v.size()) < n * 2; j++) {
      Kmer<K> k(s.substr(j, K));
      if (canonical) k = k.Canonical();
      v.push_back(k.Bits());
    }
    std::sort(v.begin(), v.end());
    v.erase(std::unique(v.begin(), v.end()), v.end());
  }
  v.resize(static_cast<std::size_t>(n));
  return KmerSet<K, N, KeyType>::FromSortedBits(std::move(v));
}

static void TestKmer() {  // test/kmer.cc
  CHECK(Kmer<5>("AGCTG").String() == "AGCTG");
  CHECK(Kmer<5>("AAAAT").Canonical().String() == "AAAAT");
  CHECK(Kmer<5>("TTTTA").Canonical().String() == "TAAAA");
  CHECK(Kmer<5>("CCCCG").Canonical().String() == "CCCCG");
  CHECK(Kmer<5>("GGGGC").Canonical().String() == "GCCCC");
  CHECK(Kmer<5>("AGCTA").Complement().String() == "TAGCT");
  CHECK(Kmer<5>("AGCTG").Next('C').String() == "GCTGC");
  CHECK(Kmer<5>("AGCTG").Prev('C').String() == "CAGCT");
  CHECK(Complement("ACGTT") == "AACGT");  // test/spss.cc:13
  for (int i = 0; i < 1000; i++) {
    Kmer<31> k(rng() & Kmer<31>::Mask());
    CHECK(k.Complement().Complement() == k);
    CHECK(Kmer<31>(k.String()) == k);
    int b; std::uint64_t key;
    std::tie(b, key) = GetBucketAndKeyFromKmer<31, 14, std::uint64_t>(k);
    CHECK((GetKmerFromBucketAndKey<31, 14, std::uint64_t>(b, key) == k));
  }
}

static void TestKmerSet() {  // test/kmer_set.cc
  constexpr int K = 5, N = 3;
  using KeyType = std::uint16_t;
  KmerSet<K, N, KeyType> s;
  s.Add(Kmer<K>("AAAAA"));
  s.Add(Kmer<K>("CCCCC"));
  s.Add(Kmer<K>("AAAAA"));
  CHECK(s.Size() == 2);
  CHECK(s.Contains(Kmer<K>("AAAAA")) && !s.Contains(Kmer<K>("GGGGG")));
  s.Remove(Kmer<K>("AAAAA"));
  CHECK(s.Size() == 1 && !s.Contains(Kmer<K>("AAAAA")));
  CHECK(s.Find([](const Kmer<K>& k) { return k.String()[1] == 'C'; }, 1).size() == 1);

  KmerSet<K, N, KeyType> s1, s2, s3;
  for (const char* x : {"AAAAA", "TTTTT", "CCCCC"}) s1.Add(Kmer<K>(x));
  for (const char* x : {"AAAAA", "TTTTT", "GGGGG"}) s2.Add(Kmer<K>(x));
  CHECK(Add(s1, s2, 1).Size() == 4);
  CHECK(Sub(s1, s2, 1).Size() == 1);
  CHECK(Sub(s2, s1, 1).Size() == 1);
  CHECK(Intersection(s2, s1, 1).Size() == 2);
  CHECK(s1.Size() == 3 && s2.Size() == 3);  // value semantics: operands untouched
  for (const char* x : {"AAAAA", "CCCCC", "GGGGG"}) s3.Add(Kmer<K>(x));
  KmerSet<K, N, KeyType> s1b;
  for (const char* x : {"CCCCC", "AAAAA", "TTTTT"}) s1b.Add(Kmer<K>(x));
  CHECK(s1.Equals(s1b, 1) && s1b.Equals(s1, 1));
  CHECK(!s1.Equals(s3, 1) && !s3.Equals(s1, 1));
  CHECK(s1.Diff(s3, 1) == 2);
  CHECK(s1.Hash(1) == s1b.Hash(1));
}

static void TestKmerCounter() {  // test/kmer_counter.cc
  CHECK(AddWithMax<std::uint8_t>(255, 1) == 255);
  constexpr int K = 5, N = 3;
  using KeyType = std::uint16_t;
  {
    KmerCounter<K, N, KeyType> c;
    c.Add(Kmer<K>("AAAAA"), 1).Add(Kmer<K>("CCCCC"), 2).Add(Kmer<K>("TTTTT"), 3).Add(Kmer<K>("AAAAA"), 1);
    CHECK(c.Get(Kmer<K>("AAAAA")) == 2 && c.Get(Kmer<K>("CCCCC")) == 2 && c.Get(Kmer<K>("TTTTT")) == 3);
  }
  {
    KmerCounter<K, N, KeyType> c;
    c.Add(Kmer<K>("AAAAA"), 3).Add(Kmer<K>("CCCCC"), 1).Add(Kmer<K>("GGGGG"), 2).Add(Kmer<K>("TTTTT"), 4);
    KmerSet<K, N, KeyType> set;
    std::int64_t cut;
    std::tie(set, cut) = c.ToKmerSet(3, 1);
    CHECK(cut == 2 && set.Size() == 2 && set.Contains(Kmer<K>("AAAAA")) && set.Contains(Kmer<K>("TTTTT")));
  }
  {
    KmerCounter<K, N, KeyType> c = KmerCounter<K, N, KeyType>::FromReads({"AACCGTT", "AACCGTA"}, false, 1);
    CHECK(c.Get(Kmer<K>("AACCG")) == 2 && c.Get(Kmer<K>("ACCGT")) == 2);
    CHECK(c.Get(Kmer<K>("CCGTT")) == 1 && c.Get(Kmer<K>("CCGTA")) == 1 && c.Size() == 4);
  }
  {
    auto bad = KmerCounter<K, N, KeyType>::FromFASTA(std::vector<std::string>{">a", "ACGT", ">b"}, true, 2);
    CHECK(!bad.ok() && bad.status().message() == "FASTA files should have an even number of lines");
    auto bad2 = KmerCounter<K, N, KeyType>::FromFASTA(std::vector<std::string>{">a", "ACGu"}, true, 2);
    CHECK(!bad2.ok() && bad2.status().message() == "invalid FASTA file");
    auto ok = KmerCounter<K, N, KeyType>::FromFASTA(std::vector<std::string>{">a", "ACGTNACGTAC", ">b", "ACGTA"}, false, 2);
    CHECK(ok.ok() && ok.value().Get(Kmer<K>("ACGTA")) == 2 && ok.value().Get(Kmer<K>("CGTAC")) == 1 && ok.value().Size() == 2);
  }
}

static void TestKmerSetCompact() {  // test/kmer_set_compact.cc
  constexpr int K = 9, N = 10;
  using KeyType = std::uint16_t;
  const int n = 100000;
  const KmerSet<K, N, KeyType> kmer_set = RandomKmerSet<K, N, KeyType>(n, true);
  const KmerSetCompact<K, N, KeyType> compact = KmerSetCompact<K, N, KeyType>::FromKmerSet(kmer_set, true, true, 4);
  CHECK(compact.Size(4) == n);
  CHECK(kmer_set.Equals(compact.ToKmerSet(true, 4), 4));
  {  // Dump -> Load
    const std::string f = (std::filesystem::temp_directory_path() / "kmsc_compact_test.txt").string();
    CHECK(compact.Dump(f, "", 4).ok());
    auto loaded = KmerSetCompact<K, N, KeyType>::Load(f, "");
    CHECK(loaded.ok() && kmer_set.Equals(loaded.value().ToKmerSet(true, 4), 4));
    CHECK(loaded.value().Weight() == compact.Weight());
    const std::string fz = f + ".gz";
    CHECK(compact.Dump(fz, "gzip -c", 4).ok());
    auto lz = KmerSetCompact<K, N, KeyType>::Load(fz, "gzip -d -c");
    CHECK(lz.ok() && kmer_set.Equals(lz.value().ToKmerSet(true, 4), 4));
    std::filesystem::remove(f);
    std::filesystem::remove(fz);
  }
  {  // GetSampledKmerSet with all buckets, reversed (test/kmer_set_compact.cc:92-129)
    std::vector<int> ids;
    for (int i = 0; i < (1 << N); i++) ids.push_back(i);
    std::reverse(ids.begin(), ids.end());
    auto sampled = compact.GetSampledKmerSet(ids, true, 4);
    KmerSet<K, N, KeyType> rebuilt;
    std::vector<std::uint64_t> bits;
    for (int i = 0; i < (1 << N); i++) {
      CHECK(std::is_sorted(sampled[i].begin(), sampled[i].end()));
      for (KeyType key : sampled[i]) bits.push_back(GetKmerFromBucketAndKey<K, N, KeyType>(ids[i], key).Bits());
    }
    std::sort(bits.begin(), bits.end());
    CHECK(kmer_set.Equals(KmerSet<K, N, KeyType>::FromSortedBits(bits), 4));
  }
  {  // SPSS validity: every k-mer exactly once (test/spss.cc:57-68, 113-124)
    std::int64_t positions = 0;
    for (const std::string& s : compact.ToStrings(1)) {
      CHECK(static_cast<int>(s.size()) >= K);
      positions += static_cast<std::int64_t>(s.size()) - K + 1;
    }
    CHECK(positions == n);
    const KmerSet<K, N, KeyType> nc = RandomKmerSet<K, N, KeyType>(20000, false);
    const auto c2 = KmerSetCompact<K, N, KeyType>::FromKmerSet(nc, false, true, 2);
    CHECK(c2.Size(1) == 20000 && nc.Equals(c2.ToKmerSet(false, 2), 2));
  }
  {  // a blank line and a line shorter than K in an SPSS file: no k-mers from them, text preserved
     // (the reference stores length - K mod 2^32 and recovers the length the same way,
     // lib/core/kmer_set_compact.h:206-287)
    const std::vector<std::string> lines = {"ACGTACGTACGT", "", "ACG", "TTTTTTTTTAC"};
    const auto c3 = KmerSetCompact<K, N, KeyType>::FromStrings(lines);
    CHECK(c3.ToStrings(1) == lines);
    CHECK(c3.Weight() == 12 + 0 + 3 + 11);
    CHECK(c3.ToKmerSet(false, 1).Size() == 4 + 3);  // 4 windows of the first line, 3 of the last, all distinct
  }
}

static void TestKmerSetSet() {  // test/kmer_set_set.cc
  constexpr int K = 9, N = 10;
  using KeyType = std::uint16_t;
  const int n = 10, m = 10000;
  std::vector<KmerSet<K, N, KeyType>> sets;
  std::vector<KmerSetCompact<K, N, KeyType>> compact;
  // related sets so the factoring has something to find
  const KmerSet<K, N, KeyType> base = RandomKmerSet<K, N, KeyType>(m, true);
  for (int i = 0; i < n; i++) {
    KmerSet<K, N, KeyType> s = Add(RandomKmerSet<K, N, KeyType>(m / 4, true), base, 1);
    if (i % 2) s = Sub(s, RandomKmerSet<K, N, KeyType>(m, true), 1);
    sets.push_back(s);
    compact.push_back(KmerSetCompact<K, N, KeyType>::FromKmerSet(s, true, true, 2));
  }
  KmerSetSetOptions opt;
  opt.seed = 7;
  KmerSetSet<K, N, KeyType> kss(compact, true, 4, opt);
  CHECK(kss.Size() > n);  // at least one merge happened
  for (int i = 0; i < n; i++) CHECK(sets[i].Equals(kss.Get(i, true, 4), 4));
  const std::string dir = (std::filesystem::temp_directory_path() / "kmsc_kss_test").string();
  std::filesystem::remove_all(dir);
  CHECK(kss.Dump(dir, "", "txt", 4).ok());
  CHECK(kss.DumpGraph(dir + "/graph.dot").ok());
  {
    auto loaded = KmerSetSet<K, N, KeyType>::Load(dir, "", "txt", 4);
    CHECK(loaded.ok() && loaded.value().Size() == kss.Size());
    for (int i = 0; i < n; i++) CHECK(sets[i].Equals(loaded.value().Get(i, true, 4), 4));
    auto reader = KmerSetSetReader<K, N, KeyType>::FromDirectory(dir, "txt", "", true);
    CHECK(reader.ok() && reader.value().Size() == kss.Size());
    for (int i = 0; i < n; i++) {
      auto s = reader.value().Get(i, 2);
      CHECK(s.ok() && sets[i].Equals(s.value(), 2));
    }
  }
  std::filesystem::remove_all(dir);
  // exact-weights variant and the mst driver
  KmerSetSetOptions ex;
  ex.exact = true;
  ex.max_iterations = 3;
  KmerSetSet<K, N, KeyType> kss2(compact, true, 4, ex);
  CHECK(kss2.Merges().size() == 3);
  for (int i = 0; i < n; i++) CHECK(sets[i].Equals(kss2.Get(i, true, 4), 4));
  auto mst = BuildMst<K, N, KeyType>(sets);
  CHECK(static_cast<int>(mst.edges.size()) == n - 1);
  for (std::size_t e = 0; e < mst.edges.size(); e++) {
    const auto& ed = mst.edges[e];
    // S_c = (S_p \ del) + add
    KmerSet<K, N, KeyType> rec = Add(Sub(sets[ed.parent], mst.del[e], 1), mst.add[e], 1);
    CHECK(rec.Equals(sets[ed.child], 1));
    CHECK(ed.distance == mst.del[e].Size() + mst.add[e].Size());
  }
}

static void TestDisjointSet() {  // test/parallel_disjoint_set.cc
  const int n = 4000;
  ParallelDisjointSet par(n);
  std::vector<int> ser(n);
  for (int i = 0; i < n; i++) ser[i] = i;
  auto find = [&](int x) { while (ser[x] != x) x = ser[x] = ser[ser[x]]; return x; };
  std::vector<std::pair<int, int>> ops;
  for (int i = 0; i < 3000; i++) ops.emplace_back(static_cast<int>(rng() % n), static_cast<int>(rng() % n));
  std::vector<std::thread> th;
  for (int t = 0; t < 8; t++)
    th.emplace_back([&, t] { for (std::size_t i = t; i < ops.size(); i += 8) par.Unite(ops[i].first, ops[i].second); });
  for (auto& t : th) t.join();
  for (const auto& o : ops) ser[find(o.first)] = find(o.second);
  for (int i = 0; i < n; i += 7)
    for (int j = 0; j < n; j += 11) CHECK(par.IsSame(i, j) == (find(i) == find(j)));
}

#define RUN(fn)                                   \
  do {                                            \
    std::printf("[ RUN ] %s\n", #fn);             \
    std::fflush(stdout);                          \
    fn();                                         \
    std::printf("[ END ] %s (failures so far: %d)\n", #fn, g_fail); \
    std::fflush(stdout);                          \
  } while (0)

// streamvbyte-0124 byte codes of the string lengths (host/kmsc/streamvbyte0124.h): round trip, and the bytes of a
// fixed vector printed for tests/test_host_facade.py to compare with the oracle's restatement
static void TestSvb0124() {
  std::vector<std::uint32_t> v;
  for (int i = 0; i < 1000; i++) {
    const int kind = static_cast<int>(rng() % 4);
    v.push_back(kind == 0 ? 0u : kind == 1 ? static_cast<std::uint32_t>(rng() & 0xff) : kind == 2 ? static_cast<std::uint32_t>(rng() & 0xffff)
                                                                                            : static_cast<std::uint32_t>(rng()));
  }
  const std::vector<std::uint8_t> enc = Svb0124Encode(v);
  CHECK(Svb0124Decode(enc, v.size()) == v);
  CHECK(Svb0124Decode(Svb0124Encode(std::vector<std::uint32_t>()), 0).empty());
  std::vector<std::uint32_t> fixed;
  for (std::uint32_t i = 0; i < 37; i++) fixed.push_back(i % 5 == 0 ? 0u : (i * 2654435761u) >> (i % 4 * 8));
  std::printf("svb0124");
  for (std::uint8_t b : Svb0124Encode(fixed)) std::printf(" %02x", b);
  std::printf("\n");
}

int main(int argc, char** argv) {
  if (argc > 1 && std::string(argv[1]) == "--host-only") {   // no device call: the CPU suite runs these
    RUN(TestKmer);
    RUN(TestDisjointSet);
    RUN(TestSvb0124);
    if (g_fail == 0) std::printf("ALL OK\n");
    return g_fail == 0 ? 0 : 1;
  }
  RUN(TestKmer);
  RUN(TestDisjointSet);
  RUN(TestSvb0124);
  RUN(TestKmerSet);
  RUN(TestKmerCounter);
  RUN(TestKmerSetCompact);
  RUN(TestKmerSetSet);
  if (g_fail == 0) std::printf("ALL OK\n");
  return g_fail == 0 ? 0 : 1;
}
