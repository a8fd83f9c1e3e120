// ref_driver.cc -- TEST INFRASTRUCTURE ONLY (oracle/).
//
// extern "C" entry points over the reference's OWN, UNMODIFIED headers
// (/root/reference/lib/core/*.h, included at build time from where they lie; no
// reference source is copied into this repository). Built by oracle/Makefile
// into oracle/_ref/libkmsc_ref.so with the std-only shims of oracle/shim/
// standing in for Abseil / Boost.Asio / spdlog / streamvbyte, none of which is
// installable here (no network). Uses:
//   * pin the C restatement (kmsc_oracle.c) against the real reference,
//   * generate tests/golden/ *.json (tests/golden/make_golden.py),
//   * be the CPU arm of bench.py (--impl reference; cpu_baseline.kind
//     "reference").
// Template instantiations are selected by a config id:
//   0 <5,3,uint8>  1 <9,10,uint8>  2 <15,14,uint16>  3 <19,10,uint32>
//   4 <23,14,uint32>  5 <31,14,uint64>
// (2-4 are the CLI's, src/kmerset-multiple-compress.cc:149-157; 0-1 are the
// reference tests'; 5 is config C4's.)
#include <unistd.h>

#include <algorithm>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>

#include "core/kmer.h"
#include "core/kmer_counter.h"
#include "core/kmer_set.h"
#include "core/kmer_set_compact.h"
#include "core/kmer_set_set.h"
#include "core/parallel_disjoint_set.h"
#include "core/random.h"
#include "core/spss.h"

namespace {

#define KMSC_DISPATCH(cfg, CALL)                                   \
  switch (cfg) {                                                   \
    case 0: return CALL(5, 3, std::uint8_t);                       \
    case 1: return CALL(9, 10, std::uint8_t);                      \
    case 2: return CALL(15, 14, std::uint16_t);                    \
    case 3: return CALL(19, 10, std::uint32_t);                    \
    case 4: return CALL(23, 14, std::uint32_t);                    \
    case 5: return CALL(31, 14, std::uint64_t);                    \
    default: return -100;                                          \
  }

std::string TempPath(const char* tag) {
  char buf[256];
  std::snprintf(buf, sizeof(buf), "/tmp/kmsc_ref_%s_%d_XXXXXX", tag, (int)getpid());
  int fd = mkstemp(buf);
  if (fd >= 0) close(fd);
  return std::string(buf);
}

template <int K, int N, typename KeyType>
KmerSet<K, N, KeyType> MakeSet(const std::uint64_t* a, std::int64_t n) {
  KmerSet<K, N, KeyType> s;
  for (std::int64_t i = 0; i < n; i++) s.Add(Kmer<K>(a[i]));
  return s;
}

template <int K, int N, typename KeyType>
std::int64_t DumpSet(const KmerSet<K, N, KeyType>& s, int n_workers, std::uint64_t* out) {
  std::vector<Kmer<K>> v = s.Find(n_workers);
  std::vector<std::uint64_t> bits(v.size());
  for (std::size_t i = 0; i < v.size(); i++) bits[i] = v[i].Bits();
  std::sort(bits.begin(), bits.end());
  if (out) std::copy(bits.begin(), bits.end(), out);
  return (std::int64_t)bits.size();
}

template <int K, int N, typename KeyType>
absl::StatusOr<KmerSetCompact<K, N, KeyType>> CompactFromStrings(const char* const* strings,
                                                                  std::int64_t n) {
  // The string constructor is private (kmer_set_compact.h:206); Load() is the
  // public way in (kmer_set_compact.h:71-87), so go through a temp file.
  const std::string path = TempPath("spss");
  {
    std::vector<std::string> lines(strings, strings + n);
    absl::Status st = WriteLines(path, "", lines);
    if (!st.ok()) return st;
  }
  auto r = KmerSetCompact<K, N, KeyType>::Load(path, "");
  std::remove(path.c_str());
  return r;
}

// ---- Kmer ----------------------------------------------------------------
template <int K, int N, typename KeyType>
int KmerOp(int op, std::uint64_t bits, const char* s, char c, std::uint64_t* out, char* sout) {
  Kmer<K> k = s ? Kmer<K>(std::string(s)) : Kmer<K>(bits);
  Kmer<K> r = k;
  switch (op) {
    case 0: break;                       // parse / identity
    case 1: r = k.Complement(); break;
    case 2: r = k.Canonical(); break;
    case 3: r = k.Next(c); break;
    case 4: r = k.Prev(c); break;
    default: return -1;
  }
  if (out) *out = r.Bits();
  if (sout) std::strcpy(sout, r.String().c_str());
  return 0;
}

template <int K, int N, typename KeyType>
int BucketKey(std::uint64_t bits, std::int32_t* bucket, std::uint64_t* key, std::uint64_t* back) {
  int b;
  KeyType k;
  std::tie(b, k) = GetBucketAndKeyFromKmer<K, N, KeyType>(Kmer<K>(bits));
  *bucket = b;
  *key = (std::uint64_t)k;
  *back = GetKmerFromBucketAndKey<K, N, KeyType>(b, k).Bits();
  return 0;
}

// ---- KmerCounter ----------------------------------------------------------
template <int K, int N, typename KeyType>
int CountReads(const char* const* reads, std::int64_t n, int canonical, int n_workers, int cutoff,
               std::uint64_t* kmers, std::uint8_t* counts, std::int64_t* n_distinct,
               std::uint64_t* kept, std::int64_t* n_kept, std::int64_t* cutoff_count) {
  std::vector<std::string> v(reads, reads + n);
  KmerCounter<K, N, KeyType> counter =
      KmerCounter<K, N, KeyType>::FromReads(std::move(v), canonical != 0, n_workers);
  KmerSet<K, N, KeyType> all;
  std::int64_t zero_cut;
  std::tie(all, zero_cut) = counter.ToKmerSet(0, n_workers);
  *n_distinct = DumpSet<K, N, KeyType>(all, n_workers, kmers);
  if (*n_distinct != counter.Size()) return -2;
  if (kmers && counts)
    for (std::int64_t i = 0; i < *n_distinct; i++) counts[i] = counter.Get(Kmer<K>(kmers[i]));
  KmerSet<K, N, KeyType> set;
  std::int64_t cut;
  std::tie(set, cut) = counter.ToKmerSet((std::uint8_t)cutoff, n_workers);
  *n_kept = DumpSet<K, N, KeyType>(set, n_workers, kept);
  *cutoff_count = cut;
  return 0;
}

template <int K, int N, typename KeyType>
int Fasta(const char* const* lines, std::int64_t n, int canonical, int n_workers,
          std::int64_t* n_distinct) {
  std::vector<std::string> v(lines, lines + n);
  auto r = KmerCounter<K, N, KeyType>::FromFASTA(std::move(v), canonical != 0, n_workers);
  if (!r.ok()) {
    if (r.status().message() == "FASTA files should have an even number of lines") return 1;
    if (r.status().message() == "invalid FASTA file") return 2;
    return 3;
  }
  *n_distinct = r.value().Size();
  return 0;
}

// ---- KmerSetCompact -------------------------------------------------------
template <int K, int N, typename KeyType>
int SampledSet(const char* const* strings, std::int64_t n, int canonical, int n_workers,
               const std::int32_t* bucket_ids, std::int32_t n_ids, std::int64_t* out_offs,
               std::uint64_t* out_keys, std::int64_t* size, std::int64_t* weight) {
  auto c = CompactFromStrings<K, N, KeyType>(strings, n);
  if (!c.ok()) return -1;
  if (size) *size = c.value().Size(n_workers);
  if (weight) *weight = c.value().Weight();
  std::vector<int> ids(bucket_ids, bucket_ids + n_ids);
  std::vector<std::vector<KeyType>> s = c.value().GetSampledKmerSet(ids, canonical != 0, n_workers);
  std::int64_t pos = 0;
  for (std::int32_t i = 0; i < n_ids; i++) {
    out_offs[i] = pos;
    for (KeyType k : s[i]) out_keys[pos++] = (std::uint64_t)k;
  }
  out_offs[n_ids] = pos;
  return 0;
}

template <int K, int N, typename KeyType>
int SetFromSpss(const char* const* strings, std::int64_t n, int canonical, int n_workers,
                std::uint64_t* out, std::int64_t* n_out, std::uint64_t* hash) {
  auto c = CompactFromStrings<K, N, KeyType>(strings, n);
  if (!c.ok()) return -1;
  KmerSet<K, N, KeyType> s = c.value().ToKmerSet(canonical != 0, n_workers);
  *n_out = DumpSet<K, N, KeyType>(s, n_workers, out);
  if (hash) *hash = s.Hash(n_workers);
  if (*n_out != s.Size()) return -2;
  return 0;
}

// SPSS of a k-mer set through the reference's builder (spss.h:1836-1858) and
// its text dump; *text receives a malloc'd '\n'-joined buffer.
template <int K, int N, typename KeyType>
int SpssFromSet(const std::uint64_t* kmers, std::int64_t n, int canonical, int fast, int n_workers,
                char** text, std::int64_t* n_strings, std::int64_t* weight) {
  KmerSet<K, N, KeyType> s = MakeSet<K, N, KeyType>(kmers, n);
  KmerSetCompact<K, N, KeyType> c =
      KmerSetCompact<K, N, KeyType>::FromKmerSet(s, canonical != 0, fast != 0, n_workers);
  const std::string path = TempPath("dump");
  if (!c.Dump(path, "", n_workers).ok()) return -1;
  auto lines = ReadLines(path, "");
  std::remove(path.c_str());
  if (!lines.ok()) return -1;
  std::string joined;
  for (const std::string& l : lines.value()) { joined += l; joined += '\n'; }
  *text = (char*)std::malloc(joined.size() + 1);
  std::memcpy(*text, joined.c_str(), joined.size() + 1);
  *n_strings = (std::int64_t)lines.value().size();
  *weight = c.Weight();
  return 0;
}

// ---- KmerSet algebra ------------------------------------------------------
template <int K, int N, typename KeyType>
int SetOp(int op, const std::uint64_t* a, std::int64_t na, const std::uint64_t* b, std::int64_t nb,
          int n_workers, std::uint64_t* out, std::int64_t* n_out, std::uint64_t* scalar) {
  KmerSet<K, N, KeyType> A = MakeSet<K, N, KeyType>(a, na);
  KmerSet<K, N, KeyType> B = MakeSet<K, N, KeyType>(b, nb);
  switch (op) {
    case 0: *n_out = DumpSet<K, N, KeyType>(Add(A, B, n_workers), n_workers, out); break;
    case 1: *n_out = DumpSet<K, N, KeyType>(Sub(A, B, n_workers), n_workers, out); break;
    case 2: *n_out = DumpSet<K, N, KeyType>(Intersection(A, B, n_workers), n_workers, out); break;
    case 3: *scalar = (std::uint64_t)A.Diff(B, n_workers); break;
    case 4: *scalar = (std::uint64_t)A.Hash(n_workers); break;
    case 5: *scalar = A.Equals(B, n_workers) ? 1 : 0; break;
    default: return -1;
  }
  return 0;
}

// ---- KmerSetSet -----------------------------------------------------------
struct LogCapture {
  std::mutex mu;
  std::string text;            // every debug line, '\n'-joined
  std::vector<double> stamps;  // seconds since start, one per line
  std::chrono::steady_clock::time_point t0;
  const char* stop_after = nullptr;
  int stop_after_merges = 0;   // > 0: leave the greedy loop when its (m+1)-th iteration announces its pair
  int merges_seen = 0;
};
LogCapture* g_capture = nullptr;
struct StopRun {};

void Hook(const char* fmt, const char* formatted) {
  LogCapture* c = g_capture;
  if (!c) return;
  {
    std::lock_guard<std::mutex> l(c->mu);
    c->text += formatted;
    c->text += '\n';
    c->stamps.push_back(
        std::chrono::duration<double>(std::chrono::steady_clock::now() - c->t0).count());
  }
  if (c->stop_after && std::strcmp(fmt, c->stop_after) == 0) throw StopRun{};
  // the per-iteration line of the greedy loop, kmer_set_set.h:324: "j = {}, k = {}, weight = {}"
  if (c->stop_after_merges > 0 && std::strncmp(formatted, "j = ", 4) == 0 && ++c->merges_seen > c->stop_after_merges) throw StopRun{};
}

// Runs the reference's KmerSetSet constructor (kmer_set_set.h:109-427) on SPSS
// files. stop_after_weights != 0 leaves the constructor (via the log hook)
// right after "calculated initial weights" (:221): the sampled-set phase
// (:138-153) and the all-pairs GetEdgeWeight phase (:187-219) have then run
// exactly as in the reference and their wall-clock is reported in phase_s[0..1].
// Otherwise the whole greedy loop runs; per original set i, Get(i) is
// reconstructed and (size, XOR hash) written to sizes/hashes.
template <int K, int N, typename KeyType>
int RunKmerSetSet(const char* const* files, std::int32_t n_files, int canonical, int n_workers,
                  int stop_after_weights, double* phase_s, char** log_text, std::int64_t* sizes,
                  std::uint64_t* hashes, std::int32_t* n_nodes, const char* dump_dir) {
  std::vector<KmerSetCompact<K, N, KeyType>> sets(n_files);
  for (std::int32_t i = 0; i < n_files; i++) {
    auto r = KmerSetCompact<K, N, KeyType>::Load(files[i], "");
    if (!r.ok()) return -1;
    sets[i] = std::move(r).value();
  }
  LogCapture cap;
  cap.t0 = std::chrono::steady_clock::now();
  if (stop_after_weights == 1) cap.stop_after = "calculated initial weights";
  if (stop_after_weights < 0) cap.stop_after_merges = -stop_after_weights;  // -m: stop when merge m + 1 is announced
  g_capture = &cap;
  kmsc_shim::log_hook() = &Hook;
  int rc = 0;
  try {
    KmerSetSet<K, N, KeyType> kss(std::move(sets), canonical != 0, n_workers);
    if (n_nodes) *n_nodes = kss.Size();
    if (sizes && hashes) {
      for (std::int32_t i = 0; i < n_files; i++) {
        KmerSet<K, N, KeyType> s = kss.Get(i, canonical != 0, n_workers);
        sizes[i] = s.Size();
        hashes[i] = s.Hash(n_workers);
      }
    }
    if (dump_dir && dump_dir[0]) {
      if (!kss.Dump(dump_dir, "", "txt", n_workers).ok()) rc = -3;
    }
  } catch (const StopRun&) {
    rc = 1;
  }
  kmsc_shim::log_hook() = nullptr;
  g_capture = nullptr;
  if (phase_s) {
    // line-by-line scan for the four phase markers (:136, :155, :189, :221)
    double t[4] = {-1, -1, -1, -1};
    const char* marks[4] = {"constructing initial sampled_kmer_sets",
                            "constructed initial sampled_kmer_sets",
                            "calculating initial weights", "calculated initial weights"};
    std::size_t pos = 0, line = 0;
    while (pos < cap.text.size()) {
      std::size_t e = cap.text.find('\n', pos);
      std::string l = cap.text.substr(pos, e - pos);
      for (int m = 0; m < 4; m++)
        if (t[m] < 0 && l == marks[m]) t[m] = cap.stamps[line];
      pos = e + 1;
      line++;
    }
    phase_s[0] = t[1] - t[0];
    phase_s[1] = t[3] - t[2];
  }
  if (log_text) {
    *log_text = (char*)std::malloc(cap.text.size() + 1);
    std::memcpy(*log_text, cap.text.c_str(), cap.text.size() + 1);
  }
  return rc;
}

// The split of one greedy iteration exactly as the reference does it (kmer_set_set.h:332-343):
// ToKmerSet x2 (SPSS text -> hash sets), n = Intersection(j, k), j.Sub(n), k.Sub(n); seconds[0..2] =
// decode, intersection, the two subtractions; sizes = |j|, |k|, |n|, |j \ n|, |k \ n|.
template <int K, int N, typename KeyType>
int SplitStage(const char* file_j, const char* file_k, int canonical, int n_workers, double* seconds, std::int64_t* sizes) {
  auto lj = KmerSetCompact<K, N, KeyType>::Load(file_j, "");
  auto lk = KmerSetCompact<K, N, KeyType>::Load(file_k, "");
  if (!lj.ok() || !lk.ok()) return -1;
  auto t0 = std::chrono::steady_clock::now();
  KmerSet<K, N, KeyType> sj = lj.value().ToKmerSet(canonical != 0, n_workers);
  KmerSet<K, N, KeyType> sk = lk.value().ToKmerSet(canonical != 0, n_workers);
  auto t1 = std::chrono::steady_clock::now();
  sizes[0] = sj.Size(); sizes[1] = sk.Size();
  KmerSet<K, N, KeyType> n = Intersection(sj, sk, n_workers);
  auto t2 = std::chrono::steady_clock::now();
  sj.Sub(n, n_workers);
  sk.Sub(n, n_workers);
  auto t3 = std::chrono::steady_clock::now();
  sizes[2] = n.Size(); sizes[3] = sj.Size(); sizes[4] = sk.Size();
  seconds[0] = std::chrono::duration<double>(t1 - t0).count();
  seconds[1] = std::chrono::duration<double>(t2 - t1).count();
  seconds[2] = std::chrono::duration<double>(t3 - t2).count();
  return 0;
}

template <int K, int N, typename KeyType>
int ReaderGet(const char* dir, int canonical, int n_workers, std::int32_t i, std::int64_t* size,
              std::uint64_t* hash, std::int32_t* n_sets) {
  auto r = KmerSetSetReader<K, N, KeyType>::FromDirectory(dir, "txt", "", canonical != 0);
  if (!r.ok()) return -1;
  *n_sets = r.value().Size();
  auto s = r.value().Get(i, n_workers);
  if (!s.ok()) return -2;
  *size = s.value().Size();
  *hash = s.value().Hash(n_workers);
  return 0;
}

}  // namespace

extern "C" {

#define C_KMEROP(K, N, T) KmerOp<K, N, T>(op, bits, s, c, out, sout)
int ref_kmer_op(int cfg, int op, std::uint64_t bits, const char* s, char c, std::uint64_t* out,
                char* sout) {
  KMSC_DISPATCH(cfg, C_KMEROP)
}

#define C_BUCKETKEY(K, N, T) BucketKey<K, N, T>(bits, bucket, key, back)
int ref_bucket_key(int cfg, std::uint64_t bits, std::int32_t* bucket, std::uint64_t* key,
                   std::uint64_t* back) {
  KMSC_DISPATCH(cfg, C_BUCKETKEY)
}

int ref_add_with_max_u8(int x, int y) { return AddWithMax<std::uint8_t>((std::uint8_t)x, (std::uint8_t)y); }

#define C_COUNTREADS(K, N, T) \
  CountReads<K, N, T>(reads, n, canonical, n_workers, cutoff, kmers, counts, n_distinct, kept, n_kept, cutoff_count)
int ref_count_reads(int cfg, const char* const* reads, std::int64_t n, int canonical, int n_workers,
                    int cutoff, std::uint64_t* kmers, std::uint8_t* counts, std::int64_t* n_distinct,
                    std::uint64_t* kept, std::int64_t* n_kept, std::int64_t* cutoff_count) {
  KMSC_DISPATCH(cfg, C_COUNTREADS)
}

#define C_FASTA(K, N, T) Fasta<K, N, T>(lines, n, canonical, n_workers, n_distinct)
int ref_fasta(int cfg, const char* const* lines, std::int64_t n, int canonical, int n_workers,
              std::int64_t* n_distinct) {
  KMSC_DISPATCH(cfg, C_FASTA)
}

#define C_SAMPLED(K, N, T) \
  SampledSet<K, N, T>(strings, n, canonical, n_workers, bucket_ids, n_ids, out_offs, out_keys, size, weight)
int ref_sampled_set(int cfg, const char* const* strings, std::int64_t n, int canonical,
                    int n_workers, const std::int32_t* bucket_ids, std::int32_t n_ids,
                    std::int64_t* out_offs, std::uint64_t* out_keys, std::int64_t* size,
                    std::int64_t* weight) {
  KMSC_DISPATCH(cfg, C_SAMPLED)
}

#define C_SETFROMSPSS(K, N, T) SetFromSpss<K, N, T>(strings, n, canonical, n_workers, out, n_out, hash)
int ref_set_from_spss(int cfg, const char* const* strings, std::int64_t n, int canonical,
                      int n_workers, std::uint64_t* out, std::int64_t* n_out, std::uint64_t* hash) {
  KMSC_DISPATCH(cfg, C_SETFROMSPSS)
}

#define C_SPSSFROMSET(K, N, T) \
  SpssFromSet<K, N, T>(kmers, n, canonical, fast, n_workers, text, n_strings, weight)
int ref_spss_from_set(int cfg, const std::uint64_t* kmers, std::int64_t n, int canonical, int fast,
                      int n_workers, char** text, std::int64_t* n_strings, std::int64_t* weight) {
  KMSC_DISPATCH(cfg, C_SPSSFROMSET)
}

#define C_SETOP(K, N, T) SetOp<K, N, T>(op, a, na, b, nb, n_workers, out, n_out, scalar)
int ref_set_op(int cfg, int op, const std::uint64_t* a, std::int64_t na, const std::uint64_t* b,
               std::int64_t nb, int n_workers, std::uint64_t* out, std::int64_t* n_out,
               std::uint64_t* scalar) {
  KMSC_DISPATCH(cfg, C_SETOP)
}

#define C_KSS(K, N, T)                                                                      \
  RunKmerSetSet<K, N, T>(files, n_files, canonical, n_workers, stop_after_weights, phase_s, \
                         log_text, sizes, hashes, n_nodes, dump_dir)
int ref_kmer_set_set(int cfg, const char* const* files, std::int32_t n_files, int canonical,
                     int n_workers, int stop_after_weights, double* phase_s, char** log_text,
                     std::int64_t* sizes, std::uint64_t* hashes, std::int32_t* n_nodes,
                     const char* dump_dir) {
  KMSC_DISPATCH(cfg, C_KSS)
}

#define C_SPLIT(K, N, T) SplitStage<K, N, T>(file_j, file_k, canonical, n_workers, seconds, sizes)
int ref_split_stage(int cfg, const char* file_j, const char* file_k, int canonical, int n_workers, double* seconds,
                    std::int64_t* sizes) {
  KMSC_DISPATCH(cfg, C_SPLIT)
}

#define C_READER(K, N, T) ReaderGet<K, N, T>(dir, canonical, n_workers, i, size, hash, n_sets)
int ref_reader_get(int cfg, const char* dir, int canonical, int n_workers, std::int32_t i,
                   std::int64_t* size, std::uint64_t* hash, std::int32_t* n_sets) {
  KMSC_DISPATCH(cfg, C_READER)
}

// GetRandomInts (core/random.h:13-41) with the shim generator's instance
// counter forced to `counter` first, so a caller can replay the bucket sample a
// seeded KmerSetSet run drew (kmer_set_set.h:123-124).
int ref_random_ints(std::uint64_t counter, int n, int lo, int hi, std::int32_t* out) {
  kmsc_shim::seed_counter().store(counter);
  std::vector<int> v = GetRandomInts<int>(n, true, true, lo, hi);
  std::copy(v.begin(), v.end(), out);
  return (int)v.size();
}
std::uint64_t ref_seed_counter(void) { return kmsc_shim::seed_counter().load(); }
void ref_set_seed_counter(std::uint64_t c) { kmsc_shim::seed_counter().store(c); }

// ParallelDisjointSet (parallel_disjoint_set.h)
void* ref_dsu_new(int n) { return new ParallelDisjointSet(n); }
void ref_dsu_free(void* d) { delete (ParallelDisjointSet*)d; }
int ref_dsu_find(void* d, int x) { return ((ParallelDisjointSet*)d)->Find(x); }
int ref_dsu_same(void* d, int x, int y) { return ((ParallelDisjointSet*)d)->IsSame(x, y) ? 1 : 0; }
void ref_dsu_unite(void* d, int x, int y) { ((ParallelDisjointSet*)d)->Unite(x, y); }

void ref_free(void* p) { std::free(p); }
void ref_set_log_level(int l) { kmsc_shim::log_level() = l; }

}  // extern "C"
