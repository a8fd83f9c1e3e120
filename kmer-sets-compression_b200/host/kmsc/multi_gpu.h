// kmsc/multi_gpu.h -- the mst driver over several GPUs of one box (north_star: "the k-mer
// prefix space is partitioned over GPUs ... partial N x N intersection counts ... summed with a single
// NCCL allreduce ... difference sets are computed where their buckets live and gathered to the host").
//
// One host thread per GPU, each with its own kmsc_ctx and a communicator owned by the library
// (kmsc_comm_init). Per rank: decode its share of the sets (sets r, r + R, ...; one batched call),
// kmsc_sets_exchange re-shards them by k-mer prefix, kmsc_pair_counts returns the all-reduced exact
// matrix, rank 0 builds the tree (MstTree), every rank splits the tree edges on ITS shard
// (kmsc_pair_split_batch with the rank's partial counts as exact hints) and copies its slices of the
// difference sets to the host, where they are concatenated in rank (= bucket) order.
// The reference has nothing like it (one process, boost::asio::thread_pool); what is kept is its
// arithmetic: sum over buckets of lib/core/kmer_set_set.h:161-181, set algebra of kmer_set.h:177-187.
#ifndef KMSC_HOST_MULTI_GPU_H_
#define KMSC_HOST_MULTI_GPU_H_
#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <cstdint>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "kmsc/kmer_set_compact.h"
#include "kmsc/kmer_set_set.h"

namespace kmsc {

template <int K, int N, typename KeyType>
struct MultiGpuMst {
  std::vector<std::int64_t> W;                       // exact n x n matrix (all-reduced)
  std::vector<MstEdge> edges;
  std::vector<std::vector<std::uint64_t>> add, del;  // per edge: ascending k-mer values of S_c \ S_p, S_p \ S_c
  std::string error;                                 // empty = ok
};

namespace internal {
class Barrier {  // std::barrier is C++20
 public:
  explicit Barrier(int n) : n_(n) {}
  void Wait() {
    std::unique_lock<std::mutex> l(m_);
    const int gen = gen_;
    if (++count_ == n_) { count_ = 0; gen_++; cv_.notify_all(); }
    else cv_.wait(l, [&] { return gen != gen_; });
  }
 private:
  std::mutex m_;
  std::condition_variable cv_;
  int n_, count_ = 0, gen_ = 0;
};
}  // namespace internal

// n_gpus ranks; the number of sets must be a multiple of n_gpus (pad with empty sets otherwise).
template <int K, int N, typename KeyType>
MultiGpuMst<K, N, KeyType> BuildMstMultiGpu(const std::vector<KmerSetCompact<K, N, KeyType>>& compact, bool canonical, int n_gpus) {
  MultiGpuMst<K, N, KeyType> res;
  const int n = static_cast<int>(compact.size());
  const int R = n_gpus;
  if (R < 1 || n % R != 0) { res.error = "the number of sets must be a multiple of --gpus"; return res; }
  unsigned char id[128];
  if (kmsc_comm_unique_id(id) != KMSC_OK) { res.error = kmsc_last_error(); return res; }
  // bucket cuts by the cumulative k-mer count: canonical k-mers are skewed (44 / 31 / 19 / 6 % by first
  // base); the text weight of set 0 is the proxy here (every set of a related collection is alike)
  std::vector<std::int32_t> cuts(static_cast<std::size_t>(R) + 1, 0);
  {
    // expected share of k-mers below bucket b for canonical k-mers: cumulative over first bases A C G T
    const double share[5] = {0.0, 7.0 / 16, 12.0 / 16, 15.0 / 16, 1.0};
    const int nb = 1 << N;
    for (int r = 1; r < R; r++) {
      const double want = static_cast<double>(r) / R;
      int base = 0;
      while (base < 3 && share[base + 1] < want) base++;
      const double inside = (want - share[base]) / (share[base + 1] - share[base]);
      cuts[static_cast<std::size_t>(r)] = canonical ? static_cast<std::int32_t>((base + inside) * (nb / 4)) : nb / R * r;
    }
    cuts[static_cast<std::size_t>(R)] = nb;
  }
  std::vector<std::string> errs(static_cast<std::size_t>(R));
  internal::Barrier bar(R);
  std::vector<MstEdge> edges;
  std::vector<std::int64_t> Wfull;
  // per rank, per edge: its slice of add / del as k-mer values
  std::vector<std::vector<std::vector<std::uint64_t>>> add_r(static_cast<std::size_t>(R)), del_r(static_cast<std::size_t>(R));
  std::atomic<int> failed{0};

  auto worker = [&](int r) {
    auto fail = [&](const char* what) { errs[static_cast<std::size_t>(r)] = std::string(what) + ": " + kmsc_last_error(); failed = 1; };
    kmsc_ctx* ctx = nullptr;
    if (kmsc_ctx_create(r, nullptr, &ctx) != KMSC_OK) { fail("kmsc_ctx_create"); }
    // NCCL initialisation is collective: every rank must arrive, failed or not
    if (ctx && kmsc_comm_init(ctx, r, R, id) != KMSC_OK) fail("kmsc_comm_init");
    bar.Wait();
    if (failed) { if (ctx) kmsc_ctx_destroy(ctx); return; }
    const int m = n / R;
    std::vector<const std::uint64_t*> words(static_cast<std::size_t>(m));
    std::vector<std::vector<std::int64_t>> offs(static_cast<std::size_t>(m));
    std::vector<const std::int64_t*> offp(static_cast<std::size_t>(m));
    std::vector<std::int64_t> nstr(static_cast<std::size_t>(m));
    for (int j = 0; j < m; j++) {
      const auto& c = compact[static_cast<std::size_t>(r + j * R)];
      words[static_cast<std::size_t>(j)] = c.PackedWords().data();
      offs[static_cast<std::size_t>(j)] = c.StringOffsets();
      offp[static_cast<std::size_t>(j)] = offs[static_cast<std::size_t>(j)].data();
      nstr[static_cast<std::size_t>(j)] = static_cast<std::int64_t>(offs[static_cast<std::size_t>(j)].size()) - 1;
    }
    std::vector<kmsc_set*> mine(static_cast<std::size_t>(m), nullptr), all(static_cast<std::size_t>(n), nullptr);
    bool ok = kmsc_sets_from_packed_batch(ctx, K, N, static_cast<int>(sizeof(KeyType)), m, words.data(), offp.data(), nstr.data(),
                                          canonical ? 1 : 0, 1, 0, 1 << N, mine.data()) == KMSC_OK;
    if (!ok) fail("kmsc_sets_from_packed_batch");
    bar.Wait();   // collectives below: all ranks or none
    if (!failed) {
      ok = kmsc_sets_exchange(ctx, mine.data(), m, cuts.data(), all.data(), n) == KMSC_OK;
      if (!ok) fail("kmsc_sets_exchange");
    }
    for (kmsc_set* s : mine) if (s) kmsc_set_free(ctx, s);
    bar.Wait();
    std::vector<std::int64_t> W(static_cast<std::size_t>(n) * n, 0);
    if (!failed) {
      ok = kmsc_pair_counts(ctx, all.data(), n, nullptr, 0, W.data(), nullptr) == KMSC_OK;   // all-reduced inside
      if (!ok) fail("kmsc_pair_counts");
    }
    bar.Wait();
    if (!failed && r == 0) { Wfull = W; edges = MstTree(W, n); }
    bar.Wait();
    if (!failed) {
      const std::size_t ne = edges.size();
      std::vector<const kmsc_set*> js(ne), ks(ne);
      for (std::size_t e = 0; e < ne; e++) { js[e] = all[static_cast<std::size_t>(edges[e].parent)]; ks[e] = all[static_cast<std::size_t>(edges[e].child)]; }
      std::vector<kmsc_set*> jm(ne, nullptr), km(ne, nullptr);
      if (ne > 0) {
        ok = kmsc_pair_split_batch(ctx, js.data(), ks.data(), static_cast<std::int32_t>(ne), nullptr, nullptr, jm.data(), km.data()) == KMSC_OK;
        if (!ok) fail("kmsc_pair_split_batch");
      }
      auto to_kmers = [&](kmsc_set* s, std::vector<std::uint64_t>* out) {
        std::int64_t nk = 0;
        kmsc_set_size(ctx, s, &nk);
        std::vector<std::int64_t> o((std::size_t(1) << N) + 1);
        std::vector<KeyType> keys(static_cast<std::size_t>(nk));
        if (kmsc_set_to_csr(ctx, s, o.data(), keys.data()) != KMSC_OK) { fail("kmsc_set_to_csr"); return; }
        out->resize(static_cast<std::size_t>(nk));
        constexpr int kb = 2 * K - N;
        for (std::size_t b = 0; b + 1 < o.size(); b++)
          for (std::int64_t i = o[b]; i < o[b + 1]; i++)
            (*out)[static_cast<std::size_t>(i)] = (static_cast<std::uint64_t>(b) << kb) | static_cast<std::uint64_t>(keys[static_cast<std::size_t>(i)]);
      };
      add_r[static_cast<std::size_t>(r)].resize(ne);
      del_r[static_cast<std::size_t>(r)].resize(ne);
      for (std::size_t e = 0; e < ne && ok; e++) {
        to_kmers(km[e], &add_r[static_cast<std::size_t>(r)][e]);   // S_child \ S_parent on this shard
        to_kmers(jm[e], &del_r[static_cast<std::size_t>(r)][e]);   // S_parent \ S_child
      }
      for (kmsc_set* s : jm) if (s) kmsc_set_free(ctx, s);
      for (kmsc_set* s : km) if (s) kmsc_set_free(ctx, s);
    }
    for (kmsc_set* s : all) if (s) kmsc_set_free(ctx, s);
    bar.Wait();
    kmsc_ctx_destroy(ctx);
  };
  std::vector<std::thread> th;
  for (int r = 0; r < R; r++) th.emplace_back(worker, r);
  for (auto& t : th) t.join();
  for (const std::string& e : errs)
    if (!e.empty()) { res.error = e; return res; }
  res.W = std::move(Wfull);
  res.edges = edges;
  res.add.resize(edges.size());
  res.del.resize(edges.size());
  for (std::size_t e = 0; e < edges.size(); e++)
    for (int r = 0; r < R; r++) {   // rank order = bucket order: the concatenation is ascending
      res.add[e].insert(res.add[e].end(), add_r[static_cast<std::size_t>(r)][e].begin(), add_r[static_cast<std::size_t>(r)][e].end());
      res.del[e].insert(res.del[e].end(), del_r[static_cast<std::size_t>(r)][e].begin(), del_r[static_cast<std::size_t>(r)][e].end());
    }
  return res;
}

}  // namespace kmsc
#endif
