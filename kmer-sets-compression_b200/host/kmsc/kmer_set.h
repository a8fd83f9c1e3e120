// kmsc/kmer_set.h -- KmerSet<K,N,KeyType> with the reference's interface
// (lib/core/kmer_set.h:57-305) backed by a device CSR set.
//
// Representation: a sorted vector of k-mer values on the host and/or an immutable
// device set (shared). Single-k-mer mutation (Add/Remove) edits the host vector and
// drops the device copy; the bulk operations the hot path uses (Add/Sub of whole
// sets, Intersection, Diff, Equals, Hash, Size) run on the GPU through libkmsc and
// never touch the host vector. n_workers arguments are accepted for source
// compatibility and ignored (the device is the parallelism).
#ifndef KMSC_HOST_KMER_SET_H_
#define KMSC_HOST_KMER_SET_H_
#include <algorithm>
#include <cstdint>
#include <optional>
#include <vector>

#include "kmsc/device.h"
#include "kmsc/kmer.h"

namespace kmsc {

template <int K, int N, typename KeyType>
class KmerSet {
  static_assert(2 * K - N <= static_cast<int>(sizeof(KeyType) * 8), "key does not fit KeyType");
  static_assert(sizeof(KeyType) == 2 || sizeof(KeyType) == 4 || sizeof(KeyType) == 8,
                "KeyType must be uint16_t, uint32_t or uint64_t on the device path");

 public:
  KmerSet() : host_(std::vector<std::uint64_t>()) {}
  // adopt a device set (used by the other facade classes)
  explicit KmerSet(SetPtr dev) : dev_(std::move(dev)) {}
  static KmerSet FromSortedBits(std::vector<std::uint64_t> bits) {
    KmerSet s;
    s.host_ = std::move(bits);
    return s;
  }

  std::int64_t Size() const {
    if (host_) return static_cast<std::int64_t>(host_->size());
    std::int64_t n = 0;
    std::lock_guard<std::mutex> l(Device::Mu());
    Device::Check(kmsc_set_size(Device::Ctx(), dev_->set, &n), "kmsc_set_size");
    return n;
  }
  void Clear() { host_ = std::vector<std::uint64_t>(); dev_.reset(); }

  void Add(const Kmer<K>& kmer) {
    EnsureHost();
    dev_.reset();
    auto it = std::lower_bound(host_->begin(), host_->end(), kmer.Bits());
    if (it == host_->end() || *it != kmer.Bits()) host_->insert(it, kmer.Bits());
  }
  void Remove(const Kmer<K>& kmer) {
    EnsureHost();
    dev_.reset();
    auto it = std::lower_bound(host_->begin(), host_->end(), kmer.Bits());
    if (it != host_->end() && *it == kmer.Bits()) host_->erase(it);
  }
  bool Contains(const Kmer<K>& kmer) const {
    EnsureHost();
    return std::binary_search(host_->begin(), host_->end(), kmer.Bits());
  }
  void Reserve(std::int64_t n) { if (host_) host_->reserve(static_cast<std::size_t>(n)); }

  template <typename PredType>
  std::vector<Kmer<K>> Find(PredType pred, int /*n_workers*/, std::int64_t /*estimated_size*/ = 0) const {
    EnsureHost();
    std::vector<Kmer<K>> out;
    for (std::uint64_t b : *host_)
      if (pred(Kmer<K>(b))) out.emplace_back(b);
    return out;
  }
  std::vector<Kmer<K>> Find(int n_workers) const {
    return Find([](const Kmer<K>&) { return true; }, n_workers);
  }

  // union / difference with another set: device merge (kmsc_set_union / kmsc_pair_split)
  KmerSet& Add(const KmerSet& other, int /*n_workers*/) {
    const kmsc_set* both[2] = {Dev()->set, other.Dev()->set};
    kmsc_set* u = nullptr;
    {
      std::lock_guard<std::mutex> l(Device::Mu());
      Device::Check(kmsc_set_union(Device::Ctx(), both, 2, &u), "kmsc_set_union");
    }
    Adopt(u);
    return *this;
  }
  KmerSet& Sub(const KmerSet& other, int /*n_workers*/) {
    kmsc_set* minus = nullptr;
    const SetPtr a = Dev(), b = other.Dev();  // uploads take the device lock themselves
    {
      std::lock_guard<std::mutex> l(Device::Mu());
      Device::Check(kmsc_pair_split(Device::Ctx(), a->set, b->set, nullptr, &minus, nullptr), "kmsc_pair_split");
    }
    Adopt(minus);
    return *this;
  }
  std::int64_t Diff(const KmerSet& other, int /*n_workers*/) const {
    std::int64_t d = 0;
    const SetPtr a = Dev(), b = other.Dev();
    std::lock_guard<std::mutex> l(Device::Mu());
    Device::Check(kmsc_set_diff(Device::Ctx(), a->set, b->set, &d), "kmsc_set_diff");
    return d;
  }
  bool Equals(const KmerSet& other, int n_workers) const { return Diff(other, n_workers) == 0; }
  std::size_t Hash(int /*n_workers*/) const {
    std::uint64_t h = 0;
    const SetPtr a = Dev();
    std::lock_guard<std::mutex> l(Device::Mu());
    Device::Check(kmsc_set_hash(Device::Ctx(), a->set, &h), "kmsc_set_hash");
    return static_cast<std::size_t>(h);
  }

  // n = j & k, j \ n, k \ n in one device pass (reference kmer_set_set.h:332-343).
  // inter_size: |j & k| if the caller knows it exactly (an all-bucket weight), else -1.
  static void Split(const KmerSet& j, const KmerSet& k, KmerSet* inter, KmerSet* j_minus, KmerSet* k_minus,
                    std::int64_t inter_size = -1) {
    std::vector<KmerSet> a, b, c;
    SplitBatch({&j}, {&k}, inter_size >= 0 ? std::vector<std::int64_t>{inter_size} : std::vector<std::int64_t>{},
               inter ? &a : nullptr, j_minus ? &b : nullptr, k_minus ? &c : nullptr);
    if (inter) *inter = a[0];
    if (j_minus) *j_minus = b[0];
    if (k_minus) *k_minus = c[0];
  }
  // the same for m pairs in one streaming pass (kmsc_pair_split_batch); inter_sizes empty or one per pair
  static void SplitBatch(const std::vector<const KmerSet*>& js, const std::vector<const KmerSet*>& ks,
                         const std::vector<std::int64_t>& inter_sizes, std::vector<KmerSet>* inter,
                         std::vector<KmerSet>* j_minus, std::vector<KmerSet>* k_minus) {
    const std::size_t m = js.size();
    std::vector<SetPtr> hold;
    std::vector<const kmsc_set*> hj(m), hk(m);
    for (std::size_t p = 0; p < m; p++) {
      hold.push_back(js[p]->Dev()); hj[p] = hold.back()->set;
      hold.push_back(ks[p]->Dev()); hk[p] = hold.back()->set;
    }
    std::vector<kmsc_set*> a(inter ? m : 0), b(j_minus ? m : 0), c(k_minus ? m : 0);
    {
      std::lock_guard<std::mutex> l(Device::Mu());
      Device::Check(kmsc_pair_split_batch(Device::Ctx(), hj.data(), hk.data(), static_cast<std::int32_t>(m),
                                          inter_sizes.size() == m && m > 0 ? inter_sizes.data() : nullptr,
                                          inter ? a.data() : nullptr, j_minus ? b.data() : nullptr,
                                          k_minus ? c.data() : nullptr), "kmsc_pair_split_batch");
    }
    auto adopt = [m](std::vector<KmerSet>* out, std::vector<kmsc_set*>& h) {
      if (!out) return;
      out->assign(m, KmerSet());
      for (std::size_t p = 0; p < m; p++) (*out)[p].Adopt(h[p]);
    };
    adopt(inter, a); adopt(j_minus, b); adopt(k_minus, c);
  }

  // device handle, uploading the host vector if needed
  const SetPtr& Dev() const {
    if (!dev_) {
      kmsc_set* s = nullptr;
      std::lock_guard<std::mutex> l(Device::Mu());
      Device::Check(kmsc_set_from_kmers(Device::Ctx(), K, N, static_cast<int>(sizeof(KeyType)), host_->data(),
                                        static_cast<std::int64_t>(host_->size()), &s), "kmsc_set_from_kmers");
      dev_ = MakeSetPtr(s);
    }
    return dev_;
  }
  // ascending k-mer values (downloads once)
  const std::vector<std::uint64_t>& SortedBits() const { EnsureHost(); return *host_; }

 private:
  void Adopt(kmsc_set* s) { dev_ = MakeSetPtr(s); host_.reset(); }
  void EnsureHost() const {
    if (host_) return;
    std::int64_t n = 0;
    std::lock_guard<std::mutex> l(Device::Mu());
    Device::Check(kmsc_set_size(Device::Ctx(), dev_->set, &n), "kmsc_set_size");
    std::vector<std::int64_t> offs((std::size_t(1) << N) + 1);
    std::vector<KeyType> keys(static_cast<std::size_t>(n));
    Device::Check(kmsc_set_to_csr(Device::Ctx(), dev_->set, offs.data(), keys.data()), "kmsc_set_to_csr");
    std::vector<std::uint64_t> bits(static_cast<std::size_t>(n));
    constexpr int kb = 2 * K - N;
    for (std::size_t b = 0; b + 1 < offs.size(); b++)
      for (std::int64_t i = offs[b]; i < offs[b + 1]; i++)
        bits[static_cast<std::size_t>(i)] = (static_cast<std::uint64_t>(b) << kb) | static_cast<std::uint64_t>(keys[static_cast<std::size_t>(i)]);
    host_ = std::move(bits);
  }

  mutable std::optional<std::vector<std::uint64_t>> host_;
  mutable SetPtr dev_;
};

template <int K, int N, typename KeyType>
KmerSet<K, N, KeyType> Add(KmerSet<K, N, KeyType> lhs, const KmerSet<K, N, KeyType>& rhs, int n_workers) {
  return lhs.Add(rhs, n_workers);
}
template <int K, int N, typename KeyType>
KmerSet<K, N, KeyType> Sub(KmerSet<K, N, KeyType> lhs, const KmerSet<K, N, KeyType>& rhs, int n_workers) {
  return lhs.Sub(rhs, n_workers);
}
template <int K, int N, typename KeyType>
KmerSet<K, N, KeyType> Intersection(KmerSet<K, N, KeyType> lhs, const KmerSet<K, N, KeyType>& rhs, int /*n_workers*/) {
  KmerSet<K, N, KeyType> inter;
  KmerSet<K, N, KeyType>::Split(lhs, rhs, &inter, nullptr, nullptr);
  return inter;
}

}  // namespace kmsc
#endif
