// kmsc/parallel_disjoint_set.h -- wait-free union-find with the reference's interface
// and linking rule (lib/core/parallel_disjoint_set.h:15-111, after Anderson & Woll
// 1991): one atomic 64-bit word per element = rank << 32 | parent; the root with the
// lower (rank, index) is linked under the higher one; Find compresses paths by CAS.
// Serves the SPSS loop cut and the `mst` driver's Kruskal.
#ifndef KMSC_HOST_PARALLEL_DISJOINT_SET_H_
#define KMSC_HOST_PARALLEL_DISJOINT_SET_H_
#include <atomic>
#include <cstdint>
#include <memory>
#include <utility>

namespace kmsc {

class ParallelDisjointSet {
 public:
  explicit ParallelDisjointSet(int size) : n_(size), a_(new std::atomic<std::uint64_t>[size > 0 ? size : 1]) {
    for (int i = 0; i < size; i++) a_[i].store(static_cast<std::uint64_t>(i), std::memory_order_relaxed);
  }

  int Find(int x) {
    int root = x;
    for (int p = Parent(root); p != root; p = Parent(root)) root = p;
    // hang every node on the way that is still below the root directly under it
    int y = x;
    while (Less(y, root)) {
      std::uint64_t w = a_[y].load();
      const int next = static_cast<int>(w & 0xffffffffu);
      const std::uint64_t want = (w & 0xffffffff00000000ull) | static_cast<std::uint32_t>(root);
      a_[y].compare_exchange_weak(w, want);
      y = next;
    }
    return root;
  }

  bool IsSame(int x, int y) {
    for (;;) {
      x = Find(x);
      y = Find(y);
      if (x == y) return true;
      if (Parent(x) == x) return false;
    }
  }

  void Unite(int x, int y) {
    for (;;) {
      x = Find(x);
      y = Find(y);
      if (x == y) return;
      int rx = Rank(x), ry = Rank(y);
      if (rx > ry || (rx == ry && x > y)) { std::swap(x, y); std::swap(rx, ry); }
      if (!Relink(x, rx, y, rx)) continue;     // x (lower) goes under y
      if (rx == ry) Relink(y, ry, y, ry + 1);  // equal ranks: the new root grows
      return;
    }
  }

  int Size() const { return n_; }

 private:
  int Rank(int i) const { return static_cast<int>(a_[i].load() >> 32); }
  int Parent(int i) const { return static_cast<int>(a_[i].load() & 0xffffffffu); }
  bool Less(int x, int y) const {
    const int rx = Rank(x), ry = Rank(y);
    return rx != ry ? rx < ry : x < y;
  }
  // x must still be a root of rank old_rank; then parent := y, rank := new_rank
  bool Relink(int x, int old_rank, int y, int new_rank) {
    std::uint64_t expect = (static_cast<std::uint64_t>(old_rank) << 32) | static_cast<std::uint32_t>(x);
    const std::uint64_t want = (static_cast<std::uint64_t>(new_rank) << 32) | static_cast<std::uint32_t>(y);
    return a_[x].compare_exchange_strong(expect, want);
  }

  int n_;
  std::unique_ptr<std::atomic<std::uint64_t>[]> a_;
};

}  // namespace kmsc
#endif
