// Host-only: MstTree of the C++ facade (host/kmsc/kmer_set_set.h: candidate edges by (d, i, j), Kruskal with
// ParallelDisjointSet -- reference lib/core/parallel_disjoint_set.h:53-106 -- breadth-first orientation from set 0).
// Reads "n" and an n x n intersection matrix from stdin, prints one "parent child distance" line per tree edge;
// tests/test_host_mst.py compares with the oracle's restatement (kmsc_o_mst) on matrices full of ties. No device call.
#include <cstdint>
#include <cstdio>
#include <iostream>
#include <vector>

#include "kmsc/kmer_set_set.h"

int main() {
  int n = 0;
  while (std::cin >> n) {
    std::vector<std::int64_t> W(static_cast<std::size_t>(n) * static_cast<std::size_t>(n));
    for (auto& w : W) std::cin >> w;
    const std::vector<kmsc::MstEdge> edges = kmsc::MstTree(W, n);
    std::printf("tree %d %zu\n", n, edges.size());
    for (const kmsc::MstEdge& e : edges) std::printf("%d %d %lld\n", e.parent, e.child, static_cast<long long>(e.distance));
  }
  return 0;
}
