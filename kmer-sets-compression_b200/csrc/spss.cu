// spss.cu -- SPSS construction on the device (reference lib/core/spss.h:230-615 unitigs,
// :1039-1858 path cover, :1836-1858 FromKmerSet).
//
// The reference walks the de Bruijn graph with hash-set Contains() calls, builds unitigs, then stitches them
// greedily into a path cover; any output that spells every k-mer of the set exactly once is a valid SPSS
// (test/spss.cc:57-68, 113-124). Here the whole construction is data parallel:
//
//   ports     every k-mer i has a left port 2 i and a right port 2 i + 1 (of the stored -- canonical -- form).
//             cand_kernel finds the up to four k-mers of the set that extend each port (one binary search per
//             neighbour) and records the PORT they are entered through: appending to x lands on the left port of
//             the successor, or on its right port when the set stores the successor's reverse complement.
//   matching  a port may be linked to one candidate, mutually: each free port proposes to its smallest free
//             candidate, mutual proposals become links (a few rounds; the smallest free port with candidates
//             is matched every round). Ports with one candidate each (the unitig rule) pair up in round one;
//             the later rounds stitch unitigs the way the greedy path cover does. Links are symmetric, one per
//             port, never from a k-mer to itself: the result is a set of paths and (rarely) cycles.
//   ranking   a walk state 2 i + p = "k-mer i, left through port p". The successor of a state is fixed by the
//             link. Every state needs the free port its walk ends at and the distance to it. Work-efficient
//             path: a state is a splitter if its walk starts there or a hash of its id says so (1 in 64); every
//             splitter walks to the next one, pointer jumping over packed (successor, distance) words runs over
//             the splitters only (in place, no double buffering), every splitter walks its segment again and
//             writes the result of each state it passes. Each path is walked from both ends; the walk that
//             starts at the smaller free port wins, and a k-mer reads its position and its start off the losing
//             walk (which ends where the winning one begins).
//   cycles    hold no start splitter, so some state stays unranked: then all states are ranked by plain pointer
//             jumping, the ones that never reach a free port lie on cycles, a second jumping pass carries the
//             smallest state of each cycle, the link leaving that state is cut, and the ranking is redone.
//   emission  string lengths at the start k-mers -> exclusive scan -> every k-mer writes its K bases (start) or
//             its last base in walk orientation at offset(start) + K - 1 + position.
// Output is deterministic for a given set: strings ordered by their start k-mer.
#include <cstdio>
#include <cstring>

#include "kmsc_common.cuh"
#include "kmer_pipeline.cuh"
#include "scan.cuh"

namespace kmsc {
namespace {

constexpr int kFree = -1;
constexpr uint32_t kEnd = 0x80000000u;   // low word of a jump entry: the walk's end has been reached (n < 2^30 k-mers)

// position of k-mer w in the set, or -1. `fine` is the set's finest offset level (2^(N+F) + 1 entries, fine bucket
// = the top N + F bits of the k-mer): the binary search runs over the ~10 keys of one fine bucket, not over a bucket.
template <typename KeyT>
__device__ __forceinline__ int32_t find_kmer(const KeyT* __restrict__ keys, const uint32_t* __restrict__ fine,
                                             unsigned long long w, int fine_shift, unsigned long long kmask) {
  const uint32_t x = (uint32_t)(w >> fine_shift);
  const unsigned long long kq = w & kmask;
  uint32_t a = fine[x];
  const uint32_t end = fine[x + 1];
  uint32_t e = end;
  while (a < e) {
    const uint32_t mid = (a + e) >> 1;
    if ((unsigned long long)keys[mid] < kq) a = mid + 1; else e = mid;
  }
  return (a < end && (unsigned long long)keys[a] == kq) ? (int32_t)a : -1;
}

// cand[8 i + 4 p + c]: the port reached from port p of k-mer i with base c, or -1 (absent, or i itself).
template <typename KeyT>
__global__ void cand_kernel(const KeyT* __restrict__ keys, const uint32_t* __restrict__ offs, const uint32_t* __restrict__ fine,
                            int fine_shift, int n_buckets, int K, int key_bits, int canonical, int32_t* __restrict__ cand) {
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  const unsigned long long mask = K == 32 ? ~0ull : ((1ull << (2 * K)) - 1);
  const unsigned long long kmask = key_bits == 64 ? ~0ull : ((1ull << key_bits) - 1);
  for (int b = blockIdx.x * wpb + (threadIdx.x >> 5); b < n_buckets; b += gridDim.x * wpb) {
    const uint32_t lo = offs[b], hi = offs[b + 1];
    for (uint32_t i = lo + lane; i < hi; i += 32) {
      const unsigned long long v = ((unsigned long long)b << key_bits) | (unsigned long long)keys[i];
#pragma unroll
      for (int d = 0; d < 8; d++) {
        const unsigned long long c = (unsigned long long)(d & 3);
        const int right = d >= 4;   // port 1 = right: drop the first base, append c; port 0 = left: prepend c
        unsigned long long w = right ? (((v << 2) & mask) | c) : ((v >> 2) | (c << (2 * (K - 1))));
        int flip = 0;
        if (canonical) {
          const unsigned long long rc = revcomp(w, K);
          if (rc < w) { w = rc; flip = 1; }
        }
        const int32_t j = find_kmer(keys, fine, w, fine_shift, kmask);
        // a right extension enters the successor through its left port (0), a left extension through the
        // right port (1); the other one when the set holds the reverse complement of the extension
        int32_t q = -1;
        if (j >= 0 && (uint32_t)j != i) q = (j << 1) | ((right ? 0 : 1) ^ flip);
        cand[(size_t)i * 8 + d] = q;
      }
    }
  }
}

// every free port proposes to its smallest free candidate
__global__ void propose_kernel(const int32_t* __restrict__ cand, const int32_t* __restrict__ link, int64_t n_ports,
                               int32_t* __restrict__ prop) {
  const int64_t P = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (P >= n_ports) return;
  int32_t best = -1;
  if (link[P] == kFree) {
    const int4 c = reinterpret_cast<const int4*>(cand)[P];
    const int32_t q[4] = {c.x, c.y, c.z, c.w};
#pragma unroll
    for (int k = 0; k < 4; k++)
      if (q[k] >= 0 && link[q[k]] == kFree && (best < 0 || q[k] < best)) best = q[k];
  }
  prop[P] = best;
}

// mutual proposals become links
__global__ void accept_kernel(const int32_t* __restrict__ prop, int64_t n_ports, int32_t* __restrict__ link,
                              int* __restrict__ linked) {
  const int64_t P = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int32_t Q = P < n_ports ? prop[P] : -1;
  const bool join = Q >= 0 && prop[Q] == (int32_t)P;
  if (join) link[P] = Q;
  // one store per warp that linked anything, not one per port (millions of stores to one address)
  if (__any_sync(__activemask(), join) && (threadIdx.x & 31) == 0) *linked = 1;
}

// walk state s = 2 i + p (k-mer i left through port p): successor and distance, packed (dist << 32 | succ).
// A free exit port ends the walk: the state points at itself with kEnd set, and a state that has reached
// the end of its walk carries the flag too (a state on a cycle may well come to point at itself: a cycle
// whose length divides the jump -- only the flag says "ended").
__global__ void jump_init_kernel(const int32_t* __restrict__ link, int64_t n_states, unsigned long long* __restrict__ jump) {
  const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_states) return;
  const int32_t Q = link[s];
  // entering k-mer Q >> 1 through port Q & 1, the walk leaves it through the other port: state Q ^ 1
  jump[s] = Q == kFree ? (unsigned long long)((uint32_t)s | kEnd) : ((1ull << 32) | (unsigned long long)(uint32_t)(Q ^ 1));
}

__global__ void jump_kernel(unsigned long long* __restrict__ jump, int64_t n_states, int* __restrict__ changed) {
  const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_states) return;
  const unsigned long long a = jump[s];
  if ((uint32_t)a & kEnd) return;         // ended
  const unsigned long long b = jump[(uint32_t)a];   // one aligned 8-byte load: a consistent (successor, distance) pair
  jump[s] = (((a >> 32) + (b >> 32)) << 32) | (unsigned long long)(uint32_t)b;
  *changed = 1;
}

// ---- work-efficient ranking: splitters -------------------------------------------------------
// Pointer jumping over all 2 n states costs a random 8-byte load and store per state per round, ~15 rounds.
// Instead: a state is a SPLITTER if its walk starts there (the port behind it is free) or a hash of its id
// says so (1 in 64). (1) every splitter walks to the next splitter or the end of its walk (~64 dependent
// loads, hundreds of thousands of walkers in flight), (2) pointer jumping runs over the splitters only,
// (3) every splitter walks its segment again and writes the final (distance, end) of each state it passes.
// A cycle holds no start splitter: its states stay unwritten (or its sampled splitters never reach an end);
// the caller counts them and, if there are any, takes the full pointer jumping with cycle cutting below.
__device__ __forceinline__ bool sampled_splitter(uint32_t s) { return ((s * 0x9E3779B1u) >> 26) == 0u; }

__global__ void split_walk_kernel(const int32_t* __restrict__ link, int64_t n_states, unsigned long long* __restrict__ sj,
                                  uint32_t* __restrict__ list, uint32_t* __restrict__ n_list, uint32_t cap) {
  const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_states) return;
  if (link[s ^ 1] != kFree && !sampled_splitter((uint32_t)s)) return;
  uint32_t cur = (uint32_t)s, d = 0;
  unsigned long long out = 0;
  for (;;) {
    const int32_t Q = link[cur];
    if (Q == kFree) { out = ((unsigned long long)d << 32) | (unsigned long long)(cur | kEnd); break; }
    const uint32_t nxt = (uint32_t)(Q ^ 1);
    d++;
    if (sampled_splitter(nxt)) { out = ((unsigned long long)d << 32) | (unsigned long long)nxt; break; }
    if (d > (1u << 24)) { out = ((unsigned long long)d << 32) | (unsigned long long)nxt; break; }   // (a long splitter-free cycle: left unresolved)
    cur = nxt;
  }
  sj[s] = out;
  const uint32_t at = atomicAdd(n_list, 1u);
  if (at < cap) list[at] = (uint32_t)s;
}

__global__ void split_jump_kernel(const uint32_t* __restrict__ list, const uint32_t* __restrict__ n_list, uint32_t cap,
                                  unsigned long long* __restrict__ sj, int* __restrict__ changed) {
  const uint32_t n = min(*n_list, cap);
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  const uint32_t s = list[t];
  const unsigned long long a = sj[s];
  if ((uint32_t)a & kEnd) return;
  const unsigned long long b = sj[(uint32_t)a];
  sj[s] = (((a >> 32) + (b >> 32)) << 32) | (unsigned long long)(uint32_t)b;
  *changed = 1;
}

__global__ void split_fill_kernel(const int32_t* __restrict__ link, const uint32_t* __restrict__ list,
                                  const uint32_t* __restrict__ n_list, uint32_t cap, const unsigned long long* __restrict__ sj,
                                  unsigned long long* __restrict__ jump) {
  const uint32_t n = min(*n_list, cap);
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  const uint32_t s = list[t];
  const unsigned long long a = sj[s];
  if (!((uint32_t)a & kEnd)) return;     // on a cycle
  const uint32_t E = (uint32_t)a;
  uint32_t D = (uint32_t)(a >> 32), cur = s;
  for (;;) {
    jump[cur] = ((unsigned long long)D << 32) | (unsigned long long)E;
    const int32_t Q = link[cur];
    if (Q == kFree) break;
    const uint32_t nxt = (uint32_t)(Q ^ 1);
    if (sampled_splitter(nxt) || D == 0) break;
    D--;
    cur = nxt;
  }
}

__global__ void count_unranked_kernel(const unsigned long long* __restrict__ jump, int64_t n_states, unsigned long long* __restrict__ n_bad) {
  const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool bad = s < n_states && !((uint32_t)jump[s] & kEnd);
  const unsigned b = __ballot_sync(0xffffffffu, bad);
  if ((threadIdx.x & 31) == 0 && b) atomicAdd(n_bad, (unsigned long long)__popc(b));
}

// states that did not reach a free port lie on cycles: (smallest state seen << 32 | successor)
__global__ void cycle_init_kernel(const int32_t* __restrict__ link, const unsigned long long* __restrict__ jump,
                                  int64_t n_states, unsigned long long* __restrict__ cyc) {
  const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_states) return;
  const bool ended = ((uint32_t)jump[s] & kEnd) != 0u;
  const int32_t Q = link[s];
  // off-cycle states are self loops carrying "no state"
  cyc[s] = (ended || Q == kFree) ? ((0xFFFFFFFFull << 32) | (unsigned long long)s)
                                 : (((unsigned long long)s << 32) | (unsigned long long)(uint32_t)(Q ^ 1));
}

__global__ void cycle_jump_kernel(unsigned long long* __restrict__ cyc, int64_t n_states) {
  const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_states) return;
  const unsigned long long a = cyc[s];
  const uint32_t q = (uint32_t)a;
  if ((int64_t)q == s) return;
  const unsigned long long b = cyc[q];
  const uint32_t m = min((uint32_t)(a >> 32), (uint32_t)(b >> 32));
  cyc[s] = ((unsigned long long)m << 32) | (unsigned long long)(uint32_t)b;
}

// the link that leaves the smallest state of a cycle is cut (both of its ends)
__global__ void cycle_cut_kernel(const unsigned long long* __restrict__ cyc, int64_t n_states, int32_t* __restrict__ link,
                                 unsigned long long* __restrict__ n_cut) {
  const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_states) return;
  if ((uint32_t)(cyc[s] >> 32) != (uint32_t)s) return;
  const int32_t Q = link[s];
  if (Q == kFree) return;
  link[s] = kFree;
  link[Q] = kFree;
  atomicAdd(n_cut, 1ull);
}

// per k-mer: which of its two walks wins, its position, its start; string length at the start k-mers
__global__ void rank_kernel(const unsigned long long* __restrict__ jump, int64_t n, int K, int canonical, uint32_t* __restrict__ info,
                            uint32_t* __restrict__ start, uint32_t* __restrict__ slen, uint32_t* __restrict__ sflag) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const unsigned long long j0 = jump[2 * i], j1 = jump[2 * i + 1];
  // walk ends (= the free port there). The walk through state s starts at the free port the OTHER walk of
  // this k-mer ends at; the walk starting at the smaller free port is the one that is written.
  // Without reverse complements (canonical == 0) every link joins a right port to a left port and only
  // the left-to-right walk spells the k-mers themselves.
  const uint32_t e0 = (uint32_t)j0 & ~kEnd, e1 = (uint32_t)j1 & ~kEnd;
  const int p = (!canonical || e0 < e1) ? 1 : 0;       // leave through the right port = forward orientation
  const unsigned long long win = p ? j1 : j0, lose = p ? j0 : j1;
  const uint32_t pos = (uint32_t)(lose >> 32);
  const uint32_t len = (uint32_t)(win >> 32) + pos + 1u;   // k-mers on the path
  info[i] = (pos << 1) | (uint32_t)p;
  start[i] = ((uint32_t)lose & ~kEnd) >> 1;
  slen[i] = pos == 0 ? len + (uint32_t)K - 1u : 0u;
  sflag[i] = pos == 0 ? 1u : 0u;
}

template <typename KeyT>
__global__ void emit_kernel(const KeyT* __restrict__ keys, const uint32_t* __restrict__ offs, int n_buckets, int K,
                            int key_bits, const uint32_t* __restrict__ info, const uint32_t* __restrict__ start,
                            const uint32_t* __restrict__ soff, const uint32_t* __restrict__ sidx, int64_t n,
                            char* __restrict__ text, long long* __restrict__ str_offs) {
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  for (int b = blockIdx.x * wpb + (threadIdx.x >> 5); b < n_buckets; b += gridDim.x * wpb) {
    const uint32_t lo = offs[b], hi = offs[b + 1];
    for (uint32_t i = lo + lane; i < hi; i += 32) {
      const unsigned long long v = ((unsigned long long)b << key_bits) | (unsigned long long)keys[i];
      const uint32_t in = info[i];
      const uint32_t pos = in >> 1;
      const unsigned long long o = (in & 1u) ? v : revcomp(v, K);
      const uint32_t off = soff[start[i]];
      if (pos == 0) {
        for (int t = 0; t < K; t++) text[(size_t)off + t] = "ACGT"[(o >> (2 * (K - 1 - t))) & 3ull];
        str_offs[sidx[i]] = (long long)off;
      } else {
        text[(size_t)off + (size_t)(K - 1) + pos] = "ACGT"[o & 3ull];
      }
    }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) str_offs[sidx[n]] = (long long)soff[n];
}

// text (one byte per base, strings back to back) -> 2 bits per base, 32 bases per word, first base in the top
// bits: the container KmerSetCompact holds (kmsc_set_from_packed reads the same layout)
__global__ void pack_text_kernel(const char* __restrict__ text, long long n_chars, unsigned long long* __restrict__ words,
                                 long long n_words) {
  const long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= n_words) return;
  unsigned long long v = 0;
  const long long base = w * 32;
#pragma unroll 8
  for (int j = 0; j < 32; j++) {
    unsigned long long code = 0;
    if (base + j < n_chars) {
      const unsigned ch = (unsigned char)text[base + j];
      code = (ch >> 1) & 3u;      // A 0, C 1, T 2, G 3 ...
      code ^= code >> 1;          // ... -> A 0, C 1, G 2, T 3
    }
    v = (v << 2) | code;
  }
  words[w] = v;
}

}  // namespace
}  // namespace kmsc

using namespace kmsc;

extern "C" int kmsc_spss_build(kmsc_ctx* ctx, const kmsc_set* set, int canonical, int rounds, int64_t* n_strings,
                               int64_t* n_chars) {
  if (!ctx || !set || !n_strings || !n_chars) { set_error("NULL argument"); return KMSC_E_INVALID; }
  if (set->n_keys >= ((int64_t)1 << 30)) { set_error("set too large for SPSS construction (%lld keys)", (long long)set->n_keys); return KMSC_E_INVALID; }
  if (rounds <= 0) rounds = 8;
  ctx->spss_strings = 0;
  ctx->spss_chars = 0;
  *n_strings = 0;
  *n_chars = 0;
  const int64_t n = set->n_keys;
  if (n == 0) return KMSC_OK;
  KMSC_TRY(set_ensure_levels(ctx, set));
  if (set->has_dups < 0) KMSC_TRY(set_check_dups(ctx, const_cast<kmsc_set*>(set)));
  if (set->has_dups == 1) { set_error("the set holds duplicate keys (an SPSS spells every k-mer once: build it from a true set)"); return KMSC_E_INVALID; }
  if ((double)n * set->K >= 4.0e9) { set_error("SPSS text may exceed 2^32 characters"); return KMSC_E_INVALID; }
  KMSC_CUDA(cudaSetDevice(ctx->device));
  const int64_t np = 2 * n;
  const size_t scan_e = scan_scratch_entries((uint64_t)n + 1);
  size_t off = 0;
  auto take = [&](size_t bytes) { const size_t o = off; off += (bytes + 255) & ~(size_t)255; return o; };
  const size_t o_cand = take((size_t)n * 8 * 4), o_link = take((size_t)np * 4), o_prop = take((size_t)np * 4),
               o_jump = take((size_t)np * 8), o_info = take((size_t)n * 4), o_start = take((size_t)n * 4),
               o_slen = take((size_t)(n + 1) * 4), o_sflag = take((size_t)(n + 1) * 4), o_soff = take((size_t)(n + 2) * 4),
               o_sidx = take((size_t)(n + 2) * 4), o_bsum = take(scan_e * 4), o_misc = take(256);
  KMSC_TRY(ctx->work3.reserve(off));
  unsigned char* base = (unsigned char*)ctx->work3.p;
  int32_t* d_cand = (int32_t*)(base + o_cand);
  int32_t* d_link = (int32_t*)(base + o_link);
  int32_t* d_prop = (int32_t*)(base + o_prop);
  unsigned long long* d_jump = (unsigned long long*)(base + o_jump);
  unsigned long long* d_cyc = (unsigned long long*)(base + o_cand);   // the candidate table is dead by then
  uint32_t* d_info = (uint32_t*)(base + o_info);
  uint32_t* d_start = (uint32_t*)(base + o_start);
  uint32_t* d_slen = (uint32_t*)(base + o_slen);
  uint32_t* d_sflag = (uint32_t*)(base + o_sflag);
  uint32_t* d_soff = (uint32_t*)(base + o_soff);
  uint32_t* d_sidx = (uint32_t*)(base + o_sidx);
  uint32_t* d_bsum = (uint32_t*)(base + o_bsum);
  int* d_changed = (int*)(base + o_misc);
  unsigned long long* d_ncut = (unsigned long long*)(base + o_misc + 8);
  uint32_t* d_totals = (uint32_t*)(base + o_misc + 16);   // [0] characters, [1] strings

  const int nb = 1 << set->N;
  int bblocks = (nb + 7) / 8;
  if (bblocks > ctx->sm_count * 16) bblocks = ctx->sm_count * 16;
  const unsigned pblocks = (unsigned)((np + 255) / 256), nblocks = (unsigned)((n + 255) / 256);
  switch (set->key_bytes) {
    case 2: cand_kernel<uint16_t><<<bblocks, 256, 0, ctx->stream>>>((const uint16_t*)set->keys, set->lev[0], set->lev[set->max_level], set->key_bits - set->max_level, nb, set->K, set->key_bits, canonical, d_cand); break;
    case 4: cand_kernel<uint32_t><<<bblocks, 256, 0, ctx->stream>>>((const uint32_t*)set->keys, set->lev[0], set->lev[set->max_level], set->key_bits - set->max_level, nb, set->K, set->key_bits, canonical, d_cand); break;
    default: cand_kernel<unsigned long long><<<bblocks, 256, 0, ctx->stream>>>((const unsigned long long*)set->keys, set->lev[0], set->lev[set->max_level], set->key_bits - set->max_level, nb, set->K, set->key_bits, canonical, d_cand); break;
  }
  count_launch(ctx);
  KMSC_CUDA(cudaMemsetAsync(d_link, 0xFF, (size_t)np * 4, ctx->stream));
  void* pin = nullptr;
  KMSC_TRY(ctx_pinned(ctx, 64, &pin));
  int* h_changed = (int*)pin;
  // the first two rounds join nearly every port that can be joined; from then on a round that made no link
  // ends the matching (read back every second round)
  for (int r = 0; r < rounds; r++) {
    if (r >= 2 && (r & 1) == 0) KMSC_CUDA(cudaMemsetAsync(d_changed, 0, 4, ctx->stream));
    propose_kernel<<<pblocks, 256, 0, ctx->stream>>>(d_cand, d_link, np, d_prop);
    accept_kernel<<<pblocks, 256, 0, ctx->stream>>>(d_prop, np, d_link, d_changed);
    count_launch(ctx, 2);
    if (r >= 3 && (r & 1) == 1 && r + 1 < rounds) {
      KMSC_CUDA(cudaMemcpyAsync(h_changed, d_changed, 4, cudaMemcpyDeviceToHost, ctx->stream));
      KMSC_CUDA(cudaStreamSynchronize(ctx->stream));
      if (*h_changed == 0) break;
    }
  }
  KMSC_CUDA(cudaGetLastError());

  unsigned long long* h_ncut = (unsigned long long*)((unsigned char*)pin + 8);
  uint32_t* h_totals = (uint32_t*)((unsigned char*)pin + 16);
  int log2n = 1;
  while ((1ll << log2n) < np) log2n++;
  // fast path: ranking through splitters; complete unless the links hold a cycle
  bool ranked = false;
  {
    unsigned long long* d_sj = (unsigned long long*)(base + o_cand);   // the candidate table is dead: (next, distance) of the splitters
    uint32_t* d_list = (uint32_t*)(base + o_prop);                     // the proposals are dead: the list of splitters
    uint32_t* d_nlist = (uint32_t*)(base + o_misc + 32);
    const uint32_t cap = (uint32_t)np;
    KMSC_CUDA(cudaMemsetAsync(d_nlist, 0, 4, ctx->stream));
    KMSC_CUDA(cudaMemsetAsync(d_jump, 0, (size_t)np * 8, ctx->stream));
    split_walk_kernel<<<pblocks, 256, 0, ctx->stream>>>(d_link, np, d_sj, d_list, d_nlist, cap);
    count_launch(ctx);
    // splitters are ~1 / 32 of the states (starts + 1 in 64): a grid over np / 8 threads covers any list
    const unsigned lblocks = (unsigned)((np / 8 + 255) / 256 + 1);
    uint32_t* h_nlist = (uint32_t*)((unsigned char*)pin + 32);
    KMSC_CUDA(cudaMemcpyAsync(h_nlist, d_nlist, 4, cudaMemcpyDeviceToHost, ctx->stream));
    KMSC_CUDA(cudaStreamSynchronize(ctx->stream));
    if ((uint64_t)*h_nlist <= (uint64_t)lblocks * 256) {
      bool converged = false;
      for (int r = 0; r < log2n + 6 && !converged; r += 4) {
        KMSC_CUDA(cudaMemsetAsync(d_changed, 0, 4, ctx->stream));
        for (int k = 0; k < 4; k++) split_jump_kernel<<<lblocks, 256, 0, ctx->stream>>>(d_list, d_nlist, cap, d_sj, d_changed);
        count_launch(ctx, 4);
        KMSC_CUDA(cudaMemcpyAsync(h_changed, d_changed, 4, cudaMemcpyDeviceToHost, ctx->stream));
        KMSC_CUDA(cudaStreamSynchronize(ctx->stream));
        converged = *h_changed == 0;
      }
      KMSC_CUDA(cudaMemsetAsync(d_ncut, 0, 8, ctx->stream));
      split_fill_kernel<<<lblocks, 256, 0, ctx->stream>>>(d_link, d_list, d_nlist, cap, d_sj, d_jump);
      count_unranked_kernel<<<pblocks, 256, 0, ctx->stream>>>(d_jump, np, d_ncut);
      count_launch(ctx, 2);
      KMSC_CUDA(cudaMemcpyAsync(h_ncut, d_ncut, 8, cudaMemcpyDeviceToHost, ctx->stream));
      KMSC_CUDA(cudaStreamSynchronize(ctx->stream));
      ranked = *h_ncut == 0;
    }
  }
  for (int attempt = 0; attempt < 3 && !ranked; attempt++) {
    jump_init_kernel<<<pblocks, 256, 0, ctx->stream>>>(d_link, np, d_jump);
    count_launch(ctx);
    bool converged = false;
    // in-place jumping at least doubles every walk's reach per round; a flag is read back every four rounds
    for (int r = 0; r < log2n + 6 && !converged; r += 4) {
      KMSC_CUDA(cudaMemsetAsync(d_changed, 0, 4, ctx->stream));
      for (int k = 0; k < 4; k++) jump_kernel<<<pblocks, 256, 0, ctx->stream>>>(d_jump, np, d_changed);
      count_launch(ctx, 4);
      KMSC_CUDA(cudaMemcpyAsync(h_changed, d_changed, 4, cudaMemcpyDeviceToHost, ctx->stream));
      KMSC_CUDA(cudaStreamSynchronize(ctx->stream));
      converged = *h_changed == 0;
    }
    if (converged) break;
    if (attempt == 2) { set_error("SPSS construction: cycles left after two cuts"); return KMSC_E_STATE; }
    // cycles: cut each at its smallest state, rank again
    KMSC_CUDA(cudaMemsetAsync(d_ncut, 0, 8, ctx->stream));
    cycle_init_kernel<<<pblocks, 256, 0, ctx->stream>>>(d_link, d_jump, np, d_cyc);
    for (int r = 0; r < log2n + 1; r++) cycle_jump_kernel<<<pblocks, 256, 0, ctx->stream>>>(d_cyc, np);
    cycle_cut_kernel<<<pblocks, 256, 0, ctx->stream>>>(d_cyc, np, d_link, d_ncut);
    count_launch(ctx, log2n + 3);
    KMSC_CUDA(cudaMemcpyAsync(h_ncut, d_ncut, 8, cudaMemcpyDeviceToHost, ctx->stream));
    KMSC_CUDA(cudaStreamSynchronize(ctx->stream));
    if (*h_ncut == 0) { set_error("SPSS construction: walks do not end and no cycle was found"); return KMSC_E_STATE; }
  }

  rank_kernel<<<nblocks, 256, 0, ctx->stream>>>(d_jump, n, set->K, canonical, d_info, d_start, d_slen, d_sflag);
  count_launch(ctx);
  KMSC_TRY(exclusive_scan_u32(ctx, d_slen, d_soff, (uint64_t)n, d_bsum, d_totals));
  KMSC_TRY(exclusive_scan_u32(ctx, d_sflag, d_sidx, (uint64_t)n, d_bsum, d_totals + 1));
  KMSC_CUDA(cudaMemcpyAsync(h_totals, d_totals, 8, cudaMemcpyDeviceToHost, ctx->stream));
  KMSC_CUDA(cudaStreamSynchronize(ctx->stream));
  const int64_t chars = h_totals[0], strings = h_totals[1];
  KMSC_TRY(ctx->spss_out.reserve((size_t)chars + 64 + (size_t)(strings + 1) * 8 + 64));
  char* d_text = (char*)ctx->spss_out.p;
  long long* d_offs = (long long*)((unsigned char*)ctx->spss_out.p + (((size_t)chars + 63) & ~(size_t)63));
  switch (set->key_bytes) {
    case 2: emit_kernel<uint16_t><<<bblocks, 256, 0, ctx->stream>>>((const uint16_t*)set->keys, set->lev[0], nb, set->K, set->key_bits, d_info, d_start, d_soff, d_sidx, n, d_text, d_offs); break;
    case 4: emit_kernel<uint32_t><<<bblocks, 256, 0, ctx->stream>>>((const uint32_t*)set->keys, set->lev[0], nb, set->K, set->key_bits, d_info, d_start, d_soff, d_sidx, n, d_text, d_offs); break;
    default: emit_kernel<unsigned long long><<<bblocks, 256, 0, ctx->stream>>>((const unsigned long long*)set->keys, set->lev[0], nb, set->K, set->key_bits, d_info, d_start, d_soff, d_sidx, n, d_text, d_offs); break;
  }
  count_launch(ctx);
  KMSC_CUDA(cudaGetLastError());
  ctx->spss_strings = strings;
  ctx->spss_chars = chars;
  *n_strings = strings;
  *n_chars = chars;
  return KMSC_OK;
}

extern "C" int kmsc_spss_fetch(kmsc_ctx* ctx, char* text, int64_t* str_offs) {
  if (!ctx || !str_offs || (ctx->spss_chars > 0 && !text)) { set_error("NULL argument"); return KMSC_E_INVALID; }
  KMSC_CUDA(cudaSetDevice(ctx->device));
  if (ctx->spss_strings == 0) { str_offs[0] = 0; return KMSC_OK; }
  const size_t chars = (size_t)ctx->spss_chars;
  const unsigned char* base = (const unsigned char*)ctx->spss_out.p;
  KMSC_CUDA(cudaMemcpyAsync(text, base, chars, cudaMemcpyDeviceToHost, ctx->stream));
  KMSC_CUDA(cudaMemcpyAsync(str_offs, base + ((chars + 63) & ~(size_t)63), (size_t)(ctx->spss_strings + 1) * 8,
                            cudaMemcpyDeviceToHost, ctx->stream));
  KMSC_CUDA(cudaStreamSynchronize(ctx->stream));
  return KMSC_OK;
}

extern "C" int kmsc_spss_fetch_packed(kmsc_ctx* ctx, uint64_t* words, int64_t* str_offs) {
  if (!ctx || !str_offs || (ctx->spss_chars > 0 && !words)) { set_error("NULL argument"); return KMSC_E_INVALID; }
  KMSC_CUDA(cudaSetDevice(ctx->device));
  if (ctx->spss_strings == 0) { str_offs[0] = 0; return KMSC_OK; }
  const size_t chars = (size_t)ctx->spss_chars;
  const long long n_words = (long long)((chars + 31) / 32);
  const unsigned char* base = (const unsigned char*)ctx->spss_out.p;
  KMSC_TRY(ctx->work2.reserve((size_t)n_words * 8 + 64));
  unsigned long long* d_words = (unsigned long long*)ctx->work2.p;
  pack_text_kernel<<<(unsigned)((n_words + 255) / 256), 256, 0, ctx->stream>>>((const char*)base, (long long)chars, d_words, n_words);
  count_launch(ctx);
  KMSC_CUDA(cudaGetLastError());
  KMSC_CUDA(cudaMemcpyAsync(words, d_words, (size_t)n_words * 8, cudaMemcpyDeviceToHost, ctx->stream));
  KMSC_CUDA(cudaMemcpyAsync(str_offs, base + ((chars + 63) & ~(size_t)63), (size_t)(ctx->spss_strings + 1) * 8,
                            cudaMemcpyDeviceToHost, ctx->stream));
  KMSC_CUDA(cudaStreamSynchronize(ctx->stream));
  return KMSC_OK;
}
