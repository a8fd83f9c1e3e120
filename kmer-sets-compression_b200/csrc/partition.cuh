// partition.cuh -- the staged two-level partition that turns k-mer positions into a sorted,
// fine-bucketed key array (P2 decode and the back end of P1 counting).
//
// Reference work replaced: the per-position insert + per-bucket std::sort of
// KmerSetCompact::GetSampledKmerSet (lib/core/kmer_set_compact.h:138-200) and the hash-set
// inserts of GetKmerSetFromSPSS (lib/core/spss.h:1898-1935).
//
// The k-mers of a sequence are pseudo-random, so a direct scatter into 2^(N+F) fine buckets is
// one 4-byte store per 32-byte sector plus one global atomic per k-mer (the r01 pipeline: 0.37 ms
// per 10 M k-mers). Here:
//   count      every CTA walks its share of the job's tiles (8192 positions each) and keeps a
//              shared-memory histogram over the top B1 bits of the 2K-bit value (B1 = 11 at 10 M
//              k-mers); it is written out per CTA: no global atomics anywhere
//   scan       one CTA per job: bin bases, the start of every (bin, CTA) slice, the largest bin
//   partition  the same CTAs walk the same tiles: extract once, histogram in shared memory, the
//              tile's k-mers are grouped by bin IN SHARED MEMORY and written as contiguous runs
//              at the CTA's running cursor of each bin
//   sort       one CTA per partition (about 5 K keys): counting sort in shared memory over
//              sub-bins (about one key each), the few keys of a sub-bin are ordered by their
//              owning thread, the keys are written once, coalesced, together with the finest
//              offset level of the set. A repeated key raises the job's flag.
// Every key is extracted twice (count + partition), written twice and read twice; all global
// traffic is coalesced runs. Jobs (sets) are batched: blockIdx.y = job.
#pragma once
#include "kmer_pipeline.cuh"

namespace kmsc {
namespace part {

constexpr int kTile = 8192;       // positions per tile (multiple of 32)
constexpr int kThreads = 512;     // count / partition kernels
constexpr int kSortThreads = 256; // sort kernel
constexpr int kMaxB1 = 13;        // bins of the first level: at most 8192
constexpr int kMaxSub = 13;       // sub-bins per partition: at most 8192
constexpr int kLongRun = 24;      // sub-bin runs longer than this are sorted by the whole CTA
constexpr int kLongCap = 128;     // such runs per partition (beyond: the owning thread heap-sorts); 1 KB: three CTAs fit an SM

struct Job {
  const unsigned long long* words;  // 2-bit codes, 32 per word, first base in the top bits
  const uint32_t* bad;              // bit p set = no k-mer starts at p
  unsigned long long n_pos;
  uint32_t* base;     // [2^B1 + 1]: exclusive bin bases (base[2^B1] = n_occ), written by the scan
  uint32_t* slice;    // [n_ctas][2^B1]: per-CTA bin counts, after the scan the start of the CTA's slice of the bin
  uint32_t* meta;     // [0] n_occ, [1] largest partition, [2] repeat flag, [3] copies dropped by the sort
  uint32_t* removed;  // [2^B1]: copies the sort dropped from each partition (Geo::dedup)
  void* tmp;          // first-level output: residuals grouped by bin
  void* keys;         // result: keys ascending inside every fine bucket
  uint32_t* fine;     // result: finest offset level, 2^(N+F) + 1 entries
};

struct Geo {
  int K, V, key_bits, canonical;
  int B1, R1;        // first-level bits, residual bits V - B1
  int FB, FB2;       // fine-bucket bits inside a partition (N + F - B1), sub-bin bits (>= FB)
  int dedup;         // the sort drops repeated keys inside its partition (a shift pass closes the gaps)
  unsigned long long bucket_lo, bucket_hi;
};

__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t t = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += t;
  }
  return v;
}

// exclusive scan of a[0..n) in shared memory by the whole block; a[n] receives the total.
// Every warp scans a contiguous chunk 32 entries at a time (lane = entry: no bank conflicts),
// then the warp totals are scanned and added. wsum: 33 words of shared memory. Ends with a barrier.
__device__ __forceinline__ uint32_t block_excl_scan(uint32_t* a, int n, uint32_t* wsum) {
  const int nw = blockDim.x >> 5, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int chunk = (((n + nw - 1) / nw) + 31) & ~31;
  const int lo = warp * chunk, hi = min(n, lo + chunk);
  uint32_t carry = 0;
  for (int i = lo; i < hi; i += 32) {
    const int idx = i + lane;
    const uint32_t v = idx < hi ? a[idx] : 0u;
    const uint32_t inc = warp_incl_scan(v, lane);
    if (idx < hi) a[idx] = carry + inc - v;
    carry += __shfl_sync(0xffffffffu, inc, 31);
  }
  if (lane == 0) wsum[warp] = carry;
  __syncthreads();
  if (warp == 0) {
    const uint32_t w = lane < nw ? wsum[lane] : 0u;
    const uint32_t wi = warp_incl_scan(w, lane);
    wsum[lane] = wi - w;
    if (lane == 31) wsum[32] = wi;
  }
  __syncthreads();
  const uint32_t add = wsum[warp];
  if (add)
    for (int i = lo + lane; i < hi; i += 32) a[i] += add;
  const uint32_t total = wsum[32];
  if (tid == 0) a[n] = total;
  __syncthreads();
  return total;
}

// The kR = kTile / kThreads consecutive k-mers starting at tile position i0 (a multiple of kR),
// rolled: the first window is cut out of the staged words, every further one shifts one base in
// (forward strand) and its complement in from the other end (reverse strand), as the reference's
// Kmer::Next does one k-mer at a time (lib/core/kmer.h:136-160); canonical = the smaller of the two
// (kmer.h:133). emit(j, v) is called for every position that starts a k-mer kept by the bucket filter.
constexpr int kR = kTile / kThreads;
static_assert(kR == 16, "one 16-bit slice of the bad-position mask per thread");

template <typename Emit>
__device__ __forceinline__ void tile_kmers(const unsigned long long* sw, const uint32_t* sbad, uint32_t i0,
                                           unsigned long long tile_base, unsigned long long n_pos, const Geo& g,
                                           Emit&& emit) {
  const unsigned long long p0 = tile_base + i0;
  if (p0 >= n_pos) return;
  uint32_t ok = ~(sbad[i0 >> 5] >> (i0 & 31)) & 0xffffu;
  if (n_pos - p0 < (unsigned long long)kR) ok &= (1u << (uint32_t)(n_pos - p0)) - 1u;
  if (!ok) return;
  const unsigned long long mask = g.V >= 64 ? ~0ull : ((1ull << g.V) - 1);
  unsigned long long fwd, rc;
  {
    const uint32_t w = i0 >> 5;
    const int o = (int)(i0 & 31) * 2;
    unsigned long long x = sw[w] << o;
    if (o) x |= sw[w + 1] >> (64 - o);
    fwd = x >> (64 - g.V);
    rc = revcomp(fwd, g.K);
  }
  uint32_t next16;  // the 16 bases that follow the first window, first one in the top two bits
  {
    const uint32_t st = i0 + (uint32_t)g.K;
    const uint32_t w = st >> 5;
    const int o = (int)(st & 31) * 2;
    unsigned long long x = sw[w] << o;
    if (o) x |= sw[w + 1] >> (64 - o);
    next16 = (uint32_t)(x >> 32);
  }
  const int top = g.V - 2;
  const bool filter = g.bucket_lo != 0 || (g.bucket_hi >> (g.V - g.key_bits)) == 0;
#pragma unroll
  for (int j = 0; j < kR; j++) {
    if (ok & (1u << j)) {
      const unsigned long long v = (g.canonical && rc < fwd) ? rc : fwd;
      bool keep = true;
      if (filter) {
        const unsigned long long b = v >> g.key_bits;
        keep = b >= g.bucket_lo && b < g.bucket_hi;
      }
      if (keep) emit(j, v);
    }
    const unsigned long long c = (next16 >> (30 - 2 * j)) & 3u;
    fwd = ((fwd << 2) | c) & mask;
    rc = (rc >> 2) | ((3ull - c) << top);
  }
}

// A tile's packed words and bad-position mask travel through one register pair per thread
// (kThreads >= kTile / 32 + 2): loaded one tile ahead, stored to shared memory when the tile's turn
// comes, so the global-load latency hides behind the previous tile's work.
struct TileRegs { unsigned long long w; uint32_t b; };
static_assert(kThreads >= kTile / 32 + 2, "one staged word per thread");

__device__ __forceinline__ TileRegs load_tile(const Job& jb, unsigned long long tile_base) {
  // the words array holds ceil(n_pos / 32) + 1 entries; positions past n_pos are never used
  TileRegs t;
  const unsigned long long w0 = tile_base >> 5;
  const unsigned long long n_words = ((jb.n_pos + 31) >> 5) + 1;
  const unsigned long long n_badw = (jb.n_pos + 31) >> 5;
  const int i = threadIdx.x;
  t.w = (i < kTile / 32 + 2 && tile_base < jb.n_pos && w0 + i < n_words) ? jb.words[w0 + i] : 0ull;
  t.b = (i < kTile / 32 && tile_base < jb.n_pos && w0 + i < n_badw) ? jb.bad[w0 + i] : 0xffffffffu;
  return t;
}

__device__ __forceinline__ void store_tile(const TileRegs& t, unsigned long long* sw, uint32_t* sbad) {
  const int i = threadIdx.x;
  if (i < kTile / 32 + 2) sw[i] = t.w;
  if (i < kTile / 32) sbad[i] = t.b;
}

// ---- level 1, pass A: bin sizes per CTA ----------------------------------------------------------
__global__ void __launch_bounds__(kThreads) count_kernel(const Job* __restrict__ jobs, Geo g) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const Job jb = jobs[blockIdx.y];
  const int bins = 1 << g.B1;
  unsigned long long* sw = (unsigned long long*)smem_raw;   // kTile / 32 + 2
  uint32_t* sbad = (uint32_t*)(sw + kTile / 32 + 2);        // kTile / 32
  uint32_t* hist = sbad + kTile / 32;                       // bins
  for (int i = threadIdx.x; i < bins; i += blockDim.x) hist[i] = 0;
  const unsigned long long n_tiles = (jb.n_pos + kTile - 1) / kTile;
  TileRegs nxt = load_tile(jb, (unsigned long long)blockIdx.x * kTile);
  for (unsigned long long t = blockIdx.x; t < n_tiles; t += gridDim.x) {
    __syncthreads();
    store_tile(nxt, sw, sbad);
    nxt = load_tile(jb, (t + gridDim.x) * kTile);
    __syncthreads();
    tile_kmers(sw, sbad, threadIdx.x * kR, t * kTile, jb.n_pos, g,
               [&](int, unsigned long long v) { atomicAdd(&hist[(uint32_t)(v >> g.R1)], 1u); });
  }
  __syncthreads();
  uint32_t* mine = jb.slice + (size_t)blockIdx.x * bins;
  for (int i = threadIdx.x; i < bins; i += blockDim.x) mine[i] = hist[i];
}

// one CTA per job: bin totals over the CTAs, exclusive scan of the bins, the start of every
// (CTA, bin) slice, n_occ, the largest bin
__global__ void __launch_bounds__(1024) scan_kernel(const Job* __restrict__ jobs, Geo g, int n_ctas) {
  __shared__ uint32_t a[(1 << kMaxB1) + 1];
  __shared__ uint32_t wsum[33];
  __shared__ uint32_t mx;
  const Job jb = jobs[blockIdx.x];
  const int bins = 1 << g.B1;
  if (threadIdx.x == 0) mx = 0;
  uint32_t m = 0;
  for (int i = threadIdx.x; i < bins; i += blockDim.x) {
    uint32_t run = 0;
    uint32_t* col = jb.slice + i;        // coalesced over the threads
    int c = 0;
    for (; c + 8 <= n_ctas; c += 8) {    // eight independent loads in flight, then the running offsets
      uint32_t v[8];
#pragma unroll
      for (int u = 0; u < 8; u++) v[u] = col[(size_t)(c + u) * bins];
#pragma unroll
      for (int u = 0; u < 8; u++) { col[(size_t)(c + u) * bins] = run; run += v[u]; }
    }
    for (; c < n_ctas; c++) {
      const uint32_t v = col[(size_t)c * bins];
      col[(size_t)c * bins] = run;       // offset of CTA c's slice inside the bin
      run += v;
    }
    a[i] = run;
    m = max(m, run);
  }
  __syncthreads();
  atomicMax(&mx, m);
  const uint32_t total = block_excl_scan(a, bins, wsum);
  for (int i = threadIdx.x; i <= bins; i += blockDim.x) jb.base[i] = a[i];
  if (threadIdx.x == 0) { jb.meta[0] = total; jb.meta[1] = mx; }
}

// ---- level 1, pass B: group every tile by bin in shared memory, write contiguous runs -------------
template <typename TmpT>
__global__ void __launch_bounds__(kThreads, 2) partition_kernel(const Job* __restrict__ jobs, Geo g) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const Job jb = jobs[blockIdx.y];
  const int bins = 1 << g.B1;
  unsigned long long* sorted = (unsigned long long*)smem_raw;           // kTile
  unsigned long long* sw = sorted + kTile;                              // kTile / 32 + 2
  uint32_t* lcur = (uint32_t*)(sw + kTile / 32 + 2);                    // bins + 1: counts -> local starts -> local ends
  uint32_t* gadj = lcur + bins + 1;                                     // bins: global index of a key = gadj[bin] + local index
  uint32_t* gcur = gadj + bins;                                         // bins: the CTA's running cursor of every bin
  uint32_t* sbad = gcur + bins;                                         // kTile / 32
  uint32_t* wsum = sbad + kTile / 32;                                   // 33
  {
    const uint32_t* mine = jb.slice + (size_t)blockIdx.x * bins;
    for (int i = threadIdx.x; i < bins; i += blockDim.x) gcur[i] = jb.base[i] + mine[i];
  }
  TmpT* tmp = (TmpT*)jb.tmp;
  const unsigned long long rmask = g.R1 >= 64 ? ~0ull : ((1ull << g.R1) - 1);
  const unsigned long long n_tiles = (jb.n_pos + kTile - 1) / kTile;
  constexpr int R = kTile / kThreads;
  TileRegs nxt = load_tile(jb, (unsigned long long)blockIdx.x * kTile);
  for (unsigned long long t = blockIdx.x; t < n_tiles; t += gridDim.x) {
    const unsigned long long tile_base = t * kTile;
    __syncthreads();   // the previous tile's write-out has finished reading sorted / gadj
    for (int i = threadIdx.x; i <= bins; i += blockDim.x) lcur[i] = 0;
    store_tile(nxt, sw, sbad);
    nxt = load_tile(jb, (t + gridDim.x) * kTile);
    __syncthreads();
    unsigned long long v[R];
    uint32_t valid = 0;
    tile_kmers(sw, sbad, threadIdx.x * kR, tile_base, jb.n_pos, g, [&](int j, unsigned long long x) {
      v[j] = x;
      valid |= 1u << j;
      atomicAdd(&lcur[(uint32_t)(x >> g.R1)], 1u);
    });
    __syncthreads();
    const uint32_t n_valid = block_excl_scan(lcur, bins, wsum);
    for (int i = threadIdx.x; i < bins; i += blockDim.x) gadj[i] = gcur[i] - lcur[i];
    __syncthreads();
#pragma unroll
    for (int r = 0; r < R; r++)
      if (valid & (1u << r)) sorted[atomicAdd(&lcur[(uint32_t)(v[r] >> g.R1)], 1u)] = v[r];
    __syncthreads();
    // lcur[bin] is now the local END of the bin: the cursor moves on by the bin's count
    for (int i = threadIdx.x; i < bins; i += blockDim.x) gcur[i] = gadj[i] + lcur[i];
    for (uint32_t i = threadIdx.x; i < n_valid; i += blockDim.x) {
      const unsigned long long x = sorted[i];
      tmp[gadj[(uint32_t)(x >> g.R1)] + i] = (TmpT)(x & rmask);
    }
  }
}

// ---- level 2: one CTA per partition ----------------------------------------------------------------
template <typename KeyT>
__device__ void heap_sort(KeyT* a, uint32_t L) {
  auto sift = [&](uint32_t start, uint32_t end) {
    uint32_t root = start;
    while (2 * root + 1 < end) {
      uint32_t c = 2 * root + 1;
      if (c + 1 < end && a[c] < a[c + 1]) c++;
      if (a[root] < a[c]) { const KeyT t = a[root]; a[root] = a[c]; a[c] = t; root = c; } else return;
    }
  };
  for (uint32_t s = L / 2; s-- > 0;) sift(s, L);
  for (uint32_t e = L; e-- > 1;) { const KeyT t = a[0]; a[0] = a[e]; a[e] = t; sift(0, e); }
}

template <typename KeyT, typename TmpT>
__global__ void __launch_bounds__(kSortThreads) sort_kernel(const Job* __restrict__ jobs, Geo g, uint32_t cap) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const Job jb = jobs[blockIdx.y];
  const uint32_t bin = blockIdx.x;
  const int bins = 1 << g.B1;
  const uint32_t base = jb.base[bin];
  const uint32_t P = jb.base[bin + 1] - base;
  const int nsub = 1 << g.FB2;
  // shared memory (offsets, so that every access stays in the shared address space):
  //   cur[nsub + 1] | wsum[33] | n_long | long_list[2 kLongCap] | B[cap] keys by slot | SB[cap] sub-bin of the slot
  //   | dropw, dropp (dedup only)
  uint32_t* cur = (uint32_t*)smem_raw;                 // counts -> exclusive starts -> run ends
  uint32_t* wsum = cur + nsub + 1;
  uint32_t* n_long = wsum + 33;
  uint32_t* long_list = n_long + 1;
  const uint32_t off_b = (uint32_t)(((nsub + 1 + 33 + 1 + 2 * kLongCap) * 4 + 15) & ~15);
  KeyT* B = (KeyT*)(smem_raw + off_b);
  const uint32_t off_sb = (off_b + (cap + 4) * (uint32_t)sizeof(KeyT) + 15u) & ~15u;
  uint16_t* SB = (uint16_t*)(smem_raw + off_sb);
  const uint32_t off_drop = (off_sb + cap * 2u + 15u) & ~15u;
  const TmpT* tmp = (const TmpT*)jb.tmp + base;
  const int sub_shift = g.R1 - g.FB2;
  if (P > cap) {  // cannot happen: the host sizes cap from the largest partition
    if (threadIdx.x == 0) atomicOr(&jb.meta[2], 2u);
    return;
  }
  for (int i = threadIdx.x; i <= nsub; i += blockDim.x) cur[i] = 0;
  if (threadIdx.x == 0) *n_long = 0;
  __syncthreads();
  // pass 1: sub-bin sizes; four independent loads in flight per thread
  for (uint32_t i0 = threadIdx.x; i0 < P; i0 += 4 * kSortThreads) {
    TmpT r[4];
#pragma unroll
    for (int u = 0; u < 4; u++) { const uint32_t i = i0 + u * kSortThreads; r[u] = i < P ? tmp[i] : (TmpT)0; }
#pragma unroll
    for (int u = 0; u < 4; u++)
      if (i0 + u * kSortThreads < P) atomicAdd(&cur[(uint32_t)((unsigned long long)r[u] >> sub_shift)], 1u);
  }
  __syncthreads();
  block_excl_scan(cur, nsub, wsum);
  // the set's finest offsets for the fine buckets of this partition (every bin writes its own,
  // empty or not); the last bin closes the array
  {
    const int S = g.FB2 - g.FB;
    const int nfine = 1 << g.FB;
    uint32_t* fine = jb.fine + ((size_t)bin << g.FB);
    for (int f = threadIdx.x; f < nfine; f += blockDim.x) fine[f] = base + cur[f << S];
    if (bin == (uint32_t)bins - 1 && threadIdx.x == 0) fine[nfine] = base + P;
  }
  __syncthreads();
  // pass 2: every key takes a slot of its sub-bin (second read: L1 / L2 hits)
  const unsigned long long kmask = g.key_bits >= 64 ? ~0ull : ((1ull << g.key_bits) - 1);
  const unsigned long long hi_part = g.R1 >= 64 ? 0ull : ((unsigned long long)bin << g.R1);
  for (uint32_t i0 = threadIdx.x; i0 < P; i0 += 4 * kSortThreads) {
    TmpT r[4];
#pragma unroll
    for (int u = 0; u < 4; u++) { const uint32_t i = i0 + u * kSortThreads; r[u] = i < P ? tmp[i] : (TmpT)0; }
#pragma unroll
    for (int u = 0; u < 4; u++)
      if (i0 + u * kSortThreads < P) {
        const unsigned long long x = (unsigned long long)r[u];
        const uint32_t sb = (uint32_t)(x >> sub_shift);
        const uint32_t slot = atomicAdd(&cur[sb], 1u);
        B[slot] = (KeyT)((hi_part | x) & kmask);
        SB[slot] = (uint16_t)sb;
      }
  }
  __syncthreads();
  // cur[s] is now the END of sub-bin s; its start is cur[s - 1]. A sub-bin holds about one key:
  // every key finds its rank among the keys of its sub-bin (equal keys: slot order) and goes
  // straight to its place in the result. Runs too long for that are left to the whole CTA.
  KeyT* out = (KeyT*)jb.keys + base;
  int dup = 0;
  for (uint32_t slot = threadIdx.x; slot < P; slot += blockDim.x) {
    const KeyT key = B[slot];
    const uint32_t sb = SB[slot];
    const uint32_t a = sb ? cur[sb - 1] : 0u, e = cur[sb];
    if (e - a > (uint32_t)kLongRun) {
      if (slot == a) {
        const uint32_t q = atomicAdd(n_long, 1u);
        if (q < (uint32_t)kLongCap) { long_list[2 * q] = a; long_list[2 * q + 1] = e - a; }
      }
      continue;
    }
    // the first four slots of the run without a loop (a sub-bin holds about one key; B has four
    // spare entries past cap); equal keys are rare and settled by slot order afterwards
    const uint32_t L = e - a;
    uint32_t lt = 0, eq = 0;
#pragma unroll
    for (uint32_t t = 0; t < 4; t++) {
      const KeyT x = B[a + t];
      const bool in = t < L;
      lt += (in && x < key) ? 1u : 0u;
      eq += (in && x == key) ? 1u : 0u;
    }
    for (uint32_t j = a + 4; j < e; j++) {
      const KeyT x = B[j];
      lt += x < key ? 1u : 0u;
      eq += x == key ? 1u : 0u;
    }
    if (eq > 1) {
      dup = 1;
      for (uint32_t j = a; j < slot; j++) lt += B[j] == key ? 1u : 0u;
    }
    out[a + lt] = key;
  }
  __syncthreads();
  // long runs (repeated k-mers, low-complexity sequence): the whole CTA sorts each in place with
  // a bitonic network whose merges all run ascending, so the virtual +inf padding never moves;
  // past kLongCap of them, one thread heap-sorts each of the rest
  const uint32_t n_l = *n_long;
  if (n_l) {
    const uint32_t nl = min(n_l, (uint32_t)kLongCap);
    for (uint32_t q = 0; q < nl; q++) {
      const uint32_t a = long_list[2 * q], L = long_list[2 * q + 1];
      uint32_t Pw = 1;
      while (Pw < L) Pw <<= 1;
      for (uint32_t k = 2; k <= Pw; k <<= 1) {
        for (uint32_t j = k >> 1; j > 0; j >>= 1) {
          for (uint32_t i = threadIdx.x; i < Pw; i += blockDim.x) {
            const uint32_t l = (j == (k >> 1)) ? (i ^ (k - 1)) : (i ^ j);
            if (l > i && l < L) {
              const KeyT x = B[a + i], y = B[a + l];
              if (x > y) { B[a + i] = y; B[a + l] = x; }
            }
          }
          __syncthreads();
        }
      }
      for (uint32_t i = threadIdx.x; i < L; i += blockDim.x) {
        out[a + i] = B[a + i];
        if (i > 0) dup |= B[a + i] == B[a + i - 1];
      }
    }
    if (n_l > (uint32_t)kLongCap) {
      // the listed ones are done; find the unlisted long runs again (their first slot) and sort them serially
      for (int sb = threadIdx.x; sb < nsub; sb += blockDim.x) {
        const uint32_t a = sb ? cur[sb - 1] : 0u, L = cur[sb] - a;
        if (L <= (uint32_t)kLongRun) continue;
        bool listed = false;
        for (uint32_t q = 0; q < nl && !listed; q++) listed = long_list[2 * q] == a;
        if (listed) continue;
        heap_sort(B + a, L);
        for (uint32_t i = 0; i < L; i++) {
          out[a + i] = B[a + i];
          if (i > 0) dup |= B[a + i] == B[a + i - 1];
        }
      }
    }
  }
  const int any_dup = __syncthreads_or(dup);
  if (any_dup && threadIdx.x == 0) atomicOr(&jb.meta[2], 1u);
  if (!(any_dup && g.dedup)) return;
  // rare: this partition holds a key more than once and the caller wants a set. Every key finds
  // its place again; copies (an equal key in an earlier slot) are marked at their place, the
  // survivors are written packed from the partition's base and the fine offsets re-derived; the
  // gaps between partitions are closed by shift_kernel once every partition's count is known.
  uint32_t* dropw = (uint32_t*)(smem_raw + off_drop);   // cap / 32 + 2 mask words over the sorted places
  uint32_t* dropp = dropw + (cap >> 5) + 2;             // their exclusive popcount prefix
  const int nwords = (int)(P >> 5) + 1;
  for (int w = threadIdx.x; w < nwords; w += blockDim.x) dropw[w] = 0;
  __syncthreads();
  auto place_of = [&](uint32_t slot, bool* copy) {   // long runs are sorted in place: slot order = key order
    const KeyT key = B[slot];
    const uint32_t sb = SB[slot];
    const uint32_t a = sb ? cur[sb - 1] : 0u, e = cur[sb];
    if (e - a > (uint32_t)kLongRun) {
      *copy = slot > a && B[slot - 1] == key;
      return slot;
    }
    uint32_t rank = 0;
    bool c = false;
    for (uint32_t j = a; j < e; j++) {
      const KeyT x = B[j];
      rank += (x < key) || (x == key && j < slot);
      c |= (x == key) && (j < slot);
    }
    *copy = c;
    return a + rank;
  };
  for (uint32_t slot = threadIdx.x; slot < P; slot += blockDim.x) {
    bool copy;
    const uint32_t q = place_of(slot, &copy);
    if (copy) atomicOr(&dropw[q >> 5], 1u << (q & 31));
  }
  __syncthreads();
  for (int w = threadIdx.x; w < nwords; w += blockDim.x) dropp[w] = __popc(dropw[w]);
  __syncthreads();
  const uint32_t removed = block_excl_scan(dropp, nwords, wsum);
  auto new_index = [&](uint32_t i) { return i - dropp[i >> 5] - __popc(dropw[i >> 5] & ((1u << (i & 31)) - 1u)); };
  {
    const int S = g.FB2 - g.FB;
    const int nfine = 1 << g.FB;
    uint32_t* fine = jb.fine + ((size_t)bin << g.FB);
    for (int f = threadIdx.x; f < nfine; f += blockDim.x) {
      const int sb = f << S;
      const uint32_t start = sb ? cur[sb - 1] : 0u;
      fine[f] = base + (start < P ? new_index(start) : P - removed);
    }
    if (bin == (uint32_t)bins - 1 && threadIdx.x == 0) fine[nfine] = base + P - removed;
  }
  for (uint32_t slot = threadIdx.x; slot < P; slot += blockDim.x) {
    bool copy;
    const uint32_t q = place_of(slot, &copy);
    if (!copy) out[new_index(q)] = B[slot];
  }
  if (threadIdx.x == 0) {
    jb.removed[bin] = removed;
    atomicAdd(&jb.meta[3], removed);
  }
}

// closes the gaps the sort left where it dropped copies: partition `bin` of the packed result
// starts sum(removed[0..bin)) keys earlier. One CTA per partition; keys and the finest offsets
// are copied into the (smaller) final set.
struct ShiftJob {
  const void* old_keys; void* new_keys;
  const uint32_t* old_fine; uint32_t* new_fine;
  const uint32_t* base; const uint32_t* removed;
};
template <typename KeyT>
__global__ void __launch_bounds__(256) shift_kernel(const ShiftJob* __restrict__ jobs, int B1, int FB) {
  __shared__ uint32_t red[32];
  const ShiftJob jb = jobs[blockIdx.y];
  const uint32_t bin = blockIdx.x;
  const int bins = 1 << B1;
  uint32_t s = 0;
  for (uint32_t i = threadIdx.x; i < bin; i += blockDim.x) s += jb.removed[i];
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  uint32_t shift = 0;
  for (int w = 0; w < (int)(blockDim.x >> 5); w++) shift += red[w];
  const uint32_t base = jb.base[bin];
  const uint32_t P = jb.base[bin + 1] - base - jb.removed[bin];
  const KeyT* src = (const KeyT*)jb.old_keys + base;
  KeyT* dst = (KeyT*)jb.new_keys + (base - shift);
  for (uint32_t i = threadIdx.x; i < P; i += blockDim.x) dst[i] = src[i];
  const int nfine = 1 << FB;
  const size_t f0 = (size_t)bin << FB;
  for (int f = threadIdx.x; f < nfine; f += blockDim.x) jb.new_fine[f0 + f] = jb.old_fine[f0 + f] - shift;
  if (bin == (uint32_t)bins - 1 && threadIdx.x == 0) jb.new_fine[f0 + nfine] = jb.old_fine[f0 + nfine] - shift;
}

// coarser offset levels of m sets from their finest ones: lev[f][x] = lev[F][x << (F - f)]
struct LevJob { uint32_t* lev_base; };
__global__ void derive_levels_batch_kernel(const LevJob* __restrict__ jobs, int N, int max_level) {
  uint32_t* lev_base = jobs[blockIdx.y].lev_base;
  const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  uint64_t fine_start = 0;
  for (int gl = 0; gl < max_level; gl++) fine_start += ((uint64_t)1 << (N + gl)) + 1;
  uint64_t start = 0;
  for (int f = 0; f < max_level; f++) {
    const uint64_t cnt = ((uint64_t)1 << (N + f)) + 1;
    if (tid < start + cnt) {
      const uint64_t x = tid - start;
      lev_base[tid] = lev_base[fine_start + (x << (max_level - f))];
      return;
    }
    start += cnt;
  }
}

}  // namespace part
}  // namespace kmsc
