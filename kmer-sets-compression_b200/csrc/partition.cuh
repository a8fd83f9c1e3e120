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
constexpr int kLongRun = 32;      // sub-bin runs longer than this are sorted by the whole CTA
constexpr int kLongCap = 512;     // such runs per partition (beyond: the owning thread heap-sorts)

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

// k-mer starting at tile position i (global position p) out of the staged words
__device__ __forceinline__ bool tile_kmer(const unsigned long long* sw, const uint32_t* sbad, uint32_t i,
                                          unsigned long long p, unsigned long long n_pos, const Geo& g,
                                          unsigned long long* out) {
  if (p >= n_pos) return false;
  if ((sbad[i >> 5] >> (i & 31)) & 1u) return false;
  const uint32_t w = i >> 5;
  const int o = (int)(i & 31) * 2;
  unsigned long long x = sw[w] << o;
  if (o) x |= sw[w + 1] >> (64 - o);
  unsigned long long v = x >> (64 - g.V);
  if (g.canonical) {
    const unsigned long long rc = revcomp(v, g.K);
    v = rc < v ? rc : v;
  }
  const unsigned long long b = v >> g.key_bits;
  if (b < g.bucket_lo || b >= g.bucket_hi) return false;
  *out = v;
  return true;
}

__device__ __forceinline__ void stage_tile(const Job& jb, unsigned long long tile_base, unsigned long long* sw,
                                           uint32_t* sbad) {
  // the words array holds ceil(n_pos / 32) + 1 entries; positions past n_pos are never used
  const unsigned long long w0 = tile_base >> 5;
  const unsigned long long n_words = ((jb.n_pos + 31) >> 5) + 1;
  for (int i = threadIdx.x; i < kTile / 32 + 1; i += blockDim.x) sw[i] = (w0 + i < n_words) ? jb.words[w0 + i] : 0ull;
  const unsigned long long n_badw = (jb.n_pos + 31) >> 5;
  for (int i = threadIdx.x; i < kTile / 32; i += blockDim.x) sbad[i] = (w0 + i < n_badw) ? jb.bad[w0 + i] : 0xffffffffu;
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
  for (unsigned long long t = blockIdx.x; t < n_tiles; t += gridDim.x) {
    __syncthreads();
    stage_tile(jb, t * kTile, sw, sbad);
    __syncthreads();
#pragma unroll 4
    for (int r = 0; r < kTile / kThreads; r++) {
      const uint32_t i = r * kThreads + threadIdx.x;
      unsigned long long v;
      if (tile_kmer(sw, sbad, i, t * kTile + i, jb.n_pos, g, &v)) atomicAdd(&hist[(uint32_t)(v >> g.R1)], 1u);
    }
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
    for (int c = 0; c < n_ctas; c++) {   // coalesced over the threads; the loads of different c are independent
      uint32_t* p = jb.slice + (size_t)c * bins + i;
      const uint32_t v = *p;
      *p = run;                          // offset of CTA c's slice inside the bin
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
__global__ void __launch_bounds__(kThreads) partition_kernel(const Job* __restrict__ jobs, Geo g) {
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
  for (unsigned long long t = blockIdx.x; t < n_tiles; t += gridDim.x) {
    const unsigned long long tile_base = t * kTile;
    __syncthreads();   // the previous tile's write-out has finished reading sorted / gadj
    for (int i = threadIdx.x; i <= bins; i += blockDim.x) lcur[i] = 0;
    stage_tile(jb, tile_base, sw, sbad);
    __syncthreads();
    unsigned long long v[R];
    uint32_t valid = 0;
#pragma unroll
    for (int r = 0; r < R; r++) {
      const uint32_t i = r * kThreads + threadIdx.x;
      v[r] = 0;
      if (tile_kmer(sw, sbad, i, tile_base + i, jb.n_pos, g, &v[r])) {
        valid |= 1u << r;
        atomicAdd(&lcur[(uint32_t)(v[r] >> g.R1)], 1u);
      }
    }
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
__device__ __forceinline__ void insertion_sort(KeyT* a, uint32_t L, bool* dup) {
  for (uint32_t i = 1; i < L; i++) {
    const KeyT x = a[i];
    uint32_t j = i;
    while (j > 0 && a[j - 1] > x) { a[j] = a[j - 1]; j--; }
    a[j] = x;
    if (j > 0 && a[j - 1] == x) *dup = true;
  }
}

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
  uint32_t* cur = (uint32_t*)smem_raw;                 // nsub + 1: counts -> exclusive starts -> run ends
  uint32_t* wsum = cur + nsub + 1;                     // 33
  uint32_t* n_long = wsum + 33;                        // 1
  uint32_t* long_list = n_long + 1;                    // 2 * kLongCap
  KeyT* B = (KeyT*)(((uintptr_t)(long_list + 2 * kLongCap) + 15) & ~(uintptr_t)15);  // cap keys
  const TmpT* tmp = (const TmpT*)jb.tmp + base;
  const int sub_shift = g.R1 - g.FB2;
  if (P > cap) {  // cannot happen: the host sizes cap from the largest partition
    if (threadIdx.x == 0) atomicOr(&jb.meta[2], 2u);
    return;
  }
  for (int i = threadIdx.x; i <= nsub; i += blockDim.x) cur[i] = 0;
  if (threadIdx.x == 0) *n_long = 0;
  __syncthreads();
  // four independent loads in flight per thread
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
  const unsigned long long kmask = g.key_bits >= 64 ? ~0ull : ((1ull << g.key_bits) - 1);
  const unsigned long long hi_part = g.R1 >= 64 ? 0ull : ((unsigned long long)bin << g.R1);
  for (uint32_t i0 = threadIdx.x; i0 < P; i0 += 4 * kSortThreads) {  // second read: L1 / L2 hits
    TmpT r[4];
#pragma unroll
    for (int u = 0; u < 4; u++) { const uint32_t i = i0 + u * kSortThreads; r[u] = i < P ? tmp[i] : (TmpT)0; }
#pragma unroll
    for (int u = 0; u < 4; u++)
      if (i0 + u * kSortThreads < P) {
        const unsigned long long x = (unsigned long long)r[u];
        const uint32_t pos = atomicAdd(&cur[(uint32_t)(x >> sub_shift)], 1u);
        B[pos] = (KeyT)((hi_part | x) & kmask);
      }
  }
  __syncthreads();
  // cur[s] is now the END of sub-bin s; its start is cur[s - 1]
  bool dup = false;
  for (int s = threadIdx.x; s < nsub; s += blockDim.x) {
    const uint32_t a = s ? cur[s - 1] : 0u, L = cur[s] - a;
    if (L < 2) continue;
    if (L <= (uint32_t)kLongRun) {
      insertion_sort(B + a, L, &dup);
    } else {
      const uint32_t slot = atomicAdd(n_long, 1u);
      if (slot < (uint32_t)kLongCap) {
        long_list[2 * slot] = a;
        long_list[2 * slot + 1] = L;
      } else {
        heap_sort(B + a, L);
        for (uint32_t i = 1; i < L; i++) dup |= B[a + i] == B[a + i - 1];
      }
    }
  }
  __syncthreads();
  // long runs (repeated k-mers, low-complexity sequence): the whole CTA sorts each with a
  // bitonic network whose merges all run ascending, so the virtual +inf padding never moves
  const uint32_t nl = min(*n_long, (uint32_t)kLongCap);
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
    for (uint32_t i = threadIdx.x + 1; i < L; i += blockDim.x) dup |= B[a + i] == B[a + i - 1];
  }
  const int any_dup = __syncthreads_or(dup ? 1 : 0);
  if (any_dup && threadIdx.x == 0) atomicOr(&jb.meta[2], 1u);
  KeyT* out = (KeyT*)jb.keys + base;
  if (!(any_dup && g.dedup)) {
    for (uint32_t i = threadIdx.x; i < P; i += blockDim.x) out[i] = B[i];
    return;
  }
  // rare: this partition holds a key more than once and the caller wants a set. Mark the copies,
  // write the survivors packed from the partition's base and re-derive its fine offsets; the
  // gaps between partitions are closed by shift_kernel once every partition's count is known.
  uint32_t* dropw = (uint32_t*)(((uintptr_t)(B + cap) + 15) & ~(uintptr_t)15);  // cap / 32 + 1 mask words
  uint32_t* dropp = dropw + (cap >> 5) + 2;                                      // their exclusive popcount prefix
  const int nwords = (int)(P >> 5) + 1;
  for (int w = threadIdx.x; w < nwords; w += blockDim.x) dropw[w] = 0;
  __syncthreads();
  for (int sb = threadIdx.x; sb < nsub; sb += blockDim.x) {
    const uint32_t a = sb ? cur[sb - 1] : 0u, e = cur[sb];
    for (uint32_t i = a + 1; i < e; i++)
      if (B[i] == B[i - 1]) atomicOr(&dropw[i >> 5], 1u << (i & 31));
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
  for (uint32_t i = threadIdx.x; i < P; i += blockDim.x)
    if (!((dropw[i >> 5] >> (i & 31)) & 1u)) out[new_index(i)] = B[i];
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
