// pair_counts.cu -- P3: all-pairs intersection counts over bucketed sorted sets.
//
// Replaces GetEdgeWeight and the all-pairs loop of KmerSetSet's constructor
// (reference lib/core/kmer_set_set.h:158-219) and the per-iteration re-weighting
// (:385-425). The reference runs one two-pointer merge per (pair, bucket):
// (n-1) * sum_i |S_i| key visits. Here every key is read ONCE:
//
//   tile   = a contiguous range of fine buckets (k-mer prefix range) holding at
//            most ~L keys summed over all n sets
//   build  = every key of every set in the tile is inserted into a shared-memory
//            hash table; slot s keeps the key and an n-bit membership mask
//   gram   = masks of the D distinct keys are bit-transposed into per-set
//            columns and W[i][j] += popc(col_i & col_j) (register-tiled 4x4)
//
// W is exact for duplicate-free sets (every KmerSet; an SPSS spells each k-mer
// once). Sets that do hold duplicate keys (possible only through
// GetSampledKmerSet on a hand-made file) are routed to a plain merge kernel
// that counts min multiplicity exactly like the reference's loop (:165-180).
//
// Algorithmic bytes per launch of the main kernel:
//   B_w = sum_i sum_{b in B} len_i[b] * sizeof(KeyType)  (+ offsets + n*n*8).
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <vector>

#include "kmsc_common.cuh"

namespace kmsc {

struct SetDesc {
  const void* keys;
  const uint32_t* lev;  // fine offsets at the level chosen for this call
};

struct PlanParams {
  uint32_t NF;        // number of fine buckets = 2^(N+f)
  int f;              // fine level
  int n_sets;
  unsigned long long L;  // target keys per tile
  uint32_t span;      // forced tile boundary every `span` fine buckets
};

// ---------------------------------------------------------------------------
// planning: offsets transposed to [fine][set], totals, prefix, tile list
// ---------------------------------------------------------------------------

// block = 256 threads handles 32 fine buckets (33 boundary rows).
__global__ void plan_gather_kernel(const SetDesc* __restrict__ sets, PlanParams pp,
                                   const uint32_t* __restrict__ sel_bitmap,
                                   uint32_t* __restrict__ offsT, uint32_t* __restrict__ totals) {
  extern __shared__ uint32_t tile[];  // [33][n_sets + 1]
  const int n = pp.n_sets;
  const int stride = n + 1;
  const uint32_t x0 = blockIdx.x * 32u;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  if (sel_bitmap && x0 < pp.NF) {
    // nothing selected in [x0, x0+32] (row x0+32 included: it may close a tile)? write zeros only
    bool any = false;
    const uint32_t xl = min(x0 + 32u, pp.NF - 1);
    for (uint32_t b = (x0 ? x0 - 1 : 0) >> pp.f; b <= (xl >> pp.f); b++)
      any |= (sel_bitmap[b >> 5] >> (b & 31)) & 1u;
    if (!any) {
      if (threadIdx.x < 32 && x0 + threadIdx.x < pp.NF) totals[x0 + threadIdx.x] = 0u;
      return;
    }
  }
  for (int s = warp; s < n; s += nw) {
    const uint32_t* lev = sets[s].lev;
    const uint32_t x = x0 + lane;
    if (x <= pp.NF) tile[lane * stride + s] = lev[x];
    if (lane == 0 && x0 + 32 <= pp.NF) tile[32 * stride + s] = lev[x0 + 32];
  }
  __syncthreads();
  // rows out (coalesced over sets)
  const uint32_t rows = min(33u, pp.NF + 1 - x0);  // last block also writes row NF
  for (uint32_t r = warp; r < rows; r += nw) {
    if (r == 32 && x0 + 32 != pp.NF) continue;  // row 32 belongs to the next block unless it is the last row
    for (int s = lane; s < n; s += 32) offsT[(size_t)(x0 + r) * n + s] = tile[r * stride + s];
  }
  // totals
  for (uint32_t r = warp; r < 32; r += nw) {
    const uint32_t x = x0 + r;
    if (x >= pp.NF) break;
    uint32_t t = 0;
    for (int s = lane; s < n; s += 32) t += tile[(r + 1) * stride + s] - tile[r * stride + s];
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    if (lane == 0) {
      const uint32_t b = x >> pp.f;
      const bool sel = !sel_bitmap || ((sel_bitmap[b >> 5] >> (b & 31)) & 1u);
      totals[x] = sel ? t : 0u;
    }
  }
}

constexpr int kScanThreads = 1024;
constexpr int kScanPer = 4;
constexpr int kScanBlock = kScanThreads * kScanPer;

__global__ void plan_block_sums_kernel(const uint32_t* __restrict__ totals, uint32_t NF,
                                       unsigned long long* __restrict__ bsum) {
  __shared__ unsigned long long red[32];
  const uint32_t base = blockIdx.x * kScanBlock + threadIdx.x * kScanPer;
  unsigned long long s = 0;
#pragma unroll
  for (int i = 0; i < kScanPer; i++)
    if (base + i < NF) s += totals[base + i];
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    s = red[threadIdx.x];
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (threadIdx.x == 0) bsum[blockIdx.x] = s;
  }
}

// single block: exclusive scan of block sums (n_blocks <= a few thousand)
__global__ void plan_scan_bsums_kernel(unsigned long long* __restrict__ bsum, int n_blocks,
                                       unsigned long long* __restrict__ grand_total) {
  if (threadIdx.x == 0) {
    unsigned long long run = 0;
    for (int i = 0; i < n_blocks; i++) {
      const unsigned long long v = bsum[i];
      bsum[i] = run;
      run += v;
    }
    *grand_total = run;
  }
}

struct Tile {
  uint32_t x0, x1;  // fine bucket range [x0, x1)
};

__device__ __forceinline__ bool plan_sel(const uint32_t* sel_bitmap, uint32_t x, int f) {
  if (!sel_bitmap) return true;
  const uint32_t b = x >> f;
  return (sel_bitmap[b >> 5] >> (b & 31)) & 1u;
}

__global__ void plan_emit_kernel(const uint32_t* __restrict__ totals, PlanParams pp,
                                 const uint32_t* __restrict__ sel_bitmap,
                                 const unsigned long long* __restrict__ bsum,
                                 Tile* __restrict__ tiles, uint32_t* __restrict__ n_tiles,
                                 uint32_t max_tiles) {
  __shared__ unsigned long long wsum[32];
  const uint32_t base = blockIdx.x * kScanBlock + threadIdx.x * kScanPer;
  uint32_t v[kScanPer];
  unsigned long long s = 0;
#pragma unroll
  for (int i = 0; i < kScanPer; i++) {
    v[i] = (base + i < pp.NF) ? totals[base + i] : 0u;
    s += v[i];
  }
  // block exclusive scan of per-thread sums
  unsigned long long inc = s;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned long long t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) wsum[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    unsigned long long w = wsum[lane];
    unsigned long long winc = w;
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned long long t = __shfl_up_sync(0xffffffffu, winc, o);
      if (lane >= o) winc += t;
    }
    wsum[lane] = winc - w;
  }
  __syncthreads();
  unsigned long long pre = bsum[blockIdx.x] + wsum[warp] + (inc - s);  // exclusive prefix at base
#pragma unroll
  for (int i = 0; i < kScanPer; i++) {
    const uint32_t x = base + i;
    if (x < pp.NF && plan_sel(sel_bitmap, x, pp.f)) {
      const unsigned long long q = pre / pp.L;
      bool start = (x == 0) || (x % pp.span == 0) || !plan_sel(sel_bitmap, x - 1, pp.f);
      if (!start) {
        const unsigned long long pprev = pre - totals[x - 1];
        start = (pprev / pp.L) != q;
      }
      if (start) {
        // walk to the end of the tile (bounded by span)
        uint32_t xe = x + 1;
        unsigned long long p = pre + v[i];
        while (xe < pp.NF && (xe % pp.span) != 0 && plan_sel(sel_bitmap, xe, pp.f) && (p / pp.L) == q) {
          p += totals[xe];
          xe++;
        }
        if (p > pre) {  // skip empty tiles
          const uint32_t slot = atomicAdd(n_tiles, 1u);
          if (slot < max_tiles) tiles[slot] = Tile{x, xe};
        }
      }
    }
    pre += v[i];
  }
}

// ---------------------------------------------------------------------------
// main kernel
// ---------------------------------------------------------------------------

template <typename KeyT> struct TableKey { using type = uint32_t; };
template <> struct TableKey<unsigned long long> { using type = unsigned long long; };

// per mask-width configuration: table slots, max distinct keys per pass, threads,
// gram chunk in 32-key words. Sized so MW<=4 fits two CTAs per SM (228 KB).
template <int MW> struct PcCfg;
// UNR = keys in flight per lane, MINB = CTAs per SM the register budget is sized for.
#ifndef PC_MW2_SMALL
#define PC_MW2_SMALL 1
#endif
template <> struct PcCfg<1> { static constexpr int S = 8192, LOG2S = 13, D = 4096, T = 256, CW = 32, UNR = 8, MINB = 2; };
#if PC_MW2_SMALL
template <> struct PcCfg<2> { static constexpr int S = 4096, LOG2S = 12, D = 1536, T = 256, CW = 8, UNR = 4, MINB = 4; };
#else
template <> struct PcCfg<2> { static constexpr int S = 8192, LOG2S = 13, D = 3072, T = 256, CW = 16, UNR = 8, MINB = 2; };
#endif
template <> struct PcCfg<4> { static constexpr int S = 4096, LOG2S = 12, D = 2048, T = 256, CW = 16, UNR = 8, MINB = 1; };
template <> struct PcCfg<8> { static constexpr int S = 4096, LOG2S = 12, D = 2048, T = 512, CW = 16, UNR = 4, MINB = 1; };

constexpr int kSeg = 256;         // keys per build work item
constexpr int kStackMax = 48;

__device__ __forceinline__ uint32_t hash_slot(uint32_t k, int log2s) {
  return (k * 0x9E3779B1u) >> (32 - log2s);
}
__device__ __forceinline__ uint32_t hash_slot(unsigned long long k, int log2s) {
  return (uint32_t)((k * 0x9E3779B97F4A7C15ull) >> (64 - log2s));
}
__device__ __forceinline__ uint32_t hash_class(uint32_t k) {
  uint32_t h = k * 0x85EBCA6Bu;
  return h ^ (h >> 15);
}
__device__ __forceinline__ uint32_t hash_class(unsigned long long k) {
  unsigned long long h = k * 0xC2B2AE3D27D4EB4Full;
  return (uint32_t)(h >> 32) ^ (uint32_t)h;
}

// 32x32 bit-matrix transpose across a warp: in = row `lane`, out = column `lane`.
__device__ __forceinline__ uint32_t warp_transpose32(uint32_t x, int lane) {
#pragma unroll
  for (int k = 16; k >= 1; k >>= 1) {
    // mask of columns c with (c & k) == 0
    const uint32_t mlow = (k == 16) ? 0x0000FFFFu : (k == 8) ? 0x00FF00FFu : (k == 4) ? 0x0F0F0F0Fu
                        : (k == 2) ? 0x33333333u : 0x55555555u;
    const uint32_t y = __shfl_xor_sync(0xffffffffu, x, k);
    if (lane & k) x = (x & ~mlow) | ((y >> k) & mlow);
    else          x = (x & mlow) | ((y << k) & ~mlow);
  }
  return x;
}

// Shared-memory layout with compile-time offsets (so every access keeps the
// shared address space: LDS / STS / ATOMS, not generic LD / ST).
template <typename TK, int MW>
struct PcLayout {
  using C = PcCfg<MW>;
  static constexpr size_t a16(size_t x) { return (x + 15) & ~(size_t)15; }
  static constexpr int NPAD = 32 * MW;
  static constexpr int NB4 = NPAD / 4;
  static constexpr size_t o_keys = 0;
  static constexpr size_t o_mask = a16((size_t)(C::S + 1) * sizeof(TK));
  static constexpr size_t o_colT = a16(o_mask + (size_t)(C::S + 1) * MW * 4);
  static constexpr size_t o_order = o_colT + (size_t)C::CW * NPAD * 4;
  static constexpr size_t o_kp = a16(o_order + (size_t)(C::D + 1) * 2);
  static constexpr size_t o_sbeg = o_kp + (size_t)NPAD * 8;
  static constexpr size_t o_send = o_sbeg + (size_t)NPAD * 4;
  static constexpr size_t o_segpre = o_send + (size_t)NPAD * 4;
  static constexpr size_t o_misc = o_segpre + (size_t)NPAD * 4;
  static constexpr size_t o_stack = o_misc + 64 * 4;
  static constexpr size_t o_blk = o_stack + (size_t)kStackMax * 8;
  static constexpr size_t total = a16(o_blk + (size_t)(NB4 * (NB4 + 1) / 2) * 2);
};

// misc[] slots
constexpr int kMiscNdist = 0, kMiscOverflow = 1, kMiscTile = 2, kMiscItems = 3, kMiscSp = 4, kMiscSpecial = 5,
              kMiscWsum = 8;

#ifndef PC_SLOW
#define PC_SLOW 0      // 0 = divergent per-lane slow path, 1 = warp-convergent resolve_row
#endif
#ifndef PC_BUCKET2
#define PC_BUCKET2 0   // 1 = two-slot buckets: the first probe reads slots h, h+1 with one 64-bit load
#endif

// Slow path, per lane (divergent): the first probe at slot h returned `c` != key.
// An EMPTY first probe goes straight to the CAS; other keys probe linearly.
template <typename TK, int S, int DMAX>
__device__ __forceinline__ uint32_t find_slot_slow(TK* skeys, uint16_t* order, int* misc, TK key, uint32_t h, TK c) {
  const TK EMPTY = (TK)~(TK)0;
  if (sizeof(TK) == 4 && key == EMPTY) {
    // the one key that collides with the empty marker lives in the extra slot S
    if (atomicExch(&misc[kMiscSpecial], 1) == 0) {
      const int r = atomicAdd(&misc[kMiscNdist], 1);
      if (r < DMAX) order[r] = (uint16_t)S; else misc[kMiscOverflow] = 1;
    }
    return S;
  }
  for (;;) {
    if (c == EMPTY) {
      const TK old = atomicCAS(&skeys[h], EMPTY, key);
      if (old == EMPTY) {
        const int r = atomicAdd(&misc[kMiscNdist], 1);
        if (r < DMAX) order[r] = (uint16_t)h; else misc[kMiscOverflow] = 1;
        return h;
      }
      if (old == key) return h;
    }
    h = (h + 1) & (S - 1);
    c = skeys[h];
    if (c == key) return h;
  }
}

// Slow path of the table lookup, warp-convergent: called by ALL 32 lanes of a warp
// for one row of keys; lanes with `pend` set missed their first probe (empty slot,
// another key, or the key equals the empty marker). Every iteration advances all
// pending lanes by one probe; new keys get their rank in `order` through one
// warp-aggregated shared atomic. Returns the slot of `key`.
template <typename TK, int S, int DMAX>
__device__ __forceinline__ uint32_t resolve_row(TK* skeys, uint16_t* order, int* misc, TK key, uint32_t h,
                                               bool pend, int lane) {
  const TK EMPTY = (TK)~(TK)0;
  bool inserted = false;
  if (sizeof(TK) == 4) {
    // the one key that collides with the empty marker lives in the extra slot S
    const bool sp = pend && key == EMPTY;
    if (__any_sync(0xffffffffu, sp)) {
      if (sp) {
        h = S;
        pend = false;
        if (atomicExch(&misc[kMiscSpecial], 1) == 0) inserted = true;
      }
    }
  }
  while (__any_sync(0xffffffffu, pend)) {
    if (pend) {
      const TK c = skeys[h];
      if (c == key) {
        pend = false;
      } else if (c == EMPTY) {
        const TK old = atomicCAS(&skeys[h], EMPTY, key);
        if (old == EMPTY) { inserted = true; pend = false; }
        else if (old == key) pend = false;
        else h = (h + 1) & (S - 1);
      } else {
        h = (h + 1) & (S - 1);
      }
    }
  }
  const unsigned im = __ballot_sync(0xffffffffu, inserted);
  if (im) {
    const int leader = __ffs(im) - 1;
    int base = 0;
    if (lane == leader) base = atomicAdd(&misc[kMiscNdist], __popc(im));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (inserted) {
      const int r = base + __popc(im & ((1u << lane) - 1u));
      if (r < DMAX) order[r] = (uint16_t)h; else misc[kMiscOverflow] = 1;
    }
  }
  return h;
}

// Insert keys [beg, end) of set s (table key = top | key) and mark membership.
// U keys per lane are handled as a batch: U independent global loads, U independent
// first probes, then the (rare) slow paths, then U membership marks.
// OWN: the mask byte (slot, s / 8) is only ever touched by the warp that owns set
// group s / 8, and the keys of one set are distinct, so a plain byte
// read-modify-write is race free (ordered across sets by __syncwarp). Otherwise a
// shared-memory atomicOr on the mask word.
template <typename KeyT, typename TK, int MW, bool OWN, int U, bool FULL>
__device__ __forceinline__ void process_batch(TK* skeys, uint32_t* smask, uint16_t* order, int* misc,
                                              const KeyT* __restrict__ kp, uint32_t i0, uint32_t end, TK top,
                                              int s, int lane, uint32_t cmask, uint32_t cp) {
  using C = PcCfg<MW>;
  const TK EMPTY = (TK)~(TK)0;
  uint8_t* maskb = reinterpret_cast<uint8_t*>(smask);
  KeyT kk[U];
#pragma unroll
  for (int u = 0; u < U; u++) {
    const uint32_t i = i0 + u * 32 + lane;
    kk[u] = (FULL || i < end) ? __ldg(kp + i) : (KeyT)0;
  }
  TK key[U];
  uint32_t h[U];
  bool act[U];
#pragma unroll
  for (int u = 0; u < U; u++) {
    const uint32_t i = i0 + u * 32 + lane;
    key[u] = top | (TK)kk[u];
    act[u] = (FULL || i < end) && (!cmask || (hash_class(key[u]) & cmask) == cp);
    h[u] = hash_slot(key[u], C::LOG2S);
  }
#if PC_BUCKET2
  // two-slot buckets: slots (h & ~1, h | 1) are read with one 64-bit load
  TK c0[U], c1[U];
#pragma unroll
  for (int u = 0; u < U; u++) {
    h[u] &= ~1u;
    if (sizeof(TK) == 4) {
      const uint2 v = *reinterpret_cast<const uint2*>(&skeys[h[u]]);
      c0[u] = (TK)v.x; c1[u] = (TK)v.y;
    } else {
      c0[u] = skeys[h[u]]; c1[u] = skeys[h[u] + 1];
    }
  }
#pragma unroll
  for (int u = 0; u < U; u++) {
    const bool sp = sizeof(TK) == 4 && key[u] == EMPTY;
    if (!sp && c1[u] == key[u]) { h[u] += 1; }
    else if (sp || c0[u] != key[u]) {
      if (act[u]) {
        // first empty of the pair, else continue after the bucket
        TK c = c0[u];
        uint32_t hh = h[u];
        if (c0[u] != EMPTY) { hh += 1; c = c1[u]; }
        h[u] = find_slot_slow<TK, C::S, C::D>(skeys, order, misc, key[u], hh, c);
      }
    }
  }
#else
  TK cur[U];
#pragma unroll
  for (int u = 0; u < U; u++) cur[u] = skeys[h[u]];
#pragma unroll
  for (int u = 0; u < U; u++) {
    const bool pend = act[u] && (cur[u] != key[u] || (sizeof(TK) == 4 && key[u] == EMPTY));
#if PC_SLOW == 1
    if (__any_sync(0xffffffffu, pend))  // warp-uniform
      h[u] = resolve_row<TK, C::S, C::D>(skeys, order, misc, key[u], h[u], pend, lane);
#else
    if (pend) h[u] = find_slot_slow<TK, C::S, C::D>(skeys, order, misc, key[u], h[u], cur[u]);
#endif
  }
#endif
  if (OWN) {
    uint8_t mv[U];
#pragma unroll
    for (int u = 0; u < U; u++) mv[u] = maskb[h[u] * (MW * 4) + (s >> 3)];
#pragma unroll
    for (int u = 0; u < U; u++)
      if (act[u]) maskb[h[u] * (MW * 4) + (s >> 3)] = (uint8_t)(mv[u] | (1u << (s & 7)));
  } else {
#pragma unroll
    for (int u = 0; u < U; u++)
      if (act[u]) atomicOr(&smask[h[u] * MW + ((uint32_t)s >> 5)], 1u << (s & 31));
  }
}

template <typename KeyT, typename TK, int MW, bool OWN, int UNR>
__device__ __forceinline__ void process_run(TK* skeys, uint32_t* smask, uint16_t* order, int* misc,
                                            const KeyT* __restrict__ kp, uint32_t beg, uint32_t end, TK top,
                                            int s, int lane, uint32_t cmask, uint32_t cp) {
  uint32_t i0 = beg;
  // full batches: no bounds checks
  for (; i0 + 32 * UNR <= end; i0 += 32 * UNR) {
    process_batch<KeyT, TK, MW, OWN, UNR, true>(skeys, smask, order, misc, kp, i0, end, top, s, lane, cmask, cp);
    if (*(volatile int*)&misc[kMiscOverflow]) return;
  }
  // tail: rows of 32 keys in decreasing power-of-two batches
  if (UNR >= 8 && i0 + 32 * 4 <= end) {
    process_batch<KeyT, TK, MW, OWN, 4, true>(skeys, smask, order, misc, kp, i0, end, top, s, lane, cmask, cp);
    i0 += 32 * 4;
  }
  if (UNR >= 4 && i0 + 32 * 2 <= end) {
    process_batch<KeyT, TK, MW, OWN, 2, true>(skeys, smask, order, misc, kp, i0, end, top, s, lane, cmask, cp);
    i0 += 32 * 2;
  }
  if (i0 + 32 <= end) {
    process_batch<KeyT, TK, MW, OWN, 1, true>(skeys, smask, order, misc, kp, i0, end, top, s, lane, cmask, cp);
    i0 += 32;
  }
  if (i0 < end) process_batch<KeyT, TK, MW, OWN, 1, false>(skeys, smask, order, misc, kp, i0, end, top, s, lane, cmask, cp);
}

template <int MW>
struct PcSmem {
  static size_t bytes(int tk_size) {
    return tk_size == 8 ? PcLayout<unsigned long long, MW>::total : PcLayout<uint32_t, MW>::total;
  }
};

template <typename KeyT, int MW>
__global__ void __launch_bounds__(PcCfg<MW>::T, PcCfg<MW>::MINB)
pair_counts_kernel(const SetDesc* __restrict__ sets, int n_sets, const uint32_t* __restrict__ offsT,
                   const Tile* __restrict__ tiles, const uint32_t* __restrict__ n_tiles_p,
                   uint32_t* __restrict__ tile_counter, unsigned long long* __restrict__ W,
                   unsigned long long* __restrict__ stats, int key_bits, int fine_level) {
  using C = PcCfg<MW>;
  using TK = typename TableKey<KeyT>::type;
  using LY = PcLayout<TK, MW>;
  constexpr int S = C::S, DMAX = C::D, T = C::T, kChunkWords = C::CW;
  constexpr int NPAD = 32 * MW;
  constexpr int NW = T / 32;
  constexpr bool OWN = MW >= 2;   // byte-ownership membership marks (no shared atomics)
  constexpr int UNR = C::UNR;     // independent keys in flight per lane
  const TK EMPTY = (TK)~(TK)0;

  extern __shared__ __align__(16) unsigned char smem_raw[];
  TK* skeys = reinterpret_cast<TK*>(smem_raw + LY::o_keys);
  uint32_t* smask = reinterpret_cast<uint32_t*>(smem_raw + LY::o_mask);
  uint32_t* colT = reinterpret_cast<uint32_t*>(smem_raw + LY::o_colT);
  uint16_t* order = reinterpret_cast<uint16_t*>(smem_raw + LY::o_order);
  const void** skp = reinterpret_cast<const void**>(smem_raw + LY::o_kp);
  uint32_t* sbeg = reinterpret_cast<uint32_t*>(smem_raw + LY::o_sbeg);
  uint32_t* send = reinterpret_cast<uint32_t*>(smem_raw + LY::o_send);
  uint32_t* segpre = reinterpret_cast<uint32_t*>(smem_raw + LY::o_segpre);
  int* misc = reinterpret_cast<int*>(smem_raw + LY::o_misc);
  uint2* stack = reinterpret_cast<uint2*>(smem_raw + LY::o_stack);
  uint8_t* blk_tab = reinterpret_cast<uint8_t*>(smem_raw + LY::o_blk);  // pairs (bi, bj)

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int NB4 = (n_sets + 3) / 4;
  const int n_blocks = NB4 * (NB4 + 1) / 2;
  constexpr int NB4MAX = NPAD / 4;
  constexpr int NBLK = (NB4MAX * (NB4MAX + 1) / 2 + T - 1) / T;

  // one-time init
  for (int i = tid; i <= S; i += T) skeys[i] = EMPTY;
  for (int i = tid; i < (S + 1) * MW; i += T) smask[i] = 0;
  for (int i = tid; i < NPAD; i += T) skp[i] = i < n_sets ? sets[i].keys : nullptr;
  if (tid == 0) {
    int q = 0;
    for (int bi = 0; bi < NB4; bi++)
      for (int bj = bi; bj < NB4; bj++) { blk_tab[2 * q] = (uint8_t)bi; blk_tab[2 * q + 1] = (uint8_t)bj; q++; }
    misc[kMiscNdist] = 0; misc[kMiscOverflow] = 0; misc[kMiscSpecial] = 0;
  }
  __syncthreads();

  int my_bi[NBLK], my_bj[NBLK];
  uint32_t acc[NBLK][16];
#pragma unroll
  for (int t = 0; t < NBLK; t++) {
    const int q = tid + t * T;
    my_bi[t] = (q < n_blocks) ? blk_tab[2 * q] : -1;
    my_bj[t] = (q < n_blocks) ? blk_tab[2 * q + 1] : -1;
#pragma unroll
    for (int e = 0; e < 16; e++) acc[t][e] = 0;
  }
  unsigned long long st_keys = 0, st_dist = 0, st_over = 0;

  const uint32_t n_tiles = *n_tiles_p;
  for (;;) {
    if (tid == 0) misc[kMiscTile] = (int)atomicAdd(tile_counter, 1u);
    __syncthreads();
    const uint32_t t_id = (uint32_t)misc[kMiscTile];
    if (t_id >= n_tiles) break;
    const Tile tl = tiles[t_id];
    const uint32_t bucket0 = tl.x0 >> fine_level;
    const int nbk = (int)(((tl.x1 - 1) >> fine_level) - bucket0) + 1;
    // per-set key ranges of this tile
    for (int s = tid; s < NPAD; s += T) {
      uint32_t b = 0, e = 0;
      if (s < n_sets) {
        b = offsT[(size_t)tl.x0 * n_sets + s];
        e = offsT[(size_t)tl.x1 * n_sets + s];
      }
      sbeg[s] = b; send[s] = e;
    }
    __syncthreads();
    if (!OWN) {
      // build work items: segments of kSeg keys; exclusive prefix over sets (NPAD <= T)
      uint32_t nseg = 0;
      if (tid < NPAD) nseg = (send[tid] - sbeg[tid] + kSeg - 1) / kSeg;
      uint32_t inc = nseg;
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
      }
      if (tid < NPAD && lane == 31) misc[kMiscWsum + warp] = (int)inc;
      __syncthreads();
      uint32_t woff = 0;
      if (tid < NPAD)
        for (int w = 0; w < warp; w++) woff += (uint32_t)misc[kMiscWsum + w];
      if (tid < NPAD) segpre[tid] = woff + inc - nseg;
      if (tid == NPAD - 1) misc[kMiscItems] = (int)(woff + inc);
    }
    if (tid == 0) { misc[kMiscSp] = 1; stack[0] = make_uint2(0u, 1u); }  // class (p=0, P=1) = everything
    __syncthreads();
    const int n_items = misc[kMiscItems];

    // process the stack of key classes (normally exactly one entry)
    for (;;) {
      __syncthreads();
      if (misc[kMiscSp] == 0) break;
      const uint2 cls = stack[misc[kMiscSp] - 1];
      __syncthreads();
      if (tid == 0) misc[kMiscSp] -= 1;
      const uint32_t cp = cls.x, cmask = cls.y - 1u;

      // ---- build -------------------------------------------------------
#define PROCESS_RUN(S_, BEG_, END_, TOP_) \
  process_run<KeyT, TK, MW, OWN, UNR>(skeys, smask, order, misc, (const KeyT*)skp[S_], (BEG_), (END_), (TOP_), (S_), lane, cmask, cp)
      if (OWN) {
        // warp w owns set groups w, w + NW, ... (8 sets each)
        const int n_groups = (n_sets + 7) >> 3;
        for (int g = warp; g < n_groups; g += NW) {
          const int s_end = min(n_sets, g * 8 + 8);
          for (int s = g * 8; s < s_end; s++) {
            if (nbk == 1) {
              PROCESS_RUN(s, sbeg[s], send[s], (TK)0);
            } else {
              for (int b = 0; b < nbk; b++) {
                const size_t row0 = (size_t)(bucket0 + (uint32_t)b) << fine_level;
                const uint32_t beg = max(offsT[row0 * n_sets + s], sbeg[s]);
                const uint32_t end = min(offsT[(row0 + ((size_t)1 << fine_level)) * n_sets + s], send[s]);
                PROCESS_RUN(s, beg, end, (TK)b << key_bits);
              }
            }
            __syncwarp();
            if (*(volatile int*)&misc[kMiscOverflow]) break;
          }
          if (*(volatile int*)&misc[kMiscOverflow]) break;
        }
      } else if (nbk == 1) {
        // tile inside one bucket: the key alone identifies the k-mer; segments of kSeg keys
        for (int item = warp; item < n_items; item += NW) {
          int lo = 0, hi = NPAD - 1;  // set owning this item: largest s with segpre[s] <= item
          while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (segpre[mid] <= (uint32_t)item) lo = mid; else hi = mid - 1;
          }
          const int s = lo;
          const uint32_t start = sbeg[s] + ((uint32_t)item - segpre[s]) * kSeg;
          PROCESS_RUN(s, start, min(start + (uint32_t)kSeg, send[s]), (TK)0);
          if (*(volatile int*)&misc[kMiscOverflow]) break;
        }
      } else {
        // tile spans several buckets: one item per (bucket, set); the table key is
        // (bucket - first bucket of the tile) << key_bits | key
        const int n_bitems = nbk * n_sets;
        for (int item = warp; item < n_bitems; item += NW) {
          const int b = item / n_sets, s = item - b * n_sets;
          const size_t row0 = (size_t)(bucket0 + (uint32_t)b) << fine_level;
          const uint32_t beg = max(offsT[row0 * n_sets + s], sbeg[s]);
          const uint32_t end = min(offsT[(row0 + ((size_t)1 << fine_level)) * n_sets + s], send[s]);
          PROCESS_RUN(s, beg, end, (TK)b << key_bits);
          if (*(volatile int*)&misc[kMiscOverflow]) break;
        }
      }
#undef PROCESS_RUN
      __syncthreads();
      const int D = misc[kMiscNdist];
      if (misc[kMiscOverflow]) {
        // too many distinct keys for one pass: wipe the table, split the class in two
        __syncthreads();
        for (int i = tid; i <= S; i += T) skeys[i] = EMPTY;
        for (int i = tid; i < (S + 1) * MW; i += T) smask[i] = 0;
        if (tid == 0) {
          misc[kMiscNdist] = 0; misc[kMiscOverflow] = 0; misc[kMiscSpecial] = 0;
          const uint32_t P = cls.y;
          if (misc[kMiscSp] + 2 <= kStackMax && P < 0x40000000u) {
            stack[misc[kMiscSp]] = make_uint2(cp, P * 2);
            stack[misc[kMiscSp] + 1] = make_uint2(cp + P, P * 2);
            misc[kMiscSp] += 2;
          } else {
            misc[kMiscTile] = -2;  // cannot split further: report failure
          }
        }
        st_over++;
        __syncthreads();
        if (misc[kMiscTile] == -2) { if (tid == 0) atomicAdd(&stats[3], 1ull); break; }
        continue;
      }
      st_dist += (tid == 0) ? (unsigned long long)D : 0ull;

      // ---- gram, CW * 32 distinct keys per chunk ---------------------------
      for (int c0 = 0; c0 < D; c0 += kChunkWords * 32) {
        const int ng = min(kChunkWords, (D - c0 + 31) >> 5);
        for (int g = warp; g < ng; g += NW) {
          const int r = c0 + g * 32 + lane;
          uint32_t m[MW];
#pragma unroll
          for (int w = 0; w < MW; w++) m[w] = 0;
          if (r < D) {
            const uint32_t slot = order[r];
#pragma unroll
            for (int w = 0; w < MW; w++) { m[w] = smask[slot * MW + w]; smask[slot * MW + w] = 0; }
            skeys[slot] = EMPTY;
          }
#pragma unroll
          for (int w = 0; w < MW; w++) colT[g * NPAD + w * 32 + lane] = warp_transpose32(m[w], lane);
        }
        __syncthreads();
#pragma unroll
        for (int t = 0; t < NBLK; t++) {
          if (my_bi[t] < 0) continue;
          const uint4* ca = reinterpret_cast<const uint4*>(colT + my_bi[t] * 4);
          const uint4* cb = reinterpret_cast<const uint4*>(colT + my_bj[t] * 4);
          for (int g = 0; g < ng; g++) {
            const uint4 a = ca[g * (NPAD / 4)];
            const uint4 b = cb[g * (NPAD / 4)];
            const uint32_t av[4] = {a.x, a.y, a.z, a.w};
            const uint32_t bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 4; i++)
#pragma unroll
              for (int j = 0; j < 4; j++) acc[t][i * 4 + j] += __popc(av[i] & bv[j]);
          }
        }
        __syncthreads();
      }
      if (tid == 0) { misc[kMiscNdist] = 0; misc[kMiscSpecial] = 0; }
    }
    if (tid < NPAD) st_keys += send[tid] - sbeg[tid];
    __syncthreads();
  }

  // flush accumulators: W is n x n, symmetric
#pragma unroll
  for (int t = 0; t < NBLK; t++) {
    if (my_bi[t] < 0) continue;
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
      for (int j = 0; j < 4; j++) {
        const int si = my_bi[t] * 4 + i, sj = my_bj[t] * 4 + j;
        const uint32_t v = acc[t][i * 4 + j];
        if (si >= n_sets || sj >= n_sets || v == 0) continue;
        atomicAdd(&W[(size_t)si * n_sets + sj], (unsigned long long)v);
        // a diagonal block holds both (i,j) and (j,i) already
        if (my_bi[t] != my_bj[t]) atomicAdd(&W[(size_t)sj * n_sets + si], (unsigned long long)v);
      }
  }
  // stats: keys processed, distinct keys, overflow retries
  for (int o = 16; o > 0; o >>= 1) st_keys += __shfl_xor_sync(0xffffffffu, st_keys, o);
  if (lane == 0 && st_keys) atomicAdd(&stats[0], st_keys);
  if (tid == 0) {
    if (st_dist) atomicAdd(&stats[1], st_dist);
    if (st_over) atomicAdd(&stats[2], st_over);
  }
}

// ---------------------------------------------------------------------------
// exact merge fallback (sets with duplicate keys), reference loop :165-180
// ---------------------------------------------------------------------------
template <typename KeyT>
__global__ void pair_counts_merge_kernel(const SetDesc* __restrict__ sets, int n_sets, int n_buckets,
                                         const uint32_t* __restrict__ sel_bitmap,
                                         unsigned long long* __restrict__ W) {
  // one block per ordered pair index p over i <= j
  int p = blockIdx.x, i = 0;
  int row = n_sets;
  while (p >= row) { p -= row; i++; row--; }
  const int j = i + p;
  const KeyT* ka = (const KeyT*)sets[i].keys;
  const KeyT* kb = (const KeyT*)sets[j].keys;
  const uint32_t* oa = sets[i].lev;
  const uint32_t* ob = sets[j].lev;
  unsigned long long c = 0;
  for (int b = threadIdx.x; b < n_buckets; b += blockDim.x) {
    if (sel_bitmap && !((sel_bitmap[b >> 5] >> (b & 31)) & 1u)) continue;
    uint32_t x = oa[b], xe = oa[b + 1], y = ob[b], ye = ob[b + 1];
    if (i == j) { c += xe - x; continue; }
    while (x < xe && y < ye) {
      const KeyT a = ka[x], bb = kb[y];
      if (a < bb) x++;
      else if (a > bb) y++;
      else { c++; x++; y++; }
    }
  }
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if ((threadIdx.x & 31) == 0 && c) {
    atomicAdd(&W[(size_t)i * n_sets + j], c);
    if (i != j) atomicAdd(&W[(size_t)j * n_sets + i], c);
  }
}

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------

template <typename KeyT, int MW>
static int launch_main(kmsc_ctx* ctx, const SetDesc* d_sets, int n_sets, const uint32_t* d_offsT,
                       const Tile* d_tiles, const uint32_t* d_ntiles, uint32_t* d_counter,
                       unsigned long long* d_W, unsigned long long* d_stats, int key_bits, int fine_level,
                       uint32_t max_tiles) {
  using TK = typename TableKey<KeyT>::type;
  const size_t smem = PcSmem<MW>::bytes((int)sizeof(TK));
  auto kern = pair_counts_kernel<KeyT, MW>;
  KMSC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int occ = 0;
  KMSC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, PcCfg<MW>::T, smem));
  if (occ < 1) { set_error("pair_counts kernel does not fit on an SM (smem %zu)", smem); return KMSC_E_CUDA; }
  long long grid = (long long)ctx->sm_count * occ;
  if (grid > (long long)max_tiles) grid = max_tiles;
  if (grid < 1) grid = 1;
  kern<<<(unsigned)grid, PcCfg<MW>::T, smem, ctx->stream>>>(d_sets, n_sets, d_offsT, d_tiles, d_ntiles,
                                                            d_counter, d_W, d_stats, key_bits, fine_level);
  count_launch(ctx);
  KMSC_CUDA(cudaGetLastError());
  return KMSC_OK;
}

template <typename KeyT>
static int launch_main_mw(kmsc_ctx* ctx, int mw, const SetDesc* d_sets, int n_sets,
                          const uint32_t* d_offsT, const Tile* d_tiles, const uint32_t* d_ntiles,
                          uint32_t* d_counter, unsigned long long* d_W, unsigned long long* d_stats,
                          int key_bits, int fine_level, uint32_t max_tiles) {
  switch (mw) {
    case 1: return launch_main<KeyT, 1>(ctx, d_sets, n_sets, d_offsT, d_tiles, d_ntiles, d_counter, d_W, d_stats, key_bits, fine_level, max_tiles);
    case 2: return launch_main<KeyT, 2>(ctx, d_sets, n_sets, d_offsT, d_tiles, d_ntiles, d_counter, d_W, d_stats, key_bits, fine_level, max_tiles);
    case 4: return launch_main<KeyT, 4>(ctx, d_sets, n_sets, d_offsT, d_tiles, d_ntiles, d_counter, d_W, d_stats, key_bits, fine_level, max_tiles);
    default: return launch_main<KeyT, 8>(ctx, d_sets, n_sets, d_offsT, d_tiles, d_ntiles, d_counter, d_W, d_stats, key_bits, fine_level, max_tiles);
  }
}

static int dmax_for(int mw) {
  switch (mw) { case 1: return PcCfg<1>::D; case 2: return PcCfg<2>::D; case 4: return PcCfg<4>::D; default: return PcCfg<8>::D; }
}

// device-side small block layout (ctx->small):
//   [0,4) n_tiles  [4,8) tile_counter  [64,96) stats (keys, distinct, retries, failures)
//   [128,136) grand total  [256, ...) SetDesc[n]  then bucket bitmap
struct PcDev {
  uint32_t* n_tiles; uint32_t* counter; unsigned long long* stats; unsigned long long* grand;
  SetDesc* sets; uint32_t* bitmap;
};

// One pass: plan tiles over the buckets selected by h_bitmap (NULL = all) with
// tile target L, then run the main kernel accumulating into d_W.
static int run_phase(kmsc_ctx* ctx, const kmsc_set* const* sets, int n, int mw,
                     const uint32_t* h_bitmap, unsigned long long L, double keys_in_phase,
                     unsigned long long* d_W, unsigned long long host_stats[4]) {
  const kmsc_set* s0 = sets[0];
  const int nb = 1 << s0->N;
  int f = 0;
  {
    const double mean_bucket = keys_in_phase;  // mean keys per selected bucket, all sets
    while (f < s0->max_level && mean_bucket / (double)(1 << f) > (double)L / 4.0) f++;
  }
  const uint32_t NF = (uint32_t)nb << f;
  const size_t sz_desc = ((size_t)n * sizeof(SetDesc) + 15) & ~(size_t)15;
  const size_t sz_bitmap = (size_t)((nb + 31) / 32) * 4;
  KMSC_TRY(ctx->small.reserve(256 + sz_desc + sz_bitmap + 64));
  unsigned char* sm = (unsigned char*)ctx->small.p;
  PcDev d;
  d.n_tiles = (uint32_t*)sm; d.counter = (uint32_t*)(sm + 4);
  d.stats = (unsigned long long*)(sm + 64); d.grand = (unsigned long long*)(sm + 128);
  d.sets = (SetDesc*)(sm + 256);
  d.bitmap = h_bitmap ? (uint32_t*)(sm + 256 + sz_desc) : nullptr;

  const uint32_t max_tiles = NF;
  const int n_scan_blocks = (int)((NF + kScanBlock - 1) / kScanBlock);
  size_t off = 0;
  const size_t o_offsT = off; off += ((size_t)(NF + 1) * n * 4 + 255) & ~(size_t)255;
  const size_t o_totals = off; off += ((size_t)NF * 4 + 255) & ~(size_t)255;
  const size_t o_bsum = off; off += ((size_t)(n_scan_blocks + 1) * 8 + 255) & ~(size_t)255;
  const size_t o_tiles = off; off += ((size_t)max_tiles * sizeof(Tile) + 255) & ~(size_t)255;
  KMSC_TRY(ctx->plan.reserve(off));
  unsigned char* pl = (unsigned char*)ctx->plan.p;
  uint32_t* d_offsT = (uint32_t*)(pl + o_offsT);
  uint32_t* d_totals = (uint32_t*)(pl + o_totals);
  unsigned long long* d_bsum = (unsigned long long*)(pl + o_bsum);
  Tile* d_tiles = (Tile*)(pl + o_tiles);

  void* pin = nullptr;
  KMSC_TRY(ctx_pinned(ctx, sz_desc + sz_bitmap + 64, &pin));
  SetDesc* h_sets = (SetDesc*)pin;
  for (int i = 0; i < n; i++) { h_sets[i].keys = sets[i]->keys; h_sets[i].lev = sets[i]->lev[f]; }
  KMSC_CUDA(cudaMemsetAsync(sm, 0, 256, ctx->stream));
  KMSC_CUDA(cudaMemcpyAsync(d.sets, h_sets, sz_desc, cudaMemcpyHostToDevice, ctx->stream));
  if (h_bitmap) {
    uint32_t* pb = (uint32_t*)((unsigned char*)pin + sz_desc);
    memcpy(pb, h_bitmap, sz_bitmap);
    KMSC_CUDA(cudaMemcpyAsync(d.bitmap, pb, sz_bitmap, cudaMemcpyHostToDevice, ctx->stream));
  }
  for (int i = 0; i < 3; i++)
    if (!ctx->pc_ev[i]) KMSC_CUDA(cudaEventCreate(&ctx->pc_ev[i]));
  KMSC_CUDA(cudaEventRecord(ctx->pc_ev[0], ctx->stream));
  PlanParams pp;
  pp.NF = NF; pp.f = f; pp.n_sets = n; pp.L = L;
  {
    // a tile may span several buckets only while (bucket - first bucket, key) fits
    // the table key: uint32 for 2/4-byte keys, uint64 for 8-byte keys
    const int tk_bits = s0->key_bytes == 8 ? 64 : 32;
    const int spare = tk_bits - s0->key_bits;
    const uint32_t bucket_span = spare >= 8 ? 256u : (1u << spare);
    const uint64_t span = (uint64_t)bucket_span << f;
    pp.span = span > 256 ? 256u : (uint32_t)span;
  }
  {
    const unsigned blocks = (NF + 32) / 32;  // covers rows 0..NF
    const size_t smem = (size_t)33 * (n + 1) * 4;
    plan_gather_kernel<<<blocks, 256, smem, ctx->stream>>>(d.sets, pp, d.bitmap, d_offsT, d_totals);
    plan_block_sums_kernel<<<n_scan_blocks, kScanThreads, 0, ctx->stream>>>(d_totals, NF, d_bsum);
    plan_scan_bsums_kernel<<<1, 32, 0, ctx->stream>>>(d_bsum, n_scan_blocks, d.grand);
    plan_emit_kernel<<<n_scan_blocks, kScanThreads, 0, ctx->stream>>>(d_totals, pp, d.bitmap, d_bsum, d_tiles,
                                                                     d.n_tiles, max_tiles);
    count_launch(ctx, 4);
    KMSC_CUDA(cudaGetLastError());
  }
  KMSC_CUDA(cudaEventRecord(ctx->pc_ev[1], ctx->stream));
  int rc;
  switch (s0->key_bytes) {
    case 2: rc = launch_main_mw<uint16_t>(ctx, mw, d.sets, n, d_offsT, d_tiles, d.n_tiles, d.counter, d_W, d.stats, s0->key_bits, f, max_tiles); break;
    case 4: rc = launch_main_mw<uint32_t>(ctx, mw, d.sets, n, d_offsT, d_tiles, d.n_tiles, d.counter, d_W, d.stats, s0->key_bits, f, max_tiles); break;
    default: rc = launch_main_mw<unsigned long long>(ctx, mw, d.sets, n, d_offsT, d_tiles, d.n_tiles, d.counter, d_W, d.stats, s0->key_bits, f, max_tiles); break;
  }
  if (rc != KMSC_OK) return rc;
  KMSC_CUDA(cudaEventRecord(ctx->pc_ev[2], ctx->stream));
  // stats come back through pinned memory; the sync also protects the staging buffer
  unsigned long long* h_stats = (unsigned long long*)((unsigned char*)pin + sz_desc + sz_bitmap);
  KMSC_CUDA(cudaMemcpyAsync(h_stats, d.stats, 32, cudaMemcpyDeviceToHost, ctx->stream));
  KMSC_CUDA(cudaStreamSynchronize(ctx->stream));
  for (int i = 0; i < 4; i++) host_stats[i] = h_stats[i];
  {
    float ms_plan = 0, ms_main = 0;
    KMSC_CUDA(cudaEventElapsedTime(&ms_plan, ctx->pc_ev[0], ctx->pc_ev[1]));
    KMSC_CUDA(cudaEventElapsedTime(&ms_main, ctx->pc_ev[1], ctx->pc_ev[2]));
    ctx->pc_plan_ms += ms_plan;
    ctx->pc_main_ms += ms_main;
    ctx->pc_main_launches += 1;
    // algorithmic bytes of this launch: every key once + the two offset rows per tile are
    // folded into (NF+1)*n*4 offsets + the n*n*8 result
    ctx->pc_algo_bytes += (double)h_stats[0] * s0->key_bytes + (double)(nb + 1) * n * 4.0;
  }
  if (host_stats[3]) { set_error("pair_counts: a tile could not be split further (%llu failures)", host_stats[3]); return KMSC_E_STATE; }
  return KMSC_OK;
}

// Core: d_W (n*n u64, device) is zeroed and filled. Synchronous at return.
int pair_counts_run(kmsc_ctx* ctx, const kmsc_set* const* sets, int32_t n,
                    const int32_t* bucket_ids, int32_t n_ids, unsigned long long* d_W) {
  if (!ctx || !sets || n < 1 || !d_W) { set_error("bad argument"); return KMSC_E_INVALID; }
  if (n > 256) { set_error("n_sets=%d > 256 not supported by this build", n); return KMSC_E_INVALID; }
  const kmsc_set* s0 = sets[0];
  if (!s0) { set_error("sets[0] is NULL"); return KMSC_E_INVALID; }
  int64_t total_keys = 0;
  bool any_dups = false;
  for (int i = 0; i < n; i++) {
    if (!sets[i]) { set_error("sets[%d] is NULL", i); return KMSC_E_INVALID; }
    if (sets[i]->K != s0->K || sets[i]->N != s0->N || sets[i]->key_bytes != s0->key_bytes) {
      set_error("sets[%d] has a different (K,N,KeyType)", i);
      return KMSC_E_INVALID;
    }
    total_keys += sets[i]->n_keys;
    if (sets[i]->has_dups < 0) KMSC_TRY(set_check_dups(ctx, const_cast<kmsc_set*>(sets[i])));
    any_dups |= sets[i]->has_dups == 1;
  }
  const int nb = 1 << s0->N;
  KMSC_CUDA(cudaSetDevice(ctx->device));
  ctx->pc_main_ms = ctx->pc_plan_ms = 0;
  ctx->pc_main_launches = 0;
  ctx->pc_algo_bytes = (double)n * n * 8.0;
  ctx->pc_last_stats[0] = ctx->pc_last_stats[1] = ctx->pc_last_stats[2] = 0;
  KMSC_CUDA(cudaMemsetAsync(d_W, 0, (size_t)n * n * 8, ctx->stream));

  // selection bitmap (host)
  const size_t words = (size_t)(nb + 31) / 32;
  std::vector<uint32_t> sel;
  int n_sel = nb;
  if (bucket_ids) {
    sel.assign(words, 0u);
    n_sel = 0;
    for (int i = 0; i < n_ids; i++) {
      const int b = bucket_ids[i];
      if (b < 0 || b >= nb) { set_error("bucket id %d out of range", b); return KMSC_E_INVALID; }
      if (!((sel[b >> 5] >> (b & 31)) & 1u)) n_sel++;
      sel[b >> 5] |= 1u << (b & 31);
    }
  }
  if (n_sel == 0 || total_keys == 0) { KMSC_CUDA(cudaStreamSynchronize(ctx->stream)); return KMSC_OK; }

  if (any_dups) {
    // exact multiset semantics (reference loop :165-180); bucket-level offsets
    const size_t sz_desc = ((size_t)n * sizeof(SetDesc) + 15) & ~(size_t)15;
    const size_t sz_bitmap = words * 4;
    KMSC_TRY(ctx->small.reserve(256 + sz_desc + sz_bitmap + 64));
    unsigned char* sm = (unsigned char*)ctx->small.p;
    SetDesc* d_sets = (SetDesc*)(sm + 256);
    uint32_t* d_bitmap = bucket_ids ? (uint32_t*)(sm + 256 + sz_desc) : nullptr;
    void* pin = nullptr;
    KMSC_TRY(ctx_pinned(ctx, sz_desc + sz_bitmap + 64, &pin));
    SetDesc* h_sets = (SetDesc*)pin;
    for (int i = 0; i < n; i++) { h_sets[i].keys = sets[i]->keys; h_sets[i].lev = sets[i]->lev[0]; }
    KMSC_CUDA(cudaMemcpyAsync(d_sets, h_sets, sz_desc, cudaMemcpyHostToDevice, ctx->stream));
    if (bucket_ids) {
      memcpy((unsigned char*)pin + sz_desc, sel.data(), sz_bitmap);
      KMSC_CUDA(cudaMemcpyAsync(d_bitmap, (unsigned char*)pin + sz_desc, sz_bitmap, cudaMemcpyHostToDevice, ctx->stream));
    }
    const int pairs = n * (n + 1) / 2;
    switch (s0->key_bytes) {
      case 2: pair_counts_merge_kernel<uint16_t><<<pairs, 256, 0, ctx->stream>>>(d_sets, n, nb, d_bitmap, d_W); break;
      case 4: pair_counts_merge_kernel<uint32_t><<<pairs, 256, 0, ctx->stream>>>(d_sets, n, nb, d_bitmap, d_W); break;
      default: pair_counts_merge_kernel<unsigned long long><<<pairs, 256, 0, ctx->stream>>>(d_sets, n, nb, d_bitmap, d_W); break;
    }
    count_launch(ctx);
    KMSC_CUDA(cudaGetLastError());
    KMSC_CUDA(cudaStreamSynchronize(ctx->stream));
    return KMSC_OK;
  }

  const int mw = n <= 32 ? 1 : n <= 64 ? 2 : n <= 128 ? 4 : 8;
  const int dmax = dmax_for(mw);
  const unsigned long long L_cons = (unsigned long long)(0.75 * dmax);
  const unsigned long long L_max = 1ull << 20;
  const double mean_bucket = (double)total_keys / (double)nb;  // all sets, per bucket
  unsigned long long st[4] = {0, 0, 0, 0};

  // redundancy rho = keys per distinct key inside a tile decides the tile size.
  // Known from the previous call on the same shape, else measured on a 1/64 slice.
  double rho = (ctx->pc_rho_n == n && ctx->pc_rho_k == s0->K) ? ctx->pc_rho : 0.0;
  std::vector<uint32_t> rest;
  const uint32_t* phase2 = bucket_ids ? sel.data() : nullptr;
  if (rho <= 0.0 && n_sel >= 2048 && (double)n_sel * mean_bucket > 64.0 * 4.0 * (double)L_cons) {
    std::vector<uint32_t> probe(words, 0u);
    rest.assign(words, 0u);
    int rank = 0;
    for (int b = 0; b < nb; b++) {
      const bool on = !bucket_ids || ((sel[b >> 5] >> (b & 31)) & 1u);
      if (!on) continue;
      if ((rank++ & 63) == 21) probe[b >> 5] |= 1u << (b & 31);
      else rest[b >> 5] |= 1u << (b & 31);
    }
    KMSC_TRY(run_phase(ctx, sets, n, mw, probe.data(), L_cons, mean_bucket, d_W, st));
    if (st[1] > 0) rho = (double)st[0] / (double)st[1];
    ctx->pc_last_stats[0] += st[0]; ctx->pc_last_stats[1] += st[1]; ctx->pc_last_stats[2] += st[2];
    phase2 = rest.data();
  }
  unsigned long long L = L_cons;
  if (rho > 1.0) {
    const unsigned long long La = (unsigned long long)(0.55 * rho * dmax);
    if (La > L) L = La;
  }
  if (L > L_max) L = L_max;
  KMSC_TRY(run_phase(ctx, sets, n, mw, phase2, L, mean_bucket, d_W, st));
  if (st[1] > 0) {
    ctx->pc_rho = (double)st[0] / (double)st[1];
    ctx->pc_rho_n = n;
    ctx->pc_rho_k = s0->K;
  }
  ctx->pc_last_stats[0] += st[0]; ctx->pc_last_stats[1] += st[1]; ctx->pc_last_stats[2] += st[2];
  ctx->pc_last_L = L;
  return KMSC_OK;
}

}  // namespace kmsc

using namespace kmsc;

extern "C" {

int kmsc_pair_counts_device(kmsc_ctx* ctx, const kmsc_set* const* sets, int32_t n,
                            const int32_t* bucket_ids, int32_t n_ids, int64_t* d_out) {
  return pair_counts_run(ctx, sets, n, bucket_ids, n_ids, (unsigned long long*)d_out);
}

int kmsc_pair_counts(kmsc_ctx* ctx, const kmsc_set* const* sets, int32_t n,
                     const int32_t* bucket_ids, int32_t n_ids, int64_t* out, int64_t* key_visits) {
  if (!ctx || !out || n < 1) { set_error("bad argument"); return KMSC_E_INVALID; }
  KMSC_TRY(ctx->work.reserve((size_t)n * n * 8));
  unsigned long long* d_W = (unsigned long long*)ctx->work.p;
  KMSC_TRY(pair_counts_run(ctx, sets, n, bucket_ids, n_ids, d_W));
  KMSC_CUDA(cudaMemcpyAsync(out, d_W, (size_t)n * n * 8, cudaMemcpyDeviceToHost, ctx->stream));
  KMSC_CUDA(cudaStreamSynchronize(ctx->stream));
  if (key_visits) {
    // sum_{i<j} (len_i + len_j) = (n - 1) * sum_i len_i over the selected buckets
    int64_t s = 0;
    for (int i = 0; i < n; i++) s += out[(size_t)i * n + i];
    *key_visits = (int64_t)(n - 1) * s;
  }
  return KMSC_OK;
}

int kmsc_pair_counts_stats(kmsc_ctx* ctx, double* out8) {
  if (!ctx || !out8) { set_error("NULL argument"); return KMSC_E_INVALID; }
  out8[0] = ctx->pc_main_ms; out8[1] = ctx->pc_plan_ms;
  out8[2] = (double)ctx->pc_last_stats[0]; out8[3] = (double)ctx->pc_last_stats[1];
  out8[4] = (double)ctx->pc_last_stats[2]; out8[5] = (double)ctx->pc_last_L;
  out8[6] = (double)ctx->pc_main_launches; out8[7] = ctx->pc_algo_bytes;
  return KMSC_OK;
}

int kmsc_pair_counts_rows(kmsc_ctx* ctx, const kmsc_set* const* sets, int32_t n,
                          const int32_t* rows, int32_t n_rows, const int32_t* bucket_ids,
                          int32_t n_ids, int64_t* out) {
  if (!ctx || !out || !rows || n < 1 || n_rows < 0) { set_error("bad argument"); return KMSC_E_INVALID; }
  for (int r = 0; r < n_rows; r++)
    if (rows[r] < 0 || rows[r] >= n) { set_error("row %d out of range", rows[r]); return KMSC_E_INVALID; }
  KMSC_TRY(ctx->work.reserve((size_t)n * n * 8));
  unsigned long long* d_W = (unsigned long long*)ctx->work.p;
  KMSC_TRY(pair_counts_run(ctx, sets, n, bucket_ids, n_ids, d_W));
  for (int r = 0; r < n_rows; r++)
    KMSC_CUDA(cudaMemcpyAsync(out + (size_t)r * n, d_W + (size_t)rows[r] * n, (size_t)n * 8,
                              cudaMemcpyDeviceToHost, ctx->stream));
  KMSC_CUDA(cudaStreamSynchronize(ctx->stream));
  return KMSC_OK;
}

}  // extern "C"
