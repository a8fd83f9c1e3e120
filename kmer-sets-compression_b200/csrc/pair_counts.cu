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
#include <cstdlib>
#include <cstring>
#include <vector>

#include "kmsc_common.cuh"
#include "umma.cuh"

namespace kmsc {

struct SetDesc {
  const void* keys;
  const uint32_t* lev;         // fine offsets at the level chosen for this call
  const uint32_t* lev_finest;  // fine offsets at the set's finest level (merge build: segment bounds)
};

struct PlanParams {
  uint32_t NF;        // number of fine buckets = 2^(N+f)
  int f;              // fine level
  int n_sets;
  unsigned long long L;  // target keys per tile
  uint32_t span;      // forced tile boundary every `span` fine buckets
  uint32_t row_lo, row_hi;  // fine buckets outside [row_lo, row_hi) hold no key of any set (a rank's prefix shard)
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
  if (x0 + 32u < pp.row_lo || x0 > pp.row_hi) {
    // outside the sets' bucket range: no keys, and no tile ever names these rows (plan_emit clamps)
    if (threadIdx.x < 32 && x0 + threadIdx.x < pp.NF) totals[x0 + threadIdx.x] = 0u;
    return;
  }
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
          if (slot < max_tiles) tiles[slot] = Tile{max(x, pp.row_lo), min(xe, pp.row_hi)};  // the rows cut off are empty
        }
      }
    }
    pre += v[i];
  }
}

// ---------------------------------------------------------------------------
// main kernel: shared-memory hash build + tensor-core Gram (tcgen05 kind::i8)
// ---------------------------------------------------------------------------
//
// One persistent CTA per (SM x occupancy) pulls tiles from an atomic counter. Per tile:
//   build    every key of every set in the tile is looked up / inserted in an open-addressing
//            table in shared memory: slot = key + NS-bit membership mask. Warp w owns the
//            sets [w * spw, (w + 1) * spw) and the mask byte w of every slot, so marking
//            membership is a plain byte read-modify-write (no shared atomics).
//   compact  the occupied slots are listed (ballot scan of the table).
//   gram     32 distinct keys at a time, the masks are expanded to 0/1 bytes in the MN-major
//            operand layout of umma.cuh (lane = key: no transpose) and W += X X^T is issued as
//            tcgen05.mma kind::i8 with int32 accumulators that stay in tensor memory for the
//            CTA's lifetime. The expansion of chunk c + 1 overlaps the MMAs of chunk c (two
//            staging buffers guarded by mbarriers); the MMAs of a tile's last chunks overlap
//            the next tile's build.
// At the end the accumulators are read back (tcgen05.ld) and added to W with 64-bit RED.

template <typename KeyT> struct TableKey { using type = uint32_t; };
template <> struct TableKey<unsigned long long> { using type = unsigned long long; };

// NS = padded set count = rows of the Gram (64 / 128 / 256); TKB = bytes of a table key.
//   S slots, D = max distinct keys per pass, T threads (one warp per 8 sets),
//   CK = K-steps (32 keys) per staging buffer, UNR = key rows in flight per warp.
template <int NS, int TKB> struct PcCfg;
#ifndef PC_D64
#define PC_D64 2560
#endif
#ifndef PC_UNR64
#define PC_UNR64 4
#endif
#ifndef PC_LFRAC
#define PC_LFRAC 0.55
#endif
#ifndef PC_LOG2S64
#define PC_LOG2S64 12
#endif
#ifndef PC_MINB64
#define PC_MINB64 3
#endif
#ifndef PC_CK64
#define PC_CK64 4
#endif
template <int TKB> struct PcCfg<64, TKB>  { static constexpr int S = 1 << PC_LOG2S64, LOG2S = PC_LOG2S64, D = PC_D64, T = 256,  CK = PC_CK64, UNR = PC_UNR64, MINB = PC_MINB64; };
template <int TKB> struct PcCfg<128, TKB> { static constexpr int S = 2048, LOG2S = 11, D = 1280, T = 512,  CK = 4, UNR = 4, MINB = 2; };
#ifndef PC_UNR256
#define PC_UNR256 2
#endif
#ifndef PC_D256
#define PC_D256 2560
#endif
template <> struct PcCfg<256, 4>          { static constexpr int S = 4096, LOG2S = 12, D = PC_D256, T = 1024, CK = 4, UNR = PC_UNR256, MINB = 1; };
template <> struct PcCfg<256, 8>          { static constexpr int S = 2048, LOG2S = 11, D = 1280, T = 1024, CK = 4, UNR = 2, MINB = 1; };

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

// Shared-memory layout with compile-time offsets (so every access keeps the
// shared address space: LDS / STS / ATOMS, not generic LD / ST).
template <typename TK, int NS>
struct PcLayout {
  using C = PcCfg<NS, (int)sizeof(TK)>;
  static constexpr int MW = NS / 32;
  static constexpr size_t al(size_t x, size_t a) { return (x + a - 1) & ~(a - 1); }
  static constexpr size_t kstep_bytes = (size_t)NS * 32;             // one K-step of the operand
  static constexpr size_t buf_bytes = kstep_bytes * C::CK;           // one staging buffer
  static constexpr size_t o_stage = 0;                               // 2 buffers, 128-byte aligned
  static constexpr size_t o_keys = al(o_stage + 2 * buf_bytes, 16);
  static constexpr size_t o_mask = al(o_keys + (size_t)(C::S + 1) * sizeof(TK), 16);
  static constexpr size_t o_list = al(o_mask + (size_t)(C::S + 1) * MW * 4, 16);
  static constexpr size_t o_kp = al(o_list + (size_t)(C::D + 2) * 2, 16);
  static constexpr size_t o_sbeg = o_kp + (size_t)NS * 8;
  static constexpr size_t o_send = o_sbeg + (size_t)NS * 4 * 2;   // two tiles: current, next
  static constexpr size_t o_misc = o_send + (size_t)NS * 4 * 2;
  static constexpr size_t o_stack = o_misc + 16 * 4;
  static constexpr size_t o_bar = al(o_stack + (size_t)kStackMax * 8, 8);
  static constexpr size_t total = al(o_bar + 2 * 8, 16);
};

// misc[] slots
constexpr int kMiscNdist = 0, kMiscOverflow = 1, kMiscTile = 2, kMiscSp = 3, kMiscSpecial = 4, kMiscTmem = 5,
              kMiscAnyMma = 6, kMiscNext = 7, kMiscIns = 8, kMiscSeg = 9;

// The table is probed two slots at a time: slots (h, h + 1) with h even are one 8 / 16-byte
// load. Slow path, per lane (divergent, rare): neither slot of the home pair held the key
// or could be claimed. Walks the following pairs; an empty slot is claimed with CAS. A probe
// sequence longer than the table means the table is full: flag overflow (the pass is
// repeated on a split class).
template <typename TK, int S>
__device__ __forceinline__ uint32_t find_slot_slow(TK* skeys, int* misc, TK key, uint32_t h, uint32_t& n_new) {
  const TK EMPTY = (TK)~(TK)0;
  for (int n = 0; n <= S; n++) {
    const TK c = skeys[h];
    if (c == key) return h;
    if (c == EMPTY) {
      const TK old = atomicCAS(&skeys[h], EMPTY, key);
      if (old == EMPTY) n_new += 1;
      if (old == EMPTY || old == key) return h;
    }
    h = (h + 1) & (S - 1);
  }
  misc[kMiscOverflow] = 1;
  return S;  // harmless: the pass is discarded
}

template <typename TK> struct SlotPair;
template <> struct SlotPair<uint32_t> {
  static __device__ __forceinline__ void load(const uint32_t* p, uint32_t& a, uint32_t& b) {
    const uint2 v = *reinterpret_cast<const uint2*>(p);
    a = v.x; b = v.y;
  }
};
template <> struct SlotPair<unsigned long long> {
  static __device__ __forceinline__ void load(const unsigned long long* p, unsigned long long& a, unsigned long long& b) {
    const ulonglong2 v = *reinterpret_cast<const ulonglong2*>(p);
    a = v.x; b = v.y;
  }
};

// One batch of the build: U rows of 32 keys of ONE set (rows past `end` are masked off).
struct PcBatch {
  const void* kp;   // keys of the set
  uint32_t i0, end; // key indices [i0, min(i0 + 32 U, end))
  uint32_t bit;     // membership bit inside the mask byte this warp owns
};

template <typename KeyT, int U>
__device__ __forceinline__ void batch_load(const PcBatch& b, int lane, KeyT (&kk)[U]) {
  const KeyT* kp = (const KeyT*)b.kp;
#pragma unroll
  for (int u = 0; u < U; u++) {
    const uint32_t i = b.i0 + u * 32 + lane;
    kk[u] = i < b.end ? __ldg(kp + i) : (KeyT)0;
  }
}

// Insert the keys of a batch (table key = top | key) and set the batch's membership bit in
// the mask byte at byte offset `mbase + 4 * slot` (masks are stored as word planes:
// word w of slot s at smask[w * (S + 1) + s], so a warp's byte accesses spread over all 32
// banks). U independent pair probes, claims of empty slots, the rare slow paths, U marks.
// The keys of one set are distinct, so the U marks of a batch never hit the same byte.
// cmask != 0: only keys of hash class (cmask, cp) take part (after a table overflow).
template <typename KeyT, typename TK, int NS, int U>
__device__ __forceinline__ void batch_process(TK* skeys, uint8_t* maskb, int* misc, const PcBatch& b, const KeyT (&kk)[U],
                                              TK top, uint32_t mbase, int lane, uint32_t cmask, uint32_t cp,
                                              uint32_t& w_new) {
  using C = PcCfg<NS, (int)sizeof(TK)>;
  const TK EMPTY = (TK)~(TK)0;
  TK key[U], c0[U], c1[U];
  uint32_t h[U];
  bool act[U];
#pragma unroll
  for (int u = 0; u < U; u++) {
    key[u] = top | (TK)kk[u];
    act[u] = b.i0 + u * 32 + lane < b.end;
    h[u] = hash_slot(key[u], C::LOG2S) & ~1u;
  }
  if (cmask) {
#pragma unroll
    for (int u = 0; u < U; u++) act[u] = act[u] && (hash_class(key[u]) & cmask) == cp;
  }
#pragma unroll
  for (int u = 0; u < U; u++) SlotPair<TK>::load(&skeys[h[u]], c0[u], c1[u]);
  bool any_sp = false;
  uint32_t n_new = 0;  // slots this lane claimed
#pragma unroll
  for (int u = 0; u < U; u++) {
    const bool hit1 = c1[u] == key[u];
    bool pend = act[u] && c0[u] != key[u] && !hit1;
    h[u] += hit1 ? 1u : 0u;
    if (sizeof(TK) == 4) {
      // the one key that collides with the empty marker lives in the extra slot S
      const bool sp = key[u] == EMPTY;
      h[u] = sp ? (uint32_t)C::S : h[u];
      pend = pend && !sp;
      any_sp |= sp && act[u];
    }
    if (pend) {
      // new key (or a collision): claim the first empty slot of the pair
      uint32_t hh = h[u];
      bool done = false;
      if (c0[u] == EMPTY) {
        const TK old = atomicCAS(&skeys[hh], EMPTY, key[u]);
        done = (old == EMPTY) || (old == key[u]);
        n_new += old == EMPTY;
      }
      if (!done && c1[u] == EMPTY) {
        hh += 1;
        const TK old = atomicCAS(&skeys[hh], EMPTY, key[u]);
        done = (old == EMPTY) || (old == key[u]);
        n_new += old == EMPTY;
      }
      if (!done) hh = find_slot_slow<TK, C::S>(skeys, misc, key[u], (h[u] + 2) & (C::S - 1), n_new);
      h[u] = hh;
    }
  }
  if (sizeof(TK) == 4 && any_sp) misc[kMiscSpecial] = 1;
  // early overflow check: the distinct keys of this pass must fit D slots (a nearly full
  // table makes every probe sequence long well before it is completely full). Claims are
  // summed per warp and published in lumps of >= 64.
  w_new += __reduce_add_sync(0xffffffffu, n_new);
  if (w_new >= 64) {
    if (lane == 0 && atomicAdd(&misc[kMiscIns], (int)w_new) + (int)w_new > C::D) misc[kMiscOverflow] = 1;
    w_new = 0;
  }
  uint8_t mv[U];
#pragma unroll
  for (int u = 0; u < U; u++) mv[u] = maskb[mbase + h[u] * 4];
#pragma unroll
  for (int u = 0; u < U; u++)
    if (act[u]) maskb[mbase + h[u] * 4] = (uint8_t)(mv[u] | b.bit);
}

template <int NS>
struct PcSmem {
  static size_t bytes(int tk_size) {
    return tk_size == 8 ? PcLayout<unsigned long long, NS>::total : PcLayout<uint32_t, NS>::total;
  }
};

// Row / column p of the Gram <-> set index. Warp w owns sets [w * spw, (w + 1) * spw) and the
// mask byte w, so set s sits at bit position 8 * (s / spw) + s % spw (spw <= 8; = s when spw = 8).
__device__ __forceinline__ int set_of_pos(int p, int spw, int n_sets) {
  const int r = p & 7;
  const int s = (p >> 3) * spw + r;
  return (r < spw && s < n_sets) ? s : -1;
}

template <typename KeyT, int NS>
__global__ void __launch_bounds__(PcCfg<NS, (int)sizeof(typename TableKey<KeyT>::type)>::T,
                                  PcCfg<NS, (int)sizeof(typename TableKey<KeyT>::type)>::MINB)
pair_counts_kernel(const SetDesc* __restrict__ sets, int n_sets, int spw, const uint32_t* __restrict__ offsT,
                   const Tile* __restrict__ tiles, const uint32_t* __restrict__ n_tiles_p,
                   uint32_t* __restrict__ tile_counter, unsigned long long* __restrict__ W,
                   unsigned long long* __restrict__ stats, int key_bits, int fine_level) {
  using TK = typename TableKey<KeyT>::type;
  using C = PcCfg<NS, (int)sizeof(TK)>;
  using LY = PcLayout<TK, NS>;
  constexpr int S = C::S, DMAX = C::D, T = C::T, CK = C::CK, UNR = C::UNR;
  constexpr int MW = NS / 32;     // mask words per slot
  constexpr int NW = T / 32;      // warps = mask bytes per slot
  static_assert(NW == NS / 8, "one warp per mask byte");
  constexpr int MMA_M = NS == 64 ? 64 : 128;
  constexpr int MMA_N = NS;       // 64 / 128 / 256
  constexpr int TMEM_COLS = NS == 256 ? 512 : NS;
  const TK EMPTY = (TK)~(TK)0;

  extern __shared__ __align__(128) unsigned char smem_raw[];
  unsigned char* stage = smem_raw + LY::o_stage;
  TK* skeys = reinterpret_cast<TK*>(smem_raw + LY::o_keys);
  uint32_t* smask = reinterpret_cast<uint32_t*>(smem_raw + LY::o_mask);
  uint8_t* maskb = reinterpret_cast<uint8_t*>(smask);
  uint16_t* list = reinterpret_cast<uint16_t*>(smem_raw + LY::o_list);
  const void** skp = reinterpret_cast<const void**>(smem_raw + LY::o_kp);
  uint32_t* sbeg = reinterpret_cast<uint32_t*>(smem_raw + LY::o_sbeg);
  uint32_t* send = reinterpret_cast<uint32_t*>(smem_raw + LY::o_send);
  int* misc = reinterpret_cast<int*>(smem_raw + LY::o_misc);
  uint2* stack = reinterpret_cast<uint2*>(smem_raw + LY::o_stack);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw + LY::o_bar);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  // one-time init
  for (int i = tid; i <= S; i += T) skeys[i] = EMPTY;
  for (int i = tid; i < (S + 1) * MW; i += T) smask[i] = 0;
  for (int i = tid; i < NS; i += T) skp[i] = i < n_sets ? sets[i].keys : nullptr;
  if (tid == 0) {
    misc[kMiscNdist] = 0; misc[kMiscIns] = 0; misc[kMiscOverflow] = 0; misc[kMiscSpecial] = 0; misc[kMiscAnyMma] = 0;
    umma::mbar_init(&bar[0], 1);
    umma::mbar_init(&bar[1], 1);
    umma::mbar_fence_init();
  }
  if (warp == 0) umma::tmem_alloc(reinterpret_cast<uint32_t*>(&misc[kMiscTmem]), TMEM_COLS);
  umma::fence_before_thread_sync();
  __syncthreads();
  umma::fence_after_thread_sync();
  const uint32_t tmem_base = (uint32_t)misc[kMiscTmem];
  const uint32_t stage_addr = umma::smem_u32(stage);
  constexpr uint32_t idesc = umma::make_idesc_u8(MMA_M, MMA_N, 1, 1);
  uint32_t uses0 = 0, uses1 = 0;   // commits issued on staging buffer 0 / 1 (uniform over the CTA)
  bool mma_started = false;        // thread 0: the accumulators hold data
  unsigned long long st_keys = 0, st_dist = 0, st_over = 0;

  // byte offset of the mask byte this warp owns in slot 0 (word plane warp / 4, byte warp % 4)
  const uint32_t mbase = (uint32_t)(warp >> 2) * (uint32_t)(S + 1) * 4u + (uint32_t)(warp & 3);
  // sets owned by this warp
  const int s_lo = warp * spw;
  const int s_hi = min(n_sets, s_lo + spw);

  const uint32_t n_tiles = *n_tiles_p;
  // The CTA always holds the tile it works on and the one after it: the key ranges of the
  // next tile are fetched and its keys are prefetched into L2 while the current tile is built.
  if (tid == 0) {
    misc[kMiscTile] = (int)atomicAdd(tile_counter, 1u);
    misc[kMiscNext] = (int)atomicAdd(tile_counter, 1u);
  }
  __syncthreads();
  {
    const uint32_t t0 = (uint32_t)misc[kMiscTile];
    if (t0 < n_tiles && tid < NS) {
      const Tile tl0 = tiles[t0];
      sbeg[tid] = tid < n_sets ? offsT[(size_t)tl0.x0 * n_sets + tid] : 0u;
      send[tid] = tid < n_sets ? offsT[(size_t)tl0.x1 * n_sets + tid] : 0u;
    }
  }
  int par = 0;  // which half of sbeg / send holds the current tile
  for (;;) {
    __syncthreads();
    const uint32_t t_id = (uint32_t)misc[kMiscTile];
    const uint32_t t_nx = (uint32_t)misc[kMiscNext];
    if (t_id >= n_tiles) break;
    const Tile tl = tiles[t_id];
    const uint32_t bucket0 = tl.x0 >> fine_level;
    const int nbk = (int)(((tl.x1 - 1) >> fine_level) - bucket0) + 1;
    const uint32_t* sb = sbeg + par * NS;
    const uint32_t* se = send + par * NS;
    // next tile: key ranges into registers now, into shared memory after the build
    uint32_t nx_b = 0, nx_e = 0;
    if (t_nx < n_tiles && tid < n_sets) {
      const Tile tn = tiles[t_nx];
      // volatile asm: the loads are issued here, not sunk to their use after the build
      asm volatile("ld.global.nc.u32 %0, [%1];" : "=r"(nx_b) : "l"(offsT + (size_t)tn.x0 * n_sets + tid));
      asm volatile("ld.global.nc.u32 %0, [%1];" : "=r"(nx_e) : "l"(offsT + (size_t)tn.x1 * n_sets + tid));
    }
    if (tid == 0) { misc[kMiscSp] = 1; stack[0] = make_uint2(0u, 1u); }  // class (p=0, P=1) = everything
    bool prefetched = false;
    __syncthreads();

    // process the stack of key classes (normally exactly one entry)
    for (uint32_t n_pass = 0;; n_pass++) {
      __syncthreads();
      if (misc[kMiscSp] == 0) break;
      if (n_pass > (1u << 20)) {  // watchdog
        if (tid == 0) { atomicAdd(&stats[3], 1ull); atomicOr(&stats[4], 4ull); }
        break;
      }
      const uint2 cls = stack[misc[kMiscSp] - 1];
      __syncthreads();
      if (tid == 0) misc[kMiscSp] -= 1;
      const uint32_t cp = cls.x, cmask = cls.y - 1u;

      // ---- build: warp w inserts the keys of its own sets, one set after the other. The
      // keys of batch i + 1 (possibly of the next set) are loaded before batch i is probed.
      uint32_t w_new = 0;  // slots claimed by this warp, not yet published
      if (nbk == 1) {
        PcBatch cur, nxt;
        KeyT ck[UNR], nk[UNR];
        int s = s_lo;
        auto first_of = [&](int from, PcBatch& o) -> int {  // first non-empty set >= from
          int t = from;
          while (t < s_hi && se[t] <= sb[t]) t++;
          if (t < s_hi) { o.kp = skp[t]; o.i0 = sb[t]; o.end = se[t]; o.bit = 1u << (t - s_lo); }
          return t;
        };
        s = first_of(s_lo, cur);
        bool have = s < s_hi;
        if (have) batch_load<KeyT, UNR>(cur, lane, ck);
        while (have) {
          // successor: next rows of the same set, else the first rows of the next set
          bool hn;
          bool new_set = false;
          if (cur.i0 + 32 * UNR < cur.end) {
            nxt = cur; nxt.i0 = cur.i0 + 32 * UNR; hn = true;
          } else {
            s = first_of(s + 1, nxt);
            hn = s < s_hi;
            new_set = true;
          }
          if (hn) batch_load<KeyT, UNR>(nxt, lane, nk);
          batch_process<KeyT, TK, NS, UNR>(skeys, maskb, misc, cur, ck, (TK)0, mbase, lane, cmask, cp, w_new);
          if (new_set) __syncwarp();  // marks of different sets may hit the same byte
          if (*(volatile int*)&misc[kMiscOverflow]) break;
          cur = nxt;
#pragma unroll
          for (int u = 0; u < UNR; u++) ck[u] = nk[u];
          have = hn;
        }
      } else {
        // tile spans several (small) buckets: table key = (bucket - first bucket of the tile) << key_bits | key
        for (int s = s_lo; s < s_hi; s++) {
          for (int b = 0; b < nbk; b++) {
            const size_t row0 = (size_t)(bucket0 + (uint32_t)b) << fine_level;
            PcBatch cur;
            cur.kp = skp[s];
            cur.bit = 1u << (s - s_lo);
            cur.end = min(offsT[(row0 + ((size_t)1 << fine_level)) * n_sets + s], se[s]);
            for (cur.i0 = max(offsT[row0 * n_sets + s], sb[s]); cur.i0 < cur.end; cur.i0 += 32 * UNR) {
              KeyT ck[UNR];
              batch_load<KeyT, UNR>(cur, lane, ck);
              batch_process<KeyT, TK, NS, UNR>(skeys, maskb, misc, cur, ck, (TK)b << key_bits, mbase, lane, cmask, cp, w_new);
            }
          }
          __syncwarp();
          if (*(volatile int*)&misc[kMiscOverflow]) break;
        }
      }
      if (w_new && lane == 0 && atomicAdd(&misc[kMiscIns], (int)w_new) + (int)w_new > C::D) misc[kMiscOverflow] = 1;
      __syncthreads();

      if (!prefetched) {
        // keys of the next tile -> L2 (one 128-byte line per request), ranges -> shared memory
        prefetched = true;
        if (tid < NS) { sbeg[(par ^ 1) * NS + tid] = nx_b; send[(par ^ 1) * NS + tid] = nx_e; }
        if (tid < n_sets && nx_e > nx_b) {
          const char* base = (const char*)skp[tid];
          const size_t lo = ((size_t)nx_b * sizeof(KeyT)) & ~(size_t)127, hi = (size_t)nx_e * sizeof(KeyT);
          for (size_t off = lo; off < hi; off += 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(base + off));
        }
      }

      // ---- compact: list the occupied slots ----------------------------------------------
      if (!misc[kMiscOverflow]) {
        for (int base = warp * 32; base <= S; base += T) {
          const int slot = base + lane;
          const bool occ = slot < S ? (skeys[slot] != EMPTY) : (slot == S && sizeof(TK) == 4 && misc[kMiscSpecial] != 0);
          const unsigned bal = __ballot_sync(0xffffffffu, occ);
          if (bal) {
            int pos = 0;
            if (lane == 0) pos = atomicAdd(&misc[kMiscNdist], __popc(bal));
            pos = __shfl_sync(0xffffffffu, pos, 0);
            if (occ) {
              const int r = pos + __popc(bal & ((1u << lane) - 1u));
              if (r < DMAX) list[r] = (uint16_t)slot;
            }
          }
        }
      }
      __syncthreads();
      const int D = misc[kMiscNdist];
      if (misc[kMiscOverflow] || D > DMAX) {
        // too many distinct keys for one pass: wipe the table, split the class in two
        __syncthreads();
        for (int i = tid; i <= S; i += T) skeys[i] = EMPTY;
        for (int i = tid; i < (S + 1) * MW; i += T) smask[i] = 0;
        if (tid == 0) {
          misc[kMiscNdist] = 0; misc[kMiscIns] = 0; misc[kMiscOverflow] = 0; misc[kMiscSpecial] = 0;
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

      // ---- gram: CK K-steps (32 distinct keys each) per staging buffer ------------------------
      for (int c0 = 0; c0 < D; c0 += CK * 32) {
        const int nk = min(CK, (D - c0 + 31) >> 5);   // K-steps in this chunk
        const uint32_t nuse = uses0 + uses1;
        const int buf = (int)(nuse & 1u);
        const uint32_t used = buf ? uses1 : uses0;
        // the MMAs that read this buffer two chunks ago must have completed
        if (used > 0 && !umma::mbar_wait_bounded(&bar[buf], (used - 1) & 1u)) {
          if (lane == 0) { atomicAdd(&stats[3], 1ull); atomicOr(&stats[4], 1ull); }
        }
        unsigned char* sb = stage + (size_t)buf * LY::buf_bytes;
        // work item = (K-step, mask word): 32 keys x 32 sets -> two 16-set blocks of 16-byte rows
        for (int item = warp; item < nk * MW; item += NW) {
          const int ks = item / MW, w = item - ks * MW;
          const int r = c0 + ks * 32 + lane;
          uint32_t m = 0;
          if (r < D) {
            const uint32_t slot = list[r];
            m = smask[w * (S + 1) + slot];
            smask[w * (S + 1) + slot] = 0;
            if (w == 0) skeys[slot] = EMPTY;
          }
          uint4 lo, hi;
          lo.x = umma::nibble_to_bytes(m);       lo.y = umma::nibble_to_bytes(m >> 4);
          lo.z = umma::nibble_to_bytes(m >> 8);  lo.w = umma::nibble_to_bytes(m >> 12);
          hi.x = umma::nibble_to_bytes(m >> 16); hi.y = umma::nibble_to_bytes(m >> 20);
          hi.z = umma::nibble_to_bytes(m >> 24); hi.w = umma::nibble_to_bytes(m >> 28);
          unsigned char* kb = sb + (size_t)ks * LY::kstep_bytes + (size_t)lane * 16;
          *reinterpret_cast<uint4*>(kb + (size_t)(2 * w) * 512) = lo;
          *reinterpret_cast<uint4*>(kb + (size_t)(2 * w + 1) * 512) = hi;
        }
        umma::fence_async_smem();
        __syncthreads();
        if (tid == 0) {
          umma::fence_after_thread_sync();
          const uint32_t a0 = stage_addr + (uint32_t)buf * (uint32_t)LY::buf_bytes;
          for (int ks = 0; ks < nk; ks++) {
            const uint32_t a = a0 + (uint32_t)ks * (uint32_t)LY::kstep_bytes;
            const uint64_t bd = umma::make_smem_desc(a, 128u, 512u);
            umma::mma_u8(tmem_base, bd, bd, idesc, mma_started ? 1u : 0u);
            if (NS == 256) {
              // rows 128..255 of the Gram: A operand starts at set block 8, accumulators at column 256
              const uint64_t ad = umma::make_smem_desc(a + 8u * 512u, 128u, 512u);
              umma::mma_u8(tmem_base + 256u, ad, bd, idesc, mma_started ? 1u : 0u);
            }
            mma_started = true;
          }
          umma::mma_commit(&bar[buf]);
          misc[kMiscAnyMma] = 1;
        }
        if (buf) uses1++; else uses0++;
      }
      if (tid == 0) { misc[kMiscNdist] = 0; misc[kMiscIns] = 0; misc[kMiscSpecial] = 0; }
    }
    if (tid < NS) st_keys += se[tid] - sb[tid];
    __syncthreads();
    if (tid == 0) {
      misc[kMiscTile] = misc[kMiscNext];
      misc[kMiscNext] = (int)atomicAdd(tile_counter, 1u);
    }
    par ^= 1;
  }

  // ---- drain the tensor pipe, read the accumulators back, add them to W ----------------------
  if ((uses0 > 0 && !umma::mbar_wait_bounded(&bar[0], (uses0 - 1) & 1u)) ||
      (uses1 > 0 && !umma::mbar_wait_bounded(&bar[1], (uses1 - 1) & 1u))) {
    if (lane == 0) { atomicAdd(&stats[3], 1ull); atomicOr(&stats[4], 2ull); }
  }
  umma::fence_after_thread_sync();
  __syncthreads();
  if (misc[kMiscAnyMma] && warp < 4) {
    // warp q reads TMEM lanes 32 q .. 32 q + 31. M = 64: row m lives in lane (m % 16) + 32 (m / 16);
    // M = 128: row m in lane m (rows 128.. of NS = 256 in columns 256..511).
    constexpr int HALVES = NS == 256 ? 2 : 1;
#pragma unroll 1
    for (int half = 0; half < HALVES; half++) {
      int row;
      if (NS == 64) row = lane < 16 ? warp * 16 + lane : -1;
      else row = half * 128 + warp * 32 + lane;
      const int si = row >= 0 ? set_of_pos(row, spw, n_sets) : -1;
#pragma unroll 1
      for (int c0 = 0; c0 < MMA_N; c0 += 32) {
        uint32_t v[32];
        umma::tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(half * 256 + c0), v);
        umma::tmem_ld_wait();
        if (si >= 0) {
#pragma unroll
          for (int j = 0; j < 32; j++) {
            const int sj = set_of_pos(c0 + j, spw, n_sets);
            if (sj >= 0 && v[j] != 0) atomicAdd(&W[(size_t)si * n_sets + sj], (unsigned long long)v[j]);
          }
        }
      }
    }
  }
  umma::fence_before_thread_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tmem_base, TMEM_COLS);

  // stats: keys processed, distinct keys, overflow retries
  for (int o = 16; o > 0; o >>= 1) st_keys += __shfl_xor_sync(0xffffffffu, st_keys, o);
  if (lane == 0 && st_keys) atomicAdd(&stats[0], st_keys);
  if (tid == 0) {
    if (st_dist) atomicAdd(&stats[1], st_dist);
    if (st_over) atomicAdd(&stats[2], st_over);
  }
}

}  // namespace kmsc
#include "pair_counts_stream.cuh"
#include "pair_counts_whash.cuh"
#include "pair_counts_lane.cuh"
namespace kmsc {

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

template <typename KeyT, int NS>
static int launch_main(kmsc_ctx* ctx, const SetDesc* d_sets, int n_sets, const uint32_t* d_offsT,
                       const Tile* d_tiles, const uint32_t* d_ntiles, uint32_t* d_counter,
                       unsigned long long* d_W, unsigned long long* d_stats, int key_bits, int fine_level,
                       uint32_t max_tiles) {
  using TK = typename TableKey<KeyT>::type;
  using C = PcCfg<NS, (int)sizeof(TK)>;
  const size_t smem = PcSmem<NS>::bytes((int)sizeof(TK));
  auto kern = pair_counts_kernel<KeyT, NS>;
  // Resident CTAs per SM from the kernel's own resources. (The occupancy API answers 1 for
  // any kernel that allocates tensor memory, whatever the column count.) Every resident CTA
  // holds its accumulators in TMEM: 512 columns per SM bound the count as well. Computed once
  // per instantiation and device.
  static int occ_cache[64] = {0};
  int occ = ctx->device < 64 ? occ_cache[ctx->device] : 0;
  if (occ == 0) {
    KMSC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaFuncAttributes fa;
    KMSC_CUDA(cudaFuncGetAttributes(&fa, kern));
    int smem_sm = 0, smem_res = 0, regs_sm = 0, thr_sm = 0;
    KMSC_CUDA(cudaDeviceGetAttribute(&smem_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, ctx->device));
    KMSC_CUDA(cudaDeviceGetAttribute(&smem_res, cudaDevAttrReservedSharedMemoryPerBlock, ctx->device));
    KMSC_CUDA(cudaDeviceGetAttribute(&regs_sm, cudaDevAttrMaxRegistersPerMultiprocessor, ctx->device));
    KMSC_CUDA(cudaDeviceGetAttribute(&thr_sm, cudaDevAttrMaxThreadsPerMultiProcessor, ctx->device));
    const int tmem_cols = NS == 256 ? 512 : NS;
    occ = (int)((size_t)smem_sm / (smem + fa.sharedSizeBytes + (size_t)smem_res));
    const int regs_per_cta = ((fa.numRegs + 7) & ~7) * C::T;
    if (regs_per_cta > 0) occ = std::min(occ, regs_sm / regs_per_cta);
    occ = std::min(occ, thr_sm / C::T);
    occ = std::min(occ, 512 / tmem_cols);
    if (occ < 1) { set_error("pair_counts kernel does not fit on an SM (smem %zu, %d regs)", smem, fa.numRegs); return KMSC_E_CUDA; }
    if (getenv("KMSC_DEBUG"))
      fprintf(stderr, "[kmsc] pair_counts NS=%d T=%d smem=%zu regs=%d -> %d CTAs per SM\n", NS, C::T, smem, fa.numRegs, occ);
    if (ctx->device < 64) occ_cache[ctx->device] = occ;
  }
  long long grid = (long long)ctx->sm_count * occ;
  if (grid > (long long)max_tiles) grid = max_tiles;
  if (grid < 1) grid = 1;
  const int n_warps = C::T / 32;
  const int spw = (n_sets + n_warps - 1) / n_warps;  // sets per warp (<= 8)
  kern<<<(unsigned)grid, C::T, smem, ctx->stream>>>(d_sets, n_sets, spw, d_offsT, d_tiles, d_ntiles,
                                                    d_counter, d_W, d_stats, key_bits, fine_level);
  count_launch(ctx);
  KMSC_CUDA(cudaGetLastError());
  return KMSC_OK;
}

// the merge build (pair_counts_stream.cuh): n <= 128 sets
template <typename KeyT, int NS>
static int launch_stream(kmsc_ctx* ctx, const SetDesc* d_sets, int n_sets, const uint32_t* d_offsT,
                         const Tile* d_tiles, const uint32_t* d_ntiles, uint32_t* d_counter,
                         unsigned long long* d_W, unsigned long long* d_stats, int fine_level, int finest_level,
                         uint32_t max_tiles) {
  using C = PmCfg<NS>;
  const size_t smem = PmLayout<KeyT, NS>::total;
  auto kern = pair_counts_stream_kernel<KeyT, NS>;
  static int occ_cache[64] = {0};
  int occ = ctx->device < 64 ? occ_cache[ctx->device] : 0;
  if (occ == 0) {
    KMSC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaFuncAttributes fa;
    KMSC_CUDA(cudaFuncGetAttributes(&fa, kern));
    int smem_sm = 0, smem_res = 0, regs_sm = 0, thr_sm = 0;
    KMSC_CUDA(cudaDeviceGetAttribute(&smem_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, ctx->device));
    KMSC_CUDA(cudaDeviceGetAttribute(&smem_res, cudaDevAttrReservedSharedMemoryPerBlock, ctx->device));
    KMSC_CUDA(cudaDeviceGetAttribute(&regs_sm, cudaDevAttrMaxRegistersPerMultiprocessor, ctx->device));
    KMSC_CUDA(cudaDeviceGetAttribute(&thr_sm, cudaDevAttrMaxThreadsPerMultiProcessor, ctx->device));
    occ = (int)((size_t)smem_sm / (smem + fa.sharedSizeBytes + (size_t)smem_res));
    const int regs_per_cta = ((fa.numRegs + 7) & ~7) * C::T;
    if (regs_per_cta > 0) occ = std::min(occ, regs_sm / regs_per_cta);
    occ = std::min(occ, thr_sm / C::T);
    occ = std::min(occ, 512 / NS);  // tensor-memory columns
    if (occ < 1) { set_error("pair_counts stream kernel does not fit on an SM (smem %zu, %d regs)", smem, fa.numRegs); return KMSC_E_CUDA; }
    if (getenv("KMSC_DEBUG"))
      fprintf(stderr, "[kmsc] pair_counts stream NS=%d T=%d smem=%zu regs=%d -> %d CTAs per SM\n", NS, C::T, smem, fa.numRegs, occ);
    if (ctx->device < 64) occ_cache[ctx->device] = occ;
  }
  long long grid = (long long)ctx->sm_count * occ;
  if (grid > (long long)max_tiles) grid = max_tiles;
  if (grid < 1) grid = 1;
  const int spw = (n_sets + NS / 8 - 1) / (NS / 8);  // sets per group of 8 Gram positions
  // prefetching the next tile into L2 (as the hash build does) costs this kernel time: its tiles
  // are several buckets long, 444 CTAs' worth of next tiles exceed the L2 and are read twice
  // (ncu r01: 5.4 GB of DRAM reads for 2.56 GB of keys with it). KMSC_P3_PF=1 turns it on.
  const int l2_prefetch = getenv("KMSC_P3_PF") ? atoi(getenv("KMSC_P3_PF")) : 0;
  kern<<<(unsigned)grid, C::T, smem, ctx->stream>>>(d_sets, n_sets, spw, d_offsT, d_tiles, d_ntiles, d_counter, d_W,
                                                    d_stats, fine_level, finest_level, l2_prefetch);
  count_launch(ctx);
  KMSC_CUDA(cudaGetLastError());
  return KMSC_OK;
}

template <typename KeyT>
static int launch_stream_ns(kmsc_ctx* ctx, int ns, const SetDesc* d_sets, int n_sets, const uint32_t* d_offsT,
                            const Tile* d_tiles, const uint32_t* d_ntiles, uint32_t* d_counter,
                            unsigned long long* d_W, unsigned long long* d_stats, int fine_level, int finest_level,
                            uint32_t max_tiles) {
  if (ns == 64) return launch_stream<KeyT, 64>(ctx, d_sets, n_sets, d_offsT, d_tiles, d_ntiles, d_counter, d_W, d_stats, fine_level, finest_level, max_tiles);
  return launch_stream<KeyT, 128>(ctx, d_sets, n_sets, d_offsT, d_tiles, d_ntiles, d_counter, d_W, d_stats, fine_level, finest_level, max_tiles);
}

// the warp-private hash build (pair_counts_whash.cuh): n <= 64 related sets
template <typename KeyT>
static int launch_whash(kmsc_ctx* ctx, const SetDesc* d_sets, int n_sets, const uint32_t* d_offsT, const Tile* d_tiles,
                        const uint32_t* d_ntiles, uint32_t* d_counter, unsigned long long* d_W, unsigned long long* d_stats,
                        int fine_level, int finest_level, uint32_t max_tiles, int rho_q) {
  constexpr int NS = 64;
  using C = WhCfg<NS>;
  const size_t smem = WhLayout<KeyT, NS>::total;
  auto kern = pair_counts_whash_kernel<KeyT, NS>;
  static int occ_cache[64] = {0};
  int occ = ctx->device < 64 ? occ_cache[ctx->device] : 0;
  if (occ == 0) {
    KMSC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaFuncAttributes fa;
    KMSC_CUDA(cudaFuncGetAttributes(&fa, kern));
    int smem_sm = 0, smem_res = 0, regs_sm = 0, thr_sm = 0;
    KMSC_CUDA(cudaDeviceGetAttribute(&smem_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, ctx->device));
    KMSC_CUDA(cudaDeviceGetAttribute(&smem_res, cudaDevAttrReservedSharedMemoryPerBlock, ctx->device));
    KMSC_CUDA(cudaDeviceGetAttribute(&regs_sm, cudaDevAttrMaxRegistersPerMultiprocessor, ctx->device));
    KMSC_CUDA(cudaDeviceGetAttribute(&thr_sm, cudaDevAttrMaxThreadsPerMultiProcessor, ctx->device));
    occ = (int)((size_t)smem_sm / (smem + fa.sharedSizeBytes + (size_t)smem_res));
    const int regs_per_cta = ((fa.numRegs + 7) & ~7) * C::T;
    if (regs_per_cta > 0) occ = std::min(occ, regs_sm / regs_per_cta);
    occ = std::min(occ, thr_sm / C::T);
    occ = std::min(occ, 512 / NS);  // tensor-memory columns
    if (occ < 1) { set_error("pair_counts whash kernel does not fit on an SM (smem %zu, %d regs)", smem, fa.numRegs); return KMSC_E_CUDA; }
    if (getenv("KMSC_DEBUG"))
      fprintf(stderr, "[kmsc] pair_counts whash NS=%d T=%d smem=%zu regs=%d -> %d CTAs per SM\n", NS, C::T, smem, fa.numRegs, occ);
    if (ctx->device < 64) occ_cache[ctx->device] = occ;
  }
  long long grid = (long long)ctx->sm_count * occ;
  if (grid > (long long)max_tiles) grid = max_tiles;
  if (grid < 1) grid = 1;
  kern<<<(unsigned)grid, C::T, smem, ctx->stream>>>(d_sets, n_sets, d_offsT, d_tiles, d_ntiles, d_counter, d_W, d_stats, fine_level,
                                                    finest_level, rho_q);
  count_launch(ctx);
  KMSC_CUDA(cudaGetLastError());
  return KMSC_OK;
}

template <typename KeyT>
static int launch_main_ns(kmsc_ctx* ctx, int ns, const SetDesc* d_sets, int n_sets,
                          const uint32_t* d_offsT, const Tile* d_tiles, const uint32_t* d_ntiles,
                          uint32_t* d_counter, unsigned long long* d_W, unsigned long long* d_stats,
                          int key_bits, int fine_level, uint32_t max_tiles) {
  switch (ns) {
    case 64: return launch_main<KeyT, 64>(ctx, d_sets, n_sets, d_offsT, d_tiles, d_ntiles, d_counter, d_W, d_stats, key_bits, fine_level, max_tiles);
    case 128: return launch_main<KeyT, 128>(ctx, d_sets, n_sets, d_offsT, d_tiles, d_ntiles, d_counter, d_W, d_stats, key_bits, fine_level, max_tiles);
    default: return launch_main<KeyT, 256>(ctx, d_sets, n_sets, d_offsT, d_tiles, d_ntiles, d_counter, d_W, d_stats, key_bits, fine_level, max_tiles);
  }
}


// ---------------------------------------------------------------------------
// the lane build (pair_counts_lane.cuh): n <= 64 sets, 2- / 4-byte keys, no planning pass
// ---------------------------------------------------------------------------
constexpr int kLaneTarget = 42;   // mean distinct keys per row the 64-slot lane tables are run at

// The level whose rows hold <= kLaneTarget distinct keys on average (rho = keys per distinct key), or -1.
// 4-byte keys that use all 32 bits need f >= 1: the table key (the key without the row's f bits) must
// never be the empty marker.
static int lane_level(const kmsc_set* s0, double total_keys, double rho, int nb_live, bool forced) {
  if (s0->key_bytes > 4) return -1;
  const int f_min = (s0->key_bytes == 4 && s0->key_bits >= 32) ? 1 : 0;
  if (f_min > s0->max_level) return -1;
  const double distinct = total_keys / (rho > 1.0 ? rho : 1.0);
  for (int f = f_min; f <= s0->max_level; f++)
    if (distinct / ((double)nb_live * (double)(1 << f)) <= (double)kLaneTarget) return f;
  return forced ? s0->max_level : -1;
}

template <typename KeyT>
static int launch_lane(kmsc_ctx* ctx, const SetDesc* d_sets, int n_sets, int f, uint32_t NF, uint32_t row_lo,
                       uint32_t row_hi, const uint32_t* d_bitmap, int tbits, float inv_rho, uint32_t* d_counter,
                       unsigned long long* d_W, unsigned long long* d_stats, uint2* d_retry,
                       uint32_t* d_retry_count, uint32_t retry_cap) {
  auto kern = pair_counts_lane_kernel<KeyT>;
  static bool attr_set[64] = {false};
  if (ctx->device >= 64 || !attr_set[ctx->device]) {
    KMSC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LnLayout::total));
    if (getenv("KMSC_DEBUG")) {
      cudaFuncAttributes fa;
      KMSC_CUDA(cudaFuncGetAttributes(&fa, kern));
      fprintf(stderr, "[kmsc] pair_counts lane T=%d smem=%zu regs=%d\n", LnCfg::T, (size_t)LnLayout::total, fa.numRegs);
    }
    if (ctx->device < 64) attr_set[ctx->device] = true;
  }
  const long long chunks = ((long long)(row_hi - (row_lo & ~31u)) + 31) / 32;
  long long grid = std::min<long long>(ctx->sm_count, (chunks + LnCfg::NW - 1) / LnCfg::NW);
  if (grid < 1) grid = 1;
  kern<<<(unsigned)grid, LnCfg::T, LnLayout::total, ctx->stream>>>(d_sets, n_sets, f, NF, row_lo, row_hi, d_bitmap,
                                                                  tbits, inv_rho, d_counter, d_W, d_stats, d_retry,
                                                                  d_retry_count, retry_cap);
  pair_counts_lane_retry_kernel<KeyT><<<ctx->sm_count, 128, 0, ctx->stream>>>(d_sets, n_sets, NF, tbits, d_retry,
                                                                             d_retry_count, retry_cap, d_W, d_stats);
  count_launch(ctx, 2);
  KMSC_CUDA(cudaGetLastError());
  return KMSC_OK;
}

constexpr int kLaneRedo = 1001;  // internal: the retry list filled up, redo the call with another build

// One pass of the lane build over the buckets selected by h_bitmap (NULL = all) at level f.
static int run_phase_lane(kmsc_ctx* ctx, const kmsc_set* const* sets, int n, int f, double rho, const uint32_t* h_bitmap,
                          unsigned long long* d_W, unsigned long long host_stats[4]) {
  const kmsc_set* s0 = sets[0];
  for (int i = 0; i < n; i++) KMSC_TRY(set_ensure_levels(ctx, sets[i]));
  const int nb = 1 << s0->N;
  const uint32_t NF = (uint32_t)nb << f;
  const size_t sz_desc = ((size_t)n * sizeof(SetDesc) + 15) & ~(size_t)15;
  const size_t sz_bitmap = (size_t)((nb + 31) / 32) * 4;
  KMSC_TRY(ctx->small.reserve(256 + sz_desc + sz_bitmap + 64));
  unsigned char* sm = (unsigned char*)ctx->small.p;
  uint32_t* d_counter = (uint32_t*)(sm + 4);
  uint32_t* d_retry_count = (uint32_t*)(sm + 8);
  unsigned long long* d_stats = (unsigned long long*)(sm + 64);
  SetDesc* d_sets = (SetDesc*)(sm + 256);
  uint32_t* d_bitmap = h_bitmap ? (uint32_t*)(sm + 256 + sz_desc) : nullptr;
  const uint32_t retry_cap = std::max<uint32_t>(4096u, NF / 4u);
  KMSC_TRY(ctx->plan.reserve((size_t)retry_cap * 8));
  uint2* d_retry = (uint2*)ctx->plan.p;

  void* pin = nullptr;
  KMSC_TRY(ctx_pinned(ctx, sz_desc + sz_bitmap + 64, &pin));
  SetDesc* h_sets = (SetDesc*)pin;
  for (int i = 0; i < n; i++) { h_sets[i].keys = sets[i]->keys; h_sets[i].lev = sets[i]->lev[f]; h_sets[i].lev_finest = sets[i]->lev[sets[i]->max_level]; }
  KMSC_CUDA(cudaMemsetAsync(sm, 0, 256, ctx->stream));
  KMSC_CUDA(cudaMemcpyAsync(d_sets, h_sets, sz_desc, cudaMemcpyHostToDevice, ctx->stream));
  if (h_bitmap) {
    uint32_t* pb = (uint32_t*)((unsigned char*)pin + sz_desc);
    memcpy(pb, h_bitmap, sz_bitmap);
    KMSC_CUDA(cudaMemcpyAsync(d_bitmap, pb, sz_bitmap, cudaMemcpyHostToDevice, ctx->stream));
  }
  for (int i = 0; i < 3; i++)
    if (!ctx->pc_ev[i]) KMSC_CUDA(cudaEventCreate(&ctx->pc_ev[i]));
  int span_lo = nb, span_hi = 0;
  for (int i = 0; i < n; i++) {
    span_lo = std::min(span_lo, sets[i]->b_lo < 0 ? 0 : sets[i]->b_lo);
    span_hi = std::max(span_hi, sets[i]->b_hi < 0 ? nb : sets[i]->b_hi);
  }
  if (span_lo >= span_hi) { span_lo = 0; span_hi = nb; }
  const uint32_t row_lo = (uint32_t)span_lo << f, row_hi = (uint32_t)span_hi << f;
  const int tbits = s0->key_bits - f;   // bits of a table key
  const float inv_rho = (float)(1.0 / (rho > 1.0 ? rho : 1.0));
  KMSC_CUDA(cudaEventRecord(ctx->pc_ev[1], ctx->stream));
  int rc;
  if (s0->key_bytes == 2)
    rc = launch_lane<uint16_t>(ctx, d_sets, n, f, NF, row_lo, row_hi, d_bitmap, tbits, inv_rho, d_counter, d_W, d_stats, d_retry, d_retry_count, retry_cap);
  else
    rc = launch_lane<uint32_t>(ctx, d_sets, n, f, NF, row_lo, row_hi, d_bitmap, tbits, inv_rho, d_counter, d_W, d_stats, d_retry, d_retry_count, retry_cap);
  if (rc != KMSC_OK) return rc;
  KMSC_CUDA(cudaEventRecord(ctx->pc_ev[2], ctx->stream));
  unsigned long long* h_stats = (unsigned long long*)((unsigned char*)pin + sz_desc + sz_bitmap);
  KMSC_CUDA(cudaMemcpyAsync(h_stats, d_stats, 64, cudaMemcpyDeviceToHost, ctx->stream));
  KMSC_CUDA(cudaStreamSynchronize(ctx->stream));
  for (int i = 0; i < 4; i++) host_stats[i] = h_stats[i];
  {
    float ms_main = 0;
    KMSC_CUDA(cudaEventElapsedTime(&ms_main, ctx->pc_ev[1], ctx->pc_ev[2]));
    ctx->pc_main_ms += ms_main;
    ctx->pc_main_launches += 1;
    // every key once + one offset per (row, set) at level f + the bucket-level rows of the formula
    ctx->pc_algo_bytes += (double)h_stats[0] * s0->key_bytes + (double)(nb + 1) * n * 4.0;
  }
  if (h_stats[4] & 16ull) return kLaneRedo;
  if (host_stats[3]) {
    set_error("pair_counts (lane build): %llu failures (watchdog code %llu)", host_stats[3], h_stats[4]);
    return KMSC_E_STATE;
  }
  return KMSC_OK;
}

constexpr int kWhashOverflow = 1000;  // internal: run_phase's "redo with another build"

static int dmax_for(int ns, int tk_bytes) {
  if (ns == 64) return PcCfg<64, 4>::D;
  if (ns == 128) return PcCfg<128, 4>::D;
  return tk_bytes == 8 ? PcCfg<256, 8>::D : PcCfg<256, 4>::D;
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
static int run_phase(kmsc_ctx* ctx, const kmsc_set* const* sets, int n, int ns, int build,
                     const uint32_t* h_bitmap, unsigned long long L, double keys_in_phase,
                     unsigned long long* d_W, unsigned long long host_stats[4], int rho_q = 1) {
  const bool merge_build = build == 1;   // 0 hash table per CTA, 1 warp-wide merge, 2 warp-private hash tables
  const kmsc_set* s0 = sets[0];
  for (int i = 0; i < n; i++) KMSC_TRY(set_ensure_levels(ctx, sets[i]));
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
  for (int i = 0; i < n; i++) { h_sets[i].keys = sets[i]->keys; h_sets[i].lev = sets[i]->lev[f]; h_sets[i].lev_finest = sets[i]->lev[sets[i]->max_level]; }
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
    int span_lo = nb, span_hi = 0;
    for (int i = 0; i < n; i++) {
      span_lo = std::min(span_lo, sets[i]->b_lo < 0 ? 0 : sets[i]->b_lo);
      span_hi = std::max(span_hi, sets[i]->b_hi < 0 ? nb : sets[i]->b_hi);
    }
    if (span_lo >= span_hi) { span_lo = 0; span_hi = nb; }
    pp.row_lo = (uint32_t)span_lo << f;
    pp.row_hi = (uint32_t)span_hi << f;
  }
  {
    // a tile may span several buckets only while (bucket - first bucket, key) fits
    // the table key: uint32 for 2/4-byte keys, uint64 for 8-byte keys
    const int tk_bits = s0->key_bytes == 8 ? 64 : 32;
    const int spare = tk_bits - s0->key_bits;
    // (the merge build has no table key: its tiles may always span buckets)
    const uint32_t bucket_span = (spare >= 8 || build != 0) ? 256u : (1u << spare);
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
  if (build == 2) {
    switch (s0->key_bytes) {
      case 2: rc = launch_whash<uint16_t>(ctx, d.sets, n, d_offsT, d_tiles, d.n_tiles, d.counter, d_W, d.stats, f, s0->max_level, max_tiles, rho_q); break;
      case 4: rc = launch_whash<uint32_t>(ctx, d.sets, n, d_offsT, d_tiles, d.n_tiles, d.counter, d_W, d.stats, f, s0->max_level, max_tiles, rho_q); break;
      default: rc = launch_whash<unsigned long long>(ctx, d.sets, n, d_offsT, d_tiles, d.n_tiles, d.counter, d_W, d.stats, f, s0->max_level, max_tiles, rho_q); break;
    }
  } else if (merge_build) {
    switch (s0->key_bytes) {
      case 2: rc = launch_stream_ns<uint16_t>(ctx, ns, d.sets, n, d_offsT, d_tiles, d.n_tiles, d.counter, d_W, d.stats, f, s0->max_level, max_tiles); break;
      case 4: rc = launch_stream_ns<uint32_t>(ctx, ns, d.sets, n, d_offsT, d_tiles, d.n_tiles, d.counter, d_W, d.stats, f, s0->max_level, max_tiles); break;
      default: rc = launch_stream_ns<unsigned long long>(ctx, ns, d.sets, n, d_offsT, d_tiles, d.n_tiles, d.counter, d_W, d.stats, f, s0->max_level, max_tiles); break;
    }
  } else
  switch (s0->key_bytes) {
    case 2: rc = launch_main_ns<uint16_t>(ctx, ns, d.sets, n, d_offsT, d_tiles, d.n_tiles, d.counter, d_W, d.stats, s0->key_bits, f, max_tiles); break;
    case 4: rc = launch_main_ns<uint32_t>(ctx, ns, d.sets, n, d_offsT, d_tiles, d.n_tiles, d.counter, d_W, d.stats, s0->key_bits, f, max_tiles); break;
    default: rc = launch_main_ns<unsigned long long>(ctx, ns, d.sets, n, d_offsT, d_tiles, d.n_tiles, d.counter, d_W, d.stats, s0->key_bits, f, max_tiles); break;
  }
  if (rc != KMSC_OK) return rc;
  KMSC_CUDA(cudaEventRecord(ctx->pc_ev[2], ctx->stream));
  // stats come back through pinned memory; the sync also protects the staging buffer
  unsigned long long* h_stats = (unsigned long long*)((unsigned char*)pin + sz_desc + sz_bitmap);
  KMSC_CUDA(cudaMemcpyAsync(h_stats, d.stats, 64, cudaMemcpyDeviceToHost, ctx->stream));
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
  if (host_stats[3]) {
    if (build == 2 && (h_stats[4] & 8ull) && !(h_stats[4] & 7ull)) return kWhashOverflow;  // a warp table filled up: the caller redoes the call
    set_error("pair_counts: %llu failures (tile could not be split further, or watchdog code %llu)", host_stats[3], h_stats[4]);
    return KMSC_E_STATE;
  }
  return KMSC_OK;
}

// W[map[a]][map[b]] = T[a][b]: places the matrix of a sub-run (m sets) into the n x n result
__global__ void place_block_kernel(const unsigned long long* __restrict__ T, int m, const int* __restrict__ map,
                                   unsigned long long* __restrict__ W, int n) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= m * m) return;
  const int a = t / m, b = t - a * m;
  W[(size_t)map[a] * n + map[b]] = T[t];
}

static int pair_counts_run_256(kmsc_ctx* ctx, const kmsc_set* const* sets, int32_t n,
                               const int32_t* bucket_ids, int32_t n_ids, unsigned long long* d_W, bool no_whash = false);

// Core: d_W (n*n u64, device) is zeroed and filled. Synchronous at return.
// One kernel pass holds the membership of up to 256 sets (the accumulator tile of the Gram).
// More sets are covered by groups of 128: every pair of groups (g, h) is one 256-set run that
// yields the blocks (g,g), (g,h), (h,g), (h,h); a block computed by several runs gets the same
// values every time.
int pair_counts_run(kmsc_ctx* ctx, const kmsc_set* const* sets, int32_t n,
                    const int32_t* bucket_ids, int32_t n_ids, unsigned long long* d_W) {
  if (!ctx || !sets || n < 1 || !d_W) { set_error("bad argument"); return KMSC_E_INVALID; }
  if (n <= 256) return pair_counts_run_256(ctx, sets, n, bucket_ids, n_ids, d_W);
  KMSC_CUDA(cudaSetDevice(ctx->device));
  const int G = 128, n_groups = (n + G - 1) / G;
  KMSC_TRY(ctx->work3.reserve((size_t)256 * 256 * 8 + 256 * sizeof(int)));
  unsigned long long* d_T = (unsigned long long*)ctx->work3.p;
  int* d_map = (int*)((unsigned char*)ctx->work3.p + (size_t)256 * 256 * 8);
  double main_ms = 0, plan_ms = 0, algo = 0;
  int launches = 0;
  unsigned long long st3[3] = {0, 0, 0};
  std::vector<const kmsc_set*> sub;
  std::vector<int> map;
  for (int g = 0; g < n_groups; g++)
    for (int h = g + 1; h < n_groups; h++) {
      sub.clear(); map.clear();
      for (int i = g * G; i < std::min(n, (g + 1) * G); i++) { sub.push_back(sets[i]); map.push_back(i); }
      for (int i = h * G; i < std::min(n, (h + 1) * G); i++) { sub.push_back(sets[i]); map.push_back(i); }
      const int m = (int)sub.size();
      KMSC_TRY(pair_counts_run_256(ctx, sub.data(), m, bucket_ids, n_ids, d_T));
      KMSC_CUDA(cudaMemcpyAsync(d_map, map.data(), (size_t)m * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
      place_block_kernel<<<(m * m + 255) / 256, 256, 0, ctx->stream>>>(d_T, m, d_map, d_W, n);
      count_launch(ctx);
      KMSC_CUDA(cudaGetLastError());
      KMSC_CUDA(cudaStreamSynchronize(ctx->stream));  // `map` is reused
      main_ms += ctx->pc_main_ms; plan_ms += ctx->pc_plan_ms; algo += ctx->pc_algo_bytes; launches += ctx->pc_main_launches;
      for (int q = 0; q < 3; q++) st3[q] += ctx->pc_last_stats[q];
    }
  ctx->pc_main_ms = main_ms; ctx->pc_plan_ms = plan_ms; ctx->pc_algo_bytes = algo; ctx->pc_main_launches = launches;
  for (int q = 0; q < 3; q++) ctx->pc_last_stats[q] = st3[q];
  return KMSC_OK;
}

static int pair_counts_run_256(kmsc_ctx* ctx, const kmsc_set* const* sets, int32_t n,
                               const int32_t* bucket_ids, int32_t n_ids, unsigned long long* d_W, bool no_whash) {
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
    for (int i = 0; i < n; i++) { h_sets[i].keys = sets[i]->keys; h_sets[i].lev = sets[i]->lev[0]; h_sets[i].lev_finest = sets[i]->lev[sets[i]->max_level]; }
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

  const int ns = n <= 64 ? 64 : n <= 128 ? 128 : 256;  // rows of the tensor-core Gram
  const int dmax = dmax_for(ns, s0->key_bytes == 8 ? 8 : 4);
  const unsigned long long L_cons = (unsigned long long)(0.75 * dmax);
  const unsigned long long L_max = 1ull << 20;
  // all sets, per bucket that can hold keys: the sets of a rank's prefix shard have all their
  // keys in [b_lo, b_hi) (8 shards = 8 x the density the whole bucket space would suggest; the
  // tile granularity and the table sizing below go by this)
  int span_lo = nb, span_hi = 0;
  for (int i = 0; i < n; i++) {
    const int lo_i = sets[i]->b_lo < 0 ? 0 : sets[i]->b_lo, hi_i = sets[i]->b_hi < 0 ? nb : sets[i]->b_hi;
    span_lo = std::min(span_lo, lo_i);
    span_hi = std::max(span_hi, hi_i);
  }
  const int nb_live = std::max(1, span_hi - span_lo);
  const double mean_bucket = (double)total_keys / (double)nb_live;
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
    KMSC_TRY(run_phase(ctx, sets, n, ns, 0, probe.data(), L_cons, mean_bucket, d_W, st));
    if (st[1] > 0) rho = (double)st[0] / (double)st[1];
    ctx->pc_last_stats[0] += st[0]; ctx->pc_last_stats[1] += st[1]; ctx->pc_last_stats[2] += st[2];
    phase2 = rest.data();
  }
  unsigned long long L = L_cons;
  if (rho > 1.0) {
    const unsigned long long La = (unsigned long long)(PC_LFRAC * rho * dmax);
    if (La > L) L = La;
  }
  // Build choice: the hash build costs ~3.7 warp instructions per key, the merge build ~51 per
  // distinct key at ns = 64 (two sets per lane) = 51 (ns / 64) / rho per key (ncu r01,
  // pair_counts_stream.cuh): break-even near rho = 14 ns / 64.
  // KMSC_P3_BUILD=hash|merge overrides (tests run both).
  bool merge_build = ns <= 128 && rho >= 14.0 * (ns / 64);
  // warp-private hash tables (pair_counts_whash.cuh): every lane inserts its own key, no shared
  // atomics on the hot path; tables sized by the distinct keys of a segment, so related sets only.
  // Measured r02u on C2: 6.3 ms against 3.07 ms for the merge build (a CTA-wide flush per 2 K distinct
  // keys), so it is never chosen by the library; KMSC_P3_BUILD=whash keeps it under the parity tests.
  bool whash_build = false;
  // lane-private tables (pair_counts_lane.cuh): one fine bucket (or a 1 / 2^q part of it) per lane, no bank
  // conflicts, no atomics, no compaction pass. Measured r02w on C2: 7.0 ms against 3.08 ms for the merge
  // build -- the 32 runs a warp walks per set differ in length (Poisson, ~10 keys), so 38 % of the lanes
  // work per iteration (ncu: 3.2e9 warp instructions, 43 % issue utilisation at 8 warps per SM). Never
  // chosen by the library; KMSC_P3_BUILD=lane keeps it under the parity tests.
  int lane_f = -1;
  bool lane_build = false;
  if (const char* e = getenv("KMSC_P3_BUILD")) {
    if (!strcmp(e, "hash")) { merge_build = false; whash_build = false; lane_build = false; }
    else if (!strcmp(e, "merge")) { merge_build = ns <= 128; whash_build = false; lane_build = false; }
    else if (!strcmp(e, "whash")) { whash_build = ns == 64 && !no_whash; lane_build = false; }
    else if (!strcmp(e, "lane")) {
      lane_build = false;
      if (ns == 64 && !no_whash) {
        lane_f = lane_level(s0, (double)total_keys, rho, nb_live, true);
        lane_build = lane_f >= 0;
      }
    }
  }
  if (lane_build) {
    ctx->pc_last_build = 3;
    const int rc = run_phase_lane(ctx, sets, n, lane_f, rho, phase2, d_W, st);
    if (rc == kLaneRedo) return pair_counts_run_256(ctx, sets, n, bucket_ids, n_ids, d_W, true);  // from scratch, without this build
    if (rc != KMSC_OK) return rc;
    if (st[1] > 0) {
      ctx->pc_rho = (double)st[0] / (double)st[1];
      ctx->pc_rho_n = n;
      ctx->pc_rho_k = s0->K;
    }
    ctx->pc_last_stats[0] += st[0]; ctx->pc_last_stats[1] += st[1]; ctx->pc_last_stats[2] += st[2];
    ctx->pc_last_L = (unsigned long long)lane_f;
    return KMSC_OK;
  }
  if (whash_build) {
    // tiles of several table fills of all warps; the kernel cuts them into segments
    const double r = rho > 1.0 ? rho : 1.0;
    double fills = 16.0;
    if (const char* e = getenv("KMSC_P3_FILLS")) { const double v = atof(e); if (v > 0.1 && v < 256) fills = v; }
    L = (unsigned long long)(fills * r * (WhCfg<64>::TS / 4) * (WhCfg<64>::T / 32));
    if (L < L_cons) L = L_cons;
  } else if (merge_build) {
    // no table to overflow (a full mask arena is flushed and the merge resumes): tiles of a few
    // arena fills keep the per-set slices long (less block over-read at their ends) and amortise
    // the end-of-tile imbalance between the warps
    const double r = rho > 1.0 ? rho : 1.0;
    double fills = 8.0;
    if (const char* e = getenv("KMSC_P3_FILLS")) { const double v = atof(e); if (v > 0.1 && v < 64) fills = v; }
    L = (unsigned long long)(fills * r * PmCfg<64>::SS);
    if (L < L_cons) L = L_cons;
  }
  if (L > L_max) L = L_max;
  const int build = whash_build ? 2 : merge_build ? 1 : 0;
  ctx->pc_last_build = build;
  {
    const int rho_q = std::max(1, (int)(0.75 * (rho > 1.0 ? rho : 1.0)));
    const int rc = run_phase(ctx, sets, n, ns, build, phase2, L, mean_bucket, d_W, st, rho_q);
    if (rc == kWhashOverflow) return pair_counts_run_256(ctx, sets, n, bucket_ids, n_ids, d_W, true);  // from scratch, without this build
    if (rc != KMSC_OK) return rc;
  }
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
  KMSC_TRY(pair_counts_run(ctx, sets, n, bucket_ids, n_ids, (unsigned long long*)d_out));
  // a rank of a prefix-sharded job: the partial matrices are summed here, one all-reduce
  return comm_allreduce_u64(ctx, (unsigned long long*)d_out, (size_t)n * n);
}

int kmsc_pair_counts_partial(kmsc_ctx* ctx, const kmsc_set* const* sets, int32_t n,
                             const int32_t* bucket_ids, int32_t n_ids, int64_t* d_out) {
  return pair_counts_run(ctx, sets, n, bucket_ids, n_ids, (unsigned long long*)d_out);
}

int kmsc_pair_counts(kmsc_ctx* ctx, const kmsc_set* const* sets, int32_t n,
                     const int32_t* bucket_ids, int32_t n_ids, int64_t* out, int64_t* key_visits) {
  if (!ctx || !out || n < 1) { set_error("bad argument"); return KMSC_E_INVALID; }
  KMSC_TRY(ctx->work.reserve((size_t)n * n * 8));
  unsigned long long* d_W = (unsigned long long*)ctx->work.p;
  KMSC_TRY(pair_counts_run(ctx, sets, n, bucket_ids, n_ids, d_W));
  KMSC_TRY(comm_allreduce_u64(ctx, d_W, (size_t)n * n));
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

int kmsc_pair_counts_build(kmsc_ctx* ctx) { return ctx ? ctx->pc_last_build : -1; }

}  // extern "C"
