// pair_counts_stream.cuh -- P3, second build: a warp-wide multiway MERGE instead of the hash table.
// Included by pair_counts.cu (shares SetDesc, Tile, set_of_pos and the tcgen05 Gram scheme).
//
// The hash build costs about 3.7 warp instructions per KEY (probe, claim, mark; ncu r01). When
// the sets are related (rho = keys per distinct key >> 1: one genome and its mutated copies) it
// is cheaper to pay per DISTINCT key: the runs of all sets are already sorted, so a warp holds
// the head key of NS / 32 sets per lane, takes the minimum over the warp with one
// redux.sync.min, and the ballots of "my head equals the minimum" ARE the membership mask of
// that distinct key, straight in the bit order the Gram expects. No table, no CAS, no compaction
// scan. Measured (ncu r01, C2, rho = 16.4): 51 warp instructions per distinct key = 3.1 per key
// against 3.7 for the hash build, 3.23 ms against 3.58 ms; the break-even is near rho = 14.
//
//   tile     as planned by plan_* (a fine-bucket range holding ~L keys over all sets)
//   segment  the tile's finest-level range cut into about four pieces per warp (never across a
//            bucket: keys ascend inside a bucket only; a bucket's pieces follow its share of the
//            tile's keys), handed out by a counter; the piece of every set is one contiguous slice
//   ring     per (lane, set) a 16-key window in shared memory, filled by cp.async in 8-key
//            blocks (global -> shared without registers). Every 4 iterations all lanes top their
//            rings up in uniform code; the blocks land while the keys before them are merged
//   masks    written to a per-warp arena of slots; when a warp's arena is full the CTA flushes
//            (Gram of everything collected so far) and the merge resumes from its registers
//   gram     unchanged: 32 masks per K-step expanded to 0/1 bytes, tcgen05.mma kind::i8 into TMEM
//
// Used for n <= 128 sets when the measured rho makes it cheaper than the hash build
// (pair_counts_run_256); results are identical (tests/test_gpu_pair_counts.py runs both).
#pragma once

namespace kmsc {

template <int NS> struct PmCfg;
template <> struct PmCfg<64>  { static constexpr int T = 256, SS = 2048, CK = 4, MINB = 3; };
template <> struct PmCfg<128> { static constexpr int T = 256, SS = 2048, CK = 4, MINB = 1; };

template <typename KeyT, int NS>
struct PmLayout {
  using C = PmCfg<NS>;
  static constexpr int MW = NS / 32, NWARP = C::T / 32;
  static constexpr size_t al(size_t x, size_t a) { return (x + a - 1) & ~(a - 1); }
  static constexpr size_t kstep_bytes = (size_t)NS * 32;
  static constexpr size_t buf_bytes = kstep_bytes * C::CK;
  static constexpr size_t slot_bytes = 16 * sizeof(KeyT);               // ring window of one (lane, set)
  static constexpr size_t o_stage = 0;
  static constexpr size_t o_mask = al(o_stage + 2 * buf_bytes, 16);
  static constexpr size_t o_list = al(o_mask + (size_t)(C::SS + 1) * MW * 4, 16);
  static constexpr size_t o_ring = al(o_list + (size_t)(C::SS + 2) * 2, 128);
  static constexpr size_t o_kp = o_ring + (size_t)NWARP * NS * slot_bytes;
  static constexpr size_t o_lev = o_kp + (size_t)NS * 8;
  static constexpr size_t o_sbeg = o_lev + (size_t)NS * 8;
  static constexpr size_t o_send = o_sbeg + (size_t)NS * 4 * 2;
  static constexpr size_t o_misc = o_send + (size_t)NS * 4 * 2;
  static constexpr size_t o_wcnt = o_misc + 16 * 4;
  static constexpr size_t o_bar = al(o_wcnt + (size_t)NWARP * 4, 8);
  static constexpr size_t total = al(o_bar + 2 * 8, 16);
};

__device__ __forceinline__ void cp_async16(uint32_t smem_dst, const void* gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(smem_dst), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// warp-wide minimum of one key per lane
__device__ __forceinline__ uint32_t warp_min_key(uint32_t v) { return __reduce_min_sync(0xffffffffu, v); }
__device__ __forceinline__ unsigned long long warp_min_key(unsigned long long v) {
  const uint32_t hi = __reduce_min_sync(0xffffffffu, (uint32_t)(v >> 32));
  const uint32_t lo = __reduce_min_sync(0xffffffffu, (uint32_t)(v >> 32) == hi ? (uint32_t)v : 0xffffffffu);
  return ((unsigned long long)hi << 32) | lo;
}

template <typename KeyT, int NS>
__global__ void __launch_bounds__(PmCfg<NS>::T, PmCfg<NS>::MINB)
pair_counts_stream_kernel(const SetDesc* __restrict__ sets, int n_sets, int spw, const uint32_t* __restrict__ offsT,
                          const Tile* __restrict__ tiles, const uint32_t* __restrict__ n_tiles_p,
                          uint32_t* __restrict__ tile_counter, unsigned long long* __restrict__ W,
                          unsigned long long* __restrict__ stats, int fine_level, int finest_level, int l2_prefetch) {
  using C = PmCfg<NS>;
  using LY = PmLayout<KeyT, NS>;
  using CmpT = typename TableKey<KeyT>::type;  // uint32 for 2/4-byte keys, uint64 for 8-byte keys
  constexpr int SS = C::SS, T = C::T, CK = C::CK;
  constexpr int MW = NS / 32;   // mask words per distinct key = sets per lane
  constexpr int NW = T / 32;
  constexpr int DW = SS / NW;   // mask slots per warp between two flushes
  static_assert(DW % 4 == 0, "the arena-full test rides on the ring top-up");
  constexpr int MMA_M = NS == 64 ? 64 : 128;
  constexpr int MMA_N = NS;
  constexpr int TMEM_COLS = NS;
  constexpr uint32_t SLOTB = (uint32_t)LY::slot_bytes;
  constexpr uint32_t NCH = SLOTB / 16;                    // 16-byte chunks per ring slot
  constexpr uint32_t BLKB = 8 * sizeof(KeyT);             // bytes of one 8-key block
  const CmpT KMAX = (CmpT)~(CmpT)0;

  extern __shared__ __align__(128) unsigned char smem_raw[];
  unsigned char* stage = smem_raw + LY::o_stage;
  uint32_t* smask = reinterpret_cast<uint32_t*>(smem_raw + LY::o_mask);
  uint16_t* list = reinterpret_cast<uint16_t*>(smem_raw + LY::o_list);
  const void** skp = reinterpret_cast<const void**>(smem_raw + LY::o_kp);
  const uint32_t** slev = reinterpret_cast<const uint32_t**>(smem_raw + LY::o_lev);
  uint32_t* sbeg = reinterpret_cast<uint32_t*>(smem_raw + LY::o_sbeg);
  uint32_t* send = reinterpret_cast<uint32_t*>(smem_raw + LY::o_send);
  int* misc = reinterpret_cast<int*>(smem_raw + LY::o_misc);
  uint32_t* wcnt = reinterpret_cast<uint32_t*>(smem_raw + LY::o_wcnt);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw + LY::o_bar);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  __shared__ uint32_t s_segstart[260];  // first segment of every bucket of the tile (a tile spans <= 256 buckets)

  for (int i = tid; i < (SS + 1) * MW; i += T) smask[i] = 0;
  for (int i = tid; i < NS; i += T) {
    skp[i] = i < n_sets ? sets[i].keys : nullptr;
    slev[i] = i < n_sets ? sets[i].lev_finest : nullptr;
  }
  if (tid == 0) {
    misc[kMiscAnyMma] = 0;
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
  const uint32_t smask_addr = umma::smem_u32(smask);
  constexpr uint32_t idesc = umma::make_idesc_u8(MMA_M, MMA_N, 1, 1);
  uint32_t uses0 = 0, uses1 = 0;
  bool mma_started = false;
  unsigned long long st_keys = 0, st_dist = 0, st_flush = 0;

  // the sets this lane merges: Gram position p = 32 j + lane (word j, bit lane of the mask)
  int sidx[MW];
  uint32_t ring_addr[MW];   // shared address of the (lane, set) ring slot
  uint32_t swz[MW];
#pragma unroll
  for (int j = 0; j < MW; j++) {
    const int p = 32 * j + lane;
    sidx[j] = set_of_pos(p, spw, n_sets);
    const uint32_t slot = (uint32_t)p;
    ring_addr[j] = umma::smem_u32(smem_raw + LY::o_ring) + ((uint32_t)warp * NS + slot) * SLOTB;
    swz[j] = ((slot >> 1) & (NCH - 1)) << 4;
  }
  // byte offset of key index g inside its ring slot (16-byte chunks swizzled by slot: swz = chunk xor << 4)
  auto ring_off = [&](uint32_t g, uint32_t sw) -> uint32_t {
    return ((g & 15u) * (uint32_t)sizeof(KeyT)) ^ sw;
  };
  auto ring_read = [&](int j, uint32_t g) -> CmpT {
    KeyT v;
    const uint32_t a = ring_addr[j] + ring_off(g, swz[j]);
    if (sizeof(KeyT) == 2) { unsigned short t; asm volatile("ld.shared.u16 %0, [%1];" : "=h"(t) : "r"(a)); v = (KeyT)t; }
    else if (sizeof(KeyT) == 4) { uint32_t t; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(t) : "r"(a)); v = (KeyT)t; }
    else { unsigned long long t; asm volatile("ld.shared.u64 %0, [%1];" : "=l"(t) : "r"(a)); v = (KeyT)t; }
    return (CmpT)v;
  };
  // issue the loads of the 8-key block starting at key index gb (a multiple of 8)
  auto ring_fetch = [&](int j, const char* kp, uint32_t gb) {
    const uint32_t o = (gb & 15u) * (uint32_t)sizeof(KeyT);  // chunk-aligned: low 4 bits are 0
#pragma unroll
    for (uint32_t q = 0; q < BLKB / 16; q++)
      cp_async16(ring_addr[j] + ((o + 16 * q) ^ swz[j]), kp + (size_t)gb * sizeof(KeyT) + 16 * q);
  };

  const uint32_t n_tiles = *n_tiles_p;
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
  int par = 0;
  const int up = finest_level - fine_level;  // tile coordinates -> finest-level coordinates
  for (;;) {
    __syncthreads();
    const uint32_t t_id = (uint32_t)misc[kMiscTile];
    const uint32_t t_nx = (uint32_t)misc[kMiscNext];
    if (t_id >= n_tiles) break;
    const Tile tl = tiles[t_id];
    const uint32_t* sb = sbeg + par * NS;
    const uint32_t* se = send + par * NS;
    // next tile: key ranges -> shared memory, keys -> L2 (one 128-byte line per request)
    if (t_nx < n_tiles && tid < NS) {
      uint32_t nx_b = 0, nx_e = 0;
      if (tid < n_sets) {
        const Tile tn = tiles[t_nx];
        nx_b = offsT[(size_t)tn.x0 * n_sets + tid];
        nx_e = offsT[(size_t)tn.x1 * n_sets + tid];
        if (l2_prefetch) {
          const char* base = (const char*)skp[tid];
          const size_t lo = ((size_t)nx_b * sizeof(KeyT)) & ~(size_t)127, hi = (size_t)nx_e * sizeof(KeyT);
          for (size_t off = lo; off < hi; off += 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(base + off));
        }
      }
      sbeg[(par ^ 1) * NS + tid] = nx_b;
      send[(par ^ 1) * NS + tid] = nx_e;
    }
    // segments of the tile at the finest level: one bucket -> NW pieces; several buckets -> one each
    const uint32_t X0 = tl.x0 << up, X1 = tl.x1 << up;
    const uint32_t bucket0 = X0 >> finest_level;
    const uint32_t nbk = ((X1 - 1) >> finest_level) - bucket0 + 1;
    // The tile is cut into about four segments per warp, handed out by a counter. One bucket:
    // P equal pieces. Several buckets: bucket b gets pieces in proportion to its share of the
    // tile's keys in set 0 (the sets of one job are related, any of them shows where the keys
    // are) -- a tile may hold one heavy bucket next to many empty ones (the edge of a rank's
    // prefix range), and a bucket must not be left to a single warp.
    uint32_t P = 1, n_seg;
    if (tid == 0) misc[kMiscSeg] = 0;
    if (nbk == 1) {
      while (P < 4u * NW && P < (1u << finest_level)) P <<= 1;
      n_seg = P;
      __syncthreads();
    } else {
      if (warp == 0) {
        const uint32_t* lv0 = slev[0];
        const unsigned long long t0 = (unsigned long long)(lv0[X1] - lv0[X0]);
        uint32_t carry = 0;
        for (uint32_t base = 0; base < nbk; base += 32) {
          const uint32_t b = base + lane;
          uint32_t pb = 0;
          if (b < nbk) {
            const uint32_t B0 = max(X0, (bucket0 + b) << finest_level), B1 = min(X1, (bucket0 + b + 1) << finest_level);
            const unsigned long long k = (unsigned long long)(lv0[B1] - lv0[B0]);
            pb = 1;
            if (t0 > 0) pb = (uint32_t)min((unsigned long long)(1u << finest_level), max(1ull, (k * 4ull * NW + t0 - 1) / t0));
          }
          uint32_t inc = pb;
#pragma unroll
          for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
          }
          if (b < nbk) s_segstart[b + 1] = carry + inc;
          carry += __shfl_sync(0xffffffffu, inc, 31);
        }
        if (lane == 0) s_segstart[0] = 0;
      }
      __syncthreads();
      n_seg = s_segstart[nbk];
    }
    auto take_seg = [&]() -> uint32_t {
      uint32_t q = 0;
      if (lane == 0) q = (uint32_t)atomicAdd(&misc[kMiscSeg], 1);
      return __shfl_sync(0xffffffffu, q, 0);
    };
    uint32_t seg_next = take_seg();
    bool have_seg = false;
    // merge state of this warp's current segment (kept across flushes)
    uint32_t ci[MW], ce[MW];   // next key index / end of the slice, per set of this lane
    uint32_t clim[MW];         // keys below clim are in the ring or on their way (a multiple of 8)
    const char* cptr[MW];      // address of key clim in the set's key array
    CmpT ck[MW];               // head key; KMAX once the slice is exhausted
    uint32_t cnt = 0;          // masks in this warp's arena
    for (uint32_t round = 0;; round++) {
      bool full = false;
      while (!full) {
        if (!have_seg) {
          if (seg_next >= n_seg) break;
          uint32_t bk = bucket0, piece = seg_next, Pb = P;
          if (nbk > 1) {  // the bucket whose piece range holds this segment
            uint32_t lo = 0, hi = nbk;
            while (hi - lo > 1) {
              const uint32_t mid = (lo + hi) >> 1;
              if (s_segstart[mid] <= seg_next) lo = mid; else hi = mid;
            }
            bk = bucket0 + lo;
            piece = seg_next - s_segstart[lo];
            Pb = s_segstart[lo + 1] - s_segstart[lo];
          }
          const uint32_t B0 = max(X0, bk << finest_level), B1 = min(X1, (bk + 1) << finest_level);
          const uint32_t xa = B0 + (uint32_t)(((unsigned long long)(B1 - B0) * piece) / Pb);
          const uint32_t xb = B0 + (uint32_t)(((unsigned long long)(B1 - B0) * (piece + 1)) / Pb);
          seg_next = take_seg();
          if (xa >= xb) continue;
          cp_async_wait<0>();  // nothing of the previous segment may still land in the rings
#pragma unroll
          for (int j = 0; j < MW; j++) {
            ci[j] = 0; ce[j] = 0; ck[j] = KMAX; clim[j] = 0xffffffffu; cptr[j] = nullptr;  // (an empty slice never asks for keys)
            if (sidx[j] >= 0) {
              const uint32_t* lv = slev[sidx[j]];
              ci[j] = lv[xa];
              ce[j] = lv[xb];
              if (ci[j] < ce[j]) {
                const char* kp = (const char*)skp[sidx[j]];
                const uint32_t g0 = ci[j] & ~7u;
                ring_fetch(j, kp, g0);
                if (g0 + 8 < ce[j]) ring_fetch(j, kp, g0 + 8);
                clim[j] = g0 + 16;
                cptr[j] = kp + (size_t)(g0 + 16) * sizeof(KeyT);
              }
            }
          }
          cp_async_commit();
          cp_async_wait<0>();
#pragma unroll
          for (int j = 0; j < MW; j++)
            if (ci[j] < ce[j]) ck[j] = ring_read(j, ci[j]);
          have_seg = true;
        }
        // ---- the merge: one distinct key per iteration ------------------------------------------
        // Exhausted slices hold KMAX, so the warp minimum is the next distinct key; only when it
        // IS KMAX (the end, or a real all-ones key) the lanes look at their indices. Every 4
        // iterations all lanes top their rings up in uniform code (no per-advance refill branch).
        // A lane advances at most once per iteration; after a top-up its ring reaches >= 9 keys
        // ahead, of which >= 5 were asked for at an earlier top-up and have landed (wait_group 1),
        // so the block asked for now is not read before the next top-up has waited for it.
        for (;;) {
          if ((cnt & 3u) == 0u) {
            if (cnt == (uint32_t)DW) { full = true; break; }  // (DW is a multiple of 4)
#pragma unroll
            for (int j = 0; j < MW; j++) {
              const uint32_t want = (ci[j] & ~7u) + 16u;  // the ring covers [block of ci, +16)
              if (clim[j] < want && clim[j] < ce[j]) {
                const uint32_t o = (clim[j] & 8u) * (uint32_t)sizeof(KeyT);  // first or second half of the slot
#pragma unroll
                for (uint32_t q = 0; q < BLKB / 16; q++) cp_async16(ring_addr[j] + ((o + 16 * q) ^ swz[j]), cptr[j] + 16 * q);
                clim[j] += 8;
                cptr[j] += BLKB;
              }
            }
            cp_async_commit();
            cp_async_wait<1>();  // the blocks asked for at the previous top-up have landed
          }
          CmpT v = ck[0];
#pragma unroll
          for (int j = 1; j < MW; j++) v = ck[j] < v ? ck[j] : v;
          const CmpT g = warp_min_key(v);
          bool live[MW];
#pragma unroll
          for (int j = 0; j < MW; j++) live[j] = true;
          if (g == KMAX) {
            bool act = false;
#pragma unroll
            for (int j = 0; j < MW; j++) { live[j] = ci[j] < ce[j]; act |= live[j]; }
            if (!__any_sync(0xffffffffu, act)) { have_seg = false; break; }
          }
          const uint32_t mslot = smask_addr + ((uint32_t)warp * DW + cnt) * 4u;
#pragma unroll
          for (int j = 0; j < MW; j++) {
            const bool hit = live[j] && ck[j] == g;
            const uint32_t bal = __ballot_sync(0xffffffffu, hit);
            if (lane == 0) asm volatile("st.shared.u32 [%0], %1;" ::"r"(mslot + (uint32_t)j * (uint32_t)(SS + 1) * 4u), "r"(bal) : "memory");
            if (hit) {
              const uint32_t i = ++ci[j];
              const CmpT nk = ring_read(j, i);
              ck[j] = i < ce[j] ? nk : KMAX;
            }
          }
          cnt++;
        }
      }
      // ---- flush: Gram of the masks collected so far -------------------------------------------
      if (lane == 0) wcnt[warp] = cnt;
      __syncthreads();
      uint32_t pw = 0, D = 0;
#pragma unroll
      for (int w = 0; w < NW; w++) {
        const uint32_t c = wcnt[w];
        if (w < warp) pw += c;
        D += c;
      }
      for (uint32_t i = lane; i < cnt; i += 32) list[pw + i] = (uint16_t)(warp * DW + i);
      const int more = __syncthreads_or((have_seg || seg_next < n_seg) ? 1 : 0);
      cnt = 0;
      st_dist += (tid == 0) ? (unsigned long long)D : 0ull;
      if (round > 0) st_flush += (tid == 0) ? 1ull : 0ull;
      for (int c0 = 0; c0 < (int)D; c0 += CK * 32) {
        const int nk = min(CK, ((int)D - c0 + 31) >> 5);
        const uint32_t nuse = uses0 + uses1;
        const int buf = (int)(nuse & 1u);
        const uint32_t used = buf ? uses1 : uses0;
        if (used > 0 && !umma::mbar_wait_bounded(&bar[buf], (used - 1) & 1u)) {
          if (lane == 0) { atomicAdd(&stats[3], 1ull); atomicOr(&stats[4], 1ull); }
        }
        unsigned char* sbuf = stage + (size_t)buf * LY::buf_bytes;
        for (int item = warp; item < nk * MW; item += NW) {
          const int ks = item / MW, w = item - ks * MW;
          const int r = c0 + ks * 32 + lane;
          uint32_t m = 0;
          if (r < (int)D) {
            const uint32_t slot = list[r];
            m = smask[w * (SS + 1) + slot];
            smask[w * (SS + 1) + slot] = 0;
          }
          uint4 lo, hi;
          lo.x = umma::nibble_to_bytes(m);       lo.y = umma::nibble_to_bytes(m >> 4);
          lo.z = umma::nibble_to_bytes(m >> 8);  lo.w = umma::nibble_to_bytes(m >> 12);
          hi.x = umma::nibble_to_bytes(m >> 16); hi.y = umma::nibble_to_bytes(m >> 20);
          hi.z = umma::nibble_to_bytes(m >> 24); hi.w = umma::nibble_to_bytes(m >> 28);
          unsigned char* kb = sbuf + (size_t)ks * LY::kstep_bytes + (size_t)lane * 16;
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
            mma_started = true;
          }
          umma::mma_commit(&bar[buf]);
          misc[kMiscAnyMma] = 1;
        }
        if (buf) uses1++; else uses0++;
      }
      if (!more) break;
      if (round > (1u << 24)) {  // watchdog
        if (tid == 0) { atomicAdd(&stats[3], 1ull); atomicOr(&stats[4], 4ull); }
        break;
      }
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
    // M = 64: row m lives in TMEM lane (m % 16) + 32 (m / 16); M = 128: row m in lane m
    int row;
    if (NS == 64) row = lane < 16 ? warp * 16 + lane : -1;
    else row = warp * 32 + lane;
    const int si = row >= 0 ? set_of_pos(row, spw, n_sets) : -1;
#pragma unroll 1
    for (int c0 = 0; c0 < MMA_N; c0 += 32) {
      uint32_t v[32];
      umma::tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
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
  umma::fence_before_thread_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tmem_base, TMEM_COLS);

  for (int o = 16; o > 0; o >>= 1) st_keys += __shfl_xor_sync(0xffffffffu, st_keys, o);
  if (lane == 0 && st_keys) atomicAdd(&stats[0], st_keys);
  if (tid == 0) {
    if (st_dist) atomicAdd(&stats[1], st_dist);
    if (st_flush) atomicAdd(&stats[2], st_flush);
  }
}

}  // namespace kmsc
