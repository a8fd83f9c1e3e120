// pair_counts_whash.cuh -- P3, third build: WARP-PRIVATE hash tables, no shared atomics on the hot path.
// Included by pair_counts.cu (shares SetDesc, Tile, the planning and the tcgen05 Gram scheme).
//
// What the two older builds pay (ncu, C2): the CTA-wide hash build 3.7 warp instructions per key (a
// fixed cost per (set, tile) slice of ~90 keys, byte planes, overflow classes), the warp-wide merge 51
// per distinct key = 3.2 per key (a merge step advances ~16 of the 64 set heads a warp holds: a quarter
// of the lanes work). Here every lane works on its own key all the time:
//
//   segment  a piece of ONE bucket (a finest-level fine-bucket range), handed out by a counter as in the
//            merge build; the slice of every set in it is contiguous
//   table    per warp: TS key slots in shared memory. The slot index IS the index into the warp's part
//            of the mask arena, so there is no value array and no compaction of values
//   insert   a warp is four groups of 8 lanes; in round r group g takes set 4 r + g: its 8 lanes read 8
//            consecutive keys of that set's slice (one 32-byte sector) and probe the table with a plain
//            LDS; a new key claims its slot with one CAS (1 key in rho). Membership is a plain 16-bit
//            read-modify-write of the group's OWN quarter of the slot's 64-bit mask: the keys of one set
//            are distinct, so the lanes of a group never meet in a slot, and the four groups own
//            different 16-bit planes -- no atomics. Gram position of set 4 r + g = 16 g + r.
//   flush    when a warp's table is about half full, or its next segment lies in another bucket (equal
//            keys of different buckets are different k-mers), the CTA flushes: occupied slots are listed
//            (ballot scan), their masks go through the same tcgen05 Gram as the other builds, keys and
//            masks are wiped on the way.
// Chosen for n <= 64 related sets (rho >= 4: the table is sized by the distinct keys of a segment);
// a table that fills up raises a flag and the call is redone with the merge / hash build.
#pragma once

namespace kmsc {

template <int NS> struct WhCfg;
template <> struct WhCfg<64> { static constexpr int T = 256, TS = 512, LOG2TS = 9, CK = 3, MINB = 3; };  // CK = 3: 72 KB, three CTAs per SM

template <typename KeyT, int NS>
struct WhLayout {
  using C = WhCfg<NS>;
  using TK = typename TableKey<KeyT>::type;
  static constexpr int MW = NS / 32, NWARP = C::T / 32;
  static constexpr int DW = C::TS + 8;            // per-warp arena stride; slot TS = the all-ones key
  static constexpr int SS = NWARP * DW;
  static constexpr size_t al(size_t x, size_t a) { return (x + a - 1) & ~(a - 1); }
  static constexpr size_t kstep_bytes = (size_t)NS * 32;
  static constexpr size_t buf_bytes = kstep_bytes * C::CK;
  static constexpr size_t o_stage = 0;
  static constexpr size_t o_mask = al(o_stage + 2 * buf_bytes, 16);
  static constexpr size_t o_tkeys = al(o_mask + (size_t)(SS + 1) * MW * 4, 16);
  static constexpr size_t o_list = al(o_tkeys + (size_t)NWARP * C::TS * sizeof(TK), 16);
  static constexpr size_t o_kp = al(o_list + (size_t)(SS + 2) * 2, 16);
  static constexpr size_t o_lev = o_kp + (size_t)NS * 8;
  static constexpr size_t o_sbeg = o_lev + (size_t)NS * 8;
  static constexpr size_t o_send = o_sbeg + (size_t)NS * 4 * 2;
  static constexpr size_t o_misc = o_send + (size_t)NS * 4 * 2;
  static constexpr size_t o_wcnt = o_misc + 16 * 4;
  static constexpr size_t o_bar = al(o_wcnt + (size_t)NWARP * 4, 8);
  static constexpr size_t total = al(o_bar + 2 * 8, 16);
};

template <typename KeyT, int NS>
__global__ void __launch_bounds__(WhCfg<NS>::T, WhCfg<NS>::MINB)
pair_counts_whash_kernel(const SetDesc* __restrict__ sets, int n_sets, const uint32_t* __restrict__ offsT,
                         const Tile* __restrict__ tiles, const uint32_t* __restrict__ n_tiles_p,
                         uint32_t* __restrict__ tile_counter, unsigned long long* __restrict__ W,
                         unsigned long long* __restrict__ stats, int fine_level, int finest_level, int rho_q) {
  using C = WhCfg<NS>;
  using LY = WhLayout<KeyT, NS>;
  using TK = typename LY::TK;
  static_assert(NS == 64, "16-bit membership planes: four groups x 16 sets");
  constexpr int T = C::T, CK = C::CK, TS = C::TS, DW = LY::DW, SS = LY::SS;
  constexpr int MW = NS / 32, NW = T / 32;
  constexpr int SPG = NS / 4;   // sets per 8-lane group = rounds per segment
  constexpr int MMA_M = 64, MMA_N = NS, TMEM_COLS = NS;
  const TK EMPTY = (TK)~(TK)0;

  extern __shared__ __align__(128) unsigned char smem_raw[];
  unsigned char* stage = smem_raw + LY::o_stage;
  uint32_t* smask = reinterpret_cast<uint32_t*>(smem_raw + LY::o_mask);
  uint16_t* smask16 = reinterpret_cast<uint16_t*>(smem_raw + LY::o_mask);
  TK* tkeys = reinterpret_cast<TK*>(smem_raw + LY::o_tkeys);
  uint16_t* list = reinterpret_cast<uint16_t*>(smem_raw + LY::o_list);
  const void** skp = reinterpret_cast<const void**>(smem_raw + LY::o_kp);
  const uint32_t** slev = reinterpret_cast<const uint32_t**>(smem_raw + LY::o_lev);
  uint32_t* sbeg = reinterpret_cast<uint32_t*>(smem_raw + LY::o_sbeg);
  uint32_t* send = reinterpret_cast<uint32_t*>(smem_raw + LY::o_send);
  int* misc = reinterpret_cast<int*>(smem_raw + LY::o_misc);
  uint32_t* wcnt = reinterpret_cast<uint32_t*>(smem_raw + LY::o_wcnt);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw + LY::o_bar);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int grp = lane >> 3, sub = lane & 7;
  __shared__ uint32_t s_segstart[260];
  __shared__ uint32_t s_tk[2];

  for (int i = tid; i < (SS + 1) * MW; i += T) smask[i] = 0;
  for (int i = tid; i < NW * TS; i += T) tkeys[i] = EMPTY;
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
  constexpr uint32_t idesc = umma::make_idesc_u8(MMA_M, MMA_N, 1, 1);
  uint32_t uses0 = 0, uses1 = 0;
  bool mma_started = false;
  unsigned long long st_keys = 0, st_dist = 0, st_flush = 0;

  TK* const tk = tkeys + (size_t)warp * TS;                 // this warp's table
  const uint32_t arena0 = (uint32_t)warp * DW;              // ... and its part of the mask arena
  // this group's 16-bit plane of a slot's mask: Gram positions [16 grp, 16 grp + 16) = half (grp & 1) of word (grp >> 1)
  uint16_t* const plane = smask16 + ((size_t)(grp >> 1) * (SS + 1) + arena0) * 2 + (grp & 1);

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
  const int up = finest_level - fine_level;
  uint32_t cnt = 0;        // slots claimed in this warp's table since the last flush (warp-uniform)
  int tbucket = -1;        // the bucket its keys belong to
  for (;;) {
    __syncthreads();
    const uint32_t t_id = (uint32_t)misc[kMiscTile];
    const uint32_t t_nx = (uint32_t)misc[kMiscNext];
    if (t_id >= n_tiles) break;
    const Tile tl = tiles[t_id];
    const uint32_t* sb = sbeg + par * NS;
    const uint32_t* se = send + par * NS;
    if (t_nx < n_tiles && tid < NS) {
      uint32_t nx_b = 0, nx_e = 0;
      if (tid < n_sets) {
        const Tile tn = tiles[t_nx];
        nx_b = offsT[(size_t)tn.x0 * n_sets + tid];
        nx_e = offsT[(size_t)tn.x1 * n_sets + tid];
      }
      sbeg[(par ^ 1) * NS + tid] = nx_b;
      send[(par ^ 1) * NS + tid] = nx_e;
    }
    // segments: as in the merge build, but finer (a segment's distinct keys must fit the table)
    const uint32_t X0 = tl.x0 << up, X1 = tl.x1 << up;
    const uint32_t bucket0 = X0 >> finest_level;
    const uint32_t nbk = ((X1 - 1) >> finest_level) - bucket0 + 1;
    // keys of the tile over all sets / rho_q = distinct keys expected; pieces of <= TS / 4 of them
    unsigned long long tile_keys = 0;
    if (tid < NS) tile_keys = se[tid] - sb[tid];
    {
      for (int o = 16; o > 0; o >>= 1) tile_keys += __shfl_xor_sync(0xffffffffu, tile_keys, o);
      if (lane == 0 && warp < 2) s_tk[warp] = (uint32_t)tile_keys;
    }
    uint32_t P = 1, n_seg;
    if (tid == 0) misc[kMiscSeg] = 0;
    __syncthreads();
    const unsigned long long tkeys_all = (unsigned long long)s_tk[0] + s_tk[1];
    const uint32_t want_pieces = (uint32_t)min(4096ull, 1ull + tkeys_all / (unsigned long long)(max(1, rho_q) * (TS / 4)));
    if (nbk == 1) {
      while (P < want_pieces && P < (1u << finest_level)) P <<= 1;
      if (P < (uint32_t)NW) P = min((uint32_t)NW, 1u << finest_level);
      n_seg = P;
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
            if (t0 > 0) pb = (uint32_t)min((unsigned long long)(1u << finest_level), max(1ull, (k * (unsigned long long)max(want_pieces, (uint32_t)NW) + t0 - 1) / t0));
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
    for (uint32_t round = 0;; round++) {
      bool full = false;
      while (!full && seg_next < n_seg) {
        uint32_t bk = bucket0, piece = seg_next, Pb = P;
        if (nbk > 1) {
          uint32_t lo = 0, hi = nbk;
          while (hi - lo > 1) {
            const uint32_t mid = (lo + hi) >> 1;
            if (s_segstart[mid] <= seg_next) lo = mid; else hi = mid;
          }
          bk = bucket0 + lo;
          piece = seg_next - s_segstart[lo];
          Pb = s_segstart[lo + 1] - s_segstart[lo];
        }
        // a table holds the keys of one bucket only, and is flushed before it is half full
        if (cnt > 0 && ((int)bk != tbucket || cnt > (uint32_t)(TS * 3 / 8))) { full = true; break; }
        const uint32_t B0 = max(X0, bk << finest_level), B1 = min(X1, (bk + 1) << finest_level);
        const uint32_t xa = B0 + (uint32_t)(((unsigned long long)(B1 - B0) * piece) / Pb);
        const uint32_t xb = B0 + (uint32_t)(((unsigned long long)(B1 - B0) * (piece + 1)) / Pb);
        seg_next = take_seg();
        if (xa >= xb) continue;
        tbucket = (int)bk;
        // slice of set p (= Gram position p) in the segment: lane l keeps sets l and l + 32
        uint32_t ci[MW], ce[MW];
#pragma unroll
        for (int j = 0; j < MW; j++) {
          const int p = 32 * j + lane;
          ci[j] = 0; ce[j] = 0;
          if (p < n_sets) { const uint32_t* lv = slev[p]; ci[j] = lv[xa]; ce[j] = lv[xb]; }
        }
        uint32_t n_new = 0;
        bool overflow = false;
#pragma unroll 1
        for (int r = 0; r < SPG; r++) {
          const int s = 4 * r + grp;                 // this group's set in round r; Gram position 16 grp + r
          const uint32_t a0 = __shfl_sync(0xffffffffu, ci[0], s & 31), a1 = __shfl_sync(0xffffffffu, ci[1], s & 31);
          const uint32_t e0 = __shfl_sync(0xffffffffu, ce[0], s & 31), e1 = __shfl_sync(0xffffffffu, ce[1], s & 31);
          const uint32_t a = s < 32 ? a0 : a1, e = s < 32 ? e0 : e1;
          const KeyT* kp = (const KeyT*)skp[s];
          const uint16_t bit = (uint16_t)(1u << r);
          for (uint32_t i0 = a + sub; i0 < e; i0 += 32) {
            KeyT kk[4];
#pragma unroll
            for (int u = 0; u < 4; u++) { const uint32_t i = i0 + 8 * u; kk[u] = i < e ? kp[i] : (KeyT)0; }
#pragma unroll
            for (int u = 0; u < 4; u++) {
              if (i0 + 8 * u >= e) break;
              const TK key = (TK)kk[u];
              uint32_t h;
              if (sizeof(TK) == 4 && sizeof(KeyT) == 4 && key == EMPTY) {
                h = (uint32_t)TS;                    // the one key that looks like the empty marker: its own slot
              } else {
                h = hash_slot(key, C::LOG2TS);
                uint32_t probes = 0;
                for (;;) {
                  const TK c = tk[h];
                  if (c == key) break;
                  if (c == EMPTY) {
                    const TK old = atomicCAS(&tk[h], EMPTY, key);
                    if (old == EMPTY) { n_new++; break; }
                    if (old == key) break;
                  }
                  h = (h + 1) & (uint32_t)(TS - 1);
                  if (++probes > (uint32_t)TS) { overflow = true; break; }
                }
              }
              uint16_t* m = plane + (size_t)h * 2;
              *m = (uint16_t)(*m | bit);
            }
          }
          __syncwarp();   // no lane may run ahead into the next set: a group's plane has one writer per slot at a time
        }
        cnt += __reduce_add_sync(0xffffffffu, n_new);
        if (__any_sync(0xffffffffu, overflow) && lane == 0) { atomicAdd(&stats[3], 1ull); atomicOr(&stats[4], 8ull); }
      }
      // ---- flush: list the occupied slots, Gram of their masks, wipe keys and masks ---------------
      uint32_t occ_n = 0;
      for (uint32_t i0 = 0; i0 < (uint32_t)DW; i0 += 32) {   // uniform trip count: ballots inside
        const uint32_t i = i0 + lane;
        uint32_t m = 0;
        if (i < (uint32_t)DW) {
#pragma unroll
          for (int w = 0; w < MW; w++) m |= smask[w * (SS + 1) + arena0 + i];
        }
        occ_n += __popc(__ballot_sync(0xffffffffu, m != 0));
      }
      if (lane == 0) wcnt[warp] = occ_n;
      __syncthreads();
      uint32_t pw = 0, D = 0;
#pragma unroll
      for (int w = 0; w < NW; w++) {
        const uint32_t c = wcnt[w];
        if (w < warp) pw += c;
        D += c;
      }
      {
        uint32_t at = pw;
        for (uint32_t i0 = 0; i0 < (uint32_t)DW; i0 += 32) {
          const uint32_t i = i0 + lane;
          uint32_t m = 0;
          if (i < (uint32_t)DW) {
#pragma unroll
            for (int w = 0; w < MW; w++) m |= smask[w * (SS + 1) + arena0 + i];
          }
          const uint32_t bal = __ballot_sync(0xffffffffu, m != 0);
          if (m != 0) list[at + __popc(bal & ((1u << lane) - 1u))] = (uint16_t)(arena0 + i);
          at += __popc(bal);
        }
      }
      const int more = __syncthreads_or(seg_next < n_seg ? 1 : 0);
      cnt = 0;
      tbucket = -1;
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
            if (w == 0) {
              const uint32_t wq = slot / (uint32_t)DW, sl = slot - wq * (uint32_t)DW;
              if (sl < (uint32_t)TS) tkeys[(size_t)wq * TS + sl] = EMPTY;
            }
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
      // the wiped keys and masks must be in place before any warp inserts again
      __syncthreads();
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

  // ---- drain the tensor pipe, read the accumulators back, add them to W (position p = set p) ----
  if ((uses0 > 0 && !umma::mbar_wait_bounded(&bar[0], (uses0 - 1) & 1u)) ||
      (uses1 > 0 && !umma::mbar_wait_bounded(&bar[1], (uses1 - 1) & 1u))) {
    if (lane == 0) { atomicAdd(&stats[3], 1ull); atomicOr(&stats[4], 2ull); }
  }
  umma::fence_after_thread_sync();
  __syncthreads();
  if (misc[kMiscAnyMma] && warp < 4) {
    // M = 64: row m lives in TMEM lane (m % 16) + 32 (m / 16)
    const int row = lane < 16 ? warp * 16 + lane : -1;
    // Gram position q holds set 4 (q % 16) + q / 16
    auto set_of = [&](int q) { const int s = 4 * (q & 15) + (q >> 4); return s < n_sets ? s : -1; };
    const int si = row >= 0 ? set_of(row) : -1;
#pragma unroll 1
    for (int c0 = 0; c0 < MMA_N; c0 += 32) {
      uint32_t v[32];
      umma::tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
      umma::tmem_ld_wait();
      if (si >= 0) {
#pragma unroll
        for (int j = 0; j < 32; j++) {
          const int sj = set_of(c0 + j);
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
