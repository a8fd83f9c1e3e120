// pair_counts_lane.cuh -- P3, fourth build: LANE-PRIVATE tables (one fine bucket per lane).
// Included by pair_counts.cu (shares SetDesc, plan_sel and the tcgen05 Gram scheme of umma.cuh).
//
// What the older builds pay (ncu, C2): the CTA-wide hash build 3.7 warp instructions and 0.8
// shared-memory wavefronts per key (random probes and byte marks are 3-6-way bank conflicted, one
// CAS per new key, every rare path is taken by some lane of almost every warp row); the warp-wide
// merge 51 instructions per distinct key = 3.2 per key (a merge step advances ~16 of the 64 heads a
// warp holds). Both are bound by instruction issue / shared-memory bandwidth at 11-13 % of HBM.
//
// Here a lane owns ONE row (a fine bucket at level f, chosen so that a row holds ~40 distinct keys)
// and a private 64-slot table for it:
//   layout   lane-interleaved: key pair p of lane l at p * 256 + l * 8, mask word w of slot q of
//            lane l at 8192 (1 + w) + q * 128 + l * 4. Whatever slots the 32 lanes probe, lane l
//            only touches banks 2 l, 2 l + 1 (keys) or l (masks): NO bank conflicts, and since no
//            other thread ever touches the table, no atomics, no CAS, no claim protocol.
//   walk     the warp goes through the sets one by one; the 32 runs of a set (rows x0 .. x0 + 31)
//            are one contiguous slice of its key array, staged by cp.async (two buffers per warp:
//            set s + 1 lands while set s is walked, its offsets are loaded one set earlier still).
//            Every lane walks its run as a little state machine: one slot-pair probe per
//            iteration (one 8-byte LDS), hit / claim / next pair decided by predicates, the
//            membership bit set by a plain read-modify-write off the critical path, the next key
//            fetched one iteration ahead. No lane waits for another lane's probe chain.
//   gram     slot q of the 32 lanes IS one K-step of the Gram in the "lane = key" operand layout:
//            no compaction scan, no slot list; empty slots contribute zero columns. Every warp
//            expands and issues its own tcgen05.mma (kind::i8, M = N = 64) into its own 64
//            columns of tensor memory (8 warps x 64 = the SM's 512 columns); the staging buffers
//            are the warp's key buffers (idle during the flush). No CTA-wide barrier anywhere in
//            the loop: warps are independent workers pulling 32-row chunks from one counter.
//   overflow a lane whose table reaches kLimit keys gives the row up: its masks are dropped and the
//            row goes to a list that pair_counts_lane_retry_kernel (a plain warp-wide merge with
//            CUDA-core counting, any row size) works off afterwards. Poisson tail: ~3 rows in 10^4
//            at C2.
// n <= 64 sets, 2- and 4-byte keys. Results are identical to the other builds
// (tests/test_gpu_pair_counts.py runs every case with each).
#pragma once

namespace kmsc {

struct LnCfg {
  static constexpr int T = 256, NW = 8;
  static constexpr int PAIRS = 32, SLOTS = 64;
  static constexpr int kTarget = 34;   // mean distinct keys per lane the chunk's split factor aims at
  static constexpr int kLimit = 58;   // keys a table may hold (probe chains stay short, one pair always ends a miss)
  static constexpr int STG = 2048;    // bytes of one staging buffer = one K-step of the Gram (64 sets x 32 keys)
};

struct LnLayout {
  static constexpr size_t tab_bytes = 3 * 8192;   // keys | mask word 0 | mask word 1, per warp
  static constexpr size_t o_tab = 0;
  static constexpr size_t o_stage = o_tab + (size_t)LnCfg::NW * tab_bytes;
  static constexpr size_t o_kp = o_stage + (size_t)LnCfg::NW * 2 * LnCfg::STG;
  static constexpr size_t o_lev = o_kp + 64 * 8;
  static constexpr size_t o_bar = o_lev + 64 * 8;
  static constexpr size_t o_misc = o_bar + (size_t)LnCfg::NW * 2 * 8;
  static constexpr size_t total = o_misc + 64;
};

namespace ln {
__device__ __forceinline__ uint32_t lds32(uint32_t a) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ uint32_t lds16(uint32_t a) { unsigned short v; asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(a)); return (uint32_t)v; }
__device__ __forceinline__ uint2 lds64(uint32_t a) { uint2 v; asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a)); return v; }
__device__ __forceinline__ void sts32(uint32_t a, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void sts64(uint32_t a, uint32_t x, uint32_t y) { asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(a), "r"(x), "r"(y) : "memory"); }
__device__ __forceinline__ void sts128(uint32_t a, uint4 v) { asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory"); }
template <typename KeyT> __device__ __forceinline__ uint32_t lds_key(uint32_t a) { return sizeof(KeyT) == 2 ? lds16(a) : lds32(a); }
// pair index of a table key: the top 5 bits of a Fibonacci product, as a byte offset (x 256)
__device__ __forceinline__ uint32_t pair_off(uint32_t t) { return ((t * 0x9E3779B1u) >> 19) & 0x1F00u; }
}  // namespace ln

// The walk of one lane's run: a branch-free state machine, one slot-pair probe per iteration.
// DIRECT = 0: keys staged in shared memory; 1: straight from global memory (a slice that outgrew the
// staging buffer). `left` is the whole state of the run (0 = done); a table that reaches kLimit keys
// ends the walk (the caller reads that off cnt).
template <typename KeyT, int DIRECT>
__device__ __forceinline__ void lane_walk(uint32_t pos_s, const KeyT* pos_g, uint32_t n, uint32_t lowmask,
                                          uint32_t tk_lane, uint32_t mdelta, uint32_t bit, uint32_t& cnt) {
  constexpr uint32_t SZ = (uint32_t)sizeof(KeyT);
  constexpr uint32_t EMPTY = 0xFFFFFFFFu;
  uint32_t left = n;
  uint32_t raw = 0;
  if (DIRECT) { if (left) raw = (uint32_t)__ldg(pos_g); }
  else raw = ln::lds_key<KeyT>(pos_s);
  uint32_t t = raw & lowmask;
  uint32_t hp = tk_lane + ln::pair_off(t);   // shared address of the pair being probed
  const uint32_t tk_end = tk_lane + 8192u;
  while (__any_sync(0xffffffffu, left != 0u)) {
    // the key after this one, one iteration ahead (a read past the run's end is harmless: staged
    // bytes are followed by shared memory, key arrays are padded)
    uint32_t rawn;
    if (DIRECT) rawn = left > 1u ? (uint32_t)__ldg(pos_g + 1) : 0u;
    else rawn = ln::lds_key<KeyT>(pos_s + SZ);
    const uint2 c = ln::lds64(hp);
    const bool act = left != 0u;
    const bool hit0 = c.x == t, e0 = c.x == EMPTY, hit1 = c.y == t, e1 = c.y == EMPTY;
    const bool use0 = hit0 | e0;
    const bool res = act & (use0 | hit1 | e1);    // this pair settles the key
    const bool ins = act & (e0 | (!use0 & !hit1 & e1));
    if (ins) {
      ln::sts32(hp + (use0 ? 0u : 4u), t);
      cnt++;
    }
    // membership bit: an unconditional read-modify-write (a lane that is not settled ORs in nothing)
    const uint32_t ma = hp + mdelta + (use0 ? 0u : 128u);
    const uint32_t m = ln::lds32(ma);
    ln::sts32(ma, m | (res ? bit : 0u));
    // next state
    const uint32_t tn = rawn & lowmask;
    uint32_t hadv = hp + 256u;
    if (hadv >= tk_end) hadv -= 8192u;
    hp = res ? tk_lane + ln::pair_off(tn) : hadv;
    t = res ? tn : t;
    if (res) {
      left--;
      if (DIRECT) pos_g++; else pos_s += SZ;
    }
    if (cnt >= (uint32_t)LnCfg::kLimit) left = 0u;
  }
}

template <typename KeyT>
__global__ void __launch_bounds__(LnCfg::T, 1)
pair_counts_lane_kernel(const SetDesc* __restrict__ sets, int n_sets, int f, uint32_t NF, uint32_t row_lo,
                        uint32_t row_hi, const uint32_t* __restrict__ sel_bitmap, int tbits, float inv_rho,
                        uint32_t* __restrict__ chunk_counter, unsigned long long* __restrict__ W,
                        unsigned long long* __restrict__ stats, uint2* __restrict__ retry_units,
                        uint32_t* __restrict__ retry_count, uint32_t retry_cap) {
  using C = LnCfg;
  using LY = LnLayout;
  constexpr uint32_t SZ = (uint32_t)sizeof(KeyT);
  constexpr uint32_t DIRECT = 0xFFFFFFFFu;
  constexpr uint32_t idesc = umma::make_idesc_u8(64, 64, 1, 1);

  extern __shared__ __align__(128) unsigned char smem_raw[];
  const void** skp = reinterpret_cast<const void**>(smem_raw + LY::o_kp);
  const uint32_t** slev = reinterpret_cast<const uint32_t**>(smem_raw + LY::o_lev);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw + LY::o_bar);
  int* misc = reinterpret_cast<int*>(smem_raw + LY::o_misc);   // [0] tmem base, [1 + w] warp w issued an MMA

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t smem0 = umma::smem_u32(smem_raw);
  const uint32_t tab = smem0 + (uint32_t)LY::o_tab + (uint32_t)warp * (uint32_t)LY::tab_bytes;
  const uint32_t tk_lane = tab + (uint32_t)lane * 8u;            // key pair p at + p * 256
  const uint32_t mk_lane = tab + 8192u + (uint32_t)lane * 4u;    // mask word w of slot q of this lane at + w * 8192 + q * 128
  const uint32_t stg = smem0 + (uint32_t)LY::o_stage + (uint32_t)warp * 2u * (uint32_t)C::STG;
  uint64_t* wbar = bar + warp * 2;
  const uint32_t lowmask = tbits >= 32 ? 0xFFFFFFFFu : ((1u << tbits) - 1u);

  // one-time init: empty tables
#pragma unroll 4
  for (int p = 0; p < C::PAIRS; p++) ln::sts64(tk_lane + (uint32_t)p * 256u, 0xFFFFFFFFu, 0xFFFFFFFFu);
#pragma unroll 4
  for (int q = 0; q < 2 * C::SLOTS; q++) ln::sts32(mk_lane + (uint32_t)q * 128u, 0u);
  for (int i = tid; i < 64; i += C::T) {
    skp[i] = i < n_sets ? sets[i].keys : nullptr;
    slev[i] = i < n_sets ? sets[i].lev : nullptr;
  }
  if (tid < C::NW * 2) umma::mbar_init(&bar[tid], 1);
  if (tid < 1 + C::NW) misc[1 + tid] = 0;
  if (tid == 0) umma::mbar_fence_init();
  if (warp == 0) umma::tmem_alloc(reinterpret_cast<uint32_t*>(&misc[0]), 512);
  umma::fence_before_thread_sync();
  __syncthreads();
  umma::fence_after_thread_sync();
  const uint32_t tmem_acc = (uint32_t)misc[0] + (uint32_t)warp * 64u;   // this warp's accumulator columns

  uint32_t uses0 = 0, uses1 = 0;      // commits on this warp's staging buffer 0 / 1 (uniform over the warp)
  // wait until the Gram has read staging buffer b for the last time (bounded: a protocol bug is reported, not hung on)
  auto wait_buf = [&](int b, uint32_t used) {
    if (used > 0 && !umma::mbar_wait_bounded(&wbar[b], (used - 1) & 1u)) {
      if (lane == 0) { atomicAdd(&stats[3], 1ull); atomicOr(&stats[4], 2ull); }
    }
  };
  // stage the slice [A, E) of set s into buffer b: 16-byte chunks from the aligned address below A
  auto stage_set = [&](int s, uint32_t A, uint32_t E, int b) -> uint32_t {
    const uintptr_t g0 = (uintptr_t)skp[s] + (size_t)A * SZ;
    const uintptr_t ga = g0 & ~(uintptr_t)15;
    const uint32_t shift = (uint32_t)(g0 - ga);
    const uint32_t bytes = shift + (E - A) * SZ;
    if (bytes > (uint32_t)C::STG) return DIRECT;
    for (uint32_t c = (uint32_t)lane * 16u; c < bytes; c += 512u)
      cp_async16(stg + (uint32_t)b * (uint32_t)C::STG + c, (const void*)(ga + c));
    return shift;
  };
  bool mma_started = false;
  unsigned long long st_keys = 0, st_dist = 0, st_retry = 0;
  const uint32_t c_lo = row_lo >> 5;

  for (;;) {
    uint32_t ch = 0;
    if (lane == 0) ch = atomicAdd(chunk_counter, 1u);
    ch = __shfl_sync(0xffffffffu, ch, 0);
    const uint32_t x0 = (c_lo + ch) << 5;
    if (x0 >= row_hi) break;
    {
      const uint32_t xl = x0 + (uint32_t)lane;
      if (!__any_sync(0xffffffffu, xl >= row_lo && xl < row_hi && plan_sel(sel_bitmap, xl, f))) continue;
    }
    // How many lanes a row needs: the keys of all sets in the chunk / rho = its distinct keys; a lane
    // takes 1 / 2^q of a row's key range so that its table gets ~kTarget of them (the prefix space is
    // not uniform: canonical k-mers starting with A are 7 x denser than those starting with T).
    int q = 0;
    {
      uint32_t tot = 0;
      const uint32_t xa = min(x0, NF), xb = min(x0 + 32u, NF);
      for (int s = lane; s < n_sets; s += 32) tot += slev[s][xb] - slev[s][xa];
      tot = __reduce_add_sync(0xffffffffu, tot);
      if (tot == 0u) continue;
      float d_row = (float)tot * inv_rho * (1.0f / 32.0f);
      while (d_row > (float)C::kTarget && q < 5 && q < tbits) { d_row *= 0.5f; q++; }
    }
    const uint32_t rows_pp = 32u >> q;                  // rows per pass
    const uint32_t sub = (uint32_t)lane & ((1u << q) - 1u);
    const uint32_t cut = q ? sub << (tbits - q) : 0u;   // first table key of this lane's part of the row

    for (uint32_t pass = 0; pass < (1u << q); pass++) {
      const uint32_t xb0 = x0 + pass * rows_pp;
      const uint32_t x = xb0 + ((uint32_t)lane >> q);
      const bool sel = x >= row_lo && x < row_hi && plan_sel(sel_bitmap, x, f);
      if (!__any_sync(0xffffffffu, sel)) continue;
      const uint32_t xi = min(x, NF), xe = min(x + 1u, NF);

      uint32_t a0 = slev[0][xi], e0 = slev[0][xe], a1 = 0, e1 = 0;
      if (n_sets > 1) { a1 = slev[1][xi]; e1 = slev[1][xe]; }
      // the Gram of the previous pass may still read the staging buffers
      wait_buf(0, uses0);
      wait_buf(1, uses1);
      uint32_t A0 = __shfl_sync(0xffffffffu, a0, 0);
      uint32_t sh0 = stage_set(0, A0, __shfl_sync(0xffffffffu, e0, 31), 0);
      cp_async_commit();
      uint32_t cnt = 0, row_keys = 0;
      for (int s = 0; s < n_sets; s++) {
        uint32_t a2 = 0, e2 = 0;
        if (s + 2 < n_sets) { a2 = slev[s + 2][xi]; e2 = slev[s + 2][xe]; }
        uint32_t A1 = 0, sh1 = 0;
        if (s + 1 < n_sets) {
          A1 = __shfl_sync(0xffffffffu, a1, 0);
          sh1 = stage_set(s + 1, A1, __shfl_sync(0xffffffffu, e1, 31), (s + 1) & 1);
        }
        cp_async_commit();
        cp_async_wait<1>();
        __syncwarp();
        const uint32_t sbase = stg + (uint32_t)(s & 1) * (uint32_t)C::STG + sh0;   // shared address of key A0
        const KeyT* gk = reinterpret_cast<const KeyT*>(skp[s]);
        uint32_t start = a0, end = e0;
        if (q) {
          // this lane's part of the row's run: lower bound of its first table key (the next lane's is its end)
          uint32_t lo = a0, len = sub ? e0 - a0 : 0u;
          while (__any_sync(0xffffffffu, len != 0u)) {
            const uint32_t half = len >> 1;
            uint32_t k = 0;
            if (len) k = (sh0 != DIRECT ? ln::lds_key<KeyT>(sbase + (lo + half - A0) * SZ) : (uint32_t)gk[lo + half]) & lowmask;
            if (len && k < cut) { lo += half + 1u; len -= half + 1u; } else len = half;
          }
          start = lo;
          const uint32_t nx = __shfl_down_sync(0xffffffffu, start, 1);
          if (sub != (1u << q) - 1u) end = nx;
        }
        const uint32_t n = (sel && cnt < (uint32_t)C::kLimit) ? end - start : 0u;
        row_keys += n;
        // mask word of slot 2 p + j of this lane = pair address + mdelta + 128 j
        const uint32_t mdelta = (mk_lane - tk_lane) + (uint32_t)(s >> 5) * 8192u;
        const uint32_t bit = 1u << (s & 31);
        if (sh0 != DIRECT)
          lane_walk<KeyT, 0>(sbase + (start - A0) * SZ, nullptr, n, lowmask, tk_lane, mdelta, bit, cnt);
        else
          lane_walk<KeyT, 1>(0u, gk + start, n, lowmask, tk_lane, mdelta, bit, cnt);
        __syncwarp();
        a0 = a1; e0 = e1; A0 = A1; sh0 = sh1;
        a1 = a2; e1 = e2;
      }
      cp_async_wait<0>();
      __syncwarp();

      // parts given up: to the retry list, their masks do not reach the Gram
      const bool ovf = cnt >= (uint32_t)C::kLimit;
      if (ovf) {
        const uint32_t slot = atomicAdd(retry_count, 1u);
        if (slot < retry_cap) retry_units[slot] = make_uint2(x, ((uint32_t)q << 8) | sub);
        else atomicOr(&stats[4], 16ull);
        st_retry++;   // (its keys and distinct keys are counted by the retry kernel)
      } else {
        st_keys += row_keys;
        st_dist += cnt;
      }

      // ---- flush: slot j of the 32 lanes = one K-step of the Gram -------------------------------
#pragma unroll 1
      for (int j = 0; j < C::SLOTS; j++) {
        const uint32_t ma = mk_lane + (uint32_t)j * 128u;
        uint32_t w0 = ln::lds32(ma), w1 = ln::lds32(ma + 8192u);
        if (ovf) { w0 = 0; w1 = 0; }
        if (!__any_sync(0xffffffffu, (w0 | w1) != 0u)) continue;
        ln::sts32(ma, 0u);
        ln::sts32(ma + 8192u, 0u);
        const int b = (int)((uses0 + uses1) & 1u);
        const uint32_t used = b ? uses1 : uses0;
        wait_buf(b, used);
        const uint32_t kb = stg + (uint32_t)b * (uint32_t)C::STG + (uint32_t)lane * 16u;
        uint4 lo, hi;
        lo.x = umma::nibble_to_bytes(w0);       lo.y = umma::nibble_to_bytes(w0 >> 4);
        lo.z = umma::nibble_to_bytes(w0 >> 8);  lo.w = umma::nibble_to_bytes(w0 >> 12);
        hi.x = umma::nibble_to_bytes(w0 >> 16); hi.y = umma::nibble_to_bytes(w0 >> 20);
        hi.z = umma::nibble_to_bytes(w0 >> 24); hi.w = umma::nibble_to_bytes(w0 >> 28);
        ln::sts128(kb, lo);
        ln::sts128(kb + 512u, hi);
        lo.x = umma::nibble_to_bytes(w1);       lo.y = umma::nibble_to_bytes(w1 >> 4);
        lo.z = umma::nibble_to_bytes(w1 >> 8);  lo.w = umma::nibble_to_bytes(w1 >> 12);
        hi.x = umma::nibble_to_bytes(w1 >> 16); hi.y = umma::nibble_to_bytes(w1 >> 20);
        hi.z = umma::nibble_to_bytes(w1 >> 24); hi.w = umma::nibble_to_bytes(w1 >> 28);
        ln::sts128(kb + 1024u, lo);
        ln::sts128(kb + 1536u, hi);
        umma::fence_async_smem();
        __syncwarp();
        if (lane == 0) {
          umma::fence_after_thread_sync();
          const uint64_t bd = umma::make_smem_desc(stg + (uint32_t)b * (uint32_t)C::STG, 128u, 512u);
          umma::mma_u8(tmem_acc, bd, bd, idesc, mma_started ? 1u : 0u);
          umma::mma_commit(&wbar[b]);
        }
        mma_started = true;
        if (b) uses1++; else uses0++;
      }
#pragma unroll 4
      for (int p = 0; p < C::PAIRS; p++) ln::sts64(tk_lane + (uint32_t)p * 256u, 0xFFFFFFFFu, 0xFFFFFFFFu);
      if (ovf) {  // every mask of a given-up lane (the flush skipped what the other lanes left empty)
#pragma unroll 4
        for (int j = 0; j < 2 * C::SLOTS; j++) ln::sts32(mk_lane + (uint32_t)j * 128u, 0u);
      }
      __syncwarp();
    }
  }

  // ---- drain the tensor pipe, sum the warps' accumulators, add them to W --------------------------
  wait_buf(0, uses0);
  wait_buf(1, uses1);
  if (lane == 0 && mma_started) misc[1 + warp] = 1;
  umma::fence_before_thread_sync();
  __syncthreads();
  umma::fence_after_thread_sync();
  if (warp < 4) {
    // M = 64: row m lives in TMEM lane (m % 16) + 32 (m / 16); warp q reads lanes 32 q .. 32 q + 31
    const int row = lane < 16 ? warp * 16 + lane : -1;
#pragma unroll 1
    for (int c0 = 0; c0 < 64; c0 += 32) {
      uint32_t acc[32];
#pragma unroll
      for (int j = 0; j < 32; j++) acc[j] = 0;
#pragma unroll 1
      for (int w = 0; w < C::NW; w++) {
        if (!misc[1 + w]) continue;
        uint32_t v[32];
        umma::tmem_ld32((uint32_t)misc[0] + ((uint32_t)(warp * 32) << 16) + (uint32_t)(w * 64 + c0), v);
        umma::tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; j++) acc[j] += v[j];
      }
      if (row >= 0 && row < n_sets) {
#pragma unroll
        for (int j = 0; j < 32; j++)
          if (c0 + j < n_sets && acc[j] != 0) atomicAdd(&W[(size_t)row * n_sets + c0 + j], (unsigned long long)acc[j]);
      }
    }
  }
  umma::fence_before_thread_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc((uint32_t)misc[0], 512);

  for (int o = 16; o > 0; o >>= 1) {
    st_keys += __shfl_xor_sync(0xffffffffu, st_keys, o);
    st_dist += __shfl_xor_sync(0xffffffffu, st_dist, o);
    st_retry += __shfl_xor_sync(0xffffffffu, st_retry, o);
  }
  if (lane == 0) {
    if (st_keys) atomicAdd(&stats[0], st_keys);
    if (st_dist) atomicAdd(&stats[1], st_dist);
    if (st_retry) atomicAdd(&stats[2], st_retry);
  }
}

// The parts of rows a lane gave up: one warp per unit (row, the 1 / 2^q part `sub` of its key range), a
// plain warp-wide multiway merge (lane l holds the heads of sets l and l + 32, the warp minimum is the
// next distinct key, two ballots are its membership mask) and CUDA-core counting of the pairs in
// shared memory. No capacity anywhere: any unit works.
template <typename KeyT>
__global__ void __launch_bounds__(128)
pair_counts_lane_retry_kernel(const SetDesc* __restrict__ sets, int n_sets, uint32_t NF, int tbits,
                              const uint2* __restrict__ units, const uint32_t* __restrict__ count_p,
                              uint32_t cap, unsigned long long* __restrict__ W,
                              unsigned long long* __restrict__ stats) {
  __shared__ uint32_t pc[64 * 64];
  const uint32_t n_units = min(*count_p, cap);
  if ((uint32_t)blockIdx.x * 4u >= n_units) return;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t lowmask = tbits >= 32 ? 0xFFFFFFFFu : ((1u << tbits) - 1u);
  for (int i = tid; i < 64 * 64; i += 128) pc[i] = 0;
  __syncthreads();
  unsigned long long st_keys = 0, st_dist = 0;
  for (uint32_t r = (uint32_t)blockIdx.x * 4u + (uint32_t)warp; r < n_units; r += gridDim.x * 4u) {
    const uint2 u = units[r];
    const uint32_t x = u.x;
    const int q = (int)(u.y >> 8);
    const uint32_t sub = u.y & 255u;
    // table keys of the unit: [t_lo, t_hi]
    const uint32_t t_lo = q ? sub << (tbits - q) : 0u;
    const uint32_t t_hi = q ? (t_lo + (1u << (tbits - q)) - 1u) : lowmask;
    const KeyT* kp[2];
    uint32_t i[2], e[2], k[2];
#pragma unroll
    for (int j = 0; j < 2; j++) {
      const int s = lane + 32 * j;
      kp[j] = nullptr; i[j] = 0; e[j] = 0; k[j] = 0;
      if (s < n_sets) {
        kp[j] = reinterpret_cast<const KeyT*>(sets[s].keys);
        i[j] = sets[s].lev[min(x, NF)];
        e[j] = sets[s].lev[min(x + 1u, NF)];
        while (i[j] < e[j] && ((uint32_t)kp[j][i[j]] & lowmask) < t_lo) i[j]++;
        uint32_t z = i[j];
        while (z < e[j] && ((uint32_t)kp[j][z] & lowmask) <= t_hi) z++;
        e[j] = z;
        if (i[j] < e[j]) k[j] = (uint32_t)kp[j][i[j]];
        st_keys += e[j] - i[j];
      }
    }
    for (;;) {
      const bool l0 = i[0] < e[0], l1 = i[1] < e[1];
      if (!__any_sync(0xffffffffu, l0 | l1)) break;
      uint32_t v = 0xFFFFFFFFu;
      if (l0) v = k[0];
      if (l1 && k[1] < v) v = k[1];
      const uint32_t g = __reduce_min_sync(0xffffffffu, v);
      const bool h0 = l0 && k[0] == g, h1 = l1 && k[1] == g;
      const uint32_t m0 = __ballot_sync(0xffffffffu, h0), m1 = __ballot_sync(0xffffffffu, h1);
      unsigned long long M = ((unsigned long long)m1 << 32) | m0;
      while (M) {
        const int a = __ffsll((long long)M) - 1;
        M &= M - 1;
        if (h0) atomicAdd(&pc[a * 64 + lane], 1u);
        if (h1) atomicAdd(&pc[a * 64 + lane + 32], 1u);
      }
      if (lane == 0) st_dist++;
      if (h0) { i[0]++; if (i[0] < e[0]) k[0] = (uint32_t)kp[0][i[0]]; }
      if (h1) { i[1]++; if (i[1] < e[1]) k[1] = (uint32_t)kp[1][i[1]]; }
    }
  }
  __syncthreads();
  for (int t = tid; t < 64 * 64; t += 128) {
    const int a = t >> 6, b = t & 63;
    if (pc[t] && a < n_sets && b < n_sets) atomicAdd(&W[(size_t)a * n_sets + b], (unsigned long long)pc[t]);
  }
  for (int o = 16; o > 0; o >>= 1) {
    st_keys += __shfl_xor_sync(0xffffffffu, st_keys, o);
    st_dist += __shfl_xor_sync(0xffffffffu, st_dist, o);
  }
  if (lane == 0) {
    if (st_keys) atomicAdd(&stats[0], st_keys);
    if (st_dist) atomicAdd(&stats[1], st_dist);
  }
}

}  // namespace kmsc
