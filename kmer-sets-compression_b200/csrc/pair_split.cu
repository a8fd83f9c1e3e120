// pair_split.cu -- P4: the split step of KmerSetSet's main loop as ONE streaming pass.
//
// Replaces lib/core/kmer_set_set.h:332-343 (n = Intersection(j, k); j.Sub(n); k.Sub(n), built from
// KmerSet::Sub / Intersection, lib/core/kmer_set.h:177-187, 301-305) for a BATCH of pairs: the
// n-1 edges of the `mst` driver, or the one pair of a greedy iteration. The reference decodes
// two SPSS into hash sets and erases key by key; here both inputs are sorted CSR sets, every key
// is read from HBM once and every output key is written once.
//
// split_stream_kernel: one CTA = 256 consecutive finest-level fine buckets of one pair, one thread
// per fine bucket (about 10 + 10 keys). Units are handed out by a ticket counter, chunk-major
// (chunk c of every pair of the batch before chunk c + 1 of any).
//   1. the thread merges its two runs straight from global memory (each run is a few sectors that
//      stay in L1) and remembers which keys are common in two 32-bit masks;
//   2. a block scan of the per-bucket match counts gives the outputs' finest offsets inside the
//      chunk and the chunk's totals;
//   3. the totals go through a decoupled look-back over the earlier chunks of the same pair (one
//      64-bit word per output: flag | count), so the global position of every output key is
//      known without a counting pass and without a device-wide scan;
//   4. the thread writes its keys from the masks (runs longer than 32 keys merge again) and
//      every offset level of the up to three new sets that has an entry at its fine bucket.
// Measured alternatives (profiles/r01_p4_notes.md): a shared-memory tile version (coalesced tile
// loads, per-key binary search, warp-ballot compaction in place) was 2-3 x slower (4.3-7.2 warp
// instructions per key against 2.2 here; tools/experiments/pair_split_smem_tile.cu.txt); staging
// each warp's runs in shared memory with 16-byte loads changed nothing (not bound by L1); with
// pair-major unit order a quarter of the time went to the barrier behind the look-back (the
// predecessors were still merging), chunk-major order took that away (78 -> 52 us per edge).
//
// Output sizes are not known before the pass: with a caller-supplied |j & k| (the pair-counts
// matrix has it) the outputs are allocated exactly and written directly; without it they are
// written to an upper-bound staging area and copied into exact allocations afterwards.
//
// Algorithmic bytes: (n_j + n_k + |n| + |j\n| + |k\n|) * sizeof(Key) + 5 * (2^N + 1) * 4.
#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "kmsc_common.cuh"

namespace kmsc {

#ifndef KMSC_SPLIT_MINB
#define KMSC_SPLIT_MINB 8   // resident CTAs per SM the register budget is cut for (32 registers)
#endif
constexpr int kSpThreads = 256;

struct SplitPair {
  const void* ka; const uint32_t* la;  // keys and finest-level offsets of j
  const void* kb; const uint32_t* lb;  // ... of k
  void* out[3];                        // keys of n, j\n, k\n (nullptr: not wanted)
  uint32_t* lev[3];                    // lev_base of the new sets (nullptr: not wanted)
  uint32_t cap[3];                     // capacity of out[q] in keys
  uint32_t pad;
};

struct CopyDesc { const void* src; void* dst; unsigned long long bytes; };

__device__ __forceinline__ unsigned long long ld_state(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_state(unsigned long long* p, unsigned long long v) {
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// every offset level that has an entry at finest index x receives v
__device__ __forceinline__ void write_levels(uint32_t* __restrict__ base, int N, int F, uint32_t x, uint32_t v) {
  for (int f = F; f >= 0; f--) {
    const int sh = F - f;
    if (x & ((1u << sh) - 1u)) break;
    const size_t start = ((size_t)1 << N) * (((size_t)1 << f) - 1) + (size_t)f;
    base[start + (x >> sh)] = v;
  }
}

template <typename KeyT>
__device__ __forceinline__ uint32_t lower_bound_key(const KeyT* __restrict__ k, uint32_t lo, uint32_t hi, KeyT t) {
  while (lo < hi) {
    const uint32_t mid = (lo + hi) >> 1;
    if (k[mid] < t) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// One CTA = 256 >> qlog consecutive finest-level fine buckets of one pair, 2^qlog threads per
// fine bucket (qlog = 0 unless the sets are dense: a rank's prefix shard holds N x the keys per
// fine bucket; the threads of a fine bucket then take the 2^qlog sub-ranges of its key space,
// found by one binary search each, so that a thread still merges about 10 + 10 keys).
// Units are handed out by a ticket so that a unit only ever waits for units taken earlier.
template <typename KeyT>
__global__ void __launch_bounds__(kSpThreads, KMSC_SPLIT_MINB) split_stream_kernel(
    const SplitPair* __restrict__ pairs, uint32_t n_pairs, uint32_t chunks_per_pair, uint32_t NF, int N, int F,
    int key_bits, int qlog_live, uint32_t xlo, uint32_t xhi, unsigned long long* __restrict__ state,
    uint32_t* __restrict__ ticket, uint32_t* __restrict__ totals, int lean) {
  uint32_t* const watchdog = ticket + 1;
  __shared__ uint32_t s_w[kSpThreads / 32];
  __shared__ uint32_t s_unit;
  __shared__ uint32_t s_g[3];  // exclusive prefix of this chunk inside its pair
  const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) s_unit = atomicAdd(ticket, 1u);
  __syncthreads();
  const uint32_t u = s_unit;
  // chunk-major order: chunk c of every pair before chunk c + 1 of any. The predecessors of a
  // unit were taken n_pairs tickets earlier (their totals are long published when it looks back),
  // and a set shared by several pairs of the batch is read from DRAM once and from L2 after that
  const uint32_t c = u / n_pairs, p = u - c * n_pairs;
  const SplitPair* P = pairs + p;
  const KeyT* __restrict__ ka = (const KeyT*)P->ka;
  const KeyT* __restrict__ kb = (const KeyT*)P->kb;
  const uint32_t* __restrict__ la = P->la;
  const uint32_t* __restrict__ lb = P->lb;
  // fine buckets [xlo, xhi) (multiples of 256) can hold keys: their chunks put 2^qlog_live threads
  // on a fine bucket; the chunks before and after them (empty for a prefix shard) take 256 each
  const uint32_t c1 = xlo / kSpThreads, c2 = (xhi - xlo) / ((uint32_t)kSpThreads >> qlog_live);
  int qlog = 0;
  uint32_t x0;
  if (c < c1) x0 = c * kSpThreads;
  else if (c < c1 + c2) { qlog = qlog_live; x0 = xlo + (c - c1) * ((uint32_t)kSpThreads >> qlog_live); }
  else x0 = xhi + (c - c1 - c2) * kSpThreads;
  const uint32_t fpc = (uint32_t)kSpThreads >> qlog;  // fine buckets of this chunk
  const uint32_t x = x0 + (tid >> qlog), sub = tid & ((1u << qlog) - 1u);
  const uint32_t xe = min(x0 + fpc, NF);
  const bool live = x < NF;
  // ---- pass 1: merge the two runs, remember which keys are common (runs of <= 32 keys) -----------
  uint32_t i0 = 0, i1 = 0, j0 = 0, j1 = 0;
  if (live) { i0 = la[x]; i1 = la[x + 1]; j0 = lb[x]; j1 = lb[x + 1]; }
  if (qlog > 0) {
    // this thread's sub-range of the fine bucket's key space: [t_lo, t_lo + 2^(key_bits - F - qlog))
    const int sh = key_bits - F - qlog;
    const unsigned long long t_lo = ((unsigned long long)(x & ((1u << F) - 1u)) << (key_bits - F)) | ((unsigned long long)sub << sh);
    uint32_t lo_a = i0, lo_b = j0;
    if (live && sub > 0) {
      lo_a = lower_bound_key<KeyT>(ka, i0, i1, (KeyT)t_lo);
      lo_b = lower_bound_key<KeyT>(kb, j0, j1, (KeyT)t_lo);
    }
    // the next thread's lower bound is this thread's end (the last sub-range ends with the fine bucket)
    const uint32_t up_a = __shfl_down_sync(0xffffffffu, lo_a, 1), up_b = __shfl_down_sync(0xffffffffu, lo_b, 1);
    if (sub + 1 < (1u << qlog)) { i1 = up_a; j1 = up_b; }
    i0 = lo_a; j0 = lo_b;
  }
  const uint32_t A0 = la[x0], B0 = lb[x0], A1 = la[xe], B1 = lb[xe];
  const bool small = (i1 - i0 <= 32) && (j1 - j0 <= 32);
  uint32_t mA = 0, mB = 0;  // bit r: key r of the run is common (exact for runs of <= 32 keys)
  uint32_t nI = 0;
  {
    uint32_t i = i0, j = j0;
    while (i < i1 && j < j1) {
      const KeyT a = ka[i], b = kb[j];
      if (a == b) {
        mA |= 1u << ((i - i0) & 31);
        mB |= 1u << ((j - j0) & 31);
        nI++;
      }
      i += (a <= b);
      j += (b <= a);
    }
  }
  // ---- common keys before this fine bucket inside the chunk, chunk total --------------------------
  uint32_t inc = nI;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) s_w[warp] = inc;
  __syncthreads();
  uint32_t before = 0, totI = 0;
#pragma unroll
  for (int w = 0; w < kSpThreads / 32; w++) {
    const uint32_t v = s_w[w];
    if ((uint32_t)w < warp) before += v;
    totI += v;
  }
  const uint32_t preI = before + inc - nI;
  const uint32_t totA = (A1 - A0) - totI, totB = (B1 - B0) - totI;
  // ---- decoupled look-back over the earlier chunks of the pair (warp 0) ---------------------------
  unsigned long long* const st_me = state + ((size_t)p * chunks_per_pair + c) * 4;
  if (warp == 0) {
    if (lane == 0) {
      st_state(st_me + 0, (1ull << 32) | totI);
      st_state(st_me + 1, (1ull << 32) | totA);
      st_state(st_me + 2, (1ull << 32) | totB);
    }
    uint32_t eI = 0, eA = 0, eB = 0;
    int64_t idx = (int64_t)c - 1;
    while (idx >= 0) {
      const int64_t j = idx - (int64_t)lane;
      unsigned long long w0 = 0, w1 = 0, w2 = 0;
      if (j >= 0) {
        const unsigned long long* s = state + ((size_t)p * chunks_per_pair + (size_t)j) * 4;
        uint32_t polls = 0;
        do {
          w0 = ld_state(s); w1 = ld_state(s + 1); w2 = ld_state(s + 2);
          if (++polls > (1u << 24)) {  // watchdog: a protocol bug must end as KMSC_E_STATE, not as a hung GPU
            atomicExch(watchdog, 1u);
            w0 = w1 = w2 = 2ull << 32;
            break;
          }
        } while ((w0 >> 32) == 0 || (w1 >> 32) != (w0 >> 32) || (w2 >> 32) != (w0 >> 32));  // 0: not yet; mixed: mid-update
      }
      const bool incl = j >= 0 && (w0 >> 32) == 2;
      const uint32_t bal = __ballot_sync(0xffffffffu, incl);
      const int stop = bal ? (__ffs(bal) - 1) : 31;  // lanes 0..stop contribute
      uint32_t vI = ((int)lane <= stop && j >= 0) ? (uint32_t)w0 : 0u;
      uint32_t vA = ((int)lane <= stop && j >= 0) ? (uint32_t)w1 : 0u;
      uint32_t vB = ((int)lane <= stop && j >= 0) ? (uint32_t)w2 : 0u;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        vI += __shfl_xor_sync(0xffffffffu, vI, o);
        vA += __shfl_xor_sync(0xffffffffu, vA, o);
        vB += __shfl_xor_sync(0xffffffffu, vB, o);
      }
      eI += vI; eA += vA; eB += vB;
      if (bal) break;
      idx -= 32;
    }
    if (lane == 0) {
      st_state(st_me + 0, (2ull << 32) | (unsigned long long)(eI + totI));
      st_state(st_me + 1, (2ull << 32) | (unsigned long long)(eA + totA));
      st_state(st_me + 2, (2ull << 32) | (unsigned long long)(eB + totB));
      s_g[0] = eI; s_g[1] = eA; s_g[2] = eB;
      if (c == chunks_per_pair - 1) {
        totals[p * 3 + 0] = eI + totI; totals[p * 3 + 1] = eA + totA; totals[p * 3 + 2] = eB + totB;
        for (int q = 0; q < 3; q++)
          if (P->lev[q]) {
            const uint32_t tot_q = q == 0 ? eI + totI : q == 1 ? eA + totA : eB + totB;
            if (lean) P->lev[q][(size_t)1 << N] = tot_q;   // lean: only the bucket level is written
            else write_levels(P->lev[q], N, F, NF, tot_q);
          }
      }
    }
  }
  __syncthreads();
  if (!live) return;
  // ---- pass 2: write the keys (the runs are still in L1 / L2) and the offset levels ---------------
  KeyT* const oI = (KeyT*)P->out[0];
  KeyT* const oA = (KeyT*)P->out[1];
  KeyT* const oB = (KeyT*)P->out[2];
  const uint32_t capI = P->cap[0], capA = P->cap[1], capB = P->cap[2];
  uint32_t pI = s_g[0] + preI;
  uint32_t pA = s_g[1] + (i0 - A0) - preI;
  uint32_t pB = s_g[2] + (j0 - B0) - preI;
  {
    uint32_t* const l0 = P->lev[0];
    uint32_t* const l1 = P->lev[1];
    uint32_t* const l2 = P->lev[2];
    const int tz = sub ? -1 : x ? min(F, __ffs(x) - 1) : F;  // levels F, F-1, ..., F-tz have an entry at x
    size_t start = ((size_t)1 << N) * (((size_t)1 << F) - 1) + (size_t)F;
    if (lean) {
      // only the bucket level (the new sets' finer levels are built on first use: an output of 0.2 M
      // keys would otherwise cost 8 MB of offsets, 4 / 5 of everything this kernel writes)
      if (tz == F) {
        const size_t idx = x >> F;
        if (l0) l0[idx] = pI;
        if (l1) l1[idx] = pA;
        if (l2) l2[idx] = pB;
      }
    } else
    for (int sh = 0; sh <= tz; sh++) {
      const size_t idx = start + (x >> sh);
      if (l0) l0[idx] = pI;
      if (l1) l1[idx] = pA;
      if (l2) l2[idx] = pB;
      start -= ((size_t)1 << (N + F - sh - 1)) + 1;  // level f - 1 starts 2^(N+f-1) + 1 entries earlier
    }
  }
  if (small) {
    const uint32_t lenA = i1 - i0, lenB = j1 - j0;
    const KeyT* const srcA = ka + i0;  // (the runs are still in L1 / L2)
    const KeyT* const srcB = kb + j0;
    if (oI) for (uint32_t m = mA; m; m &= m - 1) { if (pI < capI) oI[pI] = srcA[__ffs(m) - 1]; pI++; }
    if (oA) {
      uint32_t m = ~mA & (lenA >= 32 ? ~0u : ((1u << lenA) - 1u));
      for (; m; m &= m - 1) { if (pA < capA) oA[pA] = srcA[__ffs(m) - 1]; pA++; }
    }
    if (oB) {
      uint32_t m = ~mB & (lenB >= 32 ? ~0u : ((1u << lenB) - 1u));
      for (; m; m &= m - 1) { if (pB < capB) oB[pB] = srcB[__ffs(m) - 1]; pB++; }
    }
  } else {  // a run longer than the masks: merge again
    uint32_t i = i0, j = j0;
    while (i < i1 && j < j1) {
      const KeyT a = ka[i], b = kb[j];
      if (a < b) { if (oA && pA < capA) oA[pA] = a; pA++; i++; }
      else if (b < a) { if (oB && pB < capB) oB[pB] = b; pB++; j++; }
      else { if (oI && pI < capI) oI[pI] = a; pI++; i++; j++; }
    }
    for (; i < i1; i++) { if (oA && pA < capA) oA[pA] = ka[i]; pA++; }
    for (; j < j1; j++) { if (oB && pB < capB) oB[pB] = kb[j]; pB++; }
  }
}


// staged outputs -> exact allocations: blockIdx.y = descriptor, 16-byte vectors
__global__ void split_copy_kernel(const CopyDesc* __restrict__ d) {
  const CopyDesc cd = d[blockIdx.y];
  const uint4* __restrict__ s = (const uint4*)cd.src;
  uint4* __restrict__ t = (uint4*)cd.dst;
  const unsigned long long nv = (cd.bytes + 15) / 16;
  for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < nv;
       i += (unsigned long long)gridDim.x * blockDim.x)
    t[i] = s[i];
}

static uint64_t levels_entries(int N, int max_level) {
  uint64_t t = 0;
  for (int f = 0; f <= max_level; f++) t += ((uint64_t)1 << (N + f)) + 1;
  return t;
}

// a new set without keys: levels allocated, keys attached later
static int shell_alloc(kmsc_ctx* ctx, const kmsc_set* like, kmsc_set** out) {
  kmsc_set* s = new kmsc_set();
  s->K = like->K; s->N = like->N; s->key_bytes = like->key_bytes; s->key_bits = like->key_bits;
  s->max_level = like->max_level;
  s->n_keys = 0;
  s->has_dups = 0;
  s->b_lo = like->b_lo; s->b_hi = like->b_hi;  // outputs of a rank's shard live in the same bucket range
  cudaError_t e = cudaMallocAsync((void**)&s->lev_base, levels_entries(s->N, s->max_level) * sizeof(uint32_t), ctx->stream);
  if (e != cudaSuccess) { delete s; return cuda_fail(e, "cudaMallocAsync levels", __FILE__, __LINE__); }
  uint64_t start = 0;
  for (int f = 0; f <= s->max_level; f++) {
    s->lev[f] = s->lev_base + start;
    start += ((uint64_t)1 << (s->N + f)) + 1;
  }
  *out = s;
  return KMSC_OK;
}

template <typename KeyT>
static int launch_split(kmsc_ctx* ctx, const SplitPair* d_pairs, uint32_t total_units, uint32_t cpp, uint32_t NF, int N, int F,
                        int key_bits, int qlog, uint32_t xlo, uint32_t xhi, unsigned long long* d_state, uint32_t* d_ticket,
                        uint32_t* d_totals, int lean) {
  split_stream_kernel<KeyT><<<total_units, kSpThreads, 0, ctx->stream>>>(d_pairs, total_units / cpp, cpp, NF, N, F, key_bits, qlog, xlo, xhi, d_state, d_ticket, d_totals, lean);
  count_launch(ctx);
  KMSC_CUDA(cudaGetLastError());
  return KMSC_OK;
}

struct SplitOut {           // host bookkeeping of one requested output
  kmsc_set* set = nullptr;
  bool direct = false;      // keys allocated exactly up front (hint)
  size_t stage_off = 0;     // else: byte offset in the staging area
  uint32_t cap = 0;
};

// one sub-batch: every pair's staging fits the budget
// tot_out (may be NULL): m x {|n|, |j\n|, |k\n|}
static int split_run(kmsc_ctx* ctx, const kmsc_set* const* js, const kmsc_set* const* ks, int32_t m,
                     const int64_t* hint, kmsc_set** outs[3], uint32_t* tot_out = nullptr) {
  const kmsc_set* like = js[0];
  const int kb = like->key_bytes, N = like->N, F = like->max_level;
  const uint32_t NF = (uint32_t)1 << (N + F);
  // the inputs are read at their finest level; the outputs get the bucket level only (lean) unless
  // KMSC_SPLIT_LEVELS=full: their finer levels are built when a consumer first needs them
  for (int32_t p = 0; p < m; p++) { KMSC_TRY(set_ensure_levels(ctx, js[p])); KMSC_TRY(set_ensure_levels(ctx, ks[p])); }
  const char* lv_env = getenv("KMSC_SPLIT_LEVELS");
  const int lean = (F > 0 && !(lv_env && strcmp(lv_env, "full") == 0)) ? 1 : 0;
  // threads per fine bucket: 2^qlog, so that a thread merges about 10 + 10 keys. The density is
  // taken over the bucket range the sets can hold keys in (a rank's prefix shard is N x denser)
  int qlog = 0;
  uint32_t xlo = 0, xhi = NF;  // fine buckets that can hold keys, widened to multiples of 256
  {
    double keys = 0;
    for (int32_t p = 0; p < m; p++) keys += (double)std::max(js[p]->n_keys, ks[p]->n_keys);
    const int nb = 1 << N;
    int lo = nb, hi = 0;  // union of the bucket ranges of all sets of the batch
    for (int32_t p = 0; p < m; p++)
      for (const kmsc_set* t : {js[p], ks[p]}) {
        lo = std::min(lo, t->b_lo < 0 ? 0 : (int)t->b_lo);
        hi = std::max(hi, t->b_hi < 0 ? nb : (int)t->b_hi);
      }
    if (hi <= lo) { lo = 0; hi = nb; }
    const double live_fine = (double)std::max(1, hi - lo) * (double)(1 << F);
    const double dens = keys / m / live_fine;
    while (qlog < 5 && qlog < like->key_bits - F && dens > 16.0 * (double)(1 << qlog)) qlog++;
    if (const char* e = getenv("KMSC_SPLIT_QLOG")) {
      const int v = atoi(e);
      if (v >= 0 && v <= 5 && v <= like->key_bits - F) qlog = v;
    }
    if (NF >= (uint32_t)kSpThreads) {
      xlo = ((uint32_t)lo << F) & ~(uint32_t)(kSpThreads - 1);
      xhi = std::min<uint64_t>(NF, (((uint64_t)hi << F) + kSpThreads - 1) & ~(uint64_t)(kSpThreads - 1));
    } else {
      qlog = 0;  // fewer than 256 fine buckets in all: one chunk (of the trailing kind)
      xlo = xhi = 0;
    }
  }
  const uint32_t cpp = xlo / kSpThreads + (xhi - xlo) / ((uint32_t)kSpThreads >> qlog) + (NF - xhi + kSpThreads - 1) / kSpThreads;
  if ((uint64_t)cpp * m >= 0xffffffffull) { set_error("pair_split_batch: too many work units"); return KMSC_E_INVALID; }
  const uint32_t units = cpp * (uint32_t)m;

  std::vector<SplitPair> hp(m);
  std::vector<SplitOut> ho((size_t)m * 3);
  size_t stage_bytes = 0;
  int rc = KMSC_OK;
  auto cleanup = [&]() {
    for (auto& o : ho) { kmsc_set_free(ctx, o.set); o.set = nullptr; }
  };
  for (int32_t p = 0; p < m && rc == KMSC_OK; p++) {
    const kmsc_set *a = js[p], *b = ks[p];
    SplitPair& P = hp[p];
    memset(&P, 0, sizeof(P));
    P.ka = a->keys; P.la = a->lev[F]; P.kb = b->keys; P.lb = b->lev[F];
    const int64_t h = hint ? hint[p] : -1;
    const bool exact = h >= 0 && h <= std::min(a->n_keys, b->n_keys);
    const int64_t ub[3] = {exact ? h : std::min(a->n_keys, b->n_keys), exact ? a->n_keys - h : a->n_keys,
                           exact ? b->n_keys - h : b->n_keys};
    for (int q = 0; q < 3 && rc == KMSC_OK; q++) {
      if (!outs[q]) continue;
      SplitOut& o = ho[(size_t)p * 3 + q];
      rc = shell_alloc(ctx, like, &o.set);
      if (rc != KMSC_OK) break;
      o.set->fine_ready = !lean;
      o.cap = (uint32_t)ub[q];
      o.direct = exact;
      P.lev[q] = o.set->lev_base;
      P.cap[q] = o.cap;
      if (exact) {
        cudaError_t e = cudaMallocAsync(&o.set->keys, (size_t)ub[q] * kb + 64, ctx->stream);
        if (e != cudaSuccess) { rc = cuda_fail(e, "cudaMallocAsync keys", __FILE__, __LINE__); break; }
        o.set->n_keys = ub[q];
        P.out[q] = o.set->keys;
      } else {
        o.stage_off = stage_bytes;
        stage_bytes += ((size_t)ub[q] * kb + 64 + 255) & ~(size_t)255;
      }
    }
  }
  if (rc != KMSC_OK) { cleanup(); return rc; }
  void* d_stage = nullptr;
  if (stage_bytes) {
    cudaError_t e = cudaMallocAsync(&d_stage, stage_bytes, ctx->stream);
    if (e != cudaSuccess) { cleanup(); return cuda_fail(e, "cudaMallocAsync staging", __FILE__, __LINE__); }
    for (int32_t p = 0; p < m; p++)
      for (int q = 0; q < 3; q++) {
        SplitOut& o = ho[(size_t)p * 3 + q];
        if (o.set && !o.direct) hp[p].out[q] = (char*)d_stage + o.stage_off;
      }
  }
  auto fail = [&](int code) {
    if (d_stage) cudaFreeAsync(d_stage, ctx->stream);
    cleanup();
    return code;
  };
  // device scratch: descriptors | totals | ticket | state
  const size_t desc_b = ((size_t)m * sizeof(SplitPair) + 255) & ~(size_t)255;
  const size_t tot_b = ((size_t)m * 3 * 4 + 255) & ~(size_t)255;
  const size_t state_b = (size_t)units * 32;
  rc = ctx->work2.reserve(desc_b + tot_b + 256 + state_b);
  if (rc != KMSC_OK) return fail(rc);
  char* base = (char*)ctx->work2.p;
  SplitPair* d_pairs = (SplitPair*)base;
  uint32_t* d_totals = (uint32_t*)(base + desc_b);
  uint32_t* d_ticket = (uint32_t*)(base + desc_b + tot_b);
  unsigned long long* d_state = (unsigned long long*)(base + desc_b + tot_b + 256);
  void* pin = nullptr;
  rc = ctx_pinned(ctx, std::max(desc_b, (size_t)m * 3 * sizeof(CopyDesc)) + tot_b + 64, &pin);
  if (rc != KMSC_OK) return fail(rc);
  memcpy(pin, hp.data(), (size_t)m * sizeof(SplitPair));
  cudaError_t e = cudaMemcpyAsync(d_pairs, pin, (size_t)m * sizeof(SplitPair), cudaMemcpyHostToDevice, ctx->stream);
  if (e == cudaSuccess) e = cudaMemsetAsync(d_totals, 0, tot_b + 256 + state_b, ctx->stream);
  if (e != cudaSuccess) return fail(cuda_fail(e, "pair_split setup", __FILE__, __LINE__));
  switch (kb) {
    case 2: rc = launch_split<uint16_t>(ctx, d_pairs, units, cpp, NF, N, F, like->key_bits, qlog, xlo, xhi, d_state, d_ticket, d_totals, lean); break;
    case 4: rc = launch_split<uint32_t>(ctx, d_pairs, units, cpp, NF, N, F, like->key_bits, qlog, xlo, xhi, d_state, d_ticket, d_totals, lean); break;
    default: rc = launch_split<unsigned long long>(ctx, d_pairs, units, cpp, NF, N, F, like->key_bits, qlog, xlo, xhi, d_state, d_ticket, d_totals, lean); break;
  }
  if (rc != KMSC_OK) return fail(rc);
  uint32_t* h_tot = (uint32_t*)((char*)pin + std::max(desc_b, (size_t)m * 3 * sizeof(CopyDesc)));
  // totals, then (tot_b bytes later) the ticket word and the watchdog word
  e = cudaMemcpyAsync(h_tot, d_totals, tot_b + 8, cudaMemcpyDeviceToHost, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  if (e != cudaSuccess) return fail(cuda_fail(e, "pair_split kernel", __FILE__, __LINE__));
  if (((const uint32_t*)((const char*)h_tot + tot_b))[1] != 0) {
    set_error("pair_split: look-back watchdog fired (protocol error)");
    return fail(KMSC_E_STATE);
  }
  std::vector<uint32_t> tot(h_tot, h_tot + (size_t)m * 3);
  if (tot_out) memcpy(tot_out, tot.data(), (size_t)m * 3 * sizeof(uint32_t));

  // exact allocations for the staged outputs + one gather copy; pairs whose hint was wrong are redone
  std::vector<CopyDesc> cds;
  std::vector<int32_t> redo;
  for (int32_t p = 0; p < m && rc == KMSC_OK; p++) {
    bool bad = false;
    for (int q = 0; q < 3; q++) {
      SplitOut& o = ho[(size_t)p * 3 + q];
      if (o.set && o.direct && tot[(size_t)p * 3 + q] != o.cap) bad = true;
    }
    if (bad) { redo.push_back(p); continue; }
    for (int q = 0; q < 3 && rc == KMSC_OK; q++) {
      SplitOut& o = ho[(size_t)p * 3 + q];
      if (!o.set || o.direct) continue;
      const uint32_t nk = tot[(size_t)p * 3 + q];
      cudaError_t e2 = cudaMallocAsync(&o.set->keys, (size_t)nk * kb + 64, ctx->stream);
      if (e2 != cudaSuccess) { rc = cuda_fail(e2, "cudaMallocAsync keys", __FILE__, __LINE__); break; }
      o.set->n_keys = nk;
      if (nk) cds.push_back({(const char*)d_stage + o.stage_off, o.set->keys, (unsigned long long)nk * kb});
    }
  }
  if (rc != KMSC_OK) return fail(rc);
  if (!cds.empty()) {
    CopyDesc* d_cd = (CopyDesc*)d_pairs;  // the pair descriptors are no longer needed
    memcpy(pin, cds.data(), cds.size() * sizeof(CopyDesc));
    e = cudaMemcpyAsync(d_cd, pin, cds.size() * sizeof(CopyDesc), cudaMemcpyHostToDevice, ctx->stream);
    if (e != cudaSuccess) return fail(cuda_fail(e, "pair_split copy descriptors", __FILE__, __LINE__));
    unsigned long long maxb = 0;
    for (auto& c : cds) maxb = std::max(maxb, c.bytes);
    const unsigned bx = (unsigned)std::min<unsigned long long>(std::max<unsigned long long>(1, maxb / (16 * 256 * 4)), 64);
    split_copy_kernel<<<dim3(bx, (unsigned)cds.size()), 256, 0, ctx->stream>>>(d_cd);
    count_launch(ctx);
    e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);  // pinned / scratch reuse by the next call
    if (e != cudaSuccess) return fail(cuda_fail(e, "pair_split copy", __FILE__, __LINE__));
  }
  if (d_stage) { cudaFreeAsync(d_stage, ctx->stream); d_stage = nullptr; }
  // hand the sets over
  for (int32_t p = 0; p < m; p++)
    for (int q = 0; q < 3; q++)
      if (outs[q]) { outs[q][p] = ho[(size_t)p * 3 + q].set; ho[(size_t)p * 3 + q].set = nullptr; }
  // wrong hints: free what was made for those pairs and run them again without a hint
  if (!redo.empty()) {
    std::vector<const kmsc_set*> rj, rk;
    for (int32_t p : redo) {
      for (int q = 0; q < 3; q++) if (outs[q]) { kmsc_set_free(ctx, outs[q][p]); outs[q][p] = nullptr; }
      rj.push_back(js[p]); rk.push_back(ks[p]);
    }
    std::vector<kmsc_set*> ro[3];
    kmsc_set** rp[3] = {nullptr, nullptr, nullptr};
    for (int q = 0; q < 3; q++) if (outs[q]) { ro[q].assign(redo.size(), nullptr); rp[q] = ro[q].data(); }
    rc = split_run(ctx, rj.data(), rk.data(), (int32_t)redo.size(), nullptr, rp);
    if (rc != KMSC_OK) {
      for (int32_t p = 0; p < m; p++)
        for (int q = 0; q < 3; q++) if (outs[q]) { kmsc_set_free(ctx, outs[q][p]); outs[q][p] = nullptr; }
      return rc;
    }
    for (size_t r = 0; r < redo.size(); r++)
      for (int q = 0; q < 3; q++) if (outs[q]) outs[q][redo[r]] = ro[q][r];
  }
  return KMSC_OK;
}

}  // namespace kmsc

using namespace kmsc;

extern "C" {

int kmsc_pair_split_batch(kmsc_ctx* ctx, const kmsc_set* const* js, const kmsc_set* const* ks, int32_t m,
                          const int64_t* inter_hint, kmsc_set** inter, kmsc_set** j_minus, kmsc_set** k_minus) {
  if (!ctx || !js || !ks || m < 0) { set_error("bad argument"); return KMSC_E_INVALID; }
  if (m == 0) return KMSC_OK;
  for (int32_t p = 0; p < m; p++) {
    if (!js[p] || !ks[p]) { set_error("pair %d: NULL set", p); return KMSC_E_INVALID; }
    const kmsc_set *a = js[p], *b = ks[p], *f = js[0];
    if (a->K != f->K || a->N != f->N || a->key_bytes != f->key_bytes || b->K != f->K || b->N != f->N ||
        b->key_bytes != f->key_bytes) {
      set_error("sets have different (K,N,KeyType)");
      return KMSC_E_INVALID;
    }
  }
  KMSC_CUDA(cudaSetDevice(ctx->device));
  // the merge assumes true sets: a repeated key would survive in j \ n while also being in n
  for (int32_t p = 0; p < m; p++)
    for (const kmsc_set* x : {js[p], ks[p]}) {
      KMSC_TRY(set_check_dups(ctx, const_cast<kmsc_set*>(x)));
      if (x->has_dups == 1) { set_error("pair %d: a set holds duplicate keys (split / diff need true sets)", p); return KMSC_E_INVALID; }
    }
  kmsc_set** outs[3] = {inter, j_minus, k_minus};
  for (int q = 0; q < 3; q++)
    if (outs[q]) for (int32_t p = 0; p < m; p++) outs[q][p] = nullptr;
  // sub-batches bounded by the staging the pairs need (no hint: upper-bound outputs)
  const size_t budget = (size_t)6 << 30;
  int32_t lo = 0;
  while (lo < m) {
    size_t need = 0;
    int32_t hi = lo;
    while (hi < m) {
      const size_t kb = (size_t)js[hi]->key_bytes;
      size_t add = 0;
      if (inter) add += (size_t)std::min(js[hi]->n_keys, ks[hi]->n_keys) * kb;
      if (j_minus) add += (size_t)js[hi]->n_keys * kb;
      if (k_minus) add += (size_t)ks[hi]->n_keys * kb;
      if (hi > lo && need + add > budget) break;
      need += add;
      hi++;
    }
    kmsc_set** sub[3] = {inter ? inter + lo : nullptr, j_minus ? j_minus + lo : nullptr, k_minus ? k_minus + lo : nullptr};
    int rc = split_run(ctx, js + lo, ks + lo, hi - lo, inter_hint ? inter_hint + lo : nullptr, sub);
    if (rc != KMSC_OK) {
      for (int q = 0; q < 3; q++)
        if (outs[q]) for (int32_t p = 0; p < lo; p++) { kmsc_set_free(ctx, outs[q][p]); outs[q][p] = nullptr; }
      return rc;
    }
    lo = hi;
  }
  return KMSC_OK;
}

/* KmerSet::Diff (lib/core/kmer_set.h:191-214): |a \ b| + |b \ a| = the totals of a split without outputs */
int kmsc_set_diff(kmsc_ctx* ctx, const kmsc_set* a, const kmsc_set* b, int64_t* diff) {
  if (!ctx || !diff || !a || !b) { set_error("NULL argument"); return KMSC_E_INVALID; }
  if (a->K != b->K || a->N != b->N || a->key_bytes != b->key_bytes) {
    set_error("sets have different (K,N,KeyType)");
    return KMSC_E_INVALID;
  }
  KMSC_CUDA(cudaSetDevice(ctx->device));
  for (const kmsc_set* x : {a, b}) {
    KMSC_TRY(set_check_dups(ctx, const_cast<kmsc_set*>(x)));
    if (x->has_dups == 1) { set_error("a set holds duplicate keys (split / diff need true sets)"); return KMSC_E_INVALID; }
  }
  kmsc_set** none[3] = {nullptr, nullptr, nullptr};
  uint32_t tot[3] = {0, 0, 0};
  KMSC_TRY(split_run(ctx, &a, &b, 1, nullptr, none, tot));
  *diff = (int64_t)tot[1] + (int64_t)tot[2];
  return KMSC_OK;
}

int kmsc_pair_split(kmsc_ctx* ctx, const kmsc_set* j, const kmsc_set* k, kmsc_set** inter, kmsc_set** j_minus,
                    kmsc_set** k_minus) {
  return kmsc_pair_split_batch(ctx, &j, &k, 1, nullptr, inter, j_minus, k_minus);
}

}  // extern "C"
