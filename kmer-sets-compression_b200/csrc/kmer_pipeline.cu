// kmer_pipeline.cu -- packed bases -> k-mers -> fine-bucket partition -> sort ->
// count / cutoff / dedup -> device CSR set. Shared by P2 (decode.cu) and P1
// (count.cu). Reference semantics:
//   k-mer value and canonical form      lib/core/kmer.h:22-46, 103-133
//   bucket / key split                  lib/core/kmer_set.h:22-31
//   sorted buckets, duplicates kept     lib/core/kmer_set_compact.h:145-200 (mode 0)
//   hash-set insert (dedup)             lib/core/spss.h:1903-1925           (mode 1)
//   uint8 saturating counts + cutoff    lib/core/kmer_counter.h:28-38, 94, 222-236 (mode 2)
//
// The partition is by the FINEST fine-bucket index (top N+F bits of the 2K-bit
// value), so the scan of the histogram is directly the set's finest offset level
// and the per-bucket sort only ever sees a few keys (about n / 2^(N+F)).
#include "kmer_pipeline.cuh"
#include "scan.cuh"

namespace kmsc {

namespace {

constexpr int kRankSortMax = 64;    // runs up to this length: O(L^2) rank sort by one thread
constexpr int kSmemSortMax = 4096;  // runs up to this length: bitonic sort in shared memory

struct KP {
  const unsigned long long* words;
  const uint32_t* bad;
  unsigned long long n_pos;
  int K, key_bits, fine_shift;  // fine index = value >> fine_shift
  int canonical;
  unsigned long long bucket_lo, bucket_hi;  // keep buckets in [lo, hi)
};

__device__ __forceinline__ bool kmer_at(const KP& kp, unsigned long long p, unsigned long long* v_out) {
  if (p >= kp.n_pos) return false;
  if ((kp.bad[p >> 5] >> (p & 31)) & 1u) return false;
  unsigned long long v = load_kmer(kp.words, p, kp.K);
  if (kp.canonical) {
    const unsigned long long rc = revcomp(v, kp.K);
    v = rc < v ? rc : v;
  }
  const unsigned long long b = v >> kp.key_bits;
  if (b < kp.bucket_lo || b >= kp.bucket_hi) return false;
  *v_out = v;
  return true;
}

__global__ void hist_kernel(KP kp, uint32_t* __restrict__ cnt) {
  const unsigned long long p = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  unsigned long long v;
  if (kmer_at(kp, p, &v)) atomicAdd(&cnt[v >> kp.fine_shift], 1u);
}

template <typename KeyT>
__global__ void scatter_kernel(KP kp, const uint32_t* __restrict__ offs, uint32_t* __restrict__ cursor,
                               KeyT* __restrict__ tmp) {
  const unsigned long long p = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  unsigned long long v;
  if (kmer_at(kp, p, &v)) {
    const unsigned long long x = v >> kp.fine_shift;
    const uint32_t pos = offs[x] + atomicAdd(&cursor[x], 1u);
    const unsigned long long mask = kp.key_bits == 64 ? ~0ull : ((1ull << kp.key_bits) - 1);
    tmp[pos] = (KeyT)(v & mask);
  }
}

// thread per fine bucket: short runs are rank-sorted tmp -> out, long ones queued. Runs of up
// to 16 keys (the common case: ~10 keys per finest bucket) are sorted out of registers: one
// load per key instead of one per comparison. dup_flag is raised when a run holds equal keys.
template <typename KeyT>
__global__ void sort_runs_kernel(const KeyT* __restrict__ tmp, const uint32_t* __restrict__ offs, uint32_t NF,
                                 KeyT* __restrict__ out, uint32_t* __restrict__ big_list,
                                 uint32_t* __restrict__ big_count, uint32_t* __restrict__ dup_flag) {
  const uint32_t x = blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= NF) return;
  const uint32_t a = offs[x], b = offs[x + 1], L = b - a;
  if (L == 0) return;
  if (L == 1) { out[a] = tmp[a]; return; }
  if (L > kRankSortMax) { big_list[atomicAdd(big_count, 1u)] = x; return; }
  bool dup = false;
  if (L <= 16) {
    // 16-input bitonic network in registers (80 compare-exchanges); slots past L hold the
    // largest key value and sort to the end (a real key equal to it is still among the first L)
    KeyT k[16];
#pragma unroll
    for (int i = 0; i < 16; i++) k[i] = (uint32_t)i < L ? tmp[a + i] : (KeyT)~(KeyT)0;
#pragma unroll
    for (int kk = 2; kk <= 16; kk <<= 1) {
#pragma unroll
      for (int j = kk >> 1; j > 0; j >>= 1) {
#pragma unroll
        for (int i = 0; i < 16; i++) {
          const int l = i ^ j;
          if (l > i) {
            const KeyT x = k[i], y = k[l];
            const bool up = (i & kk) == 0;
            const KeyT lo = x < y ? x : y, hi = x < y ? y : x;
            k[i] = up ? lo : hi;
            k[l] = up ? hi : lo;
          }
        }
      }
    }
#pragma unroll
    for (int i = 0; i < 16; i++) {
      if ((uint32_t)i < L) {
        out[a + i] = k[i];
        if (i > 0) dup |= k[i] == k[i - 1];
      }
    }
  } else {
    for (uint32_t i = a; i < b; i++) {
      const KeyT ki = tmp[i];
      uint32_t r = 0;
      for (uint32_t j = a; j < b; j++) {
        const KeyT kj = tmp[j];
        r += (kj < ki) || (kj == ki && j < i);
        dup |= (kj == ki) && (j < i);
      }
      out[a + r] = ki;
    }
  }
  if (dup) *dup_flag = 1u;
}

// CTA per queued long run (kRankSortMax < L <= kSmemSortMax; counting reads at 90 x coverage puts ~200 k-mer
// instances of ~60 distinct k-mers into a finest bucket): a counting sort in shared memory over the top
// kSubBits bits below the fine level -- histogram, scan, grouped copy -- then every key finds its rank
// inside its sub-bin (a handful of keys: the copies of one k-mer and a neighbour or two) and goes straight
// to its final place. Equal keys are interchangeable, so the order among them is whatever the grouped copy
// left. A sub-bin longer than kSubRankMax (a k-mer with hundreds of copies next to others) sends the run
// through the bitonic network instead; longer runs are re-queued for the host-driven fallback.
constexpr int kSubBits = 10;
constexpr int kSubRankMax = 96;

template <typename KeyT>
__global__ void __launch_bounds__(256)
sort_big_kernel(const KeyT* __restrict__ tmp, const uint32_t* __restrict__ offs,
                KeyT* __restrict__ out, const uint32_t* __restrict__ big_list,
                const uint32_t* __restrict__ big_count, uint32_t* __restrict__ huge_list,
                uint32_t* __restrict__ huge_count, uint32_t* __restrict__ dup_flag, int rem_bits) {
  extern __shared__ __align__(16) unsigned char sort_big_smem[];
  KeyT* sk = reinterpret_cast<KeyT*>(sort_big_smem);                     // the run as loaded
  KeyT* sg = sk + kSmemSortMax;                                          // grouped by sub-bin
  uint32_t* cnt = reinterpret_cast<uint32_t*>(sg + kSmemSortMax);        // [2^kSubBits + 1] counts -> starts
  uint32_t* cur = cnt + (1 << kSubBits) + 1;                             // [2^kSubBits] fill cursors
  __shared__ uint32_t s_wsum[8];
  __shared__ int s_long, s_dup;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int shift = rem_bits > kSubBits ? rem_bits - kSubBits : 0;
  const uint32_t bmask = (1u << kSubBits) - 1u;
  const uint32_t nbig = *big_count;
  for (uint32_t t = blockIdx.x; t < nbig; t += gridDim.x) {
    const uint32_t x = big_list[t];
    const uint32_t a = offs[x], L = offs[x + 1] - a;
    if (L > kSmemSortMax) {
      if (tid == 0) huge_list[atomicAdd(huge_count, 1u)] = x;
      continue;
    }
    for (int i = tid; i <= (1 << kSubBits); i += 256) cnt[i] = 0;
    if (tid == 0) { s_long = 0; s_dup = 0; }
    __syncthreads();
    for (uint32_t i = tid; i < L; i += 256) {
      const KeyT k = tmp[a + i];
      sk[i] = k;
      atomicAdd(&cnt[(uint32_t)(k >> shift) & bmask], 1u);
    }
    __syncthreads();
    // exclusive scan of the 1024 counts: 4 per thread, warp scan, warp totals
    {
      uint32_t v[4], s = 0;
#pragma unroll
      for (int q = 0; q < 4; q++) { v[q] = cnt[tid * 4 + q]; s += v[q]; }
      if (max(max(v[0], v[1]), max(v[2], v[3])) > (uint32_t)kSubRankMax) s_long = 1;
      uint32_t inc = s;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t u = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += u;
      }
      if (lane == 31) s_wsum[warp] = inc;
      __syncthreads();
      uint32_t pre = inc - s;
      for (int w = 0; w < warp; w++) pre += s_wsum[w];
#pragma unroll
      for (int q = 0; q < 4; q++) { cnt[tid * 4 + q] = pre; cur[tid * 4 + q] = pre; pre += v[q]; }
      if (tid == 255) cnt[1 << kSubBits] = pre;
    }
    __syncthreads();
    if (s_long) {
      // bitonic network over the padded run (INF padding sorts last; real keys equal to INF are still
      // among the first L)
      uint32_t P = 1;
      while (P < L) P <<= 1;
      const KeyT INF = (KeyT)~(KeyT)0;
      for (uint32_t i = L + tid; i < P; i += 256) sk[i] = INF;
      __syncthreads();
      for (uint32_t k = 2; k <= P; k <<= 1) {
        for (uint32_t j = k >> 1; j > 0; j >>= 1) {
          for (uint32_t i = tid; i < P; i += 256) {
            const uint32_t ixj = i ^ j;
            if (ixj > i) {
              const KeyT u = sk[i], w = sk[ixj];
              const bool up = (i & k) == 0;
              if ((u > w) == up) { sk[i] = w; sk[ixj] = u; }
            }
          }
          __syncthreads();
        }
      }
      for (uint32_t i = tid; i < L; i += 256) {
        out[a + i] = sk[i];
        if (i + 1 < L && sk[i] == sk[i + 1]) s_dup = 1;
      }
    } else {
      for (uint32_t i = tid; i < L; i += 256) {
        const KeyT k = sk[i];
        sg[atomicAdd(&cur[(uint32_t)(k >> shift) & bmask], 1u)] = k;
      }
      __syncthreads();
      for (uint32_t i = tid; i < L; i += 256) {
        const KeyT k = sg[i];
        const uint32_t bin = (uint32_t)(k >> shift) & bmask;
        const uint32_t b0 = cnt[bin], b1 = cnt[bin + 1];
        uint32_t r = b0;
        bool dup = false;
        for (uint32_t j = b0; j < b1; j++) {
          const KeyT kj = sg[j];
          r += (kj < k) || (kj == k && j < i);
          dup |= kj == k && j != i;
        }
        out[a + r] = k;
        if (dup) s_dup = 1;
      }
    }
    __syncthreads();
    if (tid == 0 && s_dup) *dup_flag = 1u;
  }
}

// fallback for one huge run: bitonic sort in global memory over a padded scratch
template <typename KeyT>
__global__ void bitonic_step_kernel(KeyT* __restrict__ d, uint32_t P, uint32_t k, uint32_t j) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P) return;
  const uint32_t ixj = i ^ j;
  if (ixj > i) {
    const KeyT u = d[i], w = d[ixj];
    const bool up = (i & k) == 0;
    if ((u > w) == up) { d[i] = w; d[ixj] = u; }
  }
}

template <typename KeyT>
__global__ void pad_copy_kernel(const KeyT* __restrict__ src, uint32_t L, KeyT* __restrict__ dst, uint32_t P) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < P) dst[i] = i < L ? src[i] : (KeyT)~(KeyT)0;
}

// thread per fine bucket over the SORTED run: distinct keys and keys kept by the
// cutoff (count = min(255, multiplicity) >= cutoff)
template <typename KeyT>
__global__ void unique_count_kernel(const KeyT* __restrict__ srt, const uint32_t* __restrict__ offs, uint32_t NF,
                                    int cutoff, uint32_t* __restrict__ n_kept,
                                    unsigned long long* __restrict__ total_distinct) {
  const uint32_t x = blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t distinct = 0, kept = 0;
  if (x < NF) {
    const uint32_t a = offs[x], b = offs[x + 1];
    uint32_t i = a;
    while (i < b) {
      const KeyT k = srt[i];
      uint32_t j = i + 1;
      while (j < b && srt[j] == k) j++;
      const uint32_t c = min(255u, j - i);
      distinct++;
      kept += (c >= (uint32_t)cutoff);
      i = j;
    }
    n_kept[x] = kept;
  }
  unsigned long long d = distinct;
  for (int o = 16; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
  if ((threadIdx.x & 31) == 0 && d) atomicAdd(total_distinct, d);
}

template <typename KeyT>
__global__ void unique_write_kernel(const KeyT* __restrict__ srt, const uint32_t* __restrict__ offs, uint32_t NF,
                                    int cutoff, const uint32_t* __restrict__ out_offs, KeyT* __restrict__ out,
                                    uint8_t* __restrict__ counts) {
  const uint32_t x = blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= NF) return;
  const uint32_t a = offs[x], b = offs[x + 1];
  uint32_t i = a, w = out_offs[x];
  while (i < b) {
    const KeyT k = srt[i];
    uint32_t j = i + 1;
    while (j < b && srt[j] == k) j++;
    const uint32_t c = min(255u, j - i);
    if (c >= (uint32_t)cutoff) {
      out[w] = k;
      if (counts) counts[w] = (uint8_t)c;
      w++;
    }
    i = j;
  }
}

template <typename KeyT>
int pipeline_t(kmsc_ctx* ctx, const PipelineInput& in, const PipelineOptions& opt, PipelineResult* res) {
  const int key_bits = 2 * opt.K - opt.N;
  const int F = key_bits < kMaxFineLevel ? key_bits : kMaxFineLevel;
  const uint32_t NF = (uint32_t)1 << (opt.N + F);
  KP kp;
  kp.words = in.d_words; kp.bad = in.d_bad; kp.n_pos = (unsigned long long)in.n_pos;
  kp.K = opt.K; kp.key_bits = key_bits; kp.fine_shift = key_bits - F; kp.canonical = opt.canonical;
  kp.bucket_lo = (unsigned long long)(opt.bucket_lo < 0 ? 0 : opt.bucket_lo);
  kp.bucket_hi = (unsigned long long)opt.bucket_hi;

  // scratch: cnt/offs [NF+1], cursor/n_kept [NF+1], bsum, big lists, counters
  const size_t sb = scan_scratch_entries(NF);
  const size_t ent = (size_t)NF + 1;
  const size_t big_cap = (size_t)in.n_pos / kRankSortMax + 2;
  const size_t words_needed = ent * 2 + sb + big_cap * 2 + 64;
  KMSC_TRY(ctx->work.reserve(words_needed * 4));
  uint32_t* p = (uint32_t*)ctx->work.p;
  uint32_t* d_offs = p; p += ent;
  uint32_t* d_aux = p; p += ent;      // cursor, later n_kept / out offsets
  uint32_t* d_bsum = p; p += sb;
  uint32_t* d_big = p; p += big_cap;
  uint32_t* d_huge = p; p += big_cap;
  p = (uint32_t*)(((uintptr_t)p + 15) & ~(uintptr_t)15);
  uint32_t* d_ctr = p;                // [0] total, [1] big_count, [2] huge_count, [4..5] distinct (u64), [6] kept total
  KMSC_CUDA(cudaMemsetAsync(d_offs, 0, ent * 4, ctx->stream));
  KMSC_CUDA(cudaMemsetAsync(d_aux, 0, ent * 4, ctx->stream));
  KMSC_CUDA(cudaMemsetAsync(d_ctr, 0, 64, ctx->stream));

  const int threads = 256;
  const unsigned pos_blocks = (unsigned)((in.n_pos + threads - 1) / threads);
  // the staged partition sort (partition.cu) covers the usual shapes; the general path below
  // (global histogram, scatter, per-run sorts) takes what it declines
  PartPlan plan;
  KMSC_TRY(partition_plan(ctx, &in, 1, opt, &plan));
  const bool fast = plan.feasible;
  void* pin = nullptr;
  KMSC_TRY(ctx_pinned(ctx, 64, &pin));
  int64_t n_occ = 0;
  if (fast) {
    n_occ = plan.n_occ[0];
  } else {
    if (in.n_pos > 0) {
      hist_kernel<<<pos_blocks, threads, 0, ctx->stream>>>(kp, d_offs);
      count_launch(ctx);
    }
    KMSC_TRY(exclusive_scan_u32(ctx, d_offs, d_offs, NF, d_bsum, d_ctr));
    KMSC_CUDA(cudaMemcpyAsync(pin, d_ctr, 4, cudaMemcpyDeviceToHost, ctx->stream));
    KMSC_CUDA(cudaStreamSynchronize(ctx->stream));
    n_occ = *(uint32_t*)pin;
  }
  res->n_occurrences = n_occ;

  kmsc_set* s_all = nullptr;      // mode 0 result; mode 1 result too when no key repeats
  KeyT* d_tmp = nullptr;
  KeyT* d_sorted = nullptr;
  if (!fast) {
    KMSC_TRY(ctx->work2.reserve((size_t)(n_occ + 8) * sizeof(KeyT)));
    d_tmp = (KeyT*)ctx->work2.p;
  }
  // dedup (mode 1) sorts straight into the result set: an SPSS spells every k-mer once, so the
  // sorted occurrences normally ARE the set and the run-length pass is skipped
  if (opt.mode == 0 || opt.mode == 1) {
    KMSC_TRY(set_alloc(ctx, opt.K, opt.N, opt.key_bytes, n_occ, &s_all));
    d_sorted = (KeyT*)s_all->keys;
  } else {
    int rc = ctx->work3.reserve((size_t)(n_occ + 8) * sizeof(KeyT));
    if (rc != KMSC_OK) return rc;
    d_sorted = (KeyT*)ctx->work3.p;
  }
  auto fail = [&](int rc) { if (s_all) kmsc_set_free(ctx, s_all); return rc; };
  bool has_repeats = false, repeats_known = true, levels_done = false;

  if (n_occ > 0 && fast) {
    void* kk = d_sorted;
    uint32_t* ff = d_offs;
    int rc_p = partition_run(ctx, &plan, &kk, &ff);
    if (rc_p != KMSC_OK) return fail(rc_p);
    if (s_all) {
      cudaError_t e = cudaMemcpyAsync(s_all->lev[s_all->max_level], d_offs, ent * 4, cudaMemcpyDeviceToDevice, ctx->stream);
      if (e != cudaSuccess) return fail(cuda_fail(e, "pipeline offsets", __FILE__, __LINE__));
      int rc_l = set_derive_levels(ctx, s_all);
      if (rc_l != KMSC_OK) return fail(rc_l);
      levels_done = true;
    }
    std::vector<int> rep;
    rc_p = partition_flags(ctx, &plan, &rep);
    if (rc_p != KMSC_OK) return fail(rc_p);
    has_repeats = rep[0] != 0;
  } else if (n_occ > 0) {
    scatter_kernel<KeyT><<<pos_blocks, threads, 0, ctx->stream>>>(kp, d_offs, d_aux, d_tmp);
    sort_runs_kernel<KeyT><<<(NF + 127) / 128, 128, 0, ctx->stream>>>(d_tmp, d_offs, NF, d_sorted, d_big, d_ctr + 1, d_ctr + 3);
    {
      const size_t smem_big = (size_t)2 * kSmemSortMax * sizeof(KeyT) + ((size_t)2 * (1 << kSubBits) + 1) * 4;
      // (per device, and cheap: set on every call)
      cudaError_t ea = cudaFuncSetAttribute(sort_big_kernel<KeyT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_big);
      if (ea != cudaSuccess) return fail(cuda_fail(ea, "sort_big attribute", __FILE__, __LINE__));
      sort_big_kernel<KeyT><<<ctx->sm_count * 3, 256, smem_big, ctx->stream>>>(d_tmp, d_offs, d_sorted, d_big, d_ctr + 1,
                                                                              d_huge, d_ctr + 2, d_ctr + 3, kp.fine_shift);
    }
    count_launch(ctx, 3);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(cuda_fail(e, "pipeline sort", __FILE__, __LINE__));
    // huge runs (> kSmemSortMax keys in one fine bucket): host-driven global bitonic sort.
    // The offset levels do not depend on the order inside a run: they are enqueued before the
    // one synchronisation that brings back the huge-run count and the repeat flag.
    if (s_all) {
      e = cudaMemcpyAsync(s_all->lev[s_all->max_level], d_offs, ent * 4, cudaMemcpyDeviceToDevice, ctx->stream);
      if (e != cudaSuccess) return fail(cuda_fail(e, "pipeline offsets", __FILE__, __LINE__));
      int rc_l = set_derive_levels(ctx, s_all);
      if (rc_l != KMSC_OK) return fail(rc_l);
      levels_done = true;
    }
    e = cudaMemcpyAsync(pin, d_ctr + 2, 8, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) return fail(cuda_fail(e, "pipeline huge count", __FILE__, __LINE__));
    const uint32_t n_huge = ((uint32_t*)pin)[0];
    has_repeats = ((uint32_t*)pin)[1] != 0 || n_huge > 0;  // huge runs are not checked: assume repeats
    repeats_known = n_huge == 0;
    if (n_huge > 0) {
      std::vector<uint32_t> huge(n_huge);
      e = cudaMemcpy(huge.data(), d_huge, (size_t)n_huge * 4, cudaMemcpyDeviceToHost);
      if (e != cudaSuccess) return fail(cuda_fail(e, "pipeline huge list", __FILE__, __LINE__));
      for (uint32_t t = 0; t < n_huge; t++) {
        uint32_t ab[2];
        e = cudaMemcpy(ab, d_offs + huge[t], 8, cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) return fail(cuda_fail(e, "pipeline huge offs", __FILE__, __LINE__));
        const uint32_t L = ab[1] - ab[0];
        uint32_t P = 1;
        while (P < L) P <<= 1;
        KeyT* d_pad = nullptr;
        e = cudaMalloc(&d_pad, (size_t)P * sizeof(KeyT));
        if (e != cudaSuccess) return fail(cuda_fail(e, "pipeline huge scratch", __FILE__, __LINE__));
        const unsigned blocks = (P + 255) / 256;
        pad_copy_kernel<KeyT><<<blocks, 256, 0, ctx->stream>>>(d_tmp + ab[0], L, d_pad, P);
        for (uint32_t k = 2; k <= P; k <<= 1)
          for (uint32_t j = k >> 1; j > 0; j >>= 1) bitonic_step_kernel<KeyT><<<blocks, 256, 0, ctx->stream>>>(d_pad, P, k, j);
        count_launch(ctx, 2);
        e = cudaMemcpyAsync(d_sorted + ab[0], d_pad, (size_t)L * sizeof(KeyT), cudaMemcpyDeviceToDevice, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        cudaFree(d_pad);
        if (e != cudaSuccess) return fail(cuda_fail(e, "pipeline huge sort", __FILE__, __LINE__));
      }
    }
  }

  if (opt.mode == 0 || (opt.mode == 1 && !has_repeats)) {
    if (!levels_done) {
      cudaError_t e = cudaMemcpyAsync(s_all->lev[s_all->max_level], d_offs, ent * 4, cudaMemcpyDeviceToDevice, ctx->stream);
      if (e != cudaSuccess) return fail(cuda_fail(e, "pipeline offsets", __FILE__, __LINE__));
      int rc = set_derive_levels(ctx, s_all);
      if (rc != KMSC_OK) return fail(rc);
    }
    cudaError_t e = cudaStreamSynchronize(ctx->stream);  // immediate unless a huge run was sorted after the levels
    if (e != cudaSuccess) return fail(cuda_fail(e, "pipeline finish", __FILE__, __LINE__));
    s_all->has_dups = opt.mode == 1 ? 0 : (!repeats_known ? -1 : has_repeats ? 1 : 0);
    res->set = s_all;
    res->n_distinct = opt.mode == 1 ? n_occ : -1;
    return KMSC_OK;
  }

  // modes 1 / 2: distinct keys (+ counts, cutoff)
  const int cutoff = opt.mode == 2 ? opt.cutoff : 1;
  unsigned long long* d_distinct = (unsigned long long*)(d_ctr + 4);
  KMSC_CUDA(cudaMemsetAsync(d_aux, 0, ent * 4, ctx->stream));
  unique_count_kernel<KeyT><<<(NF + 127) / 128, 128, 0, ctx->stream>>>(d_sorted, d_offs, NF, cutoff, d_aux, d_distinct);
  count_launch(ctx);
  KMSC_TRY(exclusive_scan_u32(ctx, d_aux, d_aux, NF, d_bsum, d_ctr + 6));
  KMSC_CUDA(cudaMemcpyAsync(pin, d_ctr, 32, cudaMemcpyDeviceToHost, ctx->stream));
  KMSC_CUDA(cudaStreamSynchronize(ctx->stream));
  const uint32_t* hc = (const uint32_t*)pin;
  const int64_t n_kept = hc[6];
  unsigned long long nd;
  memcpy(&nd, hc + 4, 8);
  res->n_distinct = (int64_t)nd;

  kmsc_set* s = nullptr;
  {
    int rc_alloc = set_alloc(ctx, opt.K, opt.N, opt.key_bytes, n_kept, &s);
    if (rc_alloc != KMSC_OK) return fail(rc_alloc);
  }
  uint8_t* d_counts = nullptr;
  if (opt.mode == 2) {
    cudaError_t e = cudaMalloc(&d_counts, (size_t)n_kept + 16);
    if (e != cudaSuccess) { kmsc_set_free(ctx, s); return cuda_fail(e, "cudaMalloc counts", __FILE__, __LINE__); }
  }
  unique_write_kernel<KeyT><<<(NF + 127) / 128, 128, 0, ctx->stream>>>(d_sorted, d_offs, NF, cutoff, d_aux,
                                                                      (KeyT*)s->keys, d_counts);
  count_launch(ctx);
  cudaError_t e = cudaMemcpyAsync(s->lev[s->max_level], d_aux, ent * 4, cudaMemcpyDeviceToDevice, ctx->stream);
  int rc = e == cudaSuccess ? set_derive_levels(ctx, s) : cuda_fail(e, "pipeline offsets", __FILE__, __LINE__);
  if (rc == KMSC_OK) {
    e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) rc = cuda_fail(e, "pipeline finish", __FILE__, __LINE__);
  }
  if (s_all) { kmsc_set_free(ctx, s_all); s_all = nullptr; }  // mode 1 with repeats: it was the sort buffer
  if (rc != KMSC_OK) { kmsc_set_free(ctx, s); if (d_counts) cudaFree(d_counts); return rc; }
  s->has_dups = 0;
  res->set = s;
  res->d_counts = d_counts;
  return KMSC_OK;
}

// distinct keys of a set whose buckets are sorted but hold repeats (GetKmerSetFromSPSS's hash-set
// insert, lib/core/spss.h:1925): run-length pass over every finest fine bucket into a new set
template <typename KeyT>
int dedup_t(kmsc_ctx* ctx, const kmsc_set* src, kmsc_set** out) {
  const int F = src->max_level;
  const uint32_t NF = (uint32_t)1 << (src->N + F);
  const size_t ent = (size_t)NF + 1;
  const size_t sb = scan_scratch_entries(NF);
  KMSC_TRY(ctx->work.reserve((ent + sb + 64) * 4));
  uint32_t* d_aux = (uint32_t*)ctx->work.p;
  uint32_t* d_bsum = d_aux + ent;
  uint32_t* d_ctr = (uint32_t*)(((uintptr_t)(d_bsum + sb) + 15) & ~(uintptr_t)15);  // [0..1] distinct (u64), [2] kept total
  KMSC_CUDA(cudaMemsetAsync(d_aux, 0, ent * 4, ctx->stream));
  KMSC_CUDA(cudaMemsetAsync(d_ctr, 0, 32, ctx->stream));
  unique_count_kernel<KeyT><<<(NF + 127) / 128, 128, 0, ctx->stream>>>((const KeyT*)src->keys, src->lev[F], NF, 1, d_aux,
                                                                      (unsigned long long*)d_ctr);
  count_launch(ctx);
  KMSC_TRY(exclusive_scan_u32(ctx, d_aux, d_aux, NF, d_bsum, d_ctr + 2));
  void* pin = nullptr;
  KMSC_TRY(ctx_pinned(ctx, 64, &pin));
  KMSC_CUDA(cudaMemcpyAsync(pin, d_ctr, 16, cudaMemcpyDeviceToHost, ctx->stream));
  KMSC_CUDA(cudaStreamSynchronize(ctx->stream));
  const int64_t n_kept = ((const uint32_t*)pin)[2];
  kmsc_set* s = nullptr;
  KMSC_TRY(set_alloc(ctx, src->K, src->N, src->key_bytes, n_kept, &s));
  unique_write_kernel<KeyT><<<(NF + 127) / 128, 128, 0, ctx->stream>>>((const KeyT*)src->keys, src->lev[F], NF, 1, d_aux,
                                                                      (KeyT*)s->keys, nullptr);
  count_launch(ctx);
  cudaError_t e = cudaMemcpyAsync(s->lev[F], d_aux, ent * 4, cudaMemcpyDeviceToDevice, ctx->stream);
  int rc = e == cudaSuccess ? set_derive_levels(ctx, s) : cuda_fail(e, "dedup offsets", __FILE__, __LINE__);
  if (rc == KMSC_OK) {
    e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) rc = cuda_fail(e, "dedup finish", __FILE__, __LINE__);
  }
  if (rc != KMSC_OK) { kmsc_set_free(ctx, s); return rc; }
  s->has_dups = 0;
  s->b_lo = src->b_lo; s->b_hi = src->b_hi;
  *out = s;
  return KMSC_OK;
}

}  // namespace

int dedup_sorted_set(kmsc_ctx* ctx, const kmsc_set* src, kmsc_set** out) {
  switch (src->key_bytes) {
    case 2: return dedup_t<uint16_t>(ctx, src, out);
    case 4: return dedup_t<uint32_t>(ctx, src, out);
    default: return dedup_t<unsigned long long>(ctx, src, out);
  }
}

int run_kmer_pipeline(kmsc_ctx* ctx, const PipelineInput& in, const PipelineOptions& opt, PipelineResult* res) {
  if (in.n_pos >= ((int64_t)1 << 32) - 64) { set_error("input too long for one pass (%lld bases)", (long long)in.n_pos); return KMSC_E_INVALID; }
  switch (opt.key_bytes) {
    case 2: return pipeline_t<uint16_t>(ctx, in, opt, res);
    case 4: return pipeline_t<uint32_t>(ctx, in, opt, res);
    case 8: return pipeline_t<unsigned long long>(ctx, in, opt, res);
    default: set_error("key_bytes must be 2, 4 or 8"); return KMSC_E_INVALID;
  }
}

}  // namespace kmsc
