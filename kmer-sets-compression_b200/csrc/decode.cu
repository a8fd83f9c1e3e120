// decode.cu -- P2: SPSS text -> device CSR set.
//
// Replaces KmerSetCompact::GetSampledKmerSet (reference
// lib/core/kmer_set_compact.h:120-203: every position of every string -> k-mer ->
// canonical -> (bucket, key), kept if the bucket is selected, each bucket sorted)
// and KmerSetCompact::ToKmerSet / GetKmerSetFromSPSS (lib/core/spss.h:1861-1941,
// the same decode into hash sets). The reference re-materialises ASCII strings
// and builds each k-mer from a fresh substr (kmer.h:22-46); here the text is
// packed once to 2 bits per base (the KmerSetCompact layout, kmer_set_compact.h:
// 206-255, first base in the top bits) and every position extracts its k-mer
// with two 64-bit loads.
#include "kmer_pipeline.cuh"

namespace kmsc {
namespace {

// thread per 32 bases: ASCII -> one 64-bit word of 2-bit codes; flags bad characters
__global__ void pack_ascii_kernel(const unsigned char* __restrict__ text, unsigned long long n,
                                  unsigned long long* __restrict__ words, unsigned long long n_words,
                                  int* __restrict__ bad_char) {
  const unsigned long long w = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= n_words) return;
  unsigned long long out = 0;
  const unsigned long long base = w * 32;
  int bad = 0;
#pragma unroll 8
  for (int j = 0; j < 32; j++) {
    const unsigned long long p = base + j;
    unsigned c = 0;
    if (p < n) {
      const unsigned ch = text[p];
      // A 0x41, C 0x43, G 0x47, T 0x54: (ch >> 1) & 3 = 0, 1, 3, 2; fix the G/T swap
      c = (ch >> 1) & 3u;
      c ^= c >> 1;
      bad |= !(ch == 'A' || ch == 'C' || ch == 'G' || ch == 'T');
    }
    out = (out << 2) | c;
  }
  words[w] = out;
  if (bad) atomicExch(bad_char, 1);
}

// thread per string: no k-mer starts in the last K-1 positions of a string
// (kmer_set_compact.h:149: j < length - K + 1)
__global__ void mark_string_tails_kernel(const long long* __restrict__ str_offs, long long n_strings, int K,
                                         uint32_t* __restrict__ bad) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_strings) return;
  const long long b = str_offs[i], e = str_offs[i + 1];
  long long a = e - (K - 1);
  if (a < b) a = b;
  for (long long p = a; p < e; p++) atomicOr(&bad[p >> 5], 1u << (p & 31));
}

// the same for a batch of jobs (blockIdx.y = job)
struct TailJob { const long long* str_offs; long long n_strings; uint32_t* bad; };
__global__ void mark_string_tails_batch_kernel(const TailJob* __restrict__ jobs, int K) {
  const TailJob jb = jobs[blockIdx.y];
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < jb.n_strings; i += (long long)gridDim.x * blockDim.x) {
    const long long b = jb.str_offs[i], e = jb.str_offs[i + 1];
    long long a = e - (K - 1);
    if (a < b) a = b;
    for (long long p = a; p < e; p++) atomicOr(&jb.bad[p >> 5], 1u << (p & 31));
  }
}

// ---- SPSS construction support: the de Bruijn neighbours of every k-mer of a set -----------
// out[8 i + c] (c in 0..3): the k-mer obtained by dropping the first base of k-mer i and
// appending base c (reference Kmer::Next, lib/core/kmer.h:136-160); out[8 i + 4 + c]: dropping
// the last base and prepending c (Kmer::Prev, :163-186). An entry is -1 if that k-mer (its
// canonical form when `canonical`) is not in the set, else (index << 1) | flip where index is
// its position in the set's key order and flip = 1 if the set stores its reverse complement.
// This replaces the hash-set Contains() calls that dominate the reference's unitig / path-cover
// construction (lib/core/spss.h:230-615, 1039-1858) by one binary search per neighbour.
template <typename KeyT>
__global__ void neighbors_kernel(const KeyT* __restrict__ keys, const uint32_t* __restrict__ offs, int n_buckets,
                                 int K, int key_bits, int canonical, int32_t* __restrict__ out) {
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
        unsigned long long w = d < 4 ? (((v << 2) & mask) | c) : ((v >> 2) | (c << (2 * (K - 1))));
        int flip = 0;
        if (canonical) {
          const unsigned long long rc = revcomp(w, K);
          if (rc < w) { w = rc; flip = 1; }
        }
        const uint32_t bq = (uint32_t)(w >> key_bits);
        const unsigned long long kq = w & kmask;
        uint32_t a = offs[bq], e = offs[bq + 1];
        while (a < e) {
          const uint32_t mid = (a + e) >> 1;
          if ((unsigned long long)keys[mid] < kq) a = mid + 1; else e = mid;
        }
        const bool found = a < offs[bq + 1] && (unsigned long long)keys[a] == kq;
        out[(size_t)i * 8 + d] = found ? (int32_t)((a << 1) | (uint32_t)flip) : -1;
      }
    }
  }
}

}  // namespace
}  // namespace kmsc

using namespace kmsc;

static int spss_common(kmsc_ctx* ctx, int K, int N, int key_bytes, const char* text, const uint64_t* packed,
                       const int64_t* str_offs, int64_t n_strings, int canonical, int dedup,
                       int32_t bucket_lo, int32_t bucket_hi, kmsc_set** out);

extern "C" int kmsc_set_from_spss(kmsc_ctx* ctx, int K, int N, int key_bytes, const char* text,
                                  const int64_t* str_offs, int64_t n_strings, int canonical, int dedup,
                                  int32_t bucket_lo, int32_t bucket_hi, kmsc_set** out) {
  return spss_common(ctx, K, N, key_bytes, text, nullptr, str_offs, n_strings, canonical, dedup, bucket_lo, bucket_hi, out);
}

extern "C" int kmsc_set_from_packed(kmsc_ctx* ctx, int K, int N, int key_bytes, const uint64_t* words,
                                    const int64_t* str_offs, int64_t n_strings, int canonical, int dedup,
                                    int32_t bucket_lo, int32_t bucket_hi, kmsc_set** out) {
  if (!words && str_offs && n_strings > 0 && str_offs[n_strings] > 0) { set_error("words is NULL"); return KMSC_E_INVALID; }
  static const uint64_t zero[2] = {0, 0};
  return spss_common(ctx, K, N, key_bytes, nullptr, words ? words : zero, str_offs, n_strings, canonical, dedup,
                     bucket_lo, bucket_hi, out);
}

static int spss_common(kmsc_ctx* ctx, int K, int N, int key_bytes, const char* text, const uint64_t* packed,
                       const int64_t* str_offs, int64_t n_strings, int canonical, int dedup,
                       int32_t bucket_lo, int32_t bucket_hi, kmsc_set** out) {
  if (!ctx || !out || !str_offs || n_strings < 0) { set_error("bad argument"); return KMSC_E_INVALID; }
  if (K < 1 || K > 32 || N < 0 || N > 24 || N > 2 * K || 2 * K - N > 8 * key_bytes || 2 * K - N >= 64 ||
      (key_bytes != 2 && key_bytes != 4 && key_bytes != 8)) { set_error("bad K/N/key_bytes"); return KMSC_E_INVALID; }
  if (bucket_lo < 0 || bucket_hi > (1 << N) || bucket_lo > bucket_hi) { set_error("bad bucket range"); return KMSC_E_INVALID; }
  if (str_offs[0] != 0) { set_error("str_offs[0] must be 0"); return KMSC_E_INVALID; }
  for (int64_t i = 0; i < n_strings; i++)
    if (str_offs[i + 1] < str_offs[i]) { set_error("str_offs not monotone at %lld", (long long)i); return KMSC_E_INVALID; }
  const int64_t n = str_offs[n_strings];
  if (n > 0 && !text && !packed) { set_error("text is NULL"); return KMSC_E_INVALID; }
  KMSC_CUDA(cudaSetDevice(ctx->device));

  // device staging: text | str_offs | words | bad bits | flag   (ctx->work3 is free until the pipeline's mode>0 sort)
  const size_t n_words = (size_t)(n + 31) / 32 + 2;
  const size_t n_badw = (size_t)(n + 31) / 32 + 2;
  size_t off = 0;
  const size_t o_text = off; off += ((size_t)n + 255) & ~(size_t)255;
  const size_t o_offs = off; off += (((size_t)n_strings + 1) * 8 + 255) & ~(size_t)255;
  const size_t o_words = off; off += (n_words * 8 + 255) & ~(size_t)255;
  const size_t o_bad = off; off += (n_badw * 4 + 255) & ~(size_t)255;
  const size_t o_flag = off; off += 256;
  KMSC_TRY(ctx->stage.reserve(off));
  unsigned char* base = (unsigned char*)ctx->stage.p;
  unsigned char* d_text = base + o_text;
  long long* d_offs = (long long*)(base + o_offs);
  unsigned long long* d_words = (unsigned long long*)(base + o_words);
  uint32_t* d_bad = (uint32_t*)(base + o_bad);
  int* d_flag = (int*)(base + o_flag);

  if (packed) {
    KMSC_CUDA(cudaMemcpyAsync(d_words, packed, ((size_t)(n + 31) / 32) * 8, cudaMemcpyHostToDevice, ctx->stream));
    KMSC_CUDA(cudaMemsetAsync(d_words + (n + 31) / 32, 0, 16, ctx->stream));
  } else if (n > 0) {
    KMSC_CUDA(cudaMemcpyAsync(d_text, text, (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
  }
  KMSC_CUDA(cudaMemcpyAsync(d_offs, str_offs, ((size_t)n_strings + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
  KMSC_CUDA(cudaMemsetAsync(d_bad, 0, n_badw * 4, ctx->stream));
  KMSC_CUDA(cudaMemsetAsync(d_flag, 0, 4, ctx->stream));
  if (!packed)
    pack_ascii_kernel<<<(unsigned)((n_words + 127) / 128), 128, 0, ctx->stream>>>(d_text, (unsigned long long)n, d_words,
                                                                                n_words, d_flag);
  if (n_strings > 0)
    mark_string_tails_kernel<<<(unsigned)((n_strings + 127) / 128), 128, 0, ctx->stream>>>(d_offs, n_strings, K, d_bad);
  count_launch(ctx, 2);
  KMSC_CUDA(cudaGetLastError());
  if (!packed) {
    // only ASCII input can hold a bad character (2-bit words cannot)
    void* pin = nullptr;
    KMSC_TRY(ctx_pinned(ctx, 64, &pin));
    KMSC_CUDA(cudaMemcpyAsync(pin, d_flag, 4, cudaMemcpyDeviceToHost, ctx->stream));
    KMSC_CUDA(cudaStreamSynchronize(ctx->stream));
    if (*(int*)pin) { set_error("SPSS text holds a character other than A, C, G, T"); return KMSC_E_FORMAT; }
  }

  PipelineInput in{d_words, d_bad, n};
  PipelineOptions opt{K, N, key_bytes, canonical, bucket_lo, bucket_hi, dedup ? 1 : 0, 1};
  PipelineResult res;
  KMSC_TRY(run_kmer_pipeline(ctx, in, opt, &res));
  if (res.set && (bucket_lo > 0 || bucket_hi < (1 << N))) { res.set->b_lo = bucket_lo; res.set->b_hi = bucket_hi; }
  *out = res.set;
  return KMSC_OK;
}

// Batched decode: m packed SPSS -> m device sets with ONE launch sequence per group of sets
// (count, scan, partition, sort, levels) and two host synchronisations per group instead of
// three per set. The host-to-device copies of all groups are queued first on a copy stream,
// so the copies of the later groups overlap the kernels of the earlier ones.
extern "C" int kmsc_sets_from_packed_batch(kmsc_ctx* ctx, int K, int N, int key_bytes, int32_t m,
                                           const uint64_t* const* words, const int64_t* const* str_offs,
                                           const int64_t* n_strings, int canonical, int dedup,
                                           int32_t bucket_lo, int32_t bucket_hi, kmsc_set** out) {
  if (!ctx || !out || m < 0 || (m > 0 && (!words || !str_offs || !n_strings))) { set_error("bad argument"); return KMSC_E_INVALID; }
  if (K < 1 || K > 32 || N < 0 || N > 24 || N > 2 * K || 2 * K - N > 8 * key_bytes || 2 * K - N >= 64 ||
      (key_bytes != 2 && key_bytes != 4 && key_bytes != 8)) { set_error("bad K/N/key_bytes"); return KMSC_E_INVALID; }
  if (bucket_lo < 0 || bucket_hi > (1 << N) || bucket_lo > bucket_hi) { set_error("bad bucket range"); return KMSC_E_INVALID; }
  for (int32_t j = 0; j < m; j++) {
    out[j] = nullptr;
    if (!str_offs[j] || n_strings[j] < 0 || str_offs[j][0] != 0) { set_error("job %d: bad str_offs", j); return KMSC_E_INVALID; }
    for (int64_t i = 0; i < n_strings[j]; i++)
      if (str_offs[j][i + 1] < str_offs[j][i]) { set_error("job %d: str_offs not monotone at %lld", j, (long long)i); return KMSC_E_INVALID; }
    const int64_t n = str_offs[j][n_strings[j]];
    if (n > 0 && !words[j]) { set_error("job %d: words is NULL", j); return KMSC_E_INVALID; }
    if (n >= ((int64_t)1 << 32) - 64) { set_error("job %d: input too long for one pass (%lld bases)", j, (long long)n); return KMSC_E_INVALID; }
  }
  if (m == 0) return KMSC_OK;
  KMSC_CUDA(cudaSetDevice(ctx->device));
  if (!ctx->copy_stream) {
    KMSC_CUDA(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
    for (auto& e : ctx->copy_ev) KMSC_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    KMSC_CUDA(cudaEventCreateWithFlags(&ctx->fence_ev, cudaEventDisableTiming));
  }

  // device staging of every job: words | bad bits | string offsets
  std::vector<size_t> o_words((size_t)m), o_bad((size_t)m), o_offs((size_t)m);
  std::vector<int64_t> npos((size_t)m);
  size_t off = 0;
  for (int32_t j = 0; j < m; j++) {
    const int64_t n = str_offs[j][n_strings[j]];
    npos[(size_t)j] = n;
    const size_t nw = (size_t)(n + 31) / 32;
    o_words[(size_t)j] = off; off += ((nw + 2) * 8 + 255) & ~(size_t)255;
    o_bad[(size_t)j] = off; off += ((nw + 2) * 4 + 255) & ~(size_t)255;
    o_offs[(size_t)j] = off; off += (((size_t)n_strings[j] + 1) * 8 + 255) & ~(size_t)255;
  }
  KMSC_TRY(ctx->stage.reserve(off));
  unsigned char* base = (unsigned char*)ctx->stage.p;

  // groups: at most kGroups per call (one event each), bounded by a position budget
  constexpr int kGroups = kmsc_ctx::kP2Slots;
  int per_group = (m + kGroups - 1) / kGroups;
  if (const char* e = getenv("KMSC_P2_GROUP")) per_group = std::max(1, atoi(e));
  per_group = std::max(per_group, (m + kGroups - 1) / kGroups);
  std::vector<std::pair<int, int>> groups;
  for (int a = 0; a < m; a += per_group) groups.emplace_back(a, std::min(m, a + per_group));

  // the copy stream must not overwrite staging that earlier work on the main stream still reads
  KMSC_CUDA(cudaEventRecord(ctx->fence_ev, ctx->stream));
  KMSC_CUDA(cudaStreamWaitEvent(ctx->copy_stream, ctx->fence_ev, 0));
  void* tab = nullptr;
  KMSC_TRY(ctx->tab[1].acquire((size_t)m * sizeof(TailJob), &tab));
  TailJob* ht = (TailJob*)tab;
  int64_t max_strings = 0;
  for (size_t gi = 0; gi < groups.size(); gi++) {
    for (int j = groups[gi].first; j < groups[gi].second; j++) {
      const int64_t n = npos[(size_t)j];
      const size_t nw = (size_t)(n + 31) / 32;
      unsigned long long* dw = (unsigned long long*)(base + o_words[(size_t)j]);
      if (nw) KMSC_CUDA(cudaMemcpyAsync(dw, words[j], nw * 8, cudaMemcpyHostToDevice, ctx->copy_stream));
      KMSC_CUDA(cudaMemsetAsync(dw + nw, 0, 16, ctx->copy_stream));
      KMSC_CUDA(cudaMemsetAsync(base + o_bad[(size_t)j], 0, (nw + 2) * 4, ctx->copy_stream));
      KMSC_CUDA(cudaMemcpyAsync(base + o_offs[(size_t)j], str_offs[j], ((size_t)n_strings[j] + 1) * 8, cudaMemcpyHostToDevice, ctx->copy_stream));
      ht[j].str_offs = (const long long*)(base + o_offs[(size_t)j]);
      ht[j].n_strings = n_strings[j];
      ht[j].bad = (uint32_t*)(base + o_bad[(size_t)j]);
      max_strings = std::max<int64_t>(max_strings, n_strings[j]);
    }
    KMSC_CUDA(cudaEventRecord(ctx->copy_ev[gi], ctx->copy_stream));
  }
  KMSC_TRY(ctx->tabs_dev[1].reserve((size_t)m * sizeof(TailJob)));
  TailJob* d_tail = (TailJob*)ctx->tabs_dev[1].p;
  KMSC_CUDA(cudaMemcpyAsync(d_tail, ht, (size_t)m * sizeof(TailJob), cudaMemcpyHostToDevice, ctx->stream));
  KMSC_TRY(ctx->tab[1].commit(ctx->stream));

  PipelineOptions opt{K, N, key_bytes, canonical, bucket_lo, bucket_hi, dedup ? 1 : 0, 1};
  auto cleanup = [&]() { for (int32_t j = 0; j < m; j++) if (out[j]) { kmsc_set_free(ctx, out[j]); out[j] = nullptr; } };
  std::vector<int> slow;  // jobs that go through the general pipeline (shape declined, or repeats under dedup)
  struct Pending { PartPlan plan; int a, b; };
  std::vector<Pending> pend(groups.size());
  for (size_t gi = 0; gi < groups.size(); gi++) {
    const int a = groups[gi].first, b = groups[gi].second, gm = b - a;
    KMSC_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->copy_ev[gi], 0));
    if (max_strings > 0) {
      unsigned gx = (unsigned)std::min<int64_t>((max_strings + 127) / 128, 1024);
      mark_string_tails_batch_kernel<<<dim3(gx, (unsigned)gm), 128, 0, ctx->stream>>>(d_tail + a, K);
      count_launch(ctx);
    }
    std::vector<PipelineInput> in((size_t)gm);
    for (int j = a; j < b; j++)
      in[(size_t)(j - a)] = PipelineInput{(const unsigned long long*)(base + o_words[(size_t)j]), (const uint32_t*)(base + o_bad[(size_t)j]), npos[(size_t)j]};
    Pending& pd = pend[gi];
    pd.a = a; pd.b = b;
    // every group needs its own job table while its kernels are in flight: plan + run back to back
    int rc = partition_plan(ctx, in.data(), gm, opt, &pd.plan, (int)gi);
    if (rc != KMSC_OK) { cleanup(); return rc; }
    if (!pd.plan.feasible) {
      for (int j = a; j < b; j++) slow.push_back(j);
      continue;
    }
    pd.plan.dedup_in_sort = dedup != 0;
    std::vector<void*> dk((size_t)gm);
    std::vector<uint32_t*> df((size_t)gm);
    for (int j = a; j < b; j++) {
      rc = set_alloc(ctx, K, N, key_bytes, pd.plan.n_occ[(size_t)(j - a)], &out[j]);
      if (rc != KMSC_OK) { cleanup(); return rc; }
      dk[(size_t)(j - a)] = out[j]->keys;
      df[(size_t)(j - a)] = out[j]->lev[out[j]->max_level];
    }
    rc = partition_run(ctx, &pd.plan, dk.data(), df.data());
    // the repeat flags are looked at once everything is queued: copy them out now
    if (rc == KMSC_OK) rc = partition_flags_async(ctx, &pd.plan);
    if (rc != KMSC_OK) { cleanup(); return rc; }
  }
  {
    cudaError_t e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) { cleanup(); return cuda_fail(e, "batch decode", __FILE__, __LINE__); }
  }
  for (size_t gi = 0; gi < groups.size(); gi++) {
    Pending& pd = pend[gi];
    if (!pd.plan.feasible) continue;
    for (int j = pd.a; j < pd.b; j++) {
      const uint32_t f = pd.plan.flags_host[(size_t)(j - pd.a) * 4 + 2];
      if (f & 2u) { cleanup(); set_error("partition sort: a partition outgrew its shared-memory tile"); return KMSC_E_STATE; }
      if (!(dedup && (f & 1u))) out[j]->has_dups = dedup ? 0 : (int)(f & 1u);
    }
    // a k-mer spelled more than once (a random 10 M-base sequence repeats a 23-mer about once): the
    // sort dropped the copies inside their partitions; one shift pass packs the partitions
    std::vector<int> jobs;
    std::vector<kmsc_set*> olds, news;
    for (int j = pd.a; j < pd.b; j++) {
      const uint32_t f = pd.plan.flags_host[(size_t)(j - pd.a) * 4 + 2];
      if (!(dedup && (f & 1u))) continue;
      const uint32_t removed = pd.plan.flags_host[(size_t)(j - pd.a) * 4 + 3];
      kmsc_set* d = nullptr;
      int rc = set_alloc(ctx, K, N, key_bytes, out[j]->n_keys - removed, &d);
      if (rc != KMSC_OK) { for (kmsc_set* x : news) kmsc_set_free(ctx, x); cleanup(); return rc; }
      d->has_dups = 0;
      jobs.push_back(j - pd.a); olds.push_back(out[j]); news.push_back(d);
    }
    if (!jobs.empty()) {
      int rc = partition_shift(ctx, &pd.plan, jobs, olds.data(), news.data());
      if (rc != KMSC_OK) { for (kmsc_set* x : news) kmsc_set_free(ctx, x); cleanup(); return rc; }
      for (size_t q = 0; q < jobs.size(); q++) {
        kmsc_set_free(ctx, olds[q]);   // stream-ordered: after the shift has read it
        out[pd.a + jobs[q]] = news[q];
      }
    }
  }
  {
    // coarser offset levels of every set built above, one launch; the sets are valid in stream
    // order (every later library call runs on the same stream)
    std::vector<kmsc_set*> built;
    for (int32_t j = 0; j < m; j++) if (out[j]) built.push_back(out[j]);
    int rc = derive_levels_batch(ctx, built.data(), (int)built.size());
    if (rc != KMSC_OK) { cleanup(); return rc; }
  }
  for (int j : slow) {
    PipelineInput in{(const unsigned long long*)(base + o_words[(size_t)j]), (const uint32_t*)(base + o_bad[(size_t)j]), npos[(size_t)j]};
    PipelineResult res;
    int rc = run_kmer_pipeline(ctx, in, opt, &res);
    if (rc != KMSC_OK) { cleanup(); return rc; }
    out[j] = res.set;
  }
  if (bucket_lo > 0 || bucket_hi < (1 << N))
    for (int32_t j = 0; j < m; j++) { out[j]->b_lo = bucket_lo; out[j]->b_hi = bucket_hi; }
  return KMSC_OK;
}

extern "C" int kmsc_set_neighbors(kmsc_ctx* ctx, const kmsc_set* set, int canonical, int32_t* out) {
  if (!ctx || !set || (set->n_keys > 0 && !out)) { set_error("NULL argument"); return KMSC_E_INVALID; }
  if (set->n_keys >= ((int64_t)1 << 30)) { set_error("set too large for a neighbour table (%lld keys)", (long long)set->n_keys); return KMSC_E_INVALID; }
  if (set->n_keys == 0) return KMSC_OK;
  KMSC_CUDA(cudaSetDevice(ctx->device));
  KMSC_TRY(ctx->work3.reserve((size_t)set->n_keys * 8 * sizeof(int32_t)));
  int32_t* d_out = (int32_t*)ctx->work3.p;
  const int nb = 1 << set->N;
  int blocks = (nb + 7) / 8;
  if (blocks > ctx->sm_count * 16) blocks = ctx->sm_count * 16;
  switch (set->key_bytes) {
    case 2: neighbors_kernel<uint16_t><<<blocks, 256, 0, ctx->stream>>>((const uint16_t*)set->keys, set->lev[0], nb, set->K, set->key_bits, canonical, d_out); break;
    case 4: neighbors_kernel<uint32_t><<<blocks, 256, 0, ctx->stream>>>((const uint32_t*)set->keys, set->lev[0], nb, set->K, set->key_bits, canonical, d_out); break;
    default: neighbors_kernel<unsigned long long><<<blocks, 256, 0, ctx->stream>>>((const unsigned long long*)set->keys, set->lev[0], nb, set->K, set->key_bits, canonical, d_out); break;
  }
  count_launch(ctx);
  KMSC_CUDA(cudaGetLastError());
  KMSC_CUDA(cudaMemcpyAsync(out, d_out, (size_t)set->n_keys * 8 * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
  KMSC_CUDA(cudaStreamSynchronize(ctx->stream));
  return KMSC_OK;
}
