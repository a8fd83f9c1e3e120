// kmer_pipeline.cuh -- shared device pipeline: 2-bit packed bases -> k-mers ->
// fine-bucket partition -> per-bucket sort -> (optional) count / cutoff / dedup
// -> device CSR set. Used by P2 (SPSS decode, decode.cu) and P1 (counting,
// count.cu).
#pragma once
#include "kmsc_common.cuh"

namespace kmsc {

struct PipelineInput {
  const unsigned long long* d_words;  // 2-bit codes, 32 bases per word, first base in the top bits
  const uint32_t* d_bad;              // bit p set = no k-mer starts at position p
  int64_t n_pos;                      // number of base positions
};

struct PipelineOptions {
  int K, N, key_bytes;
  int canonical;
  int32_t bucket_lo, bucket_hi;  // keep buckets in [lo, hi)
  int mode;                      // 0 keep duplicates, 1 dedup, 2 count (saturating uint8) + cutoff
  int cutoff;
};

struct PipelineResult {
  kmsc_set* set = nullptr;
  uint8_t* d_counts = nullptr;  // mode 2: counts aligned with set keys (device, cudaMalloc'd)
  int64_t n_distinct = 0;       // mode 1/2: distinct k-mers before the cutoff
  int64_t n_occurrences = 0;    // k-mer positions kept by the bucket filter
};

int run_kmer_pipeline(kmsc_ctx* ctx, const PipelineInput& in, const PipelineOptions& opt, PipelineResult* res);

// a new set holding the distinct keys of `src` (buckets sorted, repeats allowed)
int dedup_sorted_set(kmsc_ctx* ctx, const kmsc_set* src, kmsc_set** out);

// ---- staged partition sort (partition.cu): m jobs in one launch sequence -------------------------
// plan: count + scan on the device, ONE host synchronisation, then n_occ[j] (k-mer occurrences of
// job j kept by the bucket filter) is known and the caller allocates the outputs.
// run: partition + sort, asynchronous; job j's keys (ascending inside every finest-level fine
// bucket) go to d_keys[j], its finest offset level (2^(N+F)+1 entries) to d_fine[j].
// flags: one more synchronisation; repeats[j] != 0 if job j holds a k-mer more than once.
struct PartPlan {
  bool feasible = false;          // false: shape outside the fast path, use the general pipeline
  int m = 0;
  std::vector<int64_t> n_occ;
  // internal
  int B1 = 0, R1 = 0, FB = 0, FB2 = 0, tmp_bytes = 0, key_bytes = 0, n_ctas = 0;
  uint32_t part_max = 0;
  void* d_jobs = nullptr;
  void* h_jobs = nullptr;  // pinned host copy of the device job table (ctx->p2tab[slot])
  PipelineOptions opt{};
  int64_t max_pos = 0;
  int slot = 0;
  bool dedup_in_sort = false;  // set before partition_run: the sort drops repeated keys (then partition_shift)
  const uint32_t* flags_host = nullptr;  // after partition_flags_async + a stream sync: 4 words per job, [2] = flags
};
// slot: which of the context's kP2Slots table sets the plan uses (plans in flight need distinct slots)
int partition_plan(kmsc_ctx* ctx, const PipelineInput* in, int m, const PipelineOptions& opt, PartPlan* plan, int slot = 0);
int partition_run(kmsc_ctx* ctx, PartPlan* plan, void* const* d_keys, uint32_t* const* d_fine);
// queues the read-back of the jobs' flags; valid in plan->flags_host after the stream is synchronised
int partition_flags_async(kmsc_ctx* ctx, PartPlan* plan);
int partition_flags(kmsc_ctx* ctx, PartPlan* plan, std::vector<int>* repeats);
// after a run with dedup_in_sort: jobs[q] (index into the plan) dropped flags_host[4 j + 3] copies;
// packs old_sets[q] into new_sets[q] (allocated by the caller with the smaller key count)
int partition_shift(kmsc_ctx* ctx, PartPlan* plan, const std::vector<int>& jobs, kmsc_set* const* old_sets,
                    kmsc_set* const* new_sets);
// coarser offset levels of m sets (same K, N) from their finest ones, one launch
int derive_levels_batch(kmsc_ctx* ctx, kmsc_set* const* sets, int m);

// k-mer starting at base position p of the packed stream (K <= 32)
__device__ __forceinline__ unsigned long long load_kmer(const unsigned long long* __restrict__ words,
                                                        unsigned long long p, int K) {
  const unsigned long long w = p >> 5;
  const int o = (int)(p & 31) * 2;
  const unsigned long long hi = words[w];
  unsigned long long x = hi << o;
  if (o) x |= words[w + 1] >> (64 - o);
  return x >> (64 - 2 * K);
}

// reverse complement of a 2K-bit k-mer (reference lib/core/kmer.h:103-129)
__device__ __forceinline__ unsigned long long revcomp(unsigned long long v, int K) {
  unsigned long long x = __brevll(~v);
  x = ((x & 0xAAAAAAAAAAAAAAAAull) >> 1) | ((x & 0x5555555555555555ull) << 1);
  return x >> (64 - 2 * K);
}

}  // namespace kmsc
