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
