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
  if (K < 1 || K > 32 || N < 0 || N > 2 * K || 2 * K - N > 8 * key_bytes) { set_error("bad K/N/key_bytes"); return KMSC_E_INVALID; }
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
