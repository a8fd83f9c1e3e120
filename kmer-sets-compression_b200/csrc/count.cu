// count.cu -- P1: FASTA / reads -> canonical k-mer counts with cutoff -> device set.
//
// Replaces KmerCounter::FromFASTA / FromReads / ToKmerSet / Get (reference
// lib/core/kmer_counter.h:64-133, 161-264):
//   * validation (:163-203): even number of lines; even lines start with '>';
//     odd lines hold only A, C, G, T, N -- else the reference's two messages
//   * reads are split on 'N' and every K-window inside a fragment is counted
//     (:78-96), canonical if asked; counts are uint8 and saturate at 255 (:28-38)
//   * ToKmerSet(cutoff) keeps count >= cutoff, cutoff cast to uint8 (:213-243)
// The reference builds a std::string per window and hashes it into per-thread
// maps; here the bytes are classified and packed to 2 bits per base in one pass,
// windows touching an N / newline / header byte are masked, and the k-mer
// pipeline partitions, sorts and run-length counts on the device.
#include "kmer_pipeline.cuh"
#include "scan.cuh"

namespace kmsc {
namespace {

constexpr int kFlagBadChar = 1;    // "invalid FASTA file"

// The 32 bytes of a chunk come in as two 128-bit loads (a warp reads 1 KB of text with two fully used
// load instructions instead of 32 byte loads that each touch 32 sectors); a chunk that crosses the end
// of the text, or a text pointer that is not 16-byte aligned, takes byte loads.
struct Chunk32 { unsigned char b[32]; };
__device__ __forceinline__ void load_chunk32(const unsigned char* __restrict__ text, unsigned long long base,
                                             unsigned long long n, Chunk32& c) {
  if (base + 32 <= n && ((reinterpret_cast<uintptr_t>(text) + base) & 15u) == 0) {
    const uint4 lo = __ldg(reinterpret_cast<const uint4*>(text + base));
    const uint4 hi = __ldg(reinterpret_cast<const uint4*>(text + base) + 1);
    uint32_t w[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
#pragma unroll
    for (int j = 0; j < 32; j++) c.b[j] = (unsigned char)(w[j >> 2] >> (8 * (j & 3)));
  } else {
#pragma unroll
    for (int j = 0; j < 32; j++) c.b[j] = base + j < n ? text[base + j] : (unsigned char)0;
  }
}

// thread per 32 input bytes: number of '\n'
__global__ void count_newlines_kernel(const unsigned char* __restrict__ text, unsigned long long n,
                                      uint32_t* __restrict__ nl, unsigned long long n_chunks) {
  const unsigned long long t = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_chunks) return;
  const unsigned long long base = t * 32;
  uint32_t c = 0;
  if (base + 32 <= n && ((reinterpret_cast<uintptr_t>(text) + base) & 15u) == 0) {
    const uint4 lo = __ldg(reinterpret_cast<const uint4*>(text + base));
    const uint4 hi = __ldg(reinterpret_cast<const uint4*>(text + base) + 1);
    const uint32_t w[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
#pragma unroll
    for (int j = 0; j < 8; j++) {
      // bytes equal to '\n' (0x0A): zero bytes of w ^ 0x0A0A0A0A, exact per-byte test
      const uint32_t x = w[j] ^ 0x0A0A0A0Au;
      const uint32_t z = ~(((x & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | x | 0x7F7F7F7Fu);
      c += __popc(z);
    }
  } else {
#pragma unroll 8
    for (int j = 0; j < 32; j++)
      if (base + j < n) c += text[base + j] == '\n';
  }
  nl[t] = c;
}

// thread per 32 input bytes: classify, pack to 2-bit codes, build the invalid-byte mask
__global__ void classify_pack_kernel(const unsigned char* __restrict__ text, unsigned long long n,
                                     const uint32_t* __restrict__ nl_before, int fasta,
                                     unsigned long long* __restrict__ words, uint32_t* __restrict__ inv,
                                     unsigned long long n_chunks, int* __restrict__ flags) {
  const unsigned long long t = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_chunks) return;
  const unsigned long long base = t * 32;
  uint32_t line = nl_before[t];
  unsigned prev = base > 0 ? text[base - 1] : '\n';
  Chunk32 ck;
  load_chunk32(text, base, n, ck);
  unsigned long long w = 0;
  uint32_t m = 0;
  int bad = 0;
#pragma unroll
  for (int j = 0; j < 32; j++) {
    const unsigned long long p = base + j;
    unsigned code = 0;
    bool invalid = true;
    if (p < n) {
      const unsigned ch = ck.b[j];
      const bool line_start = prev == '\n';
      const bool header = fasta && !(line & 1u);
      if (ch == '\n') {
        if (header && line_start) bad = 1;  // empty header line (kmer_counter.h:179)
        line++;
      } else if (header) {
        if (line_start && ch != '>') bad = 1;
      } else if (ch == 'A' || ch == 'C' || ch == 'G' || ch == 'T') {
        code = (ch >> 1) & 3u;
        code ^= code >> 1;
        invalid = false;
      } else if (ch != 'N') {
        bad = 1;  // kmer_counter.h:186-191
      }
      prev = ch;
    }
    w = (w << 2) | code;
    m |= (invalid ? 1u : 0u) << j;
  }
  words[t] = w;
  inv[t] = m;
  if (bad) atomicOr(flags, kFlagBadChar);
}

// thread per 32 positions: bad[p] = any invalid byte in [p, p + K)
__global__ void window_mask_kernel(const uint32_t* __restrict__ inv, unsigned long long n_chunks, int K,
                                   uint32_t* __restrict__ bad) {
  const unsigned long long t = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_chunks) return;
  const unsigned long long lo = inv[t];
  const unsigned long long hi = (t + 1 < n_chunks) ? inv[t + 1] : 0xffffffffull;
  const unsigned long long win = lo | (hi << 32);
  unsigned long long acc = 0;
  for (int i = 0; i < K; i++) acc |= win >> i;
  bad[t] = (uint32_t)acc;
}

// keep keys whose count >= cutoff: per finest-level fine bucket count / write
template <typename KeyT>
__global__ void filter_count_kernel(const uint32_t* __restrict__ offs, uint32_t NF, const uint8_t* __restrict__ counts,
                                    int cutoff, uint32_t* __restrict__ n_kept) {
  const uint32_t x = blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= NF) return;
  uint32_t k = 0;
  for (uint32_t i = offs[x]; i < offs[x + 1]; i++) k += counts[i] >= cutoff;
  n_kept[x] = k;
}

template <typename KeyT>
__global__ void filter_write_kernel(const KeyT* __restrict__ keys, const uint32_t* __restrict__ offs, uint32_t NF,
                                    const uint8_t* __restrict__ counts, int cutoff,
                                    const uint32_t* __restrict__ out_offs, KeyT* __restrict__ out) {
  const uint32_t x = blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= NF) return;
  uint32_t w = out_offs[x];
  for (uint32_t i = offs[x]; i < offs[x + 1]; i++)
    if (counts[i] >= cutoff) out[w++] = keys[i];
}

template <typename KeyT>
__global__ void lookup_kernel(const KeyT* __restrict__ keys, const uint32_t* __restrict__ offs,
                              const uint8_t* __restrict__ counts, uint32_t bucket, unsigned long long key,
                              int* __restrict__ out) {
  uint32_t lo = offs[bucket], hi = offs[bucket + 1];
  while (lo < hi) {
    const uint32_t mid = (lo + hi) >> 1;
    if ((unsigned long long)keys[mid] < key) lo = mid + 1; else hi = mid;
  }
  *out = (lo < offs[bucket + 1] && (unsigned long long)keys[lo] == key) ? (int)counts[lo] : 0;
}

template <typename KeyT>
int filter_t(kmsc_ctx* ctx, const kmsc_set* full, const uint8_t* d_counts, int cutoff, kmsc_set** out) {
  const int F = full->max_level;
  const uint32_t NF = (uint32_t)1 << (full->N + F);
  const size_t ent = (size_t)NF + 1;
  const size_t sb = scan_scratch_entries(NF);
  KMSC_TRY(ctx->work.reserve((ent + sb + 16) * 4));
  uint32_t* d_kept = (uint32_t*)ctx->work.p;
  uint32_t* d_bsum = d_kept + ent;
  uint32_t* d_total = d_bsum + sb;
  filter_count_kernel<KeyT><<<(NF + 127) / 128, 128, 0, ctx->stream>>>(full->lev[F], NF, d_counts, cutoff, d_kept);
  count_launch(ctx);
  KMSC_TRY(exclusive_scan_u32(ctx, d_kept, d_kept, NF, d_bsum, d_total));
  void* pin = nullptr;
  KMSC_TRY(ctx_pinned(ctx, 64, &pin));
  KMSC_CUDA(cudaMemcpyAsync(pin, d_total, 4, cudaMemcpyDeviceToHost, ctx->stream));
  KMSC_CUDA(cudaStreamSynchronize(ctx->stream));
  const int64_t n_kept = *(uint32_t*)pin;
  kmsc_set* s = nullptr;
  KMSC_TRY(set_alloc(ctx, full->K, full->N, full->key_bytes, n_kept, &s));
  filter_write_kernel<KeyT><<<(NF + 127) / 128, 128, 0, ctx->stream>>>((const KeyT*)full->keys, full->lev[F], NF, d_counts,
                                                                      cutoff, d_kept, (KeyT*)s->keys);
  count_launch(ctx);
  cudaError_t e = cudaMemcpyAsync(s->lev[F], d_kept, ent * 4, cudaMemcpyDeviceToDevice, ctx->stream);
  int rc = e == cudaSuccess ? set_derive_levels(ctx, s) : cuda_fail(e, "filter offsets", __FILE__, __LINE__);
  if (rc == KMSC_OK) {
    e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) rc = cuda_fail(e, "filter finish", __FILE__, __LINE__);
  }
  if (rc != KMSC_OK) { kmsc_set_free(ctx, s); return rc; }
  s->has_dups = 0;
  *out = s;
  return KMSC_OK;
}

// d_text_ready != NULL: the text is on the device already (kmsc_counter_prefetch copied it on the copy stream while
// the previous chunk was counted; the caller has made ctx->stream wait for that copy); `text` is still read on
// the host for the last byte
int count_common(kmsc_ctx* ctx, int K, int N, int key_bytes, const char* text, int64_t n, int canonical,
                 int cutoff, int fasta, kmsc_set** out, int64_t* cutoff_count, int64_t* n_distinct,
                 const unsigned char* d_text_ready = nullptr) {
  if (!ctx || !out || n < 0 || (n > 0 && !text)) { set_error("bad argument"); return KMSC_E_INVALID; }
  if (K < 1 || K > 32 || N < 0 || N > 24 || N > 2 * K || 2 * K - N > 8 * key_bytes || 2 * K - N >= 64 ||
      (key_bytes != 2 && key_bytes != 4 && key_bytes != 8)) { set_error("bad K/N/key_bytes"); return KMSC_E_INVALID; }
  KMSC_CUDA(cudaSetDevice(ctx->device));
  // drop the previous counter
  if (ctx->last_counted_owned && ctx->last_counted) kmsc_set_free(ctx, ctx->last_counted);
  ctx->last_counted = nullptr; ctx->last_counted_owned = false;
  if (ctx->last_counts) { cudaFree(ctx->last_counts); ctx->last_counts = nullptr; }

  const unsigned long long n_chunks = ((unsigned long long)n + 31) / 32;
  size_t off = 0;
  const size_t o_text = off; off += d_text_ready ? 256 : (((size_t)n + 255) & ~(size_t)255);
  const size_t o_nl = off; off += ((size_t)(n_chunks + 1) * 4 + 255) & ~(size_t)255;
  const size_t o_bsum = off; off += (scan_scratch_entries(n_chunks) * 4 + 255) & ~(size_t)255;
  const size_t o_words = off; off += ((size_t)(n_chunks + 2) * 8 + 255) & ~(size_t)255;
  const size_t o_inv = off; off += ((size_t)(n_chunks + 2) * 4 + 255) & ~(size_t)255;
  const size_t o_bad = off; off += ((size_t)(n_chunks + 2) * 4 + 255) & ~(size_t)255;
  const size_t o_flag = off; off += 256;
  KMSC_TRY(ctx->stage.reserve(off));
  unsigned char* base = (unsigned char*)ctx->stage.p;
  const unsigned char* d_text = d_text_ready ? d_text_ready : base + o_text;
  uint32_t* d_nl = (uint32_t*)(base + o_nl);
  uint32_t* d_bsum = (uint32_t*)(base + o_bsum);
  unsigned long long* d_words = (unsigned long long*)(base + o_words);
  uint32_t* d_inv = (uint32_t*)(base + o_inv);
  uint32_t* d_bad = (uint32_t*)(base + o_bad);
  int* d_flag = (int*)(base + o_flag);          // [0] flags, [1] total newlines

  if (n > 0 && !d_text_ready) KMSC_CUDA(cudaMemcpyAsync(base + o_text, text, (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
  KMSC_CUDA(cudaMemsetAsync(d_flag, 0, 16, ctx->stream));
  KMSC_CUDA(cudaMemsetAsync(d_words + n_chunks, 0, 16, ctx->stream));
  const int threads = 128;
  const unsigned blocks = (unsigned)((n_chunks + threads - 1) / threads);
  if (n_chunks > 0) {
    count_newlines_kernel<<<blocks, threads, 0, ctx->stream>>>(d_text, (unsigned long long)n, d_nl, n_chunks);
    count_launch(ctx);
  }
  KMSC_TRY(exclusive_scan_u32(ctx, d_nl, d_nl, n_chunks, d_bsum, (uint32_t*)(d_flag + 1)));
  if (n_chunks > 0) {
    classify_pack_kernel<<<blocks, threads, 0, ctx->stream>>>(d_text, (unsigned long long)n, d_nl, fasta, d_words, d_inv,
                                                            n_chunks, d_flag);
    window_mask_kernel<<<blocks, threads, 0, ctx->stream>>>(d_inv, n_chunks, K, d_bad);
    count_launch(ctx, 2);
  }
  KMSC_CUDA(cudaGetLastError());
  void* pin = nullptr;
  KMSC_TRY(ctx_pinned(ctx, 64, &pin));
  KMSC_CUDA(cudaMemcpyAsync(pin, d_flag, 8, cudaMemcpyDeviceToHost, ctx->stream));
  KMSC_CUDA(cudaStreamSynchronize(ctx->stream));
  const int flags = ((int*)pin)[0];
  const uint32_t newlines = ((uint32_t*)pin)[1];
  if (fasta) {
    // std::getline semantics (core/io.h:31-34): a trailing '\n' does not open a new line
    const uint64_t n_lines = (uint64_t)newlines + ((n > 0 && text[n - 1] != '\n') ? 1 : 0);
    if (n_lines % 2 != 0) { set_error("FASTA files should have an even number of lines"); return KMSC_E_FORMAT; }
  }
  if (flags & kFlagBadChar) { set_error(fasta ? "invalid FASTA file" : "reads may only contain A, C, G, T, N"); return KMSC_E_FORMAT; }

  PipelineInput in{d_words, d_bad, n};
  PipelineOptions opt{K, N, key_bytes, canonical, 0, 1 << N, 2, 1};
  PipelineResult res;
  KMSC_TRY(run_kmer_pipeline(ctx, in, opt, &res));
  if (n_distinct) *n_distinct = res.n_distinct;
  ctx->last_counts = res.d_counts;
  ctx->last_counted = res.set;
  const int cut8 = cutoff & 0xff;  // ValueType cutoff is uint8 (kmer_counter.h:214)
  if (cut8 <= 1) {
    // nothing to drop: hand the full set to the caller (the context keeps a borrowed view)
    ctx->last_counted_owned = false;
    *out = res.set;
    if (cutoff_count) *cutoff_count = 0;
    return KMSC_OK;
  }
  ctx->last_counted_owned = true;
  kmsc_set* kept = nullptr;
  int rc;
  switch (key_bytes) {
    case 2: rc = filter_t<uint16_t>(ctx, res.set, res.d_counts, cut8, &kept); break;
    case 4: rc = filter_t<uint32_t>(ctx, res.set, res.d_counts, cut8, &kept); break;
    default: rc = filter_t<unsigned long long>(ctx, res.set, res.d_counts, cut8, &kept); break;
  }
  if (rc != KMSC_OK) return rc;
  *out = kept;
  if (cutoff_count) *cutoff_count = res.set->n_keys - kept->n_keys;
  return KMSC_OK;
}

}  // namespace
}  // namespace kmsc

using namespace kmsc;

extern "C" {

int kmsc_count_fasta(kmsc_ctx* ctx, int K, int N, int key_bytes, const char* fasta, int64_t n_bytes,
                     int canonical, int cutoff, kmsc_set** out, int64_t* cutoff_count, int64_t* n_distinct) {
  return count_common(ctx, K, N, key_bytes, fasta, n_bytes, canonical, cutoff, 1, out, cutoff_count, n_distinct);
}

int kmsc_count_reads(kmsc_ctx* ctx, int K, int N, int key_bytes, const char* reads, int64_t n_bytes,
                     int canonical, int cutoff, kmsc_set** out, int64_t* cutoff_count, int64_t* n_distinct) {
  return count_common(ctx, K, N, key_bytes, reads, n_bytes, canonical, cutoff, 0, out, cutoff_count, n_distinct);
}

int kmsc_count_get(kmsc_ctx* ctx, uint64_t kmer, int* count) {
  if (!ctx || !count) { set_error("NULL argument"); return KMSC_E_INVALID; }
  const kmsc_set* s = ctx->last_counted;
  if (!s || !ctx->last_counts) { set_error("no counter: call kmsc_count_fasta / kmsc_count_reads first"); return KMSC_E_STATE; }
  KMSC_CUDA(cudaSetDevice(ctx->device));
  if (2 * s->K < 64 && (kmer >> (2 * s->K)) != 0) { *count = 0; return KMSC_OK; }
  const uint32_t bucket = (uint32_t)(kmer >> s->key_bits);
  const unsigned long long key = kmer & ((1ull << s->key_bits) - 1);
  KMSC_TRY(ctx->small.reserve(64));
  int* d_out = (int*)ctx->small.p;
  switch (s->key_bytes) {
    case 2: lookup_kernel<uint16_t><<<1, 1, 0, ctx->stream>>>((const uint16_t*)s->keys, s->lev[0], ctx->last_counts, bucket, key, d_out); break;
    case 4: lookup_kernel<uint32_t><<<1, 1, 0, ctx->stream>>>((const uint32_t*)s->keys, s->lev[0], ctx->last_counts, bucket, key, d_out); break;
    default: lookup_kernel<unsigned long long><<<1, 1, 0, ctx->stream>>>((const unsigned long long*)s->keys, s->lev[0], ctx->last_counts, bucket, key, d_out); break;
  }
  count_launch(ctx);
  KMSC_CUDA(cudaGetLastError());
  KMSC_CUDA(cudaMemcpyAsync(count, d_out, 4, cudaMemcpyDeviceToHost, ctx->stream));
  KMSC_CUDA(cudaStreamSynchronize(ctx->stream));
  return KMSC_OK;
}

int kmsc_count_last_counts(kmsc_ctx* ctx, uint8_t* out, int64_t n) {
  if (!ctx || (n > 0 && !out)) { set_error("NULL argument"); return KMSC_E_INVALID; }
  if (!ctx->last_counted || !ctx->last_counts) { set_error("no counter: call kmsc_count_fasta / kmsc_count_reads first"); return KMSC_E_STATE; }
  if (n != ctx->last_counted->n_keys) { set_error("n=%lld but the counter holds %lld k-mers", (long long)n, (long long)ctx->last_counted->n_keys); return KMSC_E_INVALID; }
  KMSC_CUDA(cudaSetDevice(ctx->device));
  if (n > 0) KMSC_CUDA(cudaMemcpyAsync(out, ctx->last_counts, (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
  KMSC_CUDA(cudaStreamSynchronize(ctx->stream));
  return KMSC_OK;
}

/* ---- streaming counter: chunks of a FASTA / read file are counted one at a time ----------- */

struct kmsc_counter {
  int K, N, key_bytes, canonical;
  kmsc_set* acc;        // all distinct k-mers so far
  uint8_t* counts;      // device, aligned with acc's keys
  // chunks announced by kmsc_counter_prefetch: their text on the device, copied on the copy stream. Two slots:
  // the chunk being added may itself sit in one while the next one is copied into the other
  struct Pre {
    const char* host = nullptr;
    int64_t n = 0;
  } pre[2];   // slot q's bytes live in ctx->pre_buf[q], its copy is marked by ctx->pre_ev[q]
  int pre_next = 0;
};

int kmsc_counter_create(kmsc_ctx* ctx, int K, int N, int key_bytes, int canonical, kmsc_counter** out) {
  if (!ctx || !out) { set_error("NULL argument"); return KMSC_E_INVALID; }
  if (K < 1 || K > 32 || N < 0 || N > 24 || N > 2 * K || 2 * K - N > 8 * key_bytes || 2 * K - N >= 64 ||
      (key_bytes != 2 && key_bytes != 4 && key_bytes != 8)) {
    set_error("bad K/N/key_bytes");
    return KMSC_E_INVALID;
  }
  kmsc_counter* c = new kmsc_counter();
  c->K = K; c->N = N; c->key_bytes = key_bytes; c->canonical = canonical ? 1 : 0;
  c->acc = nullptr; c->counts = nullptr;
  *out = c;
  return KMSC_OK;
}

static int counter_add(kmsc_ctx* ctx, kmsc_counter* c, const char* text, int64_t n, int fasta) {
  if (!ctx || !c) { set_error("NULL argument"); return KMSC_E_INVALID; }
  kmsc_set* s = nullptr;
  int64_t cut = 0, nd = 0;
  const unsigned char* ready = nullptr;
  for (int q = 0; q < 2; q++) {
    kmsc_counter::Pre& p = c->pre[q];
    if (p.host && p.host == text && p.n == n && n > 0 && !ready) {
      KMSC_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->pre_ev[q], 0));
      ready = (const unsigned char*)ctx->pre_buf[q].p;
      p.host = nullptr;   // consumed (the slot is reused by the prefetch after next)
    }
  }
  KMSC_TRY(count_common(ctx, c->K, c->N, c->key_bytes, text, n, c->canonical, /*cutoff=*/1, fasta, &s, &cut, &nd, ready));
  // count_common left the chunk's counter in the context (borrowed set + owned counts): take both
  uint8_t* sc = ctx->last_counts;
  ctx->last_counts = nullptr; ctx->last_counted = nullptr; ctx->last_counted_owned = false;
  if (!c->acc) { c->acc = s; c->counts = sc; return KMSC_OK; }
  kmsc_set* u = nullptr;
  uint8_t* uc = nullptr;
  const int rc = counted_union(ctx, c->acc, c->counts, s, sc, &u, &uc);
  kmsc_set_free(ctx, s);
  if (sc) cudaFree(sc);
  if (rc != KMSC_OK) return rc;
  kmsc_set_free(ctx, c->acc);
  if (c->counts) cudaFree(c->counts);
  c->acc = u; c->counts = uc;
  return KMSC_OK;
}

int kmsc_counter_prefetch(kmsc_ctx* ctx, kmsc_counter* c, const char* text, int64_t n_bytes) {
  if (!ctx || !c || n_bytes < 0 || (n_bytes > 0 && !text)) { set_error("bad argument"); return KMSC_E_INVALID; }
  if (n_bytes == 0) return KMSC_OK;
  KMSC_CUDA(cudaSetDevice(ctx->device));
  if (!ctx->copy_stream) {
    KMSC_CUDA(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
    for (auto& e : ctx->copy_ev) KMSC_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    KMSC_CUDA(cudaEventCreateWithFlags(&ctx->fence_ev, cudaEventDisableTiming));
  }
  const int q = c->pre_next;
  kmsc_counter::Pre& p = c->pre[q];
  c->pre_next ^= 1;
  p.host = nullptr;
  if (!ctx->pre_ev[q]) KMSC_CUDA(cudaEventCreateWithFlags(&ctx->pre_ev[q], cudaEventDisableTiming));
  if (ctx->pre_buf[q].cap < (size_t)n_bytes) KMSC_CUDA(cudaStreamSynchronize(ctx->stream));   // growing frees the old buffer: its readers are on the main stream
  KMSC_TRY(ctx->pre_buf[q].reserve((size_t)n_bytes));
  // the slot's last reader (the classify kernels of the chunk it held before) ran on the main stream
  KMSC_CUDA(cudaEventRecord(ctx->fence_ev, ctx->stream));
  KMSC_CUDA(cudaStreamWaitEvent(ctx->copy_stream, ctx->fence_ev, 0));
  KMSC_CUDA(cudaMemcpyAsync(ctx->pre_buf[q].p, text, (size_t)n_bytes, cudaMemcpyHostToDevice, ctx->copy_stream));
  KMSC_CUDA(cudaEventRecord(ctx->pre_ev[q], ctx->copy_stream));
  p.host = text;
  p.n = n_bytes;
  return KMSC_OK;
}

int kmsc_counter_add_fasta(kmsc_ctx* ctx, kmsc_counter* c, const char* fasta, int64_t n_bytes) {
  return counter_add(ctx, c, fasta, n_bytes, 1);
}
int kmsc_counter_add_reads(kmsc_ctx* ctx, kmsc_counter* c, const char* reads, int64_t n_bytes) {
  return counter_add(ctx, c, reads, n_bytes, 0);
}

int kmsc_counter_finish(kmsc_ctx* ctx, kmsc_counter* c, int cutoff, kmsc_set** out, int64_t* cutoff_count,
                        int64_t* n_distinct) {
  if (!ctx || !c || !out) { set_error("NULL argument"); return KMSC_E_INVALID; }
  KMSC_CUDA(cudaSetDevice(ctx->device));
  if (!c->acc) KMSC_TRY(counter_add(ctx, c, "", 0, 0));  // nothing added: the empty counter
  // the accumulated counter becomes the context's counter (kmsc_count_get / kmsc_count_last_counts)
  if (ctx->last_counted_owned && ctx->last_counted) kmsc_set_free(ctx, ctx->last_counted);
  if (ctx->last_counts) cudaFree(ctx->last_counts);
  ctx->last_counted = c->acc; ctx->last_counts = c->counts;
  kmsc_set* all = c->acc;
  c->acc = nullptr; c->counts = nullptr;
  if (n_distinct) *n_distinct = all->n_keys;
  const int cut8 = cutoff & 0xff;
  if (cut8 <= 1) {
    ctx->last_counted_owned = false;
    *out = all;
    if (cutoff_count) *cutoff_count = 0;
    return KMSC_OK;
  }
  ctx->last_counted_owned = true;
  kmsc_set* kept = nullptr;
  int rc;
  switch (all->key_bytes) {
    case 2: rc = filter_t<uint16_t>(ctx, all, ctx->last_counts, cut8, &kept); break;
    case 4: rc = filter_t<uint32_t>(ctx, all, ctx->last_counts, cut8, &kept); break;
    default: rc = filter_t<unsigned long long>(ctx, all, ctx->last_counts, cut8, &kept); break;
  }
  if (rc != KMSC_OK) return rc;
  *out = kept;
  if (cutoff_count) *cutoff_count = all->n_keys - kept->n_keys;
  return KMSC_OK;
}

void kmsc_counter_free(kmsc_ctx* ctx, kmsc_counter* c) {
  if (!c) return;
  // a copy announced but never used may still be in flight out of the caller's buffer
  if ((c->pre[0].host || c->pre[1].host) && ctx && ctx->copy_stream) { cudaSetDevice(ctx->device); cudaStreamSynchronize(ctx->copy_stream); }
  if (c->acc) kmsc_set_free(ctx, c->acc);
  if (c->counts) cudaFree(c->counts);
  delete c;
}

}  // extern "C"
