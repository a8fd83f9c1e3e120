// kmsc_api.cu -- context, error reporting and device-set management for libkmsc.
// Reference interfaces replaced: KmerSet<K,N,KeyType> construction / Size / Hash
// (lib/core/kmer_set.h:57-115, 224-244) and the bucket/key split (:22-43).
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "kmsc_common.cuh"

namespace kmsc {

static thread_local std::string g_last_error;

void set_error(const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_last_error = buf;
}

int cuda_fail(cudaError_t e, const char* what, const char* file, int line) {
  set_error("CUDA error %d (%s) at %s:%d in %s", (int)e, cudaGetErrorString(e), file, line, what);
  return e == cudaErrorMemoryAllocation ? KMSC_E_NOMEM : KMSC_E_CUDA;
}

int Scratch::reserve(size_t bytes) {
  if (bytes <= cap) return KMSC_OK;
  if (p) cudaFree(p);
  p = nullptr;
  cap = 0;
  size_t want = bytes + bytes / 4 + 256;
  KMSC_CUDA(cudaMalloc(&p, want));
  cap = want;
  return KMSC_OK;
}

void Scratch::release() {
  if (p) cudaFree(p);
  p = nullptr;
  cap = 0;
}

int PinnedTab::acquire(size_t bytes, void** out) {
  if (pending) { KMSC_CUDA(cudaEventSynchronize(ev)); pending = false; }
  if (bytes > cap) {
    if (p) cudaFreeHost(p);
    p = nullptr; cap = 0;
    const size_t want = bytes + bytes / 2 + 4096;
    KMSC_CUDA(cudaMallocHost(&p, want));
    cap = want;
  }
  if (!ev) KMSC_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
  *out = p;
  return KMSC_OK;
}

int PinnedTab::commit(cudaStream_t stream) {
  KMSC_CUDA(cudaEventRecord(ev, stream));
  pending = true;
  return KMSC_OK;
}

void PinnedTab::release() {
  if (pending && ev) cudaEventSynchronize(ev);
  if (p) cudaFreeHost(p);
  if (ev) cudaEventDestroy(ev);
  p = nullptr; cap = 0; ev = nullptr; pending = false;
}

int ctx_pinned(kmsc_ctx* ctx, size_t bytes, void** out) {
  if (bytes > ctx->pinned_cap) {
    if (ctx->pinned) cudaFreeHost(ctx->pinned);
    ctx->pinned = nullptr;
    ctx->pinned_cap = 0;
    size_t want = bytes + bytes / 4 + 4096;
    KMSC_CUDA(cudaMallocHost(&ctx->pinned, want));
    ctx->pinned_cap = want;
  }
  *out = ctx->pinned;
  return KMSC_OK;
}

// ---------------------------------------------------------------------------
// kernels
// ---------------------------------------------------------------------------

// Fine offsets for levels 1..max_level by binary search inside each bucket run.
// One thread per (level, x). Level f entry x covers keys whose top f key bits
// equal (x & (2^f - 1)) inside bucket (x >> f).
template <typename KeyT>
__global__ void build_levels_kernel(const KeyT* __restrict__ keys, const uint32_t* __restrict__ offs,
                                    uint32_t* __restrict__ lev_base, int N, int key_bits,
                                    int max_level, uint32_t n_keys) {
  // tid enumerates the entries of levels 1..max_level back to back
  uint64_t rel = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  uint64_t base = ((uint64_t)1 << N) + 1;  // level 1 starts right after level 0
  for (int f = 1; f <= max_level; f++) {
    const uint64_t cnt = ((uint64_t)1 << (N + f)) + 1;
    if (rel < cnt) {
      const uint64_t x = rel;
      uint32_t res;
      if (x == ((uint64_t)1 << (N + f))) {
        res = n_keys;
      } else {
        const uint32_t b = (uint32_t)(x >> f);
        const uint64_t sub = x & (((uint64_t)1 << f) - 1);
        const uint64_t thr = sub << (key_bits - f);
        uint32_t lo = offs[b], hi = offs[b + 1];
        while (lo < hi) {
          const uint32_t mid = (lo + hi) >> 1;
          if ((uint64_t)keys[mid] < thr) lo = mid + 1; else hi = mid;
        }
        res = lo;
      }
      lev_base[base + x] = res;
      return;
    }
    rel -= cnt;
    base += cnt;
  }
}

// coarser levels from the finest one: lev[f][x] = lev[F][x << (F - f)]
__global__ void derive_levels_kernel(uint32_t* __restrict__ lev_base, int N, int max_level) {
  const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  uint64_t fine_start = 0;
  for (int g = 0; g < max_level; g++) fine_start += ((uint64_t)1 << (N + g)) + 1;
  uint64_t start = 0;
  for (int f = 0; f < max_level; f++) {
    const uint64_t cnt = ((uint64_t)1 << (N + f)) + 1;
    if (tid >= start && tid < start + cnt) {
      const uint64_t x = tid - start;
      lev_base[tid] = lev_base[fine_start + (x << (max_level - f))];
      return;
    }
    start += cnt;
  }
}

// one warp per bucket: duplicate detection + XOR hash of full k-mer values
template <typename KeyT>
__global__ void scan_buckets_kernel(const KeyT* __restrict__ keys, const uint32_t* __restrict__ offs,
                                    int n_buckets, int key_bits, int* __restrict__ dup_flag,
                                    unsigned long long* __restrict__ hash_out) {
  const int lane = threadIdx.x & 31;
  const int warps_per_block = blockDim.x >> 5;
  unsigned long long h = 0;
  int dup = 0;
  for (int b = blockIdx.x * warps_per_block + (threadIdx.x >> 5); b < n_buckets;
       b += gridDim.x * warps_per_block) {
    const uint32_t lo = offs[b], hi = offs[b + 1];
    const unsigned long long top = (unsigned long long)b << key_bits;
    for (uint32_t i = lo + lane; i < hi; i += 32) {
      const KeyT k = keys[i];
      h ^= top | (unsigned long long)k;
      if (i + 1 < hi && keys[i + 1] == k) dup = 1;
    }
  }
  for (int o = 16; o > 0; o >>= 1) h ^= __shfl_xor_sync(0xffffffffu, h, o);
  dup = __any_sync(0xffffffffu, dup);
  if (lane == 0) {
    if (h) atomicXor(hash_out, h);
    if (dup) atomicExch(dup_flag, 1);
  }
}

template <typename KeyT>
__global__ void kmers_to_keys_kernel(const unsigned long long* __restrict__ kmers, int64_t n,
                                     int key_bits, KeyT* __restrict__ keys) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) keys[i] = (KeyT)(kmers[i] & ((key_bits == 64) ? ~0ull : ((1ull << key_bits) - 1)));
}

// bucket offsets from ascending k-mers: offs[b] = first i with (kmer[i] >> key_bits) >= b
__global__ void kmers_bucket_offsets_kernel(const unsigned long long* __restrict__ kmers, int64_t n,
                                            int key_bits, int n_buckets, uint32_t* __restrict__ offs) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b > n_buckets) return;
  int64_t lo = 0, hi = n;
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if ((kmers[mid] >> key_bits) < (unsigned long long)b) lo = mid + 1; else hi = mid;
  }
  offs[b] = (uint32_t)lo;
}

// ---- multi-GPU exchange helpers: bucket ranges of a set as (rebased offsets, keys) -----------
__global__ void gather_offsets_kernel(const uint32_t* __restrict__ offs, const int32_t* __restrict__ buckets, int n,
                                      unsigned long long* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = offs[buckets[i]];
}
__global__ void export_offsets_kernel(const uint32_t* __restrict__ offs, int lo, int n_entries, uint32_t* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_entries) out[i] = offs[lo + i] - offs[lo];
}
__global__ void import_offsets_kernel(const uint32_t* __restrict__ in, int lo, int hi, int n_buckets, uint32_t n_keys,
                                      uint32_t* __restrict__ offs) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b > n_buckets) return;
  offs[b] = b <= lo ? 0u : b >= hi ? n_keys : in[b - lo];
}

static uint64_t levels_total_entries(int N, int max_level) {
  uint64_t t = 0;
  for (int f = 0; f <= max_level; f++) t += ((uint64_t)1 << (N + f)) + 1;
  return t;
}

int set_alloc(kmsc_ctx* ctx, int K, int N, int key_bytes, int64_t n_keys, kmsc_set** out) {
  if (K < 1 || K > 32 || N < 0 || N > 24 || N > 2 * K) { set_error("bad K=%d N=%d", K, N); return KMSC_E_INVALID; }
  if (key_bytes != 2 && key_bytes != 4 && key_bytes != 8) {
    set_error("key_bytes must be 2, 4 or 8 (got %d)", key_bytes);
    return KMSC_E_INVALID;
  }
  if (2 * K - N > 8 * key_bytes) { set_error("2K-N=%d does not fit %d key bytes", 2 * K - N, key_bytes); return KMSC_E_INVALID; }
  // 64 key bits (K = 32, N = 0) would need 64-bit shifts by 64 in several kernels and has no empty
  // marker left for the hash tables; none of the reference's instantiations needs it
  if (2 * K - N >= 64) { set_error("2K-N=%d: keys of 64 bits are not supported (use N >= 1 with K = 32)", 2 * K - N); return KMSC_E_INVALID; }
  if (n_keys < 0 || n_keys >= ((int64_t)1 << 32) - 64) { set_error("n_keys=%lld out of range", (long long)n_keys); return KMSC_E_INVALID; }
  KMSC_CUDA(cudaSetDevice(ctx->device));
  kmsc_set* s = new kmsc_set();
  s->K = K; s->N = N; s->key_bytes = key_bytes; s->key_bits = 2 * K - N;
  s->max_level = s->key_bits < kMaxFineLevel ? s->key_bits : kMaxFineLevel;
  s->n_keys = n_keys;
  // stream-ordered allocations from the device's default pool (kept warm: no release
  // threshold), so building / freeing sets every step costs microseconds, not cudaMalloc.
  // 64 bytes of slack so vectorised tail loads stay in bounds
  cudaError_t e = cudaMallocAsync(&s->keys, (size_t)n_keys * key_bytes + 64, ctx->stream);
  if (e != cudaSuccess) { delete s; return cuda_fail(e, "cudaMallocAsync keys", __FILE__, __LINE__); }
  const uint64_t entries = levels_total_entries(N, s->max_level);
  e = cudaMallocAsync((void**)&s->lev_base, entries * sizeof(uint32_t), ctx->stream);
  if (e != cudaSuccess) { cudaFreeAsync(s->keys, ctx->stream); delete s; return cuda_fail(e, "cudaMallocAsync levels", __FILE__, __LINE__); }
  uint64_t start = 0;
  for (int f = 0; f <= s->max_level; f++) {
    s->lev[f] = s->lev_base + start;
    start += ((uint64_t)1 << (N + f)) + 1;
  }
  *out = s;
  return KMSC_OK;
}

template <typename KeyT>
static int build_levels_t(kmsc_ctx* ctx, kmsc_set* s) {
  if (s->max_level == 0) return KMSC_OK;
  const uint64_t entries = levels_total_entries(s->N, s->max_level) - (((uint64_t)1 << s->N) + 1);
  const int threads = 256;
  const unsigned blocks = (unsigned)((entries + threads - 1) / threads);
  build_levels_kernel<KeyT><<<blocks, threads, 0, ctx->stream>>>(
      (const KeyT*)s->keys, s->lev[0], s->lev_base, s->N, s->key_bits, s->max_level, (uint32_t)s->n_keys);
  count_launch(ctx);
  KMSC_CUDA(cudaGetLastError());
  return KMSC_OK;
}

int set_build_levels(kmsc_ctx* ctx, kmsc_set* s) {
  switch (s->key_bytes) {
    case 2: return build_levels_t<uint16_t>(ctx, s);
    case 4: return build_levels_t<uint32_t>(ctx, s);
    default: return build_levels_t<unsigned long long>(ctx, s);
  }
}

int set_ensure_levels(kmsc_ctx* ctx, const kmsc_set* cs) {
  if (!cs || cs->fine_ready) return KMSC_OK;
  kmsc_set* s = const_cast<kmsc_set*>(cs);
  KMSC_TRY(set_build_levels(ctx, s));
  s->fine_ready = true;
  return KMSC_OK;
}

int set_derive_levels(kmsc_ctx* ctx, kmsc_set* s) {
  if (s->max_level == 0) return KMSC_OK;
  uint64_t entries = levels_total_entries(s->N, s->max_level - 1);
  const int threads = 256;
  derive_levels_kernel<<<(unsigned)((entries + threads - 1) / threads), threads, 0, ctx->stream>>>(
      s->lev_base, s->N, s->max_level);
  count_launch(ctx);
  KMSC_CUDA(cudaGetLastError());
  return KMSC_OK;
}

// runs the bucket scan; fills has_dups and (optionally) the XOR hash
static int set_scan(kmsc_ctx* ctx, kmsc_set* s, uint64_t* hash) {
  KMSC_TRY(ctx->small.reserve(64));
  unsigned long long* d_hash = (unsigned long long*)ctx->small.p;
  int* d_dup = (int*)((char*)ctx->small.p + 8);
  KMSC_CUDA(cudaMemsetAsync(ctx->small.p, 0, 16, ctx->stream));
  const int nb = 1 << s->N;
  const int threads = 256;
  int blocks = (nb + 7) / 8;
  if (blocks > ctx->sm_count * 8) blocks = ctx->sm_count * 8;
  if (blocks < 1) blocks = 1;
  switch (s->key_bytes) {
    case 2: scan_buckets_kernel<uint16_t><<<blocks, threads, 0, ctx->stream>>>((const uint16_t*)s->keys, s->lev[0], nb, s->key_bits, d_dup, d_hash); break;
    case 4: scan_buckets_kernel<uint32_t><<<blocks, threads, 0, ctx->stream>>>((const uint32_t*)s->keys, s->lev[0], nb, s->key_bits, d_dup, d_hash); break;
    default: scan_buckets_kernel<unsigned long long><<<blocks, threads, 0, ctx->stream>>>((const unsigned long long*)s->keys, s->lev[0], nb, s->key_bits, d_dup, d_hash); break;
  }
  count_launch(ctx);
  KMSC_CUDA(cudaGetLastError());
  struct { unsigned long long h; int dup; int pad; } host;
  KMSC_CUDA(cudaMemcpyAsync(&host, ctx->small.p, 16, cudaMemcpyDeviceToHost, ctx->stream));
  KMSC_CUDA(cudaStreamSynchronize(ctx->stream));
  s->has_dups = host.dup ? 1 : 0;
  if (hash) *hash = host.h;
  return KMSC_OK;
}

int set_check_dups(kmsc_ctx* ctx, kmsc_set* s) {
  if (s->has_dups >= 0) return KMSC_OK;
  return set_scan(ctx, s, nullptr);
}

}  // namespace kmsc

using namespace kmsc;

extern "C" {

const char* kmsc_last_error(void) { return g_last_error.c_str(); }
const char* kmsc_version(void) { return "kmsc-b200 0.1 (sm_100a)"; }

int kmsc_ctx_create(int device, void* stream, kmsc_ctx** out) {
  if (!out) { set_error("out is NULL"); return KMSC_E_INVALID; }
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    set_error("no CUDA device available (%s); libkmsc has no CPU fallback",
              e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    return KMSC_E_CUDA;
  }
  if (device < 0 || device >= n) { set_error("device %d out of range (%d devices)", device, n); return KMSC_E_INVALID; }
  KMSC_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  KMSC_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major < 10) {
    set_error("device %d is sm_%d%d; libkmsc is built for sm_100a only", device, prop.major, prop.minor);
    return KMSC_E_CUDA;
  }
  {
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
      uint64_t keep = ~0ull;
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
  }
  kmsc_ctx* c = new kmsc_ctx();
  c->device = device;
  c->sm_count = prop.multiProcessorCount;
  if (stream) {
    c->stream = (cudaStream_t)stream;
  } else {
    e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) { delete c; return cuda_fail(e, "cudaStreamCreate", __FILE__, __LINE__); }
    c->own_stream = true;
  }
  *out = c;
  return KMSC_OK;
}

void kmsc_ctx_destroy(kmsc_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  kmsc_comm_destroy(ctx);
  ctx->plan.release(); ctx->work.release(); ctx->work2.release(); ctx->work3.release(); ctx->stage.release(); ctx->small.release(); ctx->spss_out.release(); ctx->pre_buf[0].release(); ctx->pre_buf[1].release(); for (auto& e : ctx->pre_ev) if (e) cudaEventDestroy(e); for (int i = 0; i < kmsc_ctx::kP2Slots; i++) {
    ctx->p2a[i].release(); ctx->p2b[i].release(); ctx->p2tab[i].release(); ctx->p2rb[i].release();
    if (ctx->copy_ev[i]) cudaEventDestroy(ctx->copy_ev[i]);
  }
  for (auto& t : ctx->tabs_dev) t.release();
  for (auto& t : ctx->tab) t.release();
  if (ctx->fence_ev) cudaEventDestroy(ctx->fence_ev);
  if (ctx->copy_stream) { cudaStreamSynchronize(ctx->copy_stream); cudaStreamDestroy(ctx->copy_stream); }
  if (ctx->pinned) cudaFreeHost(ctx->pinned);
  for (int i = 0; i < 3; i++)
    if (ctx->pc_ev[i]) cudaEventDestroy(ctx->pc_ev[i]);
  if (ctx->last_counted_owned && ctx->last_counted) { kmsc_set* lc = ctx->last_counted; ctx->last_counted = nullptr; kmsc_set_free(ctx, lc); }
  if (ctx->last_counts) cudaFree(ctx->last_counts);
  if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
}

int kmsc_ctx_sync(kmsc_ctx* ctx) {
  if (!ctx) { set_error("ctx is NULL"); return KMSC_E_INVALID; }
  KMSC_CUDA(cudaStreamSynchronize(ctx->stream));
  return KMSC_OK;
}

void* kmsc_ctx_stream(kmsc_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }
int64_t kmsc_ctx_launch_count(kmsc_ctx* ctx) { return ctx ? ctx->launches : 0; }

int kmsc_set_from_csr(kmsc_ctx* ctx, int K, int N, int key_bytes, const int64_t* offs,
                      const void* keys, kmsc_set** out) {
  if (!ctx || !offs || !out) { set_error("NULL argument"); return KMSC_E_INVALID; }
  // the shape first: offs has 2^N + 1 entries only if N is sane
  if (K < 1 || K > 32 || N < 0 || N > 24 || N > 2 * K || (key_bytes != 2 && key_bytes != 4 && key_bytes != 8) ||
      2 * K - N > 8 * key_bytes || 2 * K - N >= 64) {
    set_error("bad K=%d N=%d key_bytes=%d", K, N, key_bytes);
    return KMSC_E_INVALID;
  }
  const int64_t nb = (int64_t)1 << N;
  if (offs[0] != 0) { set_error("offs[0] must be 0"); return KMSC_E_INVALID; }
  for (int64_t b = 0; b < nb; b++)
    if (offs[b + 1] < offs[b]) { set_error("offs not monotone at bucket %lld", (long long)b); return KMSC_E_INVALID; }
  const int64_t n = offs[nb];
  if (n > 0 && !keys) { set_error("keys is NULL"); return KMSC_E_INVALID; }
  {
    // every consumer (offset levels, merges, tiles) assumes keys ascending inside each bucket and
    // below 2^(2K-N): refuse anything else here rather than return wrong matrices later
    const int key_bits = 2 * K - N;
    const uint64_t lim = key_bits >= 64 ? ~0ull : (((uint64_t)1 << key_bits) - 1);
    for (int64_t b = 0; b < nb; b++) {
      uint64_t prev = 0;
      for (int64_t i = offs[b]; i < offs[b + 1]; i++) {
        const uint64_t k = key_bytes == 2 ? ((const uint16_t*)keys)[i] : key_bytes == 4 ? ((const uint32_t*)keys)[i] : ((const uint64_t*)keys)[i];
        if (k > lim) { set_error("key %lld exceeds %d bits", (long long)i, key_bits); return KMSC_E_INVALID; }
        if (i > offs[b] && k < prev) { set_error("keys not ascending in bucket %lld", (long long)b); return KMSC_E_INVALID; }
        prev = k;
      }
    }
  }
  kmsc_set* s = nullptr;
  KMSC_TRY(set_alloc(ctx, K, N, key_bytes, n, &s));
  // stage offs as uint32 through pinned memory
  void* pin = nullptr;
  int rc = ctx_pinned(ctx, (size_t)(nb + 1) * 4, &pin);
  if (rc != KMSC_OK) { kmsc_set_free(ctx, s); return rc; }
  uint32_t* o32 = (uint32_t*)pin;
  for (int64_t b = 0; b <= nb; b++) o32[b] = (uint32_t)offs[b];
  cudaError_t e = cudaMemcpyAsync(s->lev[0], o32, (size_t)(nb + 1) * 4, cudaMemcpyHostToDevice, ctx->stream);
  if (e == cudaSuccess && n > 0)
    e = cudaMemcpyAsync(s->keys, keys, (size_t)n * key_bytes, cudaMemcpyHostToDevice, ctx->stream);
  if (e != cudaSuccess) { kmsc_set_free(ctx, s); return cuda_fail(e, "H2D set", __FILE__, __LINE__); }
  rc = set_build_levels(ctx, s);
  if (rc == KMSC_OK) rc = set_check_dups(ctx, s);  // also synchronises (pinned buffer reuse)
  if (rc != KMSC_OK) { kmsc_set_free(ctx, s); return rc; }
  *out = s;
  return KMSC_OK;
}

int kmsc_set_from_kmers(kmsc_ctx* ctx, int K, int N, int key_bytes, const uint64_t* kmers,
                        int64_t n, kmsc_set** out) {
  if (!ctx || !out || (n > 0 && !kmers)) { set_error("NULL argument"); return KMSC_E_INVALID; }
  for (int64_t i = 1; i < n; i++)
    if (kmers[i] < kmers[i - 1]) { set_error("kmers not ascending at %lld", (long long)i); return KMSC_E_INVALID; }
  if (n > 0 && 2 * K < 64 && (kmers[n - 1] >> (2 * K)) != 0) { set_error("k-mer value exceeds 2K bits"); return KMSC_E_INVALID; }
  kmsc_set* s = nullptr;
  KMSC_TRY(set_alloc(ctx, K, N, key_bytes, n, &s));
  int rc = ctx->work.reserve((size_t)(n ? n : 1) * 8);
  if (rc != KMSC_OK) { kmsc_set_free(ctx, s); return rc; }
  unsigned long long* d_kmers = (unsigned long long*)ctx->work.p;
  cudaError_t e = cudaSuccess;
  if (n > 0) e = cudaMemcpyAsync(d_kmers, kmers, (size_t)n * 8, cudaMemcpyHostToDevice, ctx->stream);
  if (e != cudaSuccess) { kmsc_set_free(ctx, s); return cuda_fail(e, "H2D kmers", __FILE__, __LINE__); }
  const int nb = 1 << N;
  kmers_bucket_offsets_kernel<<<(nb + 1 + 255) / 256, 256, 0, ctx->stream>>>(d_kmers, n, s->key_bits, nb, s->lev[0]);
  count_launch(ctx);
  if (n > 0) {
    const unsigned blocks = (unsigned)((n + 255) / 256);
    switch (key_bytes) {
      case 2: kmers_to_keys_kernel<uint16_t><<<blocks, 256, 0, ctx->stream>>>(d_kmers, n, s->key_bits, (uint16_t*)s->keys); break;
      case 4: kmers_to_keys_kernel<uint32_t><<<blocks, 256, 0, ctx->stream>>>(d_kmers, n, s->key_bits, (uint32_t*)s->keys); break;
      default: kmers_to_keys_kernel<unsigned long long><<<blocks, 256, 0, ctx->stream>>>(d_kmers, n, s->key_bits, (unsigned long long*)s->keys); break;
    }
    count_launch(ctx);
  }
  e = cudaGetLastError();
  if (e != cudaSuccess) { kmsc_set_free(ctx, s); return cuda_fail(e, "kmers->csr", __FILE__, __LINE__); }
  rc = set_build_levels(ctx, s);
  if (rc == KMSC_OK) rc = set_check_dups(ctx, s);
  if (rc != KMSC_OK) { kmsc_set_free(ctx, s); return rc; }
  *out = s;
  return KMSC_OK;
}

int kmsc_set_to_csr(kmsc_ctx* ctx, const kmsc_set* set, int64_t* offs, void* keys) {
  if (!ctx || !set) { set_error("NULL argument"); return KMSC_E_INVALID; }
  const int64_t nb = (int64_t)1 << set->N;
  if (offs) {
    void* pin = nullptr;
    KMSC_TRY(ctx_pinned(ctx, (size_t)(nb + 1) * 4, &pin));
    KMSC_CUDA(cudaMemcpyAsync(pin, set->lev[0], (size_t)(nb + 1) * 4, cudaMemcpyDeviceToHost, ctx->stream));
    KMSC_CUDA(cudaStreamSynchronize(ctx->stream));
    const uint32_t* o32 = (const uint32_t*)pin;
    for (int64_t b = 0; b <= nb; b++) offs[b] = o32[b];
  }
  if (keys && set->n_keys > 0) {
    KMSC_CUDA(cudaMemcpyAsync(keys, set->keys, (size_t)set->n_keys * set->key_bytes, cudaMemcpyDeviceToHost, ctx->stream));
    KMSC_CUDA(cudaStreamSynchronize(ctx->stream));
  }
  return KMSC_OK;
}

void kmsc_set_free(kmsc_ctx* ctx, kmsc_set* set) {
  if (!set) return;
  if (ctx) {
    cudaSetDevice(ctx->device);
    if (ctx->last_counted == set) ctx->last_counted = nullptr;
    // stream-ordered free: everything enqueued so far that reads the set completes first
    if (set->keys) cudaFreeAsync(set->keys, ctx->stream);
    if (set->lev_base) cudaFreeAsync(set->lev_base, ctx->stream);
  } else {
    if (set->keys) cudaFree(set->keys);
    if (set->lev_base) cudaFree(set->lev_base);
  }
  delete set;
}

int kmsc_set_size(kmsc_ctx* ctx, const kmsc_set* set, int64_t* size) {
  if (!ctx || !set || !size) { set_error("NULL argument"); return KMSC_E_INVALID; }
  *size = set->n_keys;
  return KMSC_OK;
}

int kmsc_set_hash(kmsc_ctx* ctx, const kmsc_set* set, uint64_t* hash) {
  if (!ctx || !set || !hash) { set_error("NULL argument"); return KMSC_E_INVALID; }
  return set_scan(ctx, const_cast<kmsc_set*>(set), hash);
}

int kmsc_set_info(const kmsc_set* set, int* K, int* N, int* key_bytes, int64_t* n_keys) {
  if (!set) { set_error("NULL argument"); return KMSC_E_INVALID; }
  if (K) *K = set->K;
  if (N) *N = set->N;
  if (key_bytes) *key_bytes = set->key_bytes;
  if (n_keys) *n_keys = set->n_keys;
  return KMSC_OK;
}

int kmsc_set_bucket_offsets(kmsc_ctx* ctx, const kmsc_set* set, const int32_t* buckets, int32_t n, int64_t* out) {
  if (!ctx || !set || !buckets || !out || n < 0) { set_error("bad argument"); return KMSC_E_INVALID; }
  const int nb = 1 << set->N;
  for (int i = 0; i < n; i++)
    if (buckets[i] < 0 || buckets[i] > nb) { set_error("bucket %d out of range", buckets[i]); return KMSC_E_INVALID; }
  if (n == 0) return KMSC_OK;
  KMSC_CUDA(cudaSetDevice(ctx->device));
  KMSC_TRY(ctx->small.reserve((size_t)n * 12 + 64));
  unsigned long long* d_out = (unsigned long long*)ctx->small.p;
  int32_t* d_b = (int32_t*)((unsigned char*)ctx->small.p + (((size_t)n * 8 + 15) & ~(size_t)15));
  KMSC_CUDA(cudaMemcpyAsync(d_b, buckets, (size_t)n * 4, cudaMemcpyHostToDevice, ctx->stream));
  gather_offsets_kernel<<<(n + 127) / 128, 128, 0, ctx->stream>>>(set->lev[0], d_b, n, d_out);
  count_launch(ctx);
  KMSC_CUDA(cudaGetLastError());
  KMSC_CUDA(cudaMemcpyAsync(out, d_out, (size_t)n * 8, cudaMemcpyDeviceToHost, ctx->stream));
  KMSC_CUDA(cudaStreamSynchronize(ctx->stream));
  return KMSC_OK;
}

int kmsc_set_export_range(kmsc_ctx* ctx, const kmsc_set* set, int32_t bucket_lo, int32_t bucket_hi,
                          int64_t key_lo, int64_t key_hi, uint32_t* d_offs, void* d_keys) {
  if (!ctx || !set || !d_offs) { set_error("NULL argument"); return KMSC_E_INVALID; }
  const int nb = 1 << set->N;
  if (bucket_lo < 0 || bucket_hi > nb || bucket_lo > bucket_hi || key_lo < 0 || key_hi < key_lo || key_hi > set->n_keys) {
    set_error("bad range");
    return KMSC_E_INVALID;
  }
  KMSC_CUDA(cudaSetDevice(ctx->device));
  const int n_entries = bucket_hi - bucket_lo + 1;
  export_offsets_kernel<<<(n_entries + 255) / 256, 256, 0, ctx->stream>>>(set->lev[0], bucket_lo, n_entries, d_offs);
  count_launch(ctx);
  KMSC_CUDA(cudaGetLastError());
  if (key_hi > key_lo) {
    if (!d_keys) { set_error("d_keys is NULL"); return KMSC_E_INVALID; }
    KMSC_CUDA(cudaMemcpyAsync(d_keys, (const unsigned char*)set->keys + (size_t)key_lo * set->key_bytes,
                              (size_t)(key_hi - key_lo) * set->key_bytes, cudaMemcpyDeviceToDevice, ctx->stream));
  }
  return KMSC_OK;
}

int kmsc_set_import_range(kmsc_ctx* ctx, int K, int N, int key_bytes, int32_t bucket_lo, int32_t bucket_hi,
                          const uint32_t* d_offs, const void* d_keys, int64_t n_keys, kmsc_set** out) {
  if (!ctx || !out || !d_offs || (n_keys > 0 && !d_keys)) { set_error("NULL argument"); return KMSC_E_INVALID; }
  if (N < 0 || N > 24 || bucket_lo < 0 || bucket_hi > (1 << N) || bucket_lo > bucket_hi) { set_error("bad range"); return KMSC_E_INVALID; }
  kmsc_set* s = nullptr;
  KMSC_TRY(set_alloc(ctx, K, N, key_bytes, n_keys, &s));
  const int nb = 1 << N;
  import_offsets_kernel<<<(nb + 1 + 255) / 256, 256, 0, ctx->stream>>>(d_offs, bucket_lo, bucket_hi, nb, (uint32_t)n_keys, s->lev[0]);
  count_launch(ctx);
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess && n_keys > 0)
    e = cudaMemcpyAsync(s->keys, d_keys, (size_t)n_keys * key_bytes, cudaMemcpyDeviceToDevice, ctx->stream);
  if (e != cudaSuccess) { kmsc_set_free(ctx, s); return cuda_fail(e, "import range", __FILE__, __LINE__); }
  int rc = set_build_levels(ctx, s);
  if (rc != KMSC_OK) { kmsc_set_free(ctx, s); return rc; }
  s->has_dups = -1;  // checked on first use by the pair counts
  s->b_lo = bucket_lo; s->b_hi = bucket_hi;
  *out = s;
  return KMSC_OK;
}

void kmsc_free_host(void* p) { free(p); }

int kmsc_host_alloc_pinned(size_t bytes, void** out) {
  if (!out) { set_error("NULL argument"); return KMSC_E_INVALID; }
  *out = nullptr;
  KMSC_CUDA(cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocPortable));
  return KMSC_OK;
}
void kmsc_host_free_pinned(void* p) {
  if (p) cudaFreeHost(p);
}

}  // extern "C"
