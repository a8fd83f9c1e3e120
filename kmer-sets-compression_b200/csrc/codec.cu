// codec.cu -- P6: bucket payload codec (delta + streamvbyte-style byte codes) on the device.
//
// The reference never serialises a binary k-mer set: KmerSetCompact keeps SPSS text and
// only streamvbyte-compresses the string lengths in memory (reference
// lib/core/kmer_set_compact.h:256-266, 269-275; lemire/streamvbyte v0.4.1, 2-bit control
// codes, control bytes first, little-endian data bytes). This container applies the same
// byte-code scheme to what the device actually holds, the sorted bucketed CSR: keys are
// delta-coded inside each bucket (first key of a bucket as is) and every 32-bit value
// (bucket sizes; key deltas, low word then high word for 8-byte keys) is stored in
// 1 / 2 / 3 / 4 bytes selected by a 2-bit code, four codes per control byte, LSB first.
//
// Container "KMSC", little-endian:
//   [0]  u32 magic 0x43534D4B ("KMSC")   [4]  u32 version = 1
//   [8]  u32 K   [12] u32 N   [16] u32 key_bytes   [20] u32 words per key (1 or 2)
//   [24] u64 n_keys   [32] u64 size_data_bytes   [40] u64 key_data_bytes
//   [48] size control bytes: ceil(2^N / 4)        then size data bytes
//        key control bytes:  ceil(n_keys * wpk / 4) then key data bytes
// The test-side CPU checker restates the same layout; parity is byte-exact encode and
// exact decode (tests/test_gpu_codec_bitmap.py).
#include <cstdlib>
#include <cstring>

#include "kmsc_common.cuh"
#include "scan.cuh"

namespace kmsc {
namespace {

constexpr uint32_t kMagic = 0x43534D4Bu;
constexpr size_t kHeader = 48;

__device__ __forceinline__ uint32_t code_of(uint32_t v) {  // bytes - 1
  return v < (1u << 8) ? 0u : v < (1u << 16) ? 1u : v < (1u << 24) ? 2u : 3u;
}

// bucket sizes as values
__global__ void sizes_kernel(const uint32_t* __restrict__ offs, uint32_t nb, uint32_t* __restrict__ vals,
                             uint32_t* __restrict__ lens) {
  const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nb) return;
  const uint32_t v = offs[b + 1] - offs[b];
  vals[b] = v;
  lens[b] = code_of(v) + 1;
}

// warp per bucket: key deltas as values (WPK words per key)
template <typename KeyT, int WPK>
__global__ void deltas_kernel(const KeyT* __restrict__ keys, const uint32_t* __restrict__ offs, uint32_t nb,
                              uint32_t* __restrict__ vals, uint32_t* __restrict__ lens) {
  const int lane = threadIdx.x & 31;
  const uint32_t wpb = blockDim.x >> 5;
  for (uint32_t b = blockIdx.x * wpb + (threadIdx.x >> 5); b < nb; b += gridDim.x * wpb) {
    const uint32_t lo = offs[b], hi = offs[b + 1];
    for (uint32_t i = lo + lane; i < hi; i += 32) {
      const unsigned long long k = (unsigned long long)keys[i];
      const unsigned long long p = i > lo ? (unsigned long long)keys[i - 1] : 0ull;
      const unsigned long long d = k - p;
      const uint32_t v0 = (uint32_t)d;
      vals[(size_t)i * WPK] = v0;
      lens[(size_t)i * WPK] = code_of(v0) + 1;
      if (WPK == 2) {
        const uint32_t v1 = (uint32_t)(d >> 32);
        vals[(size_t)i * WPK + 1] = v1;
        lens[(size_t)i * WPK + 1] = code_of(v1) + 1;
      }
    }
  }
}

// thread per control byte: four values -> control byte + their data bytes
__global__ void emit_kernel(const uint32_t* __restrict__ vals, const uint32_t* __restrict__ pos /* exclusive scan of lens */,
                            uint64_t m, uint8_t* __restrict__ ctrl, uint8_t* __restrict__ data) {
  const uint64_t c = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c * 4 >= m) return;
  uint32_t cb = 0;
#pragma unroll
  for (int j = 0; j < 4; j++) {
    const uint64_t i = c * 4 + j;
    if (i >= m) break;
    const uint32_t v = vals[i];
    const uint32_t code = code_of(v);
    cb |= code << (2 * j);
    uint8_t* d = data + pos[i];
    for (uint32_t t = 0; t <= code; t++) d[t] = (uint8_t)(v >> (8 * t));
  }
  ctrl[c] = (uint8_t)cb;
}

// decode: lengths from control bytes, then values from data bytes
__global__ void ctrl_lens_kernel(const uint8_t* __restrict__ ctrl, uint64_t m, uint32_t* __restrict__ lens) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  lens[i] = ((ctrl[i >> 2] >> (2 * (i & 3))) & 3u) + 1;
}
__global__ void gather_kernel(const uint8_t* __restrict__ data, const uint32_t* __restrict__ pos, const uint32_t* __restrict__ lens_end,
                              uint64_t m, uint64_t n_data, uint32_t* __restrict__ vals, int* __restrict__ bad) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  const uint32_t p = pos[i], len = pos[i + 1] - p;
  if ((uint64_t)p + len > n_data) { atomicExch(bad, 1); vals[i] = 0; return; }
  uint32_t v = 0;
  for (uint32_t t = 0; t < len; t++) v |= (uint32_t)data[p + t] << (8 * t);
  vals[i] = v;
  (void)lens_end;
}

// warp per bucket: inclusive prefix sum of the deltas -> keys
template <typename KeyT, int WPK>
__global__ void undelta_kernel(const uint32_t* __restrict__ vals, const uint32_t* __restrict__ offs, uint32_t nb,
                               KeyT* __restrict__ keys) {
  const int lane = threadIdx.x & 31;
  const uint32_t wpb = blockDim.x >> 5;
  for (uint32_t b = blockIdx.x * wpb + (threadIdx.x >> 5); b < nb; b += gridDim.x * wpb) {
    const uint32_t lo = offs[b], hi = offs[b + 1];
    unsigned long long carry = 0;
    for (uint32_t i0 = lo; i0 < hi; i0 += 32) {
      const uint32_t i = i0 + lane;
      unsigned long long d = 0;
      if (i < hi) {
        d = vals[(size_t)i * WPK];
        if (WPK == 2) d |= (unsigned long long)vals[(size_t)i * WPK + 1] << 32;
      }
      for (int o = 1; o < 32; o <<= 1) {
        const unsigned long long t = __shfl_up_sync(0xffffffffu, d, o);
        if (lane >= o) d += t;
      }
      d += carry;
      if (i < hi) keys[i] = (KeyT)d;
      carry = __shfl_sync(0xffffffffu, d, 31);
    }
  }
}

// values + lengths on the device -> control / data bytes appended to `out` (host);
// returns the number of data bytes
static int encode_stream(kmsc_ctx* ctx, const uint32_t* d_vals, uint32_t* d_lens, uint64_t m, uint32_t* d_scratch,
                         uint8_t* d_ctrl, uint8_t* d_data, uint64_t* n_data) {
  // exclusive scan of lens in place -> positions; total = data bytes
  uint32_t* d_total = d_scratch;
  uint32_t* d_bsum = d_scratch + 4;
  KMSC_TRY(exclusive_scan_u32(ctx, d_lens, d_lens, m, d_bsum, d_total));
  if (m > 0) {
    const uint64_t n_ctrl = (m + 3) / 4;
    emit_kernel<<<(unsigned)((n_ctrl + 255) / 256), 256, 0, ctx->stream>>>(d_vals, d_lens, m, d_ctrl, d_data);
    count_launch(ctx);
    KMSC_CUDA(cudaGetLastError());
  }
  void* pin = nullptr;
  KMSC_TRY(ctx_pinned(ctx, 64, &pin));
  KMSC_CUDA(cudaMemcpyAsync(pin, d_total, 4, cudaMemcpyDeviceToHost, ctx->stream));
  KMSC_CUDA(cudaStreamSynchronize(ctx->stream));
  *n_data = *(uint32_t*)pin;
  return KMSC_OK;
}

// decoded bucket sizes: none may exceed n_keys, and their 64-bit sum is accumulated (the 32-bit
// scan total alone would accept sizes that only add up modulo 2^32)
__global__ void check_sizes_kernel(const uint32_t* __restrict__ sizes, uint32_t nb, unsigned long long n_keys,
                                   unsigned long long* __restrict__ total64, int* __restrict__ bad) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  unsigned long long v = 0;
  if (i < nb) {
    v = sizes[i];
    if (v > n_keys) atomicExch(bad, 1);
  }
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0 && v) atomicAdd(total64, v);
}

template <typename T> static void put(uint8_t* p, T v) { memcpy(p, &v, sizeof(T)); }
template <typename T> static T get(const uint8_t* p) { T v; memcpy(&v, p, sizeof(T)); return v; }

}  // namespace
}  // namespace kmsc

using namespace kmsc;

extern "C" {

int kmsc_codec_encode(kmsc_ctx* ctx, const kmsc_set* set, uint8_t** bytes, int64_t* n_bytes) {
  if (!ctx || !set || !bytes || !n_bytes) { set_error("NULL argument"); return KMSC_E_INVALID; }
  KMSC_CUDA(cudaSetDevice(ctx->device));
  const uint32_t nb = (uint32_t)1 << set->N;
  const int wpk = set->key_bytes == 8 ? 2 : 1;
  const uint64_t m_keys = (uint64_t)set->n_keys * wpk;
  if (m_keys * 4 >= ((uint64_t)1 << 32)) { set_error("set too large for one codec block (%lld keys)", (long long)set->n_keys); return KMSC_E_INVALID; }
  const uint64_t m_max = m_keys > nb ? m_keys : nb;
  // scratch: vals | lens | scan scratch | ctrl | data
  const size_t sb = scan_scratch_entries(m_max) + 8;
  KMSC_TRY(ctx->work.reserve((m_max * 2 + sb + 16) * 4));
  KMSC_TRY(ctx->work2.reserve((m_max + 3) / 4 + m_max * 4 + 64));
  uint32_t* d_vals = (uint32_t*)ctx->work.p;
  uint32_t* d_lens = d_vals + m_max;
  uint32_t* d_scratch = d_lens + m_max + 1;
  uint8_t* d_ctrl = (uint8_t*)ctx->work2.p;
  uint8_t* d_data = d_ctrl + (((m_max + 3) / 4 + 15) & ~(uint64_t)15);

  // bucket sizes
  const uint64_t n_sctrl = ((uint64_t)nb + 3) / 4;
  sizes_kernel<<<(nb + 255) / 256, 256, 0, ctx->stream>>>(set->lev[0], nb, d_vals, d_lens);
  count_launch(ctx);
  uint64_t n_sdata = 0;
  KMSC_TRY(encode_stream(ctx, d_vals, d_lens, nb, d_scratch, d_ctrl, d_data, &n_sdata));
  uint8_t* out = (uint8_t*)malloc(kHeader + n_sctrl + n_sdata + (m_keys + 3) / 4 + m_keys * 4 + 16);
  if (!out) { set_error("out of host memory"); return KMSC_E_NOMEM; }
  size_t w = kHeader;
  cudaError_t e = cudaMemcpyAsync(out + w, d_ctrl, n_sctrl, cudaMemcpyDeviceToHost, ctx->stream);
  if (e == cudaSuccess && n_sdata) e = cudaMemcpyAsync(out + w + n_sctrl, d_data, n_sdata, cudaMemcpyDeviceToHost, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  if (e != cudaSuccess) { free(out); return cuda_fail(e, "codec D2H sizes", __FILE__, __LINE__); }
  w += n_sctrl + n_sdata;

  // key deltas
  uint64_t n_kdata = 0;
  const uint64_t n_kctrl = (m_keys + 3) / 4;
  if (m_keys > 0) {
    int blocks = (int)((nb + 7) / 8);
    if (blocks > ctx->sm_count * 16) blocks = ctx->sm_count * 16;
    switch (set->key_bytes) {
      case 2: deltas_kernel<uint16_t, 1><<<blocks, 256, 0, ctx->stream>>>((const uint16_t*)set->keys, set->lev[0], nb, d_vals, d_lens); break;
      case 4: deltas_kernel<uint32_t, 1><<<blocks, 256, 0, ctx->stream>>>((const uint32_t*)set->keys, set->lev[0], nb, d_vals, d_lens); break;
      default: deltas_kernel<unsigned long long, 2><<<blocks, 256, 0, ctx->stream>>>((const unsigned long long*)set->keys, set->lev[0], nb, d_vals, d_lens); break;
    }
    count_launch(ctx);
    int rc = encode_stream(ctx, d_vals, d_lens, m_keys, d_scratch, d_ctrl, d_data, &n_kdata);
    if (rc != KMSC_OK) { free(out); return rc; }
    e = cudaMemcpyAsync(out + w, d_ctrl, n_kctrl, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess && n_kdata) e = cudaMemcpyAsync(out + w + n_kctrl, d_data, n_kdata, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) { free(out); return cuda_fail(e, "codec D2H keys", __FILE__, __LINE__); }
    w += n_kctrl + n_kdata;
  }
  put<uint32_t>(out + 0, kMagic); put<uint32_t>(out + 4, 1u);
  put<uint32_t>(out + 8, (uint32_t)set->K); put<uint32_t>(out + 12, (uint32_t)set->N);
  put<uint32_t>(out + 16, (uint32_t)set->key_bytes); put<uint32_t>(out + 20, (uint32_t)wpk);
  put<uint64_t>(out + 24, (uint64_t)set->n_keys); put<uint64_t>(out + 32, n_sdata); put<uint64_t>(out + 40, n_kdata);
  *bytes = out;
  *n_bytes = (int64_t)w;
  return KMSC_OK;
}

int kmsc_codec_decode(kmsc_ctx* ctx, const uint8_t* bytes, int64_t n_bytes, kmsc_set** out) {
  if (!ctx || !bytes || !out) { set_error("NULL argument"); return KMSC_E_INVALID; }
  if (n_bytes < (int64_t)kHeader || get<uint32_t>(bytes) != kMagic || get<uint32_t>(bytes + 4) != 1u) {
    set_error("not a KMSC version-1 container");
    return KMSC_E_FORMAT;
  }
  const int K = (int)get<uint32_t>(bytes + 8), N = (int)get<uint32_t>(bytes + 12);
  const int key_bytes = (int)get<uint32_t>(bytes + 16), wpk = (int)get<uint32_t>(bytes + 20);
  const uint64_t n_keys = get<uint64_t>(bytes + 24), n_sdata = get<uint64_t>(bytes + 32), n_kdata = get<uint64_t>(bytes + 40);
  if (N < 0 || N > 24 || (key_bytes != 2 && key_bytes != 4 && key_bytes != 8) || wpk != (key_bytes == 8 ? 2 : 1) ||
      n_keys >= ((uint64_t)1 << 30)) {
    set_error("bad KMSC header");
    return KMSC_E_FORMAT;
  }
  const uint32_t nb = (uint32_t)1 << N;
  const uint64_t m_keys = n_keys * wpk;
  if (m_keys * 4 >= ((uint64_t)1 << 32)) { set_error("bad KMSC header (too many keys for one block)"); return KMSC_E_FORMAT; }  // encode's limit
  const uint64_t n_sctrl = ((uint64_t)nb + 3) / 4, n_kctrl = (m_keys + 3) / 4;
  // every length on its own first: the sum below cannot wrap
  if (n_sdata > (uint64_t)n_bytes || n_kdata > (uint64_t)n_bytes ||
      (uint64_t)n_bytes != kHeader + n_sctrl + n_sdata + n_kctrl + n_kdata) {
    set_error("KMSC container has the wrong length");
    return KMSC_E_FORMAT;
  }
  KMSC_CUDA(cudaSetDevice(ctx->device));
  kmsc_set* s = nullptr;
  KMSC_TRY(set_alloc(ctx, K, N, key_bytes, (int64_t)n_keys, &s));
  const uint64_t m_max = m_keys > nb ? m_keys : nb;
  const size_t sb = scan_scratch_entries(m_max) + 8;
  int rc = ctx->work.reserve((m_max * 2 + sb + 32) * 4);
  if (rc == KMSC_OK) rc = ctx->work2.reserve((size_t)n_bytes + 64);
  if (rc != KMSC_OK) { kmsc_set_free(ctx, s); return rc; }
  uint32_t* d_vals = (uint32_t*)ctx->work.p;
  uint32_t* d_lens = d_vals + m_max;         // m + 1 entries after the scan
  uint32_t* d_scratch = d_lens + m_max + 2;  // [0] total, [1] bad flag, [4..] block sums
  uint8_t* d_bytes = (uint8_t*)ctx->work2.p;
  cudaError_t e = cudaMemcpyAsync(d_bytes, bytes, (size_t)n_bytes, cudaMemcpyHostToDevice, ctx->stream);
  if (e == cudaSuccess) e = cudaMemsetAsync(d_scratch, 0, 16, ctx->stream);
  if (e != cudaSuccess) { kmsc_set_free(ctx, s); return cuda_fail(e, "codec H2D", __FILE__, __LINE__); }
  int* d_bad = (int*)(d_scratch + 1);
  auto decode_stream = [&](const uint8_t* ctrl, const uint8_t* data, uint64_t m, uint64_t n_data) -> int {
    if (m == 0) return KMSC_OK;
    ctrl_lens_kernel<<<(unsigned)((m + 255) / 256), 256, 0, ctx->stream>>>(ctrl, m, d_lens);
    KMSC_TRY(exclusive_scan_u32(ctx, d_lens, d_lens, m, d_scratch + 4, d_scratch));
    gather_kernel<<<(unsigned)((m + 255) / 256), 256, 0, ctx->stream>>>(data, d_lens, nullptr, m, n_data, d_vals, d_bad);
    count_launch(ctx, 2);
    KMSC_CUDA(cudaGetLastError());
    return KMSC_OK;
  };
  // bucket sizes -> lev[0]
  const uint8_t* p = d_bytes + kHeader;
  rc = decode_stream(p, p + n_sctrl, nb, n_sdata);
  unsigned long long* d_total64 = nullptr;
  if (rc == KMSC_OK) rc = ctx->small.reserve(64);
  if (rc == KMSC_OK) {
    d_total64 = (unsigned long long*)ctx->small.p;
    e = cudaMemsetAsync(d_total64, 0, 8, ctx->stream);
    if (e != cudaSuccess) rc = cuda_fail(e, "codec decode", __FILE__, __LINE__);
  }
  if (rc == KMSC_OK) {
    check_sizes_kernel<<<(nb + 255) / 256, 256, 0, ctx->stream>>>(d_vals, nb, n_keys, d_total64, d_bad);
    count_launch(ctx);
    rc = exclusive_scan_u32(ctx, d_vals, s->lev[0], nb, d_scratch + 4, d_scratch + 2);
  }
  p += n_sctrl + n_sdata;
  // the decoded bucket sizes must add up to n_keys before any key is written
  struct { uint32_t total, bad, keys_total, pad; } host;
  unsigned long long total64 = 0;
  if (rc == KMSC_OK) {
    e = cudaMemcpyAsync(&host, d_scratch, 16, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(&total64, d_total64, 8, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) rc = cuda_fail(e, "codec decode", __FILE__, __LINE__);
    else if (host.bad || total64 != n_keys || host.keys_total != (uint32_t)n_keys) { set_error("corrupt KMSC container (bucket sizes)"); rc = KMSC_E_FORMAT; }
  }
  // key deltas -> keys
  if (rc == KMSC_OK && m_keys > 0) {
    rc = decode_stream(p, p + n_kctrl, m_keys, n_kdata);
    if (rc == KMSC_OK) {
      int blocks = (int)((nb + 7) / 8);
      if (blocks > ctx->sm_count * 16) blocks = ctx->sm_count * 16;
      switch (key_bytes) {
        case 2: undelta_kernel<uint16_t, 1><<<blocks, 256, 0, ctx->stream>>>(d_vals, s->lev[0], nb, (uint16_t*)s->keys); break;
        case 4: undelta_kernel<uint32_t, 1><<<blocks, 256, 0, ctx->stream>>>(d_vals, s->lev[0], nb, (uint32_t*)s->keys); break;
        default: undelta_kernel<unsigned long long, 2><<<blocks, 256, 0, ctx->stream>>>(d_vals, s->lev[0], nb, (unsigned long long*)s->keys); break;
      }
      count_launch(ctx);
    }
  }
  // no value may run past its stream
  if (rc == KMSC_OK) {
    e = cudaMemcpyAsync(&host, d_scratch, 16, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) rc = cuda_fail(e, "codec decode", __FILE__, __LINE__);
    else if (host.bad) { set_error("corrupt KMSC container (key data)"); rc = KMSC_E_FORMAT; }
  }
  if (rc == KMSC_OK) rc = set_build_levels(ctx, s);
  if (rc == KMSC_OK) rc = set_check_dups(ctx, s);
  if (rc != KMSC_OK) { kmsc_set_free(ctx, s); return rc; }
  *out = s;
  return KMSC_OK;
}



}  // extern "C"
