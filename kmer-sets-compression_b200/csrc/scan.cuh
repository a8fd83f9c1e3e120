// scan.cuh -- small device-wide exclusive scan (uint32 in, uint32 out) used by the
// CSR builders: counts per fine bucket -> offsets. Three launches: block sums,
// single-block scan of the sums, final pass. out may alias in. out has n + 1
// entries (out[n] = total). Totals must stay below 2^32 (a set holds < 2^32 keys).
#pragma once
#include "kmsc_common.cuh"

namespace kmsc {

constexpr int kSThreads = 512;
constexpr int kSPer = 8;
constexpr int kSBlock = kSThreads * kSPer;

static __global__ void scan_block_sums_kernel(const uint32_t* __restrict__ in, uint64_t n,
                                       uint32_t* __restrict__ bsum) {
  __shared__ uint32_t red[32];
  const uint64_t base = (uint64_t)blockIdx.x * kSBlock + (uint64_t)threadIdx.x * kSPer;
  uint32_t s = 0;
#pragma unroll
  for (int i = 0; i < kSPer; i++)
    if (base + i < n) s += in[base + i];
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    s = (threadIdx.x < kSThreads / 32) ? red[threadIdx.x] : 0u;
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (threadIdx.x == 0) bsum[blockIdx.x] = s;
  }
}

// one block of 1024 threads: exclusive scan of up to a few hundred thousand sums
static __global__ void scan_sums_kernel(uint32_t* __restrict__ bsum, int n_blocks, uint32_t* __restrict__ total) {
  __shared__ uint32_t wsum[32];
  __shared__ uint32_t carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int base = 0; base < n_blocks; base += 1024) {
    const int i = base + threadIdx.x;
    const uint32_t v = i < n_blocks ? bsum[i] : 0u;
    uint32_t inc = v;
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += t;
    }
    if (lane == 31) wsum[warp] = inc;
    __syncthreads();
    if (warp == 0) {
      const uint32_t w = wsum[lane];
      uint32_t winc = w;
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, winc, o);
        if (lane >= o) winc += t;
      }
      wsum[lane] = winc - w;
    }
    __syncthreads();
    const uint32_t ex = carry + wsum[warp] + inc - v;
    if (i < n_blocks) bsum[i] = ex;
    __syncthreads();
    if (threadIdx.x == 1023) carry = ex + v;
    __syncthreads();
  }
  if (threadIdx.x == 0 && total) *total = carry;
}

static __global__ void scan_final_kernel(const uint32_t* __restrict__ in, uint64_t n,
                                  const uint32_t* __restrict__ bsum, uint32_t* __restrict__ out) {
  __shared__ uint32_t wsum[32];
  const uint64_t base = (uint64_t)blockIdx.x * kSBlock + (uint64_t)threadIdx.x * kSPer;
  uint32_t v[kSPer];
  uint32_t s = 0;
#pragma unroll
  for (int i = 0; i < kSPer; i++) {
    v[i] = (base + i < n) ? in[base + i] : 0u;
    s += v[i];
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t inc = s;
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) wsum[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    const uint32_t w = (lane < kSThreads / 32) ? wsum[lane] : 0u;
    uint32_t winc = w;
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t t = __shfl_up_sync(0xffffffffu, winc, o);
      if (lane >= o) winc += t;
    }
    wsum[lane] = winc - w;
  }
  __syncthreads();
  uint32_t pre = bsum[blockIdx.x] + wsum[warp] + inc - s;
#pragma unroll
  for (int i = 0; i < kSPer; i++) {
    if (base + i < n) out[base + i] = pre;
    pre += v[i];
    if (base + i + 1 == n) out[n] = pre;
  }
}

// d_bsum: scratch of at least (n / kSBlock + 2) uint32. d_total (may be NULL)
// receives the grand total.
inline int exclusive_scan_u32(kmsc_ctx* ctx, const uint32_t* d_in, uint32_t* d_out, uint64_t n,
                              uint32_t* d_bsum, uint32_t* d_total) {
  if (n == 0) {
    KMSC_CUDA(cudaMemsetAsync(d_out, 0, 4, ctx->stream));
    if (d_total) KMSC_CUDA(cudaMemsetAsync(d_total, 0, 4, ctx->stream));
    return KMSC_OK;
  }
  const int nblk = (int)((n + kSBlock - 1) / kSBlock);
  scan_block_sums_kernel<<<nblk, kSThreads, 0, ctx->stream>>>(d_in, n, d_bsum);
  scan_sums_kernel<<<1, 1024, 0, ctx->stream>>>(d_bsum, nblk, d_total);
  scan_final_kernel<<<nblk, kSThreads, 0, ctx->stream>>>(d_in, n, d_bsum, d_out);
  count_launch(ctx, 3);
  KMSC_CUDA(cudaGetLastError());
  return KMSC_OK;
}

inline size_t scan_scratch_entries(uint64_t n) { return (size_t)(n / kSBlock + 2); }

// ---- several equally long arrays scanned by one set of launches (blockIdx.y = array) --------
constexpr int kScanMultiMax = 4;
struct ScanMulti {
  const uint32_t* in[kScanMultiMax];
  uint32_t* out[kScanMultiMax];     // n + 1 entries each (out[n] = total); may alias in
  uint32_t* bsum[kScanMultiMax];    // scratch, scan_scratch_entries(n) each
  uint32_t* total[kScanMultiMax];   // may be NULL
};

static __global__ void scan_multi_block_sums_kernel(ScanMulti m, uint64_t n) {
  __shared__ uint32_t red[32];
  const uint32_t* in = m.in[blockIdx.y];
  const uint64_t base = (uint64_t)blockIdx.x * kSBlock + (uint64_t)threadIdx.x * kSPer;
  uint32_t s = 0;
#pragma unroll
  for (int i = 0; i < kSPer; i++)
    if (base + i < n) s += in[base + i];
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    s = (threadIdx.x < kSThreads / 32) ? red[threadIdx.x] : 0u;
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (threadIdx.x == 0) m.bsum[blockIdx.y][blockIdx.x] = s;
  }
}

static __global__ void scan_multi_sums_kernel(ScanMulti m, int n_blocks) {
  __shared__ uint32_t wsum[32];
  __shared__ uint32_t carry;
  uint32_t* bsum = m.bsum[blockIdx.x];
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int base = 0; base < n_blocks; base += 1024) {
    const int i = base + threadIdx.x;
    const uint32_t v = i < n_blocks ? bsum[i] : 0u;
    uint32_t inc = v;
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += t;
    }
    if (lane == 31) wsum[warp] = inc;
    __syncthreads();
    if (warp == 0) {
      const uint32_t w = wsum[lane];
      uint32_t winc = w;
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, winc, o);
        if (lane >= o) winc += t;
      }
      wsum[lane] = winc - w;
    }
    __syncthreads();
    const uint32_t ex = carry + wsum[warp] + inc - v;
    if (i < n_blocks) bsum[i] = ex;
    __syncthreads();
    if (threadIdx.x == 1023) carry = ex + v;
    __syncthreads();
  }
  if (threadIdx.x == 0 && m.total[blockIdx.x]) *m.total[blockIdx.x] = carry;
}

static __global__ void scan_multi_final_kernel(ScanMulti m, uint64_t n) {
  __shared__ uint32_t wsum[32];
  const uint32_t* in = m.in[blockIdx.y];
  uint32_t* out = m.out[blockIdx.y];
  const uint64_t base = (uint64_t)blockIdx.x * kSBlock + (uint64_t)threadIdx.x * kSPer;
  uint32_t v[kSPer];
  uint32_t s = 0;
#pragma unroll
  for (int i = 0; i < kSPer; i++) {
    v[i] = (base + i < n) ? in[base + i] : 0u;
    s += v[i];
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t inc = s;
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) wsum[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    const uint32_t w = (lane < kSThreads / 32) ? wsum[lane] : 0u;
    uint32_t winc = w;
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t t = __shfl_up_sync(0xffffffffu, winc, o);
      if (lane >= o) winc += t;
    }
    wsum[lane] = winc - w;
  }
  __syncthreads();
  uint32_t pre = m.bsum[blockIdx.y][blockIdx.x] + wsum[warp] + inc - s;
#pragma unroll
  for (int i = 0; i < kSPer; i++) {
    if (base + i < n) out[base + i] = pre;
    pre += v[i];
    if (base + i + 1 == n) out[n] = pre;
  }
}

// scans the first `count` arrays of m (n > 0 entries each) with three launches in all
inline int exclusive_scan_multi_u32(kmsc_ctx* ctx, const ScanMulti& m, int count, uint64_t n) {
  const int nblk = (int)((n + kSBlock - 1) / kSBlock);
  scan_multi_block_sums_kernel<<<dim3(nblk, count), kSThreads, 0, ctx->stream>>>(m, n);
  scan_multi_sums_kernel<<<count, 1024, 0, ctx->stream>>>(m, nblk);
  scan_multi_final_kernel<<<dim3(nblk, count), kSThreads, 0, ctx->stream>>>(m, n);
  count_launch(ctx, 3);
  KMSC_CUDA(cudaGetLastError());
  return KMSC_OK;
}

}  // namespace kmsc
