// bitmap.cu -- P5: dense-bitmap Gram matrix for small k (2K <= 30 bits).
//
// Every set becomes a 2^(2K)-bit bitmap indexed by the k-mer value (K = 15: 128 MiB per
// set) and W[i][j] = sum over words of popc(B_i[w] & B_j[w]) -- the same all-bucket
// intersection matrix the reference's GetEdgeWeight loop yields (reference
// lib/core/kmer_set_set.h:158-219) for duplicate-free sets, without any merge.
//
// This contraction is integer-ALU bound, not HBM bound: with
// 64 x 64 set tiles each bitmap is re-read n/64 times but every loaded word feeds 64
// AND+POPC. On sm_100a `mma.sync ... b1 .and.popc` is emulated (bit-plane LOP3 + IMMA),
// so the kernel stays on CUDA-core LOP3 + POPC with a 4 x 4 register tile per thread.
#include "kmsc_common.cuh"

namespace kmsc {
namespace {

constexpr int kTile = 64;    // sets per tile side
constexpr int kWch = 32;     // words per shared-memory stage

template <typename KeyT>
__global__ void bitmap_fill_kernel(const KeyT* __restrict__ keys, const uint32_t* __restrict__ offs, int n_buckets,
                                   int key_bits, uint32_t* __restrict__ bm) {
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  for (int b = blockIdx.x * wpb + (threadIdx.x >> 5); b < n_buckets; b += gridDim.x * wpb) {
    const uint32_t lo = offs[b], hi = offs[b + 1];
    const uint32_t top = (uint32_t)b << key_bits;
    for (uint32_t i = lo + lane; i < hi; i += 32) {
      const uint32_t v = top | (uint32_t)keys[i];
      atomicOr(&bm[v >> 5], 1u << (v & 31));
    }
  }
}

// grid = (tile pairs ti <= tj, word splits). block = 256 threads = 16 x 16, each 4 x 4 sets.
__global__ void __launch_bounds__(256)
bitmap_gram_kernel(const uint32_t* __restrict__ bm, size_t words_per_set, int n_sets, int n_tiles,
                   size_t words_per_split, unsigned long long* __restrict__ W) {
  __shared__ __align__(16) uint32_t As[kWch][kTile];
  __shared__ __align__(16) uint32_t Bs[kWch][kTile];
  // decode the tile pair
  int p = blockIdx.x, ti = 0;
  int row = n_tiles;
  while (p >= row) { p -= row; ti++; row--; }
  const int tj = ti + p;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const size_t w0 = (size_t)blockIdx.y * words_per_split;
  const size_t w1 = min(words_per_set, w0 + words_per_split);
  uint32_t acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) acc[i][j] = 0;
  unsigned long long acc64[4][4];
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) acc64[i][j] = 0;
  int since_flush = 0;
  for (size_t w = w0; w < w1; w += kWch) {
    // stage kWch words of 64 + 64 sets: thread loads (set = tid / 4 .. , 8 words)
    for (int e = tid; e < kTile * kWch; e += 256) {
      const int s = e / kWch, ww = e % kWch;
      const int sa = ti * kTile + s, sb = tj * kTile + s;
      const size_t wi = w + ww;
      As[ww][s] = (sa < n_sets && wi < w1) ? bm[(size_t)sa * words_per_set + wi] : 0u;
      Bs[ww][s] = (sb < n_sets && wi < w1) ? bm[(size_t)sb * words_per_set + wi] : 0u;
    }
    __syncthreads();
#pragma unroll 4
    for (int ww = 0; ww < kWch; ww++) {
      const uint4 a = *reinterpret_cast<const uint4*>(&As[ww][ty * 4]);
      const uint4 b = *reinterpret_cast<const uint4*>(&Bs[ww][tx * 4]);
      const uint32_t av[4] = {a.x, a.y, a.z, a.w};
      const uint32_t bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j] += __popc(av[i] & bv[j]);
    }
    __syncthreads();
    // 32 words x 32 bits per stage: flush the 32-bit accumulators well before they can wrap
    if (++since_flush == (1 << 20)) {
#pragma unroll
      for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) { acc64[i][j] += acc[i][j]; acc[i][j] = 0; }
      since_flush = 0;
    }
  }
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const unsigned long long v = acc64[i][j] + acc[i][j];
      const int si = ti * kTile + ty * 4 + i, sj = tj * kTile + tx * 4 + j;
      if (v == 0 || si >= n_sets || sj >= n_sets) continue;
      if (ti == tj) {
        atomicAdd(&W[(size_t)si * n_sets + sj], v);  // the diagonal tile computes both (i,j) and (j,i)
      } else {
        atomicAdd(&W[(size_t)si * n_sets + sj], v);
        atomicAdd(&W[(size_t)sj * n_sets + si], v);
      }
    }
}

}  // namespace
}  // namespace kmsc

using namespace kmsc;

extern "C" int kmsc_bitmap_gram(kmsc_ctx* ctx, const kmsc_set* const* sets, int32_t n, int64_t* out) {
  if (!ctx || !sets || n < 1 || !out) { set_error("bad argument"); return KMSC_E_INVALID; }
  const kmsc_set* s0 = sets[0];
  if (!s0) { set_error("sets[0] is NULL"); return KMSC_E_INVALID; }
  for (int i = 0; i < n; i++) {
    if (!sets[i] || sets[i]->K != s0->K || sets[i]->N != s0->N || sets[i]->key_bytes != s0->key_bytes) {
      set_error("sets[%d] missing or of a different (K,N,KeyType)", i);
      return KMSC_E_INVALID;
    }
  }
  if (2 * s0->K > 30) { set_error("bitmap path needs 2K <= 30 (K <= 15), got K=%d", s0->K); return KMSC_E_INVALID; }
  KMSC_CUDA(cudaSetDevice(ctx->device));
  const size_t words_per_set = ((size_t)1 << (2 * s0->K)) / 32 > 0 ? ((size_t)1 << (2 * s0->K)) / 32 : 1;
  uint32_t* d_bm = nullptr;
  unsigned long long* d_W = nullptr;
  KMSC_CUDA(cudaMallocAsync((void**)&d_bm, words_per_set * 4 * (size_t)n, ctx->stream));
  cudaError_t e = cudaMallocAsync((void**)&d_W, (size_t)n * n * 8, ctx->stream);
  if (e != cudaSuccess) { cudaFreeAsync(d_bm, ctx->stream); return cuda_fail(e, "cudaMallocAsync W", __FILE__, __LINE__); }
  auto cleanup = [&]() { cudaFreeAsync(d_bm, ctx->stream); cudaFreeAsync(d_W, ctx->stream); };
  e = cudaMemsetAsync(d_bm, 0, words_per_set * 4 * (size_t)n, ctx->stream);
  if (e == cudaSuccess) e = cudaMemsetAsync(d_W, 0, (size_t)n * n * 8, ctx->stream);
  if (e != cudaSuccess) { cleanup(); return cuda_fail(e, "memset bitmaps", __FILE__, __LINE__); }
  const int nb = 1 << s0->N;
  for (int i = 0; i < n; i++) {
    uint32_t* bm = d_bm + (size_t)i * words_per_set;
    int blocks = (nb + 7) / 8;
    if (blocks > ctx->sm_count * 16) blocks = ctx->sm_count * 16;
    switch (s0->key_bytes) {
      case 2: bitmap_fill_kernel<uint16_t><<<blocks, 256, 0, ctx->stream>>>((const uint16_t*)sets[i]->keys, sets[i]->lev[0], nb, s0->key_bits, bm); break;
      case 4: bitmap_fill_kernel<uint32_t><<<blocks, 256, 0, ctx->stream>>>((const uint32_t*)sets[i]->keys, sets[i]->lev[0], nb, s0->key_bits, bm); break;
      default: bitmap_fill_kernel<unsigned long long><<<blocks, 256, 0, ctx->stream>>>((const unsigned long long*)sets[i]->keys, sets[i]->lev[0], nb, s0->key_bits, bm); break;
    }
    count_launch(ctx);
  }
  const int n_tiles = (n + kTile - 1) / kTile;
  const int tile_pairs = n_tiles * (n_tiles + 1) / 2;
  // split the word range so the grid covers the chip a few times over
  size_t splits = ((size_t)ctx->sm_count * 4 + tile_pairs - 1) / tile_pairs;
  const size_t max_splits = (words_per_set + kWch - 1) / kWch;
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  size_t words_per_split = (words_per_set + splits - 1) / splits;
  words_per_split = (words_per_split + kWch - 1) / kWch * kWch;
  splits = (words_per_set + words_per_split - 1) / words_per_split;
  bitmap_gram_kernel<<<dim3((unsigned)tile_pairs, (unsigned)splits), 256, 0, ctx->stream>>>(d_bm, words_per_set, n, n_tiles,
                                                                                      words_per_split, d_W);
  count_launch(ctx);
  e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaMemcpyAsync(out, d_W, (size_t)n * n * 8, cudaMemcpyDeviceToHost, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  cleanup();
  if (e != cudaSuccess) return cuda_fail(e, "bitmap gram", __FILE__, __LINE__);
  return KMSC_OK;
}
