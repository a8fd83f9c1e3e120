// bitmap.cu -- P5: dense-bitmap Gram matrix for small k (2K <= 30 bits) on the tensor cores.
//
// Every set becomes a 2^(2K)-bit bitmap indexed by the k-mer value (K = 15: 128 MiB per
// set) and W[i][j] = sum over positions of B_i[p] & B_j[p] -- the same all-bucket
// intersection matrix the reference's GetEdgeWeight loop yields (reference
// lib/core/kmer_set_set.h:158-219) for duplicate-free sets, without any merge.
//
// The contraction is integer-ALU bound on CUDA cores (n^2 / 2 AND + POPC per 32 positions:
// SURVEY.md Appendix D; `mma.sync ... b1 .and.popc` is emulated on sm_100a). Here it runs
// as the same tcgen05 Gram as P3 (umma.cuh): a warp loads 16 bytes of 32 sets' bitmaps,
// every bitmap word (32 positions of one set) is expanded to 32 bytes of 0 / 1 -- exactly one
// row of the K-major operand of one K-step -- and D[sets x sets] += X X^T is issued as
// tcgen05.mma kind::i8 with int32 accumulators in tensor memory. Sets are processed in
// blocks of up to 256 (the accumulator tile); n > 256 loops over block pairs.
#include "kmsc_common.cuh"
#include "umma.cuh"

namespace kmsc {
namespace {

constexpr int kBgThreads = 288;   // 8 expander warps + 1 warp whose lane 0 issues the MMAs
constexpr int kBgExpanders = 256;
constexpr int kBgKS = 4;          // K-steps (bitmap words per set) per stage: one 16-byte load per lane

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

// One stage = kBgKS K-steps of a row block (A) and, for an off-diagonal block pair, a column
// block (B). K-major operand layout of one K-step (NS sets x 32 positions, umma.cuh):
//   [position half kh][set group of 8][8 sets][16 positions], LBO = NS * 16, SBO = 128.
template <int NS>
struct BgLayout {
  static constexpr uint32_t kstep_bytes = NS * 32;
  static constexpr uint32_t block_bytes = kstep_bytes * kBgKS;   // one block (A or B) of one stage
  static constexpr uint32_t stage_bytes = block_bytes * 2;       // A + B
  static constexpr uint32_t total = stage_bytes * 2 + 64;        // two stages + barriers / tmem slot
};

// warp wp covers sets [32 wp, 32 wp + 32) of a block (NS / 32 <= 8 warps take part); lane = set.
// The 16 bytes (4 bitmap words) of a stage are loaded one stage ahead of their expansion.
// one unit = 8 bitmap words (32 bytes: a full DRAM sector) per set = two stages
struct BgUnit { uint4 lo, hi; };
template <int NS>
__device__ __forceinline__ BgUnit bg_load(const uint32_t* __restrict__ bm, size_t words_per_set, int set0, int n_sets,
                                          size_t w0, int warp, int lane) {
  BgUnit v;
  v.lo = make_uint4(0u, 0u, 0u, 0u);
  v.hi = v.lo;
  const int s = set0 + warp * 32 + lane;
  if (warp < NS / 32 && s < n_sets && w0 + 8 <= words_per_set) {
    const uint4* p = reinterpret_cast<const uint4*>(bm + (size_t)s * words_per_set + w0);
    v.lo = __ldg(p);
    v.hi = __ldg(p + 1);
  }
  return v;
}

template <int NS>
__device__ __forceinline__ void bg_expand(unsigned char* dst, uint4 v, int warp, int lane) {
  if (warp >= NS / 32) return;
  const int sl = warp * 32 + lane;  // set inside the block
  const uint32_t wv[4] = {v.x, v.y, v.z, v.w};
  unsigned char* row = dst + (size_t)(sl >> 3) * 128 + (size_t)(sl & 7) * 16;
#pragma unroll
  for (int ks = 0; ks < kBgKS; ks++) {
    const uint32_t m = wv[ks];
    uint4 lo, hi;
    lo.x = umma::nibble_to_bytes(m);       lo.y = umma::nibble_to_bytes(m >> 4);
    lo.z = umma::nibble_to_bytes(m >> 8);  lo.w = umma::nibble_to_bytes(m >> 12);
    hi.x = umma::nibble_to_bytes(m >> 16); hi.y = umma::nibble_to_bytes(m >> 20);
    hi.z = umma::nibble_to_bytes(m >> 24); hi.w = umma::nibble_to_bytes(m >> 28);
    unsigned char* k0 = row + (size_t)ks * BgLayout<NS>::kstep_bytes;
    *reinterpret_cast<uint4*>(k0) = lo;                      // positions 0..15
    *reinterpret_cast<uint4*>(k0 + (size_t)NS * 16) = hi;    // positions 16..31
  }
}

// grid: persistent CTAs over stages of kBgKS words; (row_set0, col_set0) = the block pair
template <int NS>
__global__ void __launch_bounds__(kBgThreads)
bitmap_gram_tc_kernel(const uint32_t* __restrict__ bm, size_t words_per_set, int n_sets, int row_set0, int col_set0,
                      unsigned long long* __restrict__ W) {
  using LY = BgLayout<NS>;
  constexpr int MMA_M = NS == 64 ? 64 : 128;
  constexpr int TMEM_COLS = NS == 256 ? 512 : NS;
  extern __shared__ __align__(128) unsigned char bg_smem[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(bg_smem + LY::stage_bytes * 2);   // full[2], empty[2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bg_smem + LY::stage_bytes * 2 + 32);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bool diag = row_set0 == col_set0;
  // full[b]: every expander thread arrives once its rows of stage buffer b are written (and
  // fenced for the async proxy); empty[b]: tcgen05.commit arrives when the MMAs that read
  // buffer b are done. No block-wide barrier inside the loop: expansion and issue overlap.
  uint64_t* full = bar;
  uint64_t* empty = bar + 2;
  if (tid == 0) {
    umma::mbar_init(&full[0], kBgExpanders);
    umma::mbar_init(&full[1], kBgExpanders);
    umma::mbar_init(&empty[0], 1);
    umma::mbar_init(&empty[1], 1);
    umma::mbar_fence_init();
  }
  if (warp == 0) umma::tmem_alloc(tmem_slot, TMEM_COLS);
  umma::fence_before_thread_sync();
  __syncthreads();
  umma::fence_after_thread_sync();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t smem_addr = umma::smem_u32(bg_smem);
  constexpr uint32_t idesc = umma::make_idesc_u8(MMA_M, NS, 0, 0);
  const size_t n_units = words_per_set / (2 * kBgKS);   // words_per_set is a multiple of 8
  const size_t G = gridDim.x;
  uint32_t n_mine = 0;  // stages of this CTA (two per unit)
  if (warp < kBgExpanders / 32) {
    // ---- expanders: the units of this CTA are loaded two units ahead of their expansion -----
    BgUnit a0u, a1u, b0u, b1u;
    a0u = bg_load<NS>(bm, words_per_set, row_set0, n_sets, (size_t)blockIdx.x * 8, warp, lane);
    a1u = bg_load<NS>(bm, words_per_set, row_set0, n_sets, ((size_t)blockIdx.x + G) * 8, warp, lane);
    b0u = a0u; b1u = a1u;
    if (!diag) {
      b0u = bg_load<NS>(bm, words_per_set, col_set0, n_sets, (size_t)blockIdx.x * 8, warp, lane);
      b1u = bg_load<NS>(bm, words_per_set, col_set0, n_sets, ((size_t)blockIdx.x + G) * 8, warp, lane);
    }
    for (size_t u = blockIdx.x; u < n_units; u += G) {
      const BgUnit ca = a0u, cb = b0u;
      a0u = a1u; b0u = b1u;
      a1u = bg_load<NS>(bm, words_per_set, row_set0, n_sets, (u + 2 * G) * 8, warp, lane);
      if (!diag) b1u = bg_load<NS>(bm, words_per_set, col_set0, n_sets, (u + 2 * G) * 8, warp, lane);
#pragma unroll
      for (int h = 0; h < 2; h++, n_mine++) {
        const int buf = (int)(n_mine & 1u);
        const uint32_t use = n_mine >> 1;           // how often this buffer was used before
        if (use > 0) umma::mbar_wait(&empty[buf], (use - 1) & 1u);
        unsigned char* sa = bg_smem + (size_t)buf * LY::stage_bytes;
        bg_expand<NS>(sa, h ? ca.hi : ca.lo, warp, lane);
        if (!diag) bg_expand<NS>(sa + LY::block_bytes, h ? cb.hi : cb.lo, warp, lane);
        umma::fence_async_smem();
        umma::mbar_arrive(&full[buf]);
      }
    }
  } else if (lane == 0) {
    // ---- MMA issuer (one thread) ------------------------------------------------------------
    bool started = false;
    size_t my_units = 0;
    for (size_t u = blockIdx.x; u < n_units; u += G) my_units++;
    for (size_t st = 0; st < 2 * my_units; st++, n_mine++) {
      const int buf = (int)(n_mine & 1u);
      const uint32_t use = n_mine >> 1;
      umma::mbar_wait(&full[buf], use & 1u);
      umma::fence_after_thread_sync();
      const uint32_t a0 = smem_addr + (uint32_t)buf * LY::stage_bytes;
      const uint32_t b0 = diag ? a0 : a0 + LY::block_bytes;
      const uint64_t ad0 = umma::make_smem_desc(a0, NS * 16u, 128u);
      const uint64_t bd0 = umma::make_smem_desc(b0, NS * 16u, 128u);
#pragma unroll
      for (int ks = 0; ks < kBgKS; ks++) {
        // the start-address field (16-byte units) is the low 14 bits: step it by one K-step
        const uint64_t ad = ad0 + (uint64_t)((ks * LY::kstep_bytes) >> 4);
        const uint64_t bd = bd0 + (uint64_t)((ks * LY::kstep_bytes) >> 4);
        umma::mma_u8(tmem_base, ad, bd, idesc, started ? 1u : 0u);
        if (NS == 256) {
          // rows 128..255: the A operand starts 16 set groups (2048 bytes) further, accumulators at column 256
          umma::mma_u8(tmem_base + 256u, ad + (uint64_t)((16u * 128u) >> 4), bd, idesc, started ? 1u : 0u);
        }
        started = true;
      }
      umma::mma_commit(&empty[buf]);
    }
    // drain: the last commit on each buffer
    const uint32_t u0 = (n_mine + 1) >> 1, u1 = n_mine >> 1;
    if (u0 > 0) umma::mbar_wait(&empty[0], (u0 - 1) & 1u);
    if (u1 > 0) umma::mbar_wait(&empty[1], (u1 - 1) & 1u);
  }
  umma::fence_before_thread_sync();
  __syncthreads();
  umma::fence_after_thread_sync();
  const uint32_t uses0 = (uint32_t)((n_units > blockIdx.x) ? 1 : 0), uses1 = 0;
  if ((uses0 + uses1) > 0 && warp < 4) {
    constexpr int HALVES = NS == 256 ? 2 : 1;
#pragma unroll 1
    for (int half = 0; half < HALVES; half++) {
      int row;
      if (NS == 64) row = lane < 16 ? warp * 16 + lane : -1;
      else row = half * 128 + warp * 32 + lane;
      const int si = row >= 0 ? row_set0 + row : -1;
#pragma unroll 1
      for (int c0 = 0; c0 < NS; c0 += 32) {
        uint32_t v[32];
        umma::tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(half * 256 + c0), v);
        umma::tmem_ld_wait();
        if (si >= 0 && si < n_sets) {
#pragma unroll
          for (int j = 0; j < 32; j++) {
            const int sj = col_set0 + c0 + j;
            if (sj < n_sets && v[j] != 0) {
              atomicAdd(&W[(size_t)si * n_sets + sj], (unsigned long long)v[j]);
              if (!diag) atomicAdd(&W[(size_t)sj * n_sets + si], (unsigned long long)v[j]);
            }
          }
        }
      }
    }
  }
  umma::fence_before_thread_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tmem_base, TMEM_COLS);
}

template <int NS>
static int launch_bitmap_gram(kmsc_ctx* ctx, const uint32_t* d_bm, size_t words_per_set, int n, int row0, int col0,
                              unsigned long long* d_W) {
  const size_t smem = BgLayout<NS>::total;
  auto kern = bitmap_gram_tc_kernel<NS>;
  KMSC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  // resident CTAs per SM are bounded by shared memory and by 512 TMEM columns
  const int tmem_cols = NS == 256 ? 512 : NS;
  int per_sm = (int)((size_t)227 * 1024 / (smem + 1024));
  if (per_sm > 512 / tmem_cols) per_sm = 512 / tmem_cols;
  if (per_sm > 2) per_sm = 2;
  if (per_sm < 1) per_sm = 1;
  const size_t n_units = words_per_set / (2 * kBgKS);
  size_t grid = (size_t)ctx->sm_count * per_sm;
  if (grid > n_units) grid = n_units;
  kern<<<(unsigned)grid, kBgThreads, smem, ctx->stream>>>(d_bm, words_per_set, n, row0, col0, d_W);
  count_launch(ctx);
  KMSC_CUDA(cudaGetLastError());
  return KMSC_OK;
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
  // words per set, padded to a multiple of 8 so every 32-byte unit is aligned and in bounds
  size_t words_per_set = ((size_t)1 << (2 * s0->K)) / 32;
  if (words_per_set < 8) words_per_set = 8;
  words_per_set = (words_per_set + 7) & ~(size_t)7;
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
  int rc = KMSC_OK;
  if (n <= 64) rc = launch_bitmap_gram<64>(ctx, d_bm, words_per_set, n, 0, 0, d_W);
  else if (n <= 128) rc = launch_bitmap_gram<128>(ctx, d_bm, words_per_set, n, 0, 0, d_W);
  else {
    // blocks of 256 sets; the diagonal pairs give both triangles, off-diagonal pairs are mirrored
    for (int r0 = 0; r0 < n && rc == KMSC_OK; r0 += 256)
      for (int c0 = r0; c0 < n && rc == KMSC_OK; c0 += 256) rc = launch_bitmap_gram<256>(ctx, d_bm, words_per_set, n, r0, c0, d_W);
  }
  if (rc != KMSC_OK) { cleanup(); return rc; }
  e = cudaMemcpyAsync(out, d_W, (size_t)n * n * 8, cudaMemcpyDeviceToHost, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  cleanup();
  if (e != cudaSuccess) return cuda_fail(e, "bitmap gram", __FILE__, __LINE__);
  return KMSC_OK;
}
