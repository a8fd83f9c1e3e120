// pair_rows.cu -- P3 row mode: weight(rows[r], l) for a few row sets against all n sets.
//
// Replaces the 3n-2 GetEdgeWeight calls that follow every merge of the reference's greedy loop
// (lib/core/kmer_set_set.h:385-425; the merge itself is :158-184): rows j, k and the new node
// against every node. The reference runs one two-pointer merge per (row, column) pair and bucket;
// a full n x n matrix pass (the round-1 implementation of this entry point) reads and masks every
// key of every set for n^2 results of which 3n are wanted.
//
// Here every key of every column set is read ONCE and merged against the row runs, which are
// staged in shared memory and shared by all columns: a CTA walks sub-chunks of 128 fine buckets;
// per sub-chunk the row sets' key slices (contiguous: the sets are sorted by fine bucket) are copied
// to shared memory, then warp w takes the columns w, w + W, ...: lane = fine bucket, the lane merges
// the column's run (about 10 keys, straight from global memory) with each row's run in shared
// memory -- the reference's own loop, so duplicate keys count with min multiplicity exactly as
// there. Counts are kept per (row, column) in shared memory (one warp owns a column: no atomics)
// and added to the result once per CTA. Roofline: HBM, n * keys * sizeof(KeyType) bytes.
#include <cstring>

#include "kmsc_common.cuh"

namespace kmsc {
namespace {

constexpr int kRowsMax = 4;      // row sets per launch
constexpr int kSubFb = 128;      // fine buckets per sub-chunk
constexpr int kRowsThreads = 256;

struct RowSet { const void* keys; const uint32_t* lev; };

template <typename KeyT>
__global__ void __launch_bounds__(kRowsThreads) rows_kernel(const RowSet* __restrict__ cols, int n, const RowSet* __restrict__ rows,
                                                           int R, uint32_t NF, int f, const uint32_t* __restrict__ sel_bitmap,
                                                           uint32_t cap, unsigned long long* __restrict__ out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  // acc[kRowsMax][n] u64 | soff[kRowsMax][kSubFb + 1] u32 | staged[kRowsMax] | srow[kRowsMax][cap] keys
  unsigned long long* acc = (unsigned long long*)smem_raw;
  uint32_t* soff = (uint32_t*)(acc + (size_t)kRowsMax * n);
  uint32_t* rbase = soff + kRowsMax * (kSubFb + 1);   // first key index of the sub-chunk in each row set
  uint32_t* staged = rbase + kRowsMax;
  KeyT* srow = (KeyT*)(smem_raw + (((size_t)kRowsMax * n * 8 + (kRowsMax * (kSubFb + 1) + 2 * kRowsMax) * 4 + 15) & ~(size_t)15));
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int i = threadIdx.x; i < kRowsMax * n; i += blockDim.x) acc[i] = 0;
  const uint32_t n_sub = (NF + kSubFb - 1) / kSubFb;
  for (uint32_t sc = blockIdx.x; sc < n_sub; sc += gridDim.x) {
    const uint32_t x0 = sc * kSubFb;
    const uint32_t nx = min((uint32_t)kSubFb, NF - x0);
    // any selected bucket in this sub-chunk? (uniform over the CTA)
    if (sel_bitmap) {
      const uint32_t b0 = x0 >> f, b1 = (x0 + nx - 1) >> f;
      bool any = false;
      for (uint32_t b = b0; b <= b1 && !any; b++) any = (sel_bitmap[b >> 5] >> (b & 31)) & 1u;
      if (!any) continue;
    }
    __syncthreads();  // the previous sub-chunk's merges are done with soff / srow
    for (int r = 0; r < R; r++) {
      const uint32_t* lev = rows[r].lev;
      for (uint32_t i = threadIdx.x; i <= nx; i += blockDim.x) soff[r * (kSubFb + 1) + i] = lev[x0 + i];
    }
    __syncthreads();
    for (int r = 0; r < R; r++) {
      const uint32_t a = soff[r * (kSubFb + 1)], e = soff[r * (kSubFb + 1) + nx];
      const bool fits = e - a <= cap;
      if (threadIdx.x == 0) { staged[r] = fits ? 1u : 0u; rbase[r] = a; }
      if (fits) {
        const KeyT* src = (const KeyT*)rows[r].keys + a;
        KeyT* dst = srow + (size_t)r * cap;
        for (uint32_t i = threadIdx.x; i < e - a; i += blockDim.x) dst[i] = src[i];
      }
    }
    __syncthreads();
    for (int l = warp; l < n; l += nw) {
      const KeyT* ck = (const KeyT*)cols[l].keys;
      const uint32_t* clev = cols[l].lev;
      uint32_t cnt[kRowsMax] = {0, 0, 0, 0};
      for (uint32_t g = 0; g < nx; g += 32) {
        const uint32_t i = g + lane;
        if (i >= nx) continue;
        const uint32_t x = x0 + i;
        if (sel_bitmap && !((sel_bitmap[(x >> f) >> 5] >> ((x >> f) & 31)) & 1u)) continue;
        const uint32_t ca = clev[x], ce = clev[x + 1];
        if (ca == ce) continue;
#pragma unroll
        for (int r = 0; r < kRowsMax; r++) {
          if (r >= R) break;
          const uint32_t ra = soff[r * (kSubFb + 1) + i], re = soff[r * (kSubFb + 1) + i + 1];
          if (ra == re) continue;
          // the reference's two-pointer merge (kmer_set_set.h:165-180)
          const KeyT* rk = staged[r] ? (const KeyT*)(srow + (size_t)r * cap) - rbase[r] : (const KeyT*)rows[r].keys;
          uint32_t p = ca, q = ra, c = 0;
          KeyT u = ck[p], v = rk[q];
          while (true) {
            if (u < v) { if (++p == ce) break; u = ck[p]; }
            else if (u > v) { if (++q == re) break; v = rk[q]; }
            else { c++; ++p; ++q; if (p == ce || q == re) break; u = ck[p]; v = rk[q]; }
          }
          cnt[r] += c;
        }
      }
#pragma unroll
      for (int r = 0; r < kRowsMax; r++) {
        if (r >= R) break;
        uint32_t c = cnt[r];
        for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
        if (lane == 0 && c) acc[(size_t)r * n + l] += c;   // this warp owns column l
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < R * n; i += blockDim.x)
    if (acc[i]) atomicAdd(&out[i], acc[i]);
}

template <typename KeyT>
int launch_rows(kmsc_ctx* ctx, const RowSet* d_cols, int n, const RowSet* d_rows, int R, uint32_t NF, int f,
                const uint32_t* d_bitmap, unsigned long long* d_out) {
  // row staging: what shared memory allows next to the counters; a sub-chunk whose row slice is
  // larger reads that row from global memory instead
  const size_t fixed = (((size_t)kRowsMax * n * 8 + (kRowsMax * (kSubFb + 1) + 2 * kRowsMax) * 4 + 15) & ~(size_t)15);
  size_t budget = 96 * 1024;
  if (fixed + 4096 > budget) budget = fixed + 16 * 1024;
  uint32_t cap = (uint32_t)((budget - fixed) / (kRowsMax * sizeof(KeyT)));
  if (cap > 8192) cap = 8192;
  const size_t smem = fixed + (size_t)kRowsMax * cap * sizeof(KeyT);
  if (smem > 200 * 1024) { set_error("row mode: %d column sets need more shared memory than a CTA has", n); return KMSC_E_INVALID; }
  KMSC_CUDA(cudaFuncSetAttribute(rows_kernel<KeyT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const uint32_t n_sub = (NF + kSubFb - 1) / kSubFb;
  int per_sm = (int)((200 * 1024) / smem);
  if (per_sm > 4) per_sm = 4;
  if (per_sm < 1) per_sm = 1;
  unsigned grid = (unsigned)(ctx->sm_count * per_sm);
  if (grid > n_sub) grid = n_sub;
  rows_kernel<KeyT><<<grid, kRowsThreads, smem, ctx->stream>>>(d_cols, n, d_rows, R, NF, f, d_bitmap, cap, d_out);
  count_launch(ctx);
  KMSC_CUDA(cudaGetLastError());
  return KMSC_OK;
}

}  // namespace
}  // namespace kmsc

using namespace kmsc;

extern "C" int kmsc_pair_counts_rows(kmsc_ctx* ctx, const kmsc_set* const* sets, int32_t n, const int32_t* rows,
                                     int32_t n_rows, const int32_t* bucket_ids, int32_t n_ids, int64_t* out) {
  if (!ctx || !out || !rows || !sets || n < 1 || n_rows < 0) { set_error("bad argument"); return KMSC_E_INVALID; }
  for (int r = 0; r < n_rows; r++)
    if (rows[r] < 0 || rows[r] >= n) { set_error("row %d out of range", rows[r]); return KMSC_E_INVALID; }
  const kmsc_set* s0 = sets[0];
  if (!s0) { set_error("sets[0] is NULL"); return KMSC_E_INVALID; }
  for (int i = 0; i < n; i++) {
    if (!sets[i]) { set_error("sets[%d] is NULL", i); return KMSC_E_INVALID; }
    if (sets[i]->K != s0->K || sets[i]->N != s0->N || sets[i]->key_bytes != s0->key_bytes) {
      set_error("sets have different (K,N,KeyType)");
      return KMSC_E_INVALID;
    }
  }
  if (n_rows == 0) return KMSC_OK;
  KMSC_CUDA(cudaSetDevice(ctx->device));
  for (int i = 0; i < n; i++) KMSC_TRY(set_ensure_levels(ctx, sets[i]));
  const int nb = 1 << s0->N;
  // bucket selection (an id listed twice counts once, like the reference's map, :127-131)
  std::vector<uint32_t> bitmap;
  if (bucket_ids) {
    bitmap.assign((size_t)(nb + 31) / 32, 0u);
    for (int i = 0; i < n_ids; i++) {
      if (bucket_ids[i] < 0 || bucket_ids[i] >= nb) { set_error("bucket id %d out of range", bucket_ids[i]); return KMSC_E_INVALID; }
      bitmap[(size_t)bucket_ids[i] >> 5] |= 1u << (bucket_ids[i] & 31);
    }
  }
  // fine level: runs of about 8-16 keys per column set
  int64_t max_keys = 1;
  for (int i = 0; i < n; i++) max_keys = std::max<int64_t>(max_keys, sets[i]->n_keys);
  int f = 0;
  while (f < s0->max_level && (max_keys >> (s0->N + f)) > 12) f++;
  const uint32_t NF = (uint32_t)nb << f;

  // device tables: column descriptors | row descriptors | bitmap | out[n_rows][n]
  const size_t sz_cols = ((size_t)n * sizeof(RowSet) + 15) & ~(size_t)15;
  const size_t sz_rows = ((size_t)kRowsMax * sizeof(RowSet) + 15) & ~(size_t)15;
  const size_t sz_bitmap = (bitmap.size() * 4 + 15) & ~(size_t)15;
  const size_t sz_out = (size_t)n_rows * n * 8;
  const int n_groups = (n_rows + kRowsMax - 1) / kRowsMax;
  KMSC_TRY(ctx->work.reserve(sz_cols + sz_rows * n_groups + sz_bitmap + sz_out + 64));
  unsigned char* dv = (unsigned char*)ctx->work.p;
  void* pin = nullptr;
  KMSC_TRY(ctx_pinned(ctx, sz_cols + sz_rows * n_groups + sz_bitmap + sz_out + 64, &pin));
  unsigned char* hv = (unsigned char*)pin;
  RowSet* hc = (RowSet*)hv;
  for (int i = 0; i < n; i++) hc[i] = RowSet{sets[i]->keys, sets[i]->lev[f]};
  for (int g = 0; g < n_groups; g++) {
    RowSet* hr = (RowSet*)(hv + sz_cols + sz_rows * g);
    for (int r = 0; r < kRowsMax; r++) {
      const int q = g * kRowsMax + r;
      const kmsc_set* s = sets[rows[q < n_rows ? q : n_rows - 1]];
      hr[r] = RowSet{s->keys, s->lev[f]};
    }
  }
  if (!bitmap.empty()) memcpy(hv + sz_cols + sz_rows * n_groups, bitmap.data(), bitmap.size() * 4);
  const size_t sz_tab = sz_cols + sz_rows * n_groups + sz_bitmap;
  KMSC_CUDA(cudaMemcpyAsync(dv, hv, sz_tab, cudaMemcpyHostToDevice, ctx->stream));
  unsigned long long* d_out = (unsigned long long*)(dv + ((sz_tab + 15) & ~(size_t)15));
  KMSC_CUDA(cudaMemsetAsync(d_out, 0, sz_out, ctx->stream));
  const uint32_t* d_bitmap = bitmap.empty() ? nullptr : (const uint32_t*)(dv + sz_cols + sz_rows * n_groups);
  for (int g = 0; g < n_groups; g++) {
    const int R = std::min(kRowsMax, n_rows - g * kRowsMax);
    const RowSet* d_rows = (const RowSet*)(dv + sz_cols + sz_rows * g);
    unsigned long long* o = d_out + (size_t)g * kRowsMax * n;
    int rc;
    switch (s0->key_bytes) {
      case 2: rc = launch_rows<uint16_t>(ctx, (const RowSet*)dv, n, d_rows, R, NF, f, d_bitmap, o); break;
      case 4: rc = launch_rows<uint32_t>(ctx, (const RowSet*)dv, n, d_rows, R, NF, f, d_bitmap, o); break;
      default: rc = launch_rows<unsigned long long>(ctx, (const RowSet*)dv, n, d_rows, R, NF, f, d_bitmap, o); break;
    }
    if (rc != KMSC_OK) return rc;
  }
  KMSC_TRY(comm_allreduce_u64(ctx, d_out, (size_t)n_rows * n));  // a rank of a prefix-sharded job
  // the pinned table is read by the copy above; the result comes back through its tail
  unsigned char* h_out = hv + ((sz_tab + 15) & ~(size_t)15);
  KMSC_CUDA(cudaMemcpyAsync(h_out, d_out, sz_out, cudaMemcpyDeviceToHost, ctx->stream));
  KMSC_CUDA(cudaStreamSynchronize(ctx->stream));
  memcpy(out, h_out, sz_out);
  return KMSC_OK;
}
