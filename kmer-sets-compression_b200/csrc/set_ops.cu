// set_ops.cu -- P4: unions of device CSR sets.
//
// KmerSet::Add(other) (reference lib/core/kmer_set.h:164-174) over two or more sets -- the union
// KmerSetSet::Get / KmerSetSetReader::Get build (kmer_set_set.h:433-454, 672-755) -- and the
// union of two COUNTED sets of the streaming KmerCounter. Both inputs are sorted CSR, so one merge
// per finest-level fine bucket does it: pass 1 counts per fine bucket, a device scan turns the
// counts into the output's finest-level offsets (they ARE its fine index), pass 2 writes the keys.
// (The split n = j & k, j \ n, k \ n and KmerSet::Diff are the single-pass kernel of pair_split.cu.)
#include <algorithm>
#include <vector>

#include "kmsc_common.cuh"
#include "scan.cuh"

namespace kmsc {

// one thread per finest-level fine bucket: two-pointer classification
template <typename KeyT, bool WRITE>
__global__ void merge_classify_kernel(const KeyT* __restrict__ ka, const uint32_t* __restrict__ la,
                                      const KeyT* __restrict__ kb, const uint32_t* __restrict__ lb,
                                      uint32_t NF,
                                      uint32_t* __restrict__ cI, uint32_t* __restrict__ cA,
                                      uint32_t* __restrict__ cB,   // counts (pass 1) / offsets (pass 2)
                                      KeyT* __restrict__ oI, KeyT* __restrict__ oA, KeyT* __restrict__ oB,
                                      KeyT* __restrict__ oU, const uint32_t* __restrict__ offU) {
  const uint32_t x = blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= NF) return;
  uint32_t i = la[x], ie = la[x + 1], j = lb[x], je = lb[x + 1];
  uint32_t nI = 0, nA = 0, nB = 0;
  uint32_t pI = 0, pA = 0, pB = 0, pU = 0;
  if (WRITE) {
    if (oI) pI = cI[x];
    if (oA) pA = cA[x];
    if (oB) pB = cB[x];
    if (oU) pU = offU[x];
  }
  while (i < ie && j < je) {
    const KeyT a = ka[i], b = kb[j];
    if (a < b) {
      if (WRITE) { if (oA) oA[pA++] = a; if (oU) oU[pU++] = a; } else nA++;
      i++;
    } else if (b < a) {
      if (WRITE) { if (oB) oB[pB++] = b; if (oU) oU[pU++] = b; } else nB++;
      j++;
    } else {
      if (WRITE) { if (oI) oI[pI++] = a; if (oU) oU[pU++] = a; } else nI++;
      i++; j++;
    }
  }
  if (WRITE) {
    for (; i < ie; i++) { const KeyT a = ka[i]; if (oA) oA[pA++] = a; if (oU) oU[pU++] = a; }
    for (; j < je; j++) { const KeyT b = kb[j]; if (oB) oB[pB++] = b; if (oU) oU[pU++] = b; }
  } else {
    nA += ie - i;
    nB += je - j;
    cI[x] = nI; cA[x] = nA; cB[x] = nB;
  }
}

__global__ void add3_kernel(const uint32_t* a, const uint32_t* b, const uint32_t* c, uint32_t* out, uint32_t n) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = a[i] + b[i] + c[i];
}

struct MergePlan {
  uint32_t NF;
  int F;
  uint32_t *cI, *cA, *cB, *cU, *bsum, *totals;  // device
};

static int check_pair(const kmsc_set* a, const kmsc_set* b) {
  if (!a || !b) { set_error("NULL set"); return KMSC_E_INVALID; }
  if (a->K != b->K || a->N != b->N || a->key_bytes != b->key_bytes) {
    set_error("sets have different (K,N,KeyType)");
    return KMSC_E_INVALID;
  }
  return KMSC_OK;
}

template <typename KeyT>
static void launch_classify(kmsc_ctx* ctx, bool write, const kmsc_set* a, const kmsc_set* b, const MergePlan& mp,
                            void* oI, void* oA, void* oB, void* oU) {
  const int threads = 128;
  const unsigned blocks = (mp.NF + threads - 1) / threads;
  if (write)
    merge_classify_kernel<KeyT, true><<<blocks, threads, 0, ctx->stream>>>(
        (const KeyT*)a->keys, a->lev[mp.F], (const KeyT*)b->keys, b->lev[mp.F], mp.NF, mp.cI, mp.cA, mp.cB,
        (KeyT*)oI, (KeyT*)oA, (KeyT*)oB, (KeyT*)oU, mp.cU);
  else
    merge_classify_kernel<KeyT, false><<<blocks, threads, 0, ctx->stream>>>(
        (const KeyT*)a->keys, a->lev[mp.F], (const KeyT*)b->keys, b->lev[mp.F], mp.NF, mp.cI, mp.cA, mp.cB,
        nullptr, nullptr, nullptr, nullptr, nullptr);
  count_launch(ctx);
}

static void dispatch_classify(kmsc_ctx* ctx, bool write, const kmsc_set* a, const kmsc_set* b, const MergePlan& mp,
                              void* oI, void* oA, void* oB, void* oU) {
  switch (a->key_bytes) {
    case 2: launch_classify<uint16_t>(ctx, write, a, b, mp, oI, oA, oB, oU); break;
    case 4: launch_classify<uint32_t>(ctx, write, a, b, mp, oI, oA, oB, oU); break;
    default: launch_classify<unsigned long long>(ctx, write, a, b, mp, oI, oA, oB, oU); break;
  }
}

// pass 1 + scans; host_totals = {|A&B|, |A\B|, |B\A|}
static int merge_count(kmsc_ctx* ctx, const kmsc_set* a, const kmsc_set* b, MergePlan* mp, bool want_union,
                       uint32_t host_totals[3]) {
  mp->F = a->max_level;
  mp->NF = (uint32_t)1 << (a->N + mp->F);
  const size_t ent = (size_t)mp->NF + 1;
  const size_t sb = scan_scratch_entries(mp->NF);
  KMSC_TRY(ctx->work2.reserve((ent * 4 + sb * 4 + 16) * 4));
  uint32_t* p = (uint32_t*)ctx->work2.p;
  mp->cI = p; p += ent;
  mp->cA = p; p += ent;
  mp->cB = p; p += ent;
  mp->cU = p; p += ent;
  mp->bsum = p; p += sb * 4;
  mp->totals = p;
  dispatch_classify(ctx, false, a, b, *mp, nullptr, nullptr, nullptr, nullptr);
  KMSC_CUDA(cudaGetLastError());
  ScanMulti sm{};
  uint32_t* arr[4] = {mp->cI, mp->cA, mp->cB, mp->cU};
  int cnt = 3;
  if (want_union) {
    add3_kernel<<<(mp->NF + 255) / 256, 256, 0, ctx->stream>>>(mp->cI, mp->cA, mp->cB, mp->cU, mp->NF);
    count_launch(ctx);
    cnt = 4;
  }
  const size_t sbe = scan_scratch_entries(mp->NF);
  for (int q = 0; q < cnt; q++) {
    sm.in[q] = arr[q]; sm.out[q] = arr[q];
    sm.bsum[q] = mp->bsum + (size_t)q * sbe;
    sm.total[q] = mp->totals + q;
  }
  KMSC_TRY(exclusive_scan_multi_u32(ctx, sm, cnt, mp->NF));
  void* pin = nullptr;
  KMSC_TRY(ctx_pinned(ctx, 64, &pin));
  KMSC_CUDA(cudaMemcpyAsync(pin, mp->totals, 16, cudaMemcpyDeviceToHost, ctx->stream));
  KMSC_CUDA(cudaStreamSynchronize(ctx->stream));
  memcpy(host_totals, pin, 12);
  return KMSC_OK;
}

// all offset levels of up to three new sets from their finest-level offsets (NF + 1 entries
// each): lev[f][x >> (F - f)] = fine[x] wherever x is a multiple of 2^(F - f)
struct LevFill {
  const uint32_t* src[3];
  uint32_t* lev_base[3];
};
__global__ void fill_levels_kernel(LevFill lf, int N, int F, uint32_t NF) {
  const uint32_t x = blockIdx.x * blockDim.x + threadIdx.x;
  if (x > NF) return;
  const uint32_t v = lf.src[blockIdx.y][x];
  uint32_t* base = lf.lev_base[blockIdx.y];
  for (int f = F; f >= 0; f--) {
    const int sh = F - f;
    if (x & ((1u << sh) - 1u)) break;
    const size_t start = ((size_t)1 << N) * (((size_t)1 << f) - 1) + (size_t)f;  // levels 0..f-1 hold 2^(N+g) + 1 entries
    base[start + (x >> sh)] = v;
  }
}

// new sets (same K, N, KeyType as `like`) with n_keys[q] keys whose finest offsets are d_offs[q]
static int sets_from_fine_offsets(kmsc_ctx* ctx, const kmsc_set* like, int count, const int64_t* n_keys,
                                  uint32_t* const* d_offs, uint32_t NF, kmsc_set** out) {
  LevFill lf{};
  for (int q = 0; q < count; q++) out[q] = nullptr;
  for (int q = 0; q < count; q++) {
    int rc = set_alloc(ctx, like->K, like->N, like->key_bytes, n_keys[q], &out[q]);
    if (rc != KMSC_OK) {
      for (int r = 0; r < q; r++) { kmsc_set_free(ctx, out[r]); out[r] = nullptr; }
      return rc;
    }
    out[q]->has_dups = 0;
    lf.src[q] = d_offs[q];
    lf.lev_base[q] = out[q]->lev_base;
  }
  fill_levels_kernel<<<dim3((NF + 1 + 255) / 256, count), 256, 0, ctx->stream>>>(lf, like->N, like->max_level, NF);
  count_launch(ctx);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    for (int q = 0; q < count; q++) { kmsc_set_free(ctx, out[q]); out[q] = nullptr; }
    return cuda_fail(e, "fill levels", __FILE__, __LINE__);
  }
  return KMSC_OK;
}

// ---- union of two COUNTED sets (streaming KmerCounter: chunk results are merged) ------------
// Reference semantics: KmerCounter adds with saturation at 255 (lib/core/kmer_counter.h:28-38,
// 105-126 merges thread-local maps the same way). One thread per finest-level fine bucket.
template <typename KeyT, bool WRITE>
__global__ void merge_counts_kernel(const KeyT* __restrict__ ka, const uint32_t* __restrict__ la, const uint8_t* __restrict__ ca,
                                    const KeyT* __restrict__ kb, const uint32_t* __restrict__ lb, const uint8_t* __restrict__ cb,
                                    uint32_t NF, uint32_t* __restrict__ cU /* counts (pass 1) / offsets (pass 2) */,
                                    KeyT* __restrict__ ok, uint8_t* __restrict__ oc) {
  const uint32_t x = blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= NF) return;
  uint32_t i = la[x], ie = la[x + 1], j = lb[x], je = lb[x + 1];
  uint32_t n = 0, p = WRITE ? cU[x] : 0u;
  while (i < ie && j < je) {
    const KeyT a = ka[i], b = kb[j];
    if (a < b) {
      if (WRITE) { ok[p] = a; oc[p] = ca[i]; p++; } else n++;
      i++;
    } else if (b < a) {
      if (WRITE) { ok[p] = b; oc[p] = cb[j]; p++; } else n++;
      j++;
    } else {
      if (WRITE) { ok[p] = a; oc[p] = (uint8_t)min(255u, (uint32_t)ca[i] + (uint32_t)cb[j]); p++; } else n++;
      i++; j++;
    }
  }
  if (WRITE) {
    for (; i < ie; i++) { ok[p] = ka[i]; oc[p] = ca[i]; p++; }
    for (; j < je; j++) { ok[p] = kb[j]; oc[p] = cb[j]; p++; }
  } else {
    cU[x] = n + (ie - i) + (je - j);
  }
}

template <typename KeyT>
static int counted_union_t(kmsc_ctx* ctx, const kmsc_set* a, const uint8_t* ca, const kmsc_set* b, const uint8_t* cb,
                           kmsc_set** out, uint8_t** out_counts) {
  const int F = a->max_level;
  const uint32_t NF = (uint32_t)1 << (a->N + F);
  const size_t ent = (size_t)NF + 1, sb = scan_scratch_entries(NF);
  KMSC_TRY(ctx->work2.reserve((ent + sb + 16) * 4));
  uint32_t* cU = (uint32_t*)ctx->work2.p;
  uint32_t* bsum = cU + ent;
  uint32_t* total = bsum + sb;
  const unsigned blocks = (NF + 127) / 128;
  merge_counts_kernel<KeyT, false><<<blocks, 128, 0, ctx->stream>>>((const KeyT*)a->keys, a->lev[F], ca, (const KeyT*)b->keys,
                                                                    b->lev[F], cb, NF, cU, nullptr, nullptr);
  count_launch(ctx);
  KMSC_TRY(exclusive_scan_u32(ctx, cU, cU, NF, bsum, total));
  void* pin = nullptr;
  KMSC_TRY(ctx_pinned(ctx, 64, &pin));
  KMSC_CUDA(cudaMemcpyAsync(pin, total, 4, cudaMemcpyDeviceToHost, ctx->stream));
  KMSC_CUDA(cudaStreamSynchronize(ctx->stream));
  const int64_t nu = *(uint32_t*)pin;
  kmsc_set* u = nullptr;
  uint32_t* src[1] = {cU};
  KMSC_TRY(sets_from_fine_offsets(ctx, a, 1, &nu, src, NF, &u));
  uint8_t* d_counts = nullptr;
  cudaError_t e = cudaMalloc(&d_counts, (size_t)nu + 16);
  if (e != cudaSuccess) { kmsc_set_free(ctx, u); return cuda_fail(e, "cudaMalloc counts", __FILE__, __LINE__); }
  merge_counts_kernel<KeyT, true><<<blocks, 128, 0, ctx->stream>>>((const KeyT*)a->keys, a->lev[F], ca, (const KeyT*)b->keys,
                                                                   b->lev[F], cb, NF, cU, (KeyT*)u->keys, d_counts);
  count_launch(ctx);
  e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  if (e != cudaSuccess) { kmsc_set_free(ctx, u); cudaFree(d_counts); return cuda_fail(e, "counted union", __FILE__, __LINE__); }
  *out = u;
  *out_counts = d_counts;
  return KMSC_OK;
}

int counted_union(kmsc_ctx* ctx, const kmsc_set* a, const uint8_t* ca, const kmsc_set* b, const uint8_t* cb,
                  kmsc_set** out, uint8_t** out_counts) {
  KMSC_TRY(check_pair(a, b));
  KMSC_TRY(set_ensure_levels(ctx, a));
  KMSC_TRY(set_ensure_levels(ctx, b));
  switch (a->key_bytes) {
    case 2: return counted_union_t<uint16_t>(ctx, a, ca, b, cb, out, out_counts);
    case 4: return counted_union_t<uint32_t>(ctx, a, ca, b, cb, out, out_counts);
    default: return counted_union_t<unsigned long long>(ctx, a, ca, b, cb, out, out_counts);
  }
}

}  // namespace kmsc

using namespace kmsc;

// ---- union of up to kUnionMax sets in ONE pair of passes ---------------------------------------
// KmerSetSet::Get / KmerSetSetReader::Get unite every node reachable from a set (reference
// lib/core/kmer_set_set.h:433-454, 672-755: a fold of Add); a left fold of two-way unions reads and rewrites
// the growing result m - 1 times. Here a thread owns one finest-level fine bucket and merges the m runs
// (~10 keys each) directly: the smallest head is written once and every run holding it advances.
constexpr int kUnionMax = 16;
struct UnionArgs {
  const void* keys[kUnionMax];
  const uint32_t* lev[kUnionMax];
};

template <typename KeyT, int M, bool WRITE>
__global__ void union_multi_kernel(UnionArgs ua, int m, uint32_t NF, uint32_t* __restrict__ cnt /* counts (pass 1) / offsets (pass 2) */,
                                   KeyT* __restrict__ out) {
  const uint32_t x = blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= NF) return;
  uint32_t i[M], e[M];
  unsigned long long h[M];   // head key, or 2^64 - 1 once the run is exhausted (keys are at most 63 bits wide)
  const unsigned long long INF = ~0ull;
#pragma unroll
  for (int t = 0; t < M; t++) {
    i[t] = 0; e[t] = 0; h[t] = INF;
    if (t < m) {
      i[t] = ua.lev[t][x];
      e[t] = ua.lev[t][x + 1];
      if (i[t] < e[t]) h[t] = (unsigned long long)reinterpret_cast<const KeyT*>(ua.keys[t])[i[t]];
    }
  }
  uint32_t n = 0, p = WRITE ? cnt[x] : 0u;
  for (;;) {
    unsigned long long g = h[0];
#pragma unroll
    for (int t = 1; t < M; t++) g = h[t] < g ? h[t] : g;
    if (g == INF) break;
    if (WRITE) out[p++] = (KeyT)g; else n++;
#pragma unroll
    for (int t = 0; t < M; t++) {
      if (h[t] == g) {
        i[t]++;
        h[t] = i[t] < e[t] ? (unsigned long long)reinterpret_cast<const KeyT*>(ua.keys[t])[i[t]] : INF;
      }
    }
  }
  if (!WRITE) cnt[x] = n;
}

template <typename KeyT, bool WRITE>
static void launch_union_multi(kmsc_ctx* ctx, const UnionArgs& ua, int m, uint32_t NF, uint32_t* cnt, void* out) {
  const unsigned blocks = (NF + 127) / 128;
  if (m <= 4) union_multi_kernel<KeyT, 4, WRITE><<<blocks, 128, 0, ctx->stream>>>(ua, m, NF, cnt, (KeyT*)out);
  else if (m <= 8) union_multi_kernel<KeyT, 8, WRITE><<<blocks, 128, 0, ctx->stream>>>(ua, m, NF, cnt, (KeyT*)out);
  else union_multi_kernel<KeyT, 16, WRITE><<<blocks, 128, 0, ctx->stream>>>(ua, m, NF, cnt, (KeyT*)out);
  count_launch(ctx);
}

// union of sets[0 .. m), 2 <= m <= kUnionMax (duplicate-free inputs of one shape, levels present)
static int union_multi(kmsc_ctx* ctx, const kmsc_set* const* sets, int m, kmsc_set** out) {
  const kmsc_set* a = sets[0];
  const int F = a->max_level;
  const uint32_t NF = (uint32_t)1 << (a->N + F);
  const size_t ent = (size_t)NF + 1;
  const size_t sb = scan_scratch_entries(NF);
  KMSC_TRY(ctx->work2.reserve((ent + sb + 16) * 4));
  uint32_t* d_cnt = (uint32_t*)ctx->work2.p;
  uint32_t* d_bsum = d_cnt + ent;
  uint32_t* d_total = d_bsum + sb;
  UnionArgs ua{};
  for (int t = 0; t < m; t++) { ua.keys[t] = sets[t]->keys; ua.lev[t] = sets[t]->lev[F]; }
  switch (a->key_bytes) {
    case 2: launch_union_multi<uint16_t, false>(ctx, ua, m, NF, d_cnt, nullptr); break;
    case 4: launch_union_multi<uint32_t, false>(ctx, ua, m, NF, d_cnt, nullptr); break;
    default: launch_union_multi<unsigned long long, false>(ctx, ua, m, NF, d_cnt, nullptr); break;
  }
  KMSC_CUDA(cudaGetLastError());
  KMSC_TRY(exclusive_scan_u32(ctx, d_cnt, d_cnt, NF, d_bsum, d_total));
  void* pin = nullptr;
  KMSC_TRY(ctx_pinned(ctx, 64, &pin));
  KMSC_CUDA(cudaMemcpyAsync(pin, d_total, 4, cudaMemcpyDeviceToHost, ctx->stream));
  KMSC_CUDA(cudaStreamSynchronize(ctx->stream));
  const int64_t nu = *(uint32_t*)pin;
  kmsc_set* u = nullptr;
  uint32_t* src[1] = {d_cnt};
  KMSC_TRY(sets_from_fine_offsets(ctx, a, 1, &nu, src, NF, &u));
  switch (a->key_bytes) {
    case 2: launch_union_multi<uint16_t, true>(ctx, ua, m, NF, d_cnt, u->keys); break;
    case 4: launch_union_multi<uint32_t, true>(ctx, ua, m, NF, d_cnt, u->keys); break;
    default: launch_union_multi<unsigned long long, true>(ctx, ua, m, NF, d_cnt, u->keys); break;
  }
  cudaError_t e = cudaGetLastError();
  // the offsets in work2 are read by the write pass and the level fill: both are queued; callers that reuse
  // work2 do so on the same stream
  if (e != cudaSuccess) { kmsc_set_free(ctx, u); return cuda_fail(e, "union write", __FILE__, __LINE__); }
  *out = u;
  return KMSC_OK;
}

extern "C" {

int kmsc_set_union(kmsc_ctx* ctx, const kmsc_set* const* sets, int32_t m, kmsc_set** out) {
  if (!ctx || !sets || m < 1 || !out) { set_error("bad argument"); return KMSC_E_INVALID; }
  KMSC_CUDA(cudaSetDevice(ctx->device));
  for (int32_t t = 0; t < m; t++) {
    if (!sets[t]) { set_error("sets[%d] is NULL", t); return KMSC_E_INVALID; }
    KMSC_TRY(set_check_dups(ctx, const_cast<kmsc_set*>(sets[t])));
    if (sets[t]->has_dups == 1) { set_error("sets[%d] holds duplicate keys (union needs true sets)", t); return KMSC_E_INVALID; }
  }
  if (m >= 3) {
    // groups of up to kUnionMax sets per pass pair; the partial unions are united the same way
    for (int32_t t = 0; t < m; t++) {
      int rc = check_pair(sets[0], sets[t]);
      if (rc == KMSC_OK) rc = set_ensure_levels(ctx, sets[t]);
      if (rc != KMSC_OK) return rc;
    }
    std::vector<const kmsc_set*> cur(sets, sets + m);
    std::vector<kmsc_set*> owned;   // intermediates of the previous round
    while (cur.size() > 1) {
      std::vector<const kmsc_set*> next;
      std::vector<kmsc_set*> made;
      for (size_t a = 0; a < cur.size(); a += kUnionMax) {
        const int g = (int)std::min<size_t>(kUnionMax, cur.size() - a);
        if (g == 1) { next.push_back(cur[a]); continue; }
        kmsc_set* u = nullptr;
        const int rc = union_multi(ctx, cur.data() + a, g, &u);
        if (rc != KMSC_OK) {
          for (kmsc_set* x : made) kmsc_set_free(ctx, x);
          for (kmsc_set* x : owned) kmsc_set_free(ctx, x);
          return rc;
        }
        made.push_back(u);
        next.push_back(u);
      }
      // an intermediate that was carried over unchanged (a group of one) stays alive for the next round
      std::vector<kmsc_set*> keep;
      for (kmsc_set* x : owned) {
        bool carried = false;
        for (const kmsc_set* y : next) carried |= y == x;
        if (carried) keep.push_back(x); else kmsc_set_free(ctx, x);
      }
      owned = keep;
      owned.insert(owned.end(), made.begin(), made.end());
      cur = next;
    }
    KMSC_CUDA(cudaStreamSynchronize(ctx->stream));
    *out = const_cast<kmsc_set*>(cur[0]);
    return KMSC_OK;
  }
  // left fold of two-way unions; intermediates are freed as we go
  kmsc_set* acc = nullptr;
  for (int32_t t = 0; t < m; t++) {
    const kmsc_set* b = sets[t];
    if (!b) { set_error("sets[%d] is NULL", t); kmsc_set_free(ctx, acc); return KMSC_E_INVALID; }
    const kmsc_set* a = acc ? acc : sets[0];
    if (t == 0 && m > 1) continue;
    int rc = check_pair(a, b);
    if (rc == KMSC_OK) rc = set_ensure_levels(ctx, a);
    if (rc == KMSC_OK) rc = set_ensure_levels(ctx, b);
    MergePlan mp;
    uint32_t tot[3];
    if (rc == KMSC_OK) rc = merge_count(ctx, a, b, &mp, true, tot);
    kmsc_set* u = nullptr;
    if (rc == KMSC_OK) {
      const int64_t nu = (m == 1) ? a->n_keys : (int64_t)tot[0] + tot[1] + tot[2];
      uint32_t* src[1] = {mp.cU};
      rc = sets_from_fine_offsets(ctx, a, 1, &nu, src, mp.NF, &u);
    }
    if (rc == KMSC_OK) {
      dispatch_classify(ctx, true, a, b, mp, nullptr, nullptr, nullptr, u->keys);
      cudaError_t e = cudaGetLastError();
      if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
      if (e != cudaSuccess) rc = cuda_fail(e, "set_union write", __FILE__, __LINE__);
    }
    kmsc_set_free(ctx, acc);
    acc = nullptr;
    if (rc != KMSC_OK) { kmsc_set_free(ctx, u); return rc; }
    acc = u;
  }
  *out = acc;
  return KMSC_OK;
}

}  // extern "C"
