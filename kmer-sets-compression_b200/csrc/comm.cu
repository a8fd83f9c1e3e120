// comm.cu -- multi-GPU inside the library: one context per rank (process or thread), an NCCL
// communicator owned by the context.
//
// The reference is single-process shared-memory code; |S_i & S_j| = sum over buckets
// (lib/core/kmer_set_set.h:161-181) is what shards: the k-mer prefix space is cut into contiguous
// bucket ranges, every rank computes the partial matrix of its range and the partials are summed by
// ONE all-reduce (here, inside kmsc_pair_counts*). Getting the sets into their shards is the other
// exchange: every rank decodes WHOLE sets (its share of the input files) and kmsc_sets_exchange hands
// every rank the slice of every set inside its range -- a bucket range of a sorted set is one
// contiguous key slice, so the slices travel zero-copy (ncclSend / ncclRecv straight between the
// sets' key arrays over NVLink), one grouped call for all of them.
//
// NCCL is loaded at run time (dlopen "libnccl.so.2": the copy already in the process when the host
// program brought one, else the system's), so libkmsc has no link-time dependency on it and a
// single-GPU user never needs it.
#include <dlfcn.h>
#include <nccl.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "kmer_pipeline.cuh"

namespace kmsc {
namespace {

struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

NcclApi* nccl() {
  static NcclApi api;
  static bool tried = false;
  if (tried) return api.handle ? &api : nullptr;
  tried = true;
  // 1. KMSC_NCCL_LIB names the library (a host program that will load its own NCCL later, e.g. a
  //    Python process that imports torch after this point, must point here at that copy: two NCCL
  //    builds under one SONAME cannot live in one process); 2. a copy already in the process;
  //    3. the system's.
  void* h = nullptr;
  if (const char* e = getenv("KMSC_NCCL_LIB")) h = dlopen(e, RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL | RTLD_NOLOAD);
  if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!h) return nullptr;
#define KMSC_SYM(field, name) *(void**)(&api.field) = dlsym(h, name); if (!api.field) return nullptr;
  KMSC_SYM(GetUniqueId, "ncclGetUniqueId")
  KMSC_SYM(CommInitRank, "ncclCommInitRank")
  KMSC_SYM(CommDestroy, "ncclCommDestroy")
  KMSC_SYM(AllReduce, "ncclAllReduce")
  KMSC_SYM(AllGather, "ncclAllGather")
  KMSC_SYM(Send, "ncclSend")
  KMSC_SYM(Recv, "ncclRecv")
  KMSC_SYM(GroupStart, "ncclGroupStart")
  KMSC_SYM(GroupEnd, "ncclGroupEnd")
  KMSC_SYM(GetErrorString, "ncclGetErrorString")
#undef KMSC_SYM
  api.handle = h;
  return &api;
}

int nccl_fail(NcclApi* a, ncclResult_t r, const char* what) {
  set_error("NCCL error %d (%s) in %s", (int)r, a->GetErrorString(r), what);
  return KMSC_E_CUDA;
}
#define KMSC_NCCL(api, call)                                          \
  do {                                                                \
    ncclResult_t r__ = (call);                                        \
    if (r__ != ncclSuccess) return nccl_fail(api, r__, #call);        \
  } while (0)

// offs[q] of every (set, cut): t[j * (R + 2) + q] = lev0_j[cuts[q]], and t[j * (R + 2) + R + 1] = the
// set's duplicate flag (filled by the host side of the table)
struct CutJob { const uint32_t* lev0; uint32_t flag; };
__global__ void cut_offsets_kernel(const CutJob* __restrict__ jobs, int n_sets, const int32_t* __restrict__ cuts, int R,
                                   uint32_t* __restrict__ table) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_sets * (R + 2)) return;
  const int j = t / (R + 2), q = t % (R + 2);
  table[t] = q <= R ? jobs[j].lev0[cuts[q]] : jobs[j].flag;
}

// received raw FINEST-level offset slices -> finest level of the imported sets: rebased, empty outside
// [lo, hi) (in fine-bucket units); the coarser levels follow by striding (derive_levels_batch)
struct ImportJob { const uint32_t* in; uint32_t* fine; uint32_t n_keys; };
__global__ void import_offsets_batch_kernel(const ImportJob* __restrict__ jobs, uint32_t lo, uint32_t hi, uint32_t n_fine) {
  const ImportJob jb = jobs[blockIdx.y];
  const uint32_t x = blockIdx.x * blockDim.x + threadIdx.x;
  if (x > n_fine) return;
  jb.fine[x] = x <= lo ? 0u : x >= hi ? jb.n_keys : jb.in[x - lo] - jb.in[0];
}

}  // namespace

// used by pair_counts.cu: sums the per-rank partial matrices in place when the context has a communicator
int comm_allreduce_u64(kmsc_ctx* ctx, unsigned long long* d_buf, size_t count) {
  if (!ctx->comm || ctx->comm_ranks <= 1) return KMSC_OK;
  NcclApi* a = nccl();
  if (!a) { set_error("NCCL is not available"); return KMSC_E_STATE; }
  KMSC_NCCL(a, a->AllReduce(d_buf, d_buf, count, ncclUint64, ncclSum, (ncclComm_t)ctx->comm, ctx->stream));
  return KMSC_OK;
}

}  // namespace kmsc

using namespace kmsc;

extern "C" {

int kmsc_comm_unique_id(void* id128) {
  if (!id128) { set_error("NULL argument"); return KMSC_E_INVALID; }
  NcclApi* a = nccl();
  if (!a) { set_error("NCCL is not available (libnccl.so.2 could not be loaded)"); return KMSC_E_STATE; }
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  ncclUniqueId id;
  KMSC_NCCL(a, a->GetUniqueId(&id));
  memcpy(id128, &id, 128);
  return KMSC_OK;
}

int kmsc_comm_init(kmsc_ctx* ctx, int rank, int n_ranks, const void* id128) {
  if (!ctx || !id128 || n_ranks < 1 || rank < 0 || rank >= n_ranks) { set_error("bad argument"); return KMSC_E_INVALID; }
  if (ctx->comm) { set_error("the context already has a communicator"); return KMSC_E_STATE; }
  NcclApi* a = nccl();
  if (!a) { set_error("NCCL is not available (libnccl.so.2 could not be loaded)"); return KMSC_E_STATE; }
  KMSC_CUDA(cudaSetDevice(ctx->device));
  ncclUniqueId id;
  memcpy(&id, id128, 128);
  ncclComm_t c = nullptr;
  KMSC_NCCL(a, a->CommInitRank(&c, n_ranks, id, rank));
  ctx->comm = c;
  ctx->comm_rank = rank;
  ctx->comm_ranks = n_ranks;
  return KMSC_OK;
}

int kmsc_comm_destroy(kmsc_ctx* ctx) {
  if (!ctx) { set_error("NULL argument"); return KMSC_E_INVALID; }
  if (!ctx->comm) return KMSC_OK;
  NcclApi* a = nccl();
  if (a) {
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    a->CommDestroy((ncclComm_t)ctx->comm);
  }
  ctx->comm = nullptr; ctx->comm_rank = 0; ctx->comm_ranks = 1;
  return KMSC_OK;
}

int kmsc_comm_info(kmsc_ctx* ctx, int* rank, int* n_ranks) {
  if (!ctx) { set_error("NULL argument"); return KMSC_E_INVALID; }
  if (rank) *rank = ctx->comm ? ctx->comm_rank : 0;
  if (n_ranks) *n_ranks = ctx->comm ? ctx->comm_ranks : 1;
  return KMSC_OK;
}

int kmsc_sets_exchange(kmsc_ctx* ctx, const kmsc_set* const* mine, int32_t n_mine, const int32_t* cuts, kmsc_set** out,
                       int32_t n_total) {
  if (!ctx || !cuts || !out || n_mine < 0 || (n_mine > 0 && !mine)) { set_error("bad argument"); return KMSC_E_INVALID; }
  if (!ctx->comm) { set_error("kmsc_sets_exchange needs a communicator (kmsc_comm_init)"); return KMSC_E_STATE; }
  NcclApi* a = nccl();
  const int R = ctx->comm_ranks, me = ctx->comm_rank;
  if (n_total != n_mine * R) { set_error("every rank passes the same number of sets (n_total = n_mine * ranks)"); return KMSC_E_INVALID; }
  for (int32_t j = 0; j < n_total; j++) out[j] = nullptr;
  if (n_mine == 0) return KMSC_OK;
  const kmsc_set* s0 = mine[0];
  const int nb = 1 << s0->N;
  for (int q = 0; q <= R; q++)
    if (cuts[q] < 0 || cuts[q] > nb || (q > 0 && cuts[q] < cuts[q - 1])) { set_error("cuts must ascend inside [0, 2^N]"); return KMSC_E_INVALID; }
  if (cuts[0] != 0 || cuts[R] != nb) { set_error("cuts must cover [0, 2^N]"); return KMSC_E_INVALID; }
  for (int32_t j = 0; j < n_mine; j++) {
    if (!mine[j] || mine[j]->K != s0->K || mine[j]->N != s0->N || mine[j]->key_bytes != s0->key_bytes) {
      set_error("sets have different (K,N,KeyType)");
      return KMSC_E_INVALID;
    }
  }
  KMSC_CUDA(cudaSetDevice(ctx->device));
  const int kb = s0->key_bytes;
  const int W = R + 2;  // table row: R + 1 cut offsets, duplicate flag

  // 1. where every one of my sets is cut, gathered from all ranks: all[r][j][W]
  const size_t row_bytes = (size_t)n_mine * W * 4;
  const size_t o_jobs = 0, o_cuts = ((size_t)n_mine * sizeof(CutJob) + 255) & ~(size_t)255;
  const size_t o_mine = o_cuts + (((size_t)(R + 1) * 4 + 255) & ~(size_t)255);
  const size_t o_all = o_mine + ((row_bytes + 255) & ~(size_t)255);
  const size_t tab_end = o_all + (((size_t)R * row_bytes + 255) & ~(size_t)255);
  KMSC_TRY(ctx->work.reserve(tab_end));
  unsigned char* dv = (unsigned char*)ctx->work.p;
  void* pin = nullptr;
  KMSC_TRY(ctx_pinned(ctx, tab_end + (size_t)R * row_bytes, &pin));
  unsigned char* hv = (unsigned char*)pin;
  CutJob* hj = (CutJob*)(hv + o_jobs);
  for (int32_t j = 0; j < n_mine; j++) hj[j] = CutJob{mine[j]->lev[0], (uint32_t)(mine[j]->has_dups == 0 ? 0 : 1)};
  memcpy(hv + o_cuts, cuts, (size_t)(R + 1) * 4);
  KMSC_CUDA(cudaMemcpyAsync(dv, hv, o_mine, cudaMemcpyHostToDevice, ctx->stream));
  cut_offsets_kernel<<<(n_mine * W + 127) / 128, 128, 0, ctx->stream>>>((const CutJob*)(dv + o_jobs), n_mine, (const int32_t*)(dv + o_cuts),
                                                                     R, (uint32_t*)(dv + o_mine));
  count_launch(ctx);
  KMSC_CUDA(cudaGetLastError());
  KMSC_NCCL(a, a->AllGather(dv + o_mine, dv + o_all, row_bytes, ncclUint8, (ncclComm_t)ctx->comm, ctx->stream));
  uint32_t* h_all = (uint32_t*)(hv + tab_end);
  KMSC_CUDA(cudaMemcpyAsync(h_all, dv + o_all, (size_t)R * row_bytes, cudaMemcpyDeviceToHost, ctx->stream));
  KMSC_CUDA(cudaStreamSynchronize(ctx->stream));  // the one synchronisation: sizes of what arrives

  // 2. the sets this rank will hold: global set r + j * R (rank r's j-th set), restricted to [lo, hi)
  const int lo = cuts[me], hi = cuts[me + 1];
  const int F = s0->max_level;
  for (int32_t j = 0; j < n_mine; j++) KMSC_TRY(set_ensure_levels(ctx, mine[j]));
  // the offsets travel at the FINEST level (64 entries per bucket, +5 % on the keys): the receiver
  // rebases them and derives the coarser levels by striding instead of binary-searching every level
  const size_t n_ent = ((size_t)(hi - lo) << F) + 1;  // offset entries per set slice
  auto cleanup = [&]() { for (int32_t g = 0; g < n_total; g++) if (out[g]) { kmsc_set_free(ctx, out[g]); out[g] = nullptr; } };
  for (int r = 0; r < R; r++)
    for (int32_t j = 0; j < n_mine; j++) {
      const uint32_t* t = h_all + ((size_t)r * n_mine + j) * W;
      const int64_t nk = (int64_t)t[me + 1] - (int64_t)t[me];
      const int rc = set_alloc(ctx, s0->K, s0->N, kb, nk, &out[r + j * R]);
      if (rc != KMSC_OK) { cleanup(); return rc; }
      out[r + j * R]->has_dups = t[R + 1] ? -1 : 0;   // a slice of a duplicate-free set is duplicate-free
      out[r + j * R]->b_lo = lo; out[r + j * R]->b_hi = hi;
    }
  // raw offset slices land in scratch, then one kernel rebases them into every lev[0]
  const size_t o_in = 0, in_bytes = (((size_t)n_total * n_ent * 4) + 255) & ~(size_t)255;
  const size_t o_imp = in_bytes;
  {
    const int rc = ctx->work2.reserve(o_imp + (size_t)n_total * sizeof(ImportJob));
    if (rc != KMSC_OK) { cleanup(); return rc; }
  }
  unsigned char* d2 = (unsigned char*)ctx->work2.p;
  uint32_t* d_in = (uint32_t*)(d2 + o_in);

  // 3. one grouped exchange: keys zero-copy between the sets' arrays, offsets into the scratch
  const bool dbg = getenv("KMSC_DEBUG_COMM") != nullptr;
  cudaEvent_t ev[3] = {nullptr, nullptr, nullptr};
  if (dbg) { for (auto& x : ev) cudaEventCreate(&x); cudaEventRecord(ev[0], ctx->stream); }
  const uint32_t* my_tab = h_all + (size_t)me * n_mine * W;
  ncclResult_t gr = a->GroupStart();
  if (gr != ncclSuccess) { cleanup(); return nccl_fail(a, gr, "ncclGroupStart"); }
  for (int q = 0; q < R && gr == ncclSuccess; q++)
    for (int32_t j = 0; j < n_mine && gr == ncclSuccess; j++) {
      const uint32_t* t = my_tab + (size_t)j * W;
      const size_t nk = (size_t)t[q + 1] - t[q];
      if (nk) gr = a->Send((const unsigned char*)mine[j]->keys + (size_t)t[q] * kb, nk * kb, ncclUint8, q, (ncclComm_t)ctx->comm, ctx->stream);
      if (gr == ncclSuccess)
        gr = a->Send(mine[j]->lev[F] + ((size_t)cuts[q] << F), (((size_t)(cuts[q + 1] - cuts[q]) << F) + 1) * 4, ncclUint8, q,
                     (ncclComm_t)ctx->comm, ctx->stream);
    }
  for (int r = 0; r < R && gr == ncclSuccess; r++)
    for (int32_t j = 0; j < n_mine && gr == ncclSuccess; j++) {
      kmsc_set* s = out[r + j * R];
      if (s->n_keys) gr = a->Recv(s->keys, (size_t)s->n_keys * kb, ncclUint8, r, (ncclComm_t)ctx->comm, ctx->stream);
      if (gr == ncclSuccess)
        gr = a->Recv(d_in + (size_t)(r + j * R) * n_ent, (size_t)n_ent * 4, ncclUint8, r, (ncclComm_t)ctx->comm, ctx->stream);
    }
  {
    const ncclResult_t ge = a->GroupEnd();
    if (gr == ncclSuccess) gr = ge;
  }
  if (gr != ncclSuccess) { cleanup(); return nccl_fail(a, gr, "grouped ncclSend / ncclRecv"); }
  if (dbg) cudaEventRecord(ev[1], ctx->stream);

  // 4. offsets of every imported set in one launch, then the finer levels
  std::vector<ImportJob> imp((size_t)n_total);
  for (int32_t g = 0; g < n_total; g++) imp[(size_t)g] = ImportJob{d_in + (size_t)g * n_ent, out[g]->lev[F], (uint32_t)out[g]->n_keys};
  // (a pageable source is copied to the driver's staging before the call returns)
  cudaError_t e = cudaMemcpyAsync(d2 + o_imp, imp.data(), (size_t)n_total * sizeof(ImportJob), cudaMemcpyHostToDevice, ctx->stream);
  if (e != cudaSuccess) { cleanup(); return cuda_fail(e, "exchange", __FILE__, __LINE__); }
  const uint32_t n_fine = (uint32_t)nb << F;
  import_offsets_batch_kernel<<<dim3((n_fine + 1 + 255) / 256, (unsigned)n_total), 256, 0, ctx->stream>>>(
      (const ImportJob*)(d2 + o_imp), (uint32_t)lo << F, (uint32_t)hi << F, n_fine);
  count_launch(ctx);
  e = cudaGetLastError();
  if (e != cudaSuccess) { cleanup(); return cuda_fail(e, "exchange import", __FILE__, __LINE__); }
  {
    const int rc = derive_levels_batch(ctx, out, n_total);
    if (rc != KMSC_OK) { cleanup(); return rc; }
  }
  if (dbg) {
    cudaEventRecord(ev[2], ctx->stream);
    cudaEventSynchronize(ev[2]);
    float t01 = 0, t12 = 0;
    cudaEventElapsedTime(&t01, ev[0], ev[1]);
    cudaEventElapsedTime(&t12, ev[1], ev[2]);
    size_t sent = 0;
    for (int q = 0; q < R; q++) for (int32_t j = 0; j < n_mine; j++) sent += ((size_t)my_tab[(size_t)j * W + q + 1] - my_tab[(size_t)j * W + q]) * kb;
    fprintf(stderr, "[kmsc comm rank %d] send/recv %.3f ms (%.1f MB sent, %.1f GB/s), import + levels %.3f ms\n", me, t01, sent / 1e6,
            sent / 1e6 / t01, t12);
    for (auto& x : ev) cudaEventDestroy(x);
  }
  return KMSC_OK;
}

}  // extern "C"
