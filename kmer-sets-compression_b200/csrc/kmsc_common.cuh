// kmsc_common.cuh -- shared declarations for libkmsc (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "kmsc.h"

namespace kmsc {

constexpr int kMaxFineLevel = 6;  // fine offsets exist for levels 0..kMaxFineLevel

void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what, const char* file, int line);

#define KMSC_CUDA(call)                                                        \
  do {                                                                         \
    cudaError_t e__ = (call);                                                  \
    if (e__ != cudaSuccess) return ::kmsc::cuda_fail(e__, #call, __FILE__, __LINE__); \
  } while (0)

#define KMSC_TRY(call)          \
  do {                          \
    int rc__ = (call);          \
    if (rc__ != KMSC_OK) return rc__; \
  } while (0)

// grow-only device scratch buffer owned by a context
struct Scratch {
  void* p = nullptr;
  size_t cap = 0;
  int reserve(size_t bytes);
  void release();
};

// pinned host table that is copied to the device asynchronously: acquire() waits until the
// previous copy out of it has run, commit() marks the new one
struct PinnedTab {
  void* p = nullptr;
  size_t cap = 0;
  cudaEvent_t ev = nullptr;
  bool pending = false;
  int acquire(size_t bytes, void** out);
  int commit(cudaStream_t stream);
  void release();
};

}  // namespace kmsc

struct kmsc_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  int sm_count = 148;
  int64_t launches = 0;
  kmsc::Scratch plan;     // pair_counts planning buffers
  kmsc::Scratch work;     // generic per-call scratch
  kmsc::Scratch work2;
  kmsc::Scratch work3;
  kmsc::Scratch stage;    // input staging (text, packed bases)
  kmsc::Scratch small;    // small per-call device arrays (descriptors, ids)
  kmsc::Scratch spss_out; // result of the last kmsc_spss_build: text, then string offsets
  kmsc::Scratch pre_buf[2];   // kmsc_counter_prefetch: device copies of announced chunks (grow-only: no malloc / free per counter)
  cudaEvent_t pre_ev[2] = {nullptr, nullptr};
  int64_t spss_strings = 0, spss_chars = 0;
  // staged partition sort: up to kP2Slots groups of jobs in flight, each with its own tables
  static constexpr int kP2Slots = 4;
  kmsc::Scratch p2a[kP2Slots];      // job table, bin bases, cursors
  kmsc::Scratch p2b[kP2Slots];      // first-level output
  kmsc::PinnedTab p2tab[kP2Slots];  // host side of the job table
  kmsc::PinnedTab p2rb[kP2Slots];   // read-backs (n_occ, largest partition, repeat flags)
  kmsc::Scratch tabs_dev[2];        // device copies of small host tables (0 level jobs, 1 tail jobs)
  kmsc::PinnedTab tab[2];           // their host side
  cudaStream_t copy_stream = nullptr;  // host-to-device copies of a batch decode
  cudaEvent_t copy_ev[kP2Slots] = {};  // "group g is on the device"
  cudaEvent_t fence_ev = nullptr;
  // multi-GPU: NCCL communicator of this rank (comm.cu); NULL = single GPU
  void* comm = nullptr;
  int comm_rank = 0, comm_ranks = 1;
  void* pinned = nullptr; // pinned host staging
  size_t pinned_cap = 0;
  // pair_counts: redundancy (keys per distinct key in a tile) seen by the last call
  double pc_rho = 0.0;
  int pc_rho_n = 0, pc_rho_k = 0;
  unsigned long long pc_last_stats[3] = {0, 0, 0};
  unsigned long long pc_last_L = 0;
  int pc_last_build = 0;  // 0 hash build, 1 merge build
  cudaEvent_t pc_ev[3] = {nullptr, nullptr, nullptr};  // plan start, main start, main end
  double pc_main_ms = 0, pc_plan_ms = 0, pc_algo_bytes = 0;
  int pc_main_launches = 0;
  // last counting result (kmsc_count_get)
  kmsc_set* last_counted = nullptr;
  bool last_counted_owned = false;
  uint8_t* last_counts = nullptr;  // device, aligned with last_counted keys
};

// Device-resident bucketed set (CSR). keys are ascending inside each bucket.
// lev[f] (f = 0..max_level) are fine offsets: lev[f][x] for x in [0, 2^(N+f)]
// is the index of the first key whose (bucket << f | top f bits of key) >= x;
// lev[0] is the bucket-level CSR offs array.
struct kmsc_set {
  int K = 0, N = 0, key_bytes = 0;
  int key_bits = 0;       // 2K - N
  int max_level = 0;      // min(kMaxFineLevel, key_bits)
  int64_t n_keys = 0;
  void* keys = nullptr;   // device, n_keys * key_bytes (+ padding)
  uint32_t* lev_base = nullptr;  // device, all levels in one allocation
  uint32_t* lev[kmsc::kMaxFineLevel + 1] = {};
  int has_dups = -1;      // -1 unknown, 0 no, 1 yes
  bool fine_ready = true; // false: only lev[0] is filled (lean split output); set_ensure_levels builds the rest
  // buckets outside [b_lo, b_hi) are known to be empty (a rank's prefix shard); -1 = all buckets
  int32_t b_lo = -1, b_hi = -1;
};

namespace kmsc {

int ctx_pinned(kmsc_ctx* ctx, size_t bytes, void** out);
inline void count_launch(kmsc_ctx* ctx, int n = 1) { ctx->launches += n; }
// sums a device array over the ranks of the context's communicator, in place (no-op without one)
int comm_allreduce_u64(kmsc_ctx* ctx, unsigned long long* d_buf, size_t count);

// set construction helpers (set_build.cu)
int set_alloc(kmsc_ctx* ctx, int K, int N, int key_bytes, int64_t n_keys, kmsc_set** out);
// fills lev[1..max_level] from keys + lev[0]; also sets has_dups
int set_build_levels(kmsc_ctx* ctx, kmsc_set* s);
// given lev[max_level] (finest) already filled, derive coarser levels by striding
int set_derive_levels(kmsc_ctx* ctx, kmsc_set* s);
int set_check_dups(kmsc_ctx* ctx, kmsc_set* s);
// builds lev[1..max_level] of a set that only carries lev[0] (no-op otherwise); call before reading a finer level
int set_ensure_levels(kmsc_ctx* ctx, const kmsc_set* s);
// union of two counted sets with saturating uint8 counts (set_ops.cu); counts are device arrays
// aligned with the keys; *out_counts is cudaMalloc'd
int counted_union(kmsc_ctx* ctx, const kmsc_set* a, const uint8_t* ca, const kmsc_set* b, const uint8_t* cb,
                  kmsc_set** out, uint8_t** out_counts);

}  // namespace kmsc
