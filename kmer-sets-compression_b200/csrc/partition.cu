// partition.cu -- host side of the staged partition sort (kernels: partition.cuh).
#include <cstdlib>
#include <cstring>

#include "partition.cuh"

namespace kmsc {

namespace {

int ceil_log2_u64(uint64_t x) {
  int l = 0;
  while (((uint64_t)1 << l) < x && l < 63) l++;
  return l;
}

size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

size_t sort_smem_bytes(int FB2, uint32_t cap, int key_bytes) {
  // cur | wsum | n_long | long_list | B | SB | dropw, dropp  (the offsets of sort_kernel)
  size_t off = (((((size_t)1 << FB2) + 1 + 33 + 1 + 2 * part::kLongCap) * 4) + 15) & ~(size_t)15;
  off = (off + ((size_t)cap + 4) * key_bytes + 15) & ~(size_t)15;
  off = (off + (size_t)cap * 2 + 15) & ~(size_t)15;
  return off + 2 * (((size_t)cap >> 5) + 2) * 4 + 16;
}

size_t partition_smem_bytes(int B1) {
  const size_t bins = (size_t)1 << B1;
  return (size_t)part::kTile * 8 + (part::kTile / 32 + 2) * 8 + (bins + 1) * 4 + 2 * bins * 4 + (part::kTile / 32) * 4 + 33 * 4 + 16;
}

size_t count_smem_bytes(int B1) {
  return (part::kTile / 32 + 2) * 8 + (part::kTile / 32) * 4 + ((size_t)4 << B1) + 16;
}

template <typename K>
int set_smem_limit(K kernel, size_t bytes) {
  if (bytes > 48 * 1024) KMSC_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  return KMSC_OK;
}

part::Geo make_geo(const PartPlan& p) {
  part::Geo g;
  g.K = p.opt.K; g.V = 2 * p.opt.K; g.key_bits = 2 * p.opt.K - p.opt.N; g.canonical = p.opt.canonical;
  g.B1 = p.B1; g.R1 = p.R1; g.FB = p.FB; g.FB2 = p.FB2; g.dedup = p.dedup_in_sort ? 1 : 0;
  g.bucket_lo = (unsigned long long)(p.opt.bucket_lo < 0 ? 0 : p.opt.bucket_lo);
  g.bucket_hi = (unsigned long long)p.opt.bucket_hi;
  return g;
}

}  // namespace

int partition_plan(kmsc_ctx* ctx, const PipelineInput* in, int m, const PipelineOptions& opt, PartPlan* plan, int slot) {
  plan->feasible = false;
  plan->slot = slot;
  plan->m = m;
  plan->opt = opt;
  plan->key_bytes = opt.key_bytes;
  plan->n_occ.assign((size_t)m, 0);
  if (m <= 0 || m > 65535) return KMSC_OK;
  if (getenv("KMSC_P2_LEGACY")) return KMSC_OK;
  const int V = 2 * opt.K, key_bits = V - opt.N;
  const int F = key_bits < kMaxFineLevel ? key_bits : kMaxFineLevel;
  int64_t max_pos = 0;
  for (int j = 0; j < m; j++) max_pos = in[j].n_pos > max_pos ? in[j].n_pos : max_pos;
  plan->max_pos = max_pos;
  if (max_pos <= 0) return KMSC_OK;
  if ((uint64_t)(max_pos + part::kTile - 1) / part::kTile > 0x7fffffffull) return KMSC_OK;
  // first level: partitions of about 4-8 K occurrences; the fine buckets of one partition (and
  // the sub-bins of its counting sort) must fit the shared-memory tables
  int b1 = ceil_log2_u64((uint64_t)(max_pos + 8191) / 8192);
  if (const char* e = getenv("KMSC_P2_B1")) b1 = atoi(e);
  const int lo = std::max(V == 64 ? 1 : 0, opt.N + F - part::kMaxSub);
  const int hi = std::min(part::kMaxB1, opt.N + F);
  if (lo > hi) return KMSC_OK;
  b1 = std::min(std::max(b1, lo), hi);
  plan->B1 = b1;
  plan->R1 = V - b1;
  plan->tmp_bytes = plan->R1 <= 32 ? 4 : 8;
  plan->FB = opt.N + F - b1;
  const int sub_hi = std::min(part::kMaxSub, plan->R1);
  // sub-bins of the second-level counting sort: about one per key, at least the fine buckets
  int fb2 = ceil_log2_u64((uint64_t)(max_pos >> b1) + 1) - 1;
  if (const char* e = getenv("KMSC_P2_SUB")) fb2 = atoi(e);
  plan->FB2 = std::min(std::max(fb2, plan->FB), sub_hi);

  // CTAs per job of the count / partition kernels (both walk the same tiles): the device filled
  // about twice over all jobs, and few enough that the scan's per-bin walk over them stays short
  const uint64_t max_tiles = (uint64_t)(max_pos + part::kTile - 1) / part::kTile;
  unsigned gx = (unsigned)std::max<int64_t>(1, (int64_t)ctx->sm_count * 4 / m);
  if (const char* e = getenv("KMSC_P2_CTAS")) gx = (unsigned)std::max(1, atoi(e));
  if (gx > 256) gx = 256;
  if (gx > max_tiles) gx = (unsigned)max_tiles;
  plan->n_ctas = (int)gx;

  // device job table: Job[m] | per job base[bins + 1], meta[4], slice[gx][bins]
  const size_t bins = (size_t)1 << b1;
  const size_t per_job_words = align_up(bins + 1 + (size_t)gx * bins, 4);
  const size_t jobs_bytes = align_up((size_t)m * sizeof(part::Job), 256);
  const size_t meta_bytes = align_up((size_t)m * 16 + (size_t)m * bins * 4, 256);  // meta[m][4] | removed[m][bins]
  KMSC_TRY(ctx->p2a[slot].reserve(jobs_bytes + meta_bytes + (size_t)m * per_job_words * 4));
  unsigned char* dbase = (unsigned char*)ctx->p2a[slot].p;
  uint32_t* dmeta = (uint32_t*)(dbase + jobs_bytes);   // [m][4], contiguous: one memset, one copy back
  uint32_t* dwords = (uint32_t*)(dbase + jobs_bytes + meta_bytes);
  void* tab = nullptr;
  KMSC_TRY(ctx->p2tab[slot].acquire((size_t)m * sizeof(part::Job), &tab));
  plan->h_jobs = tab;
  part::Job* hj = (part::Job*)tab;
  memset(hj, 0, (size_t)m * sizeof(part::Job));
  for (int j = 0; j < m; j++) {
    hj[j].words = in[j].d_words;
    hj[j].bad = in[j].d_bad;
    hj[j].n_pos = (unsigned long long)(in[j].n_pos > 0 ? in[j].n_pos : 0);
    hj[j].base = dwords + (size_t)j * per_job_words;
    hj[j].meta = dmeta + (size_t)j * 4;
    hj[j].removed = dmeta + (size_t)m * 4 + (size_t)j * bins;
    hj[j].slice = hj[j].base + bins + 1;
  }
  plan->d_jobs = dbase;
  KMSC_CUDA(cudaMemcpyAsync(dbase, hj, (size_t)m * sizeof(part::Job), cudaMemcpyHostToDevice, ctx->stream));
  KMSC_TRY(ctx->p2tab[slot].commit(ctx->stream));
  KMSC_CUDA(cudaMemsetAsync(dmeta, 0, (size_t)m * 16 + (size_t)m * bins * 4, ctx->stream));  // the rest is written before it is read

  const part::Geo g = make_geo(*plan);
  const size_t csm = count_smem_bytes(b1);
  KMSC_TRY(set_smem_limit(part::count_kernel, csm));
  part::count_kernel<<<dim3(gx, (unsigned)m), part::kThreads, csm, ctx->stream>>>((const part::Job*)dbase, g);
  part::scan_kernel<<<(unsigned)m, 1024, 0, ctx->stream>>>((const part::Job*)dbase, g, (int)gx);
  count_launch(ctx, 2);
  KMSC_CUDA(cudaGetLastError());
  // one synchronisation for the whole batch: n_occ and the largest partition of every job
  void* pin = nullptr;
  KMSC_TRY(ctx->p2rb[slot].acquire((size_t)m * 32, &pin));
  KMSC_CUDA(cudaMemcpyAsync(pin, dmeta, (size_t)m * 16, cudaMemcpyDeviceToHost, ctx->stream));
  KMSC_CUDA(cudaStreamSynchronize(ctx->stream));
  uint32_t part_max = 0;
  for (int j = 0; j < m; j++) {
    const uint32_t* me = (const uint32_t*)((unsigned char*)pin + (size_t)j * 16);
    plan->n_occ[(size_t)j] = me[0];
    part_max = std::max(part_max, me[1]);
  }
  plan->part_max = part_max;
  if (part_max > 65535 || sort_smem_bytes(plan->FB2, part_max, opt.key_bytes) > 200 * 1024) return KMSC_OK;  // a partition too large for shared memory
  plan->feasible = true;
  return KMSC_OK;
}

template <typename KeyT>
static int launch_sort(kmsc_ctx* ctx, const PartPlan& p, const part::Geo& g, unsigned bins, uint32_t cap, size_t smem) {
  if (p.tmp_bytes == 4) {
    KMSC_TRY(set_smem_limit(part::sort_kernel<KeyT, uint32_t>, smem));
    part::sort_kernel<KeyT, uint32_t><<<dim3(bins, (unsigned)p.m), part::kSortThreads, smem, ctx->stream>>>((const part::Job*)p.d_jobs, g, cap);
  } else {
    KMSC_TRY(set_smem_limit(part::sort_kernel<KeyT, unsigned long long>, smem));
    part::sort_kernel<KeyT, unsigned long long><<<dim3(bins, (unsigned)p.m), part::kSortThreads, smem, ctx->stream>>>((const part::Job*)p.d_jobs, g, cap);
  }
  return KMSC_OK;
}

int partition_run(kmsc_ctx* ctx, PartPlan* plan, void* const* d_keys, uint32_t* const* d_fine) {
  if (!plan->feasible) { set_error("partition_run on an infeasible plan"); return KMSC_E_STATE; }
  const int m = plan->m;
  void* tab = nullptr;
  const int slot = plan->slot;
  KMSC_TRY(ctx->p2tab[slot].acquire((size_t)m * sizeof(part::Job), &tab));  // the plan's copy has run (it synchronised)
  if (tab != plan->h_jobs) { set_error("partition_run: the plan's job table was replaced"); return KMSC_E_STATE; }
  part::Job* hj = (part::Job*)tab;
  // first-level output of all jobs, back to back
  size_t total = 0;
  std::vector<size_t> toff((size_t)m);
  for (int j = 0; j < m; j++) { toff[(size_t)j] = total; total += align_up((size_t)plan->n_occ[(size_t)j] * plan->tmp_bytes + 16, 256); }
  KMSC_TRY(ctx->p2b[slot].reserve(total + 256));
  for (int j = 0; j < m; j++) {
    hj[j].tmp = (unsigned char*)ctx->p2b[slot].p + toff[(size_t)j];
    hj[j].keys = d_keys[j];
    hj[j].fine = d_fine[j];
  }
  KMSC_CUDA(cudaMemcpyAsync(plan->d_jobs, hj, (size_t)m * sizeof(part::Job), cudaMemcpyHostToDevice, ctx->stream));
  KMSC_TRY(ctx->p2tab[slot].commit(ctx->stream));
  const part::Geo g = make_geo(*plan);
  const unsigned max_tiles = (unsigned)plan->n_ctas;  // the CTAs of the count kernel, walking the same tiles
  const size_t psm = partition_smem_bytes(plan->B1);
  if (plan->tmp_bytes == 4) {
    KMSC_TRY(set_smem_limit(part::partition_kernel<uint32_t>, psm));
    part::partition_kernel<uint32_t><<<dim3(max_tiles, (unsigned)m), part::kThreads, psm, ctx->stream>>>((const part::Job*)plan->d_jobs, g);
  } else {
    KMSC_TRY(set_smem_limit(part::partition_kernel<unsigned long long>, psm));
    part::partition_kernel<unsigned long long><<<dim3(max_tiles, (unsigned)m), part::kThreads, psm, ctx->stream>>>((const part::Job*)plan->d_jobs, g);
  }
  const unsigned bins = 1u << plan->B1;
  const uint32_t cap = std::max<uint32_t>(plan->part_max, 1);
  const size_t ssm = sort_smem_bytes(plan->FB2, cap, plan->key_bytes);
  switch (plan->key_bytes) {
    case 2: KMSC_TRY(launch_sort<uint16_t>(ctx, *plan, g, bins, cap, ssm)); break;
    case 4: KMSC_TRY(launch_sort<uint32_t>(ctx, *plan, g, bins, cap, ssm)); break;
    default: KMSC_TRY(launch_sort<unsigned long long>(ctx, *plan, g, bins, cap, ssm)); break;
  }
  count_launch(ctx, 2);
  KMSC_CUDA(cudaGetLastError());
  return KMSC_OK;
}

int partition_flags_async(kmsc_ctx* ctx, PartPlan* plan) {
  const int m = plan->m;
  const part::Job* hj = (const part::Job*)plan->h_jobs;  // only the device pointers in it are read
  unsigned char* pin = (unsigned char*)ctx->p2rb[plan->slot].p + (size_t)m * 16;  // second half: the plan sized it
  KMSC_CUDA(cudaMemcpyAsync(pin, hj[0].meta, (size_t)m * 16, cudaMemcpyDeviceToHost, ctx->stream));  // meta[m][4] is contiguous
  KMSC_TRY(ctx->p2rb[plan->slot].commit(ctx->stream));
  plan->flags_host = (const uint32_t*)pin;
  return KMSC_OK;
}

int partition_flags(kmsc_ctx* ctx, PartPlan* plan, std::vector<int>* repeats) {
  const int m = plan->m;
  KMSC_TRY(partition_flags_async(ctx, plan));
  KMSC_CUDA(cudaStreamSynchronize(ctx->stream));
  repeats->assign((size_t)m, 0);
  for (int j = 0; j < m; j++) {
    const uint32_t f = plan->flags_host[(size_t)j * 4 + 2];
    if (f & 2u) { set_error("partition sort: a partition outgrew its shared-memory tile"); return KMSC_E_STATE; }
    (*repeats)[(size_t)j] = (int)(f & 1u);
  }
  return KMSC_OK;
}

template <typename KeyT>
static void launch_shift(kmsc_ctx* ctx, const part::ShiftJob* d_jobs, int n, int B1, int FB) {
  part::shift_kernel<KeyT><<<dim3(1u << B1, (unsigned)n), 256, 0, ctx->stream>>>(d_jobs, B1, FB);
}

int partition_shift(kmsc_ctx* ctx, PartPlan* plan, const std::vector<int>& jobs, kmsc_set* const* old_sets,
                    kmsc_set* const* new_sets) {
  const int n = (int)jobs.size();
  if (n == 0) return KMSC_OK;
  const part::Job* hj = (const part::Job*)plan->h_jobs;
  void* tab = nullptr;
  KMSC_TRY(ctx->tab[1].acquire((size_t)n * sizeof(part::ShiftJob), &tab));
  part::ShiftJob* h = (part::ShiftJob*)tab;
  for (int q = 0; q < n; q++) {
    const kmsc_set* o = old_sets[q];
    kmsc_set* s = new_sets[q];
    h[q] = part::ShiftJob{o->keys, s->keys, o->lev[o->max_level], s->lev[s->max_level], hj[jobs[(size_t)q]].base,
                                  hj[jobs[(size_t)q]].removed};
  }
  KMSC_TRY(ctx->tabs_dev[1].reserve((size_t)n * sizeof(part::ShiftJob)));
  KMSC_CUDA(cudaMemcpyAsync(ctx->tabs_dev[1].p, h, (size_t)n * sizeof(part::ShiftJob), cudaMemcpyHostToDevice, ctx->stream));
  KMSC_TRY(ctx->tab[1].commit(ctx->stream));
  const part::ShiftJob* dj = (const part::ShiftJob*)ctx->tabs_dev[1].p;
  switch (plan->key_bytes) {
    case 2: launch_shift<uint16_t>(ctx, dj, n, plan->B1, plan->FB); break;
    case 4: launch_shift<uint32_t>(ctx, dj, n, plan->B1, plan->FB); break;
    default: launch_shift<unsigned long long>(ctx, dj, n, plan->B1, plan->FB); break;
  }
  count_launch(ctx);
  KMSC_CUDA(cudaGetLastError());
  return KMSC_OK;
}

int derive_levels_batch(kmsc_ctx* ctx, kmsc_set* const* sets, int m) {
  if (m <= 0) return KMSC_OK;
  const int N = sets[0]->N, ml = sets[0]->max_level;
  if (ml == 0) return KMSC_OK;
  KMSC_TRY(ctx->tabs_dev[0].reserve((size_t)m * sizeof(part::LevJob)));
  void* tab = nullptr;
  KMSC_TRY(ctx->tab[0].acquire((size_t)m * sizeof(part::LevJob), &tab));
  part::LevJob* h = (part::LevJob*)tab;
  for (int j = 0; j < m; j++) h[j].lev_base = sets[j]->lev_base;
  KMSC_CUDA(cudaMemcpyAsync(ctx->tabs_dev[0].p, h, (size_t)m * sizeof(part::LevJob), cudaMemcpyHostToDevice, ctx->stream));
  KMSC_TRY(ctx->tab[0].commit(ctx->stream));
  uint64_t entries = 0;
  for (int f = 0; f < ml; f++) entries += ((uint64_t)1 << (N + f)) + 1;
  part::derive_levels_batch_kernel<<<dim3((unsigned)((entries + 255) / 256), (unsigned)m), 256, 0, ctx->stream>>>(
      (const part::LevJob*)ctx->tabs_dev[0].p, N, ml);
  count_launch(ctx);
  KMSC_CUDA(cudaGetLastError());
  return KMSC_OK;
}

}  // namespace kmsc
