// umma_probe.cu -- standalone B200 probe for csrc/umma.cuh: checks the tcgen05 kind::i8 Gram
// (W = X X^T over 0/1 bytes) for the shared-memory layouts / descriptor conventions the
// library relies on, and times back-to-back MMAs per shape. Not part of the product.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I kmer-sets-compression_b200/csrc \
//        tools/umma_probe.cu -o kmer-sets-compression_b200/host/bin/umma_probe
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "umma.cuh"

using namespace kmsc::umma;

struct Cfg {
  int M, N;          // MMA shape
  int mn_major;      // 0 = K-major operands, 1 = MN-major operands
  int place;         // physical placement variant (see place_addr)
  int swap;          // 1 = swap LBO/SBO in the descriptor (hypothesis test)
  int nsets;         // rows of X that hold data (<= R)
  int ksteps;        // K-steps of 32 keys
  int a_row0;        // first set of the A operand (row half for n = 256)
  int reps;          // timing: MMAs issued back to back (0 = none)
  int naccum;        // timing: independent accumulators used round-robin (1, 2 or 4)
};

__host__ __device__ inline int rows_of(const Cfg& c) { return c.M > c.N ? c.M : c.N; }

// byte offset of X[set][key] inside one K-step buffer, plus the (lbo, sbo) that describe it
__host__ __device__ inline uint32_t place_addr(const Cfg& c, int set, int key, uint32_t* lbo, uint32_t* sbo) {
  const int R = rows_of(c);
  if (!c.mn_major) {
    if (c.place == 0) {  // [key half][set group][8 sets][16 keys]
      *lbo = (uint32_t)(R / 8) * 128u; *sbo = 128u;
      return (uint32_t)(key / 16) * *lbo + (uint32_t)(set / 8) * 128u + (uint32_t)(set % 8) * 16u + (uint32_t)(key % 16);
    }
    // [set group][key half][8 sets][16 keys]
    *lbo = 128u; *sbo = 256u;
    return (uint32_t)(set / 8) * 256u + (uint32_t)(key / 16) * 128u + (uint32_t)(set % 8) * 16u + (uint32_t)(key % 16);
  }
  if (c.place == 0) {  // [set block of 16][key group of 8][8 keys][16 sets]
    *lbo = 128u; *sbo = 512u;
    return (uint32_t)(set / 16) * 512u + (uint32_t)(key / 8) * 128u + (uint32_t)(key % 8) * 16u + (uint32_t)(set % 16);
  }
  // [key group of 8][set block of 16][8 keys][16 sets]
  *lbo = (uint32_t)(R / 16) * 128u; *sbo = 128u;
  return (uint32_t)(key / 8) * *lbo + (uint32_t)(set / 16) * 128u + (uint32_t)(key % 8) * 16u + (uint32_t)(set % 16);
}

__global__ void __launch_bounds__(128) probe_kernel(Cfg c, const uint8_t* __restrict__ bits /*[ksteps*32][R]*/,
                                                    int32_t* __restrict__ raw /*[128][N]*/,
                                                    long long* __restrict__ cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int R = rows_of(c);
  const uint32_t kstep_bytes = 32u * (uint32_t)R;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (uint32_t i = tid; i < kstep_bytes * c.ksteps / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  __syncthreads();
  uint32_t lbo = 0, sbo = 0;
  for (int i = tid; i < c.ksteps * 32 * c.nsets; i += blockDim.x) {
    const int key = i / c.nsets, set = i % c.nsets;
    const uint32_t a = place_addr(c, set, key % 32, &lbo, &sbo);
    smem[(uint32_t)(key / 32) * kstep_bytes + a] = bits[(size_t)key * R + set];
  }
  place_addr(c, 0, 0, &lbo, &sbo);
  uint32_t ncols = (uint32_t)(c.N * (c.naccum > 0 ? c.naccum : 1));
  if (ncols < 32) ncols = 32;
  if (warp == 0) tmem_alloc(&tmem_base_s, ncols);
  if (tid == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
  fence_async_smem();
  fence_before_thread_sync();
  __syncthreads();
  fence_after_thread_sync();
  const uint32_t tmem_base = tmem_base_s;
  const uint32_t idesc = make_idesc_u8(c.M, c.N, c.mn_major, c.mn_major);
  const uint32_t d_lbo = c.swap ? sbo : lbo, d_sbo = c.swap ? lbo : sbo;
  // start address of the A operand's first row
  uint32_t t0, t1;
  const uint32_t a_off = place_addr(c, c.a_row0, 0, &t0, &t1);
  const uint32_t base = smem_u32(smem);
  if (tid == 0) {
    for (int ks = 0; ks < c.ksteps; ks++) {
      const uint64_t ad = make_smem_desc(base + ks * kstep_bytes + a_off, d_lbo, d_sbo);
      const uint64_t bd = make_smem_desc(base + ks * kstep_bytes, d_lbo, d_sbo);
      mma_u8(tmem_base, ad, bd, idesc, ks > 0 ? 1u : 0u);
    }
    mma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  fence_after_thread_sync();
  // raw dump: warp w reads its 32 lanes, all N columns
  for (int c0 = 0; c0 < c.N; c0 += 32) {
    uint32_t v[32];
    tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 32; j++)
      if (c0 + j < c.N) raw[(size_t)(warp * 32 + lane) * c.N + c0 + j] = (int32_t)v[j];
  }
  // timing: reps MMAs back to back over the staged K-steps (results discarded), descriptors
  // precomputed, round-robin over c.naccum independent accumulators
  if (c.reps > 0) {
    fence_before_thread_sync();
    __syncthreads();
    fence_after_thread_sync();
    long long t_start = clock64();
    if (tid == 0) {
      uint64_t ad[8], bd[8];
#pragma unroll
      for (int j = 0; j < 8; j++) {
        const int ks = j % c.ksteps;
        ad[j] = make_smem_desc(base + ks * kstep_bytes + a_off, d_lbo, d_sbo);
        bd[j] = make_smem_desc(base + ks * kstep_bytes, d_lbo, d_sbo);
      }
      const uint32_t astep = (uint32_t)c.N;
      for (int r = 0; r < c.reps; r += 8) {
#pragma unroll
        for (int j = 0; j < 8; j++) mma_u8(tmem_base + (uint32_t)(j % c.naccum) * astep, ad[j], bd[j], idesc, 1u);
      }
      mma_commit(&bar);
    }
    mbar_wait(&bar, 1);
    long long t_end = clock64();
    if (tid == 0) cycles[blockIdx.x] = t_end - t_start;
  }
  fence_before_thread_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, ncols);
}

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); return 2; } } while (0)

static int run(const Cfg& c, int grid, const char* name) {
  const int R = rows_of(c);
  const int keys = c.ksteps * 32;
  std::vector<uint8_t> bits((size_t)keys * R, 0);
  uint32_t seed = 12345u + c.M * 7 + c.N * 13 + c.mn_major * 101 + c.place * 1009;
  for (int k = 0; k < keys; k++)
    for (int s = 0; s < c.nsets; s++) {
      seed = seed * 1664525u + 1013904223u;
      bits[(size_t)k * R + s] = ((seed >> 24) % 10) < 3 ? 1 : 0;
    }
  uint8_t* d_bits; int32_t* d_raw; long long* d_cyc;
  CK(cudaMalloc(&d_bits, bits.size()));
  CK(cudaMalloc(&d_raw, (size_t)128 * c.N * 4));
  CK(cudaMalloc(&d_cyc, sizeof(long long) * grid));
  CK(cudaMemcpy(d_bits, bits.data(), bits.size(), cudaMemcpyHostToDevice));
  CK(cudaMemset(d_raw, 0xff, (size_t)128 * c.N * 4));
  CK(cudaMemset(d_cyc, 0, sizeof(long long) * grid));
  size_t smem = (size_t)32 * R * c.ksteps;
  if (smem < 100 * 1024) smem = 100 * 1024;  // head room so hypothesis descriptors stay in bounds
  CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  CK(cudaEventRecord(e0));
  probe_kernel<<<grid, 128, smem>>>(c, d_bits, d_raw, d_cyc);
  CK(cudaEventRecord(e1));
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());
  float ms = 0; CK(cudaEventElapsedTime(&ms, e0, e1));
  std::vector<int32_t> raw((size_t)128 * c.N);
  std::vector<long long> cyc(grid);
  CK(cudaMemcpy(raw.data(), d_raw, raw.size() * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(cyc.data(), d_cyc, sizeof(long long) * grid, cudaMemcpyDeviceToHost));
  // expected D[m][n] = sum_k X[a_row0 + m][k] X[n][k]
  long long bad128 = 0, bad64 = 0;
  for (int m = 0; m < c.M; m++)
    for (int n = 0; n < c.N; n++) {
      int want = 0;
      for (int k = 0; k < keys; k++) want += bits[(size_t)k * R + c.a_row0 + m] * bits[(size_t)k * R + n];
      if (c.M == 128) { if (raw[(size_t)m * c.N + n] != want) bad128++; }
      else {
        const int lane = (m % 16) + 32 * (m / 16);
        if (raw[(size_t)lane * c.N + n] != want) bad64++;
        if (raw[(size_t)m * c.N + n] != want) bad128++;
      }
    }
  long long cmax = 0;
  for (long long x : cyc) cmax = x > cmax ? x : cmax;
  const long long bad = c.M == 128 ? bad128 : bad64;
  printf("%-34s M=%3d N=%3d mn=%d place=%d swap=%d sets=%3d ksteps=%2d a_row0=%3d nacc=%d grid=%3d : %s (mismatches %lld%s)", name,
         c.M, c.N, c.mn_major, c.place, c.swap, c.nsets, c.ksteps, c.a_row0, c.naccum, grid, bad == 0 ? "OK  " : "FAIL", bad,
         c.M == 64 ? (bad128 == 0 ? ", row=lane mapping matches" : "") : "");
  if (c.reps > 0) printf("  | %d MMAs: %lld cycles max = %.1f cyc/MMA, kernel %.3f ms", c.reps, cmax, (double)cmax / c.reps, ms);
  printf("\n");
  if (bad != 0) {
    printf("    raw[0][0..7] =");
    for (int j = 0; j < 8 && j < c.N; j++) printf(" %d", raw[j]);
    int w0 = 0; for (int k = 0; k < keys; k++) w0 += bits[(size_t)k * R + c.a_row0] * bits[(size_t)k * R];
    printf("   want[0][0] = %d\n", w0);
  }
  cudaFree(d_bits); cudaFree(d_raw); cudaFree(d_cyc);
  return bad == 0 ? 0 : 1;
}

int main() {
  int fails = 0;
  // correctness over layouts / conventions
  for (int mn = 0; mn < 2; mn++)
    for (int place = 0; place < 2; place++)
      for (int swap = 0; swap < 2; swap++) {
        Cfg c{128, 128, mn, place, swap, 100, 3, 0, 0, 1};
        fails += run(c, 1, swap ? "hypothesis (LBO/SBO swapped)" : "layout as documented in umma.cuh") && !swap;
      }
  for (int mn = 0; mn < 2; mn++) {
    Cfg a{64, 64, mn, 0, 0, 64, 4, 0, 0, 1};     fails += run(a, 1, "n<=64");
    Cfg b{128, 256, mn, 0, 0, 256, 2, 0, 0, 1};  fails += run(b, 1, "n<=256 rows 0..127");
    Cfg d{128, 256, mn, 0, 0, 256, 2, 128, 0, 1}; fails += run(d, 1, "n<=256 rows 128..255");
    Cfg e{64, 64, mn, 0, 0, 37, 5, 0, 0, 1};     fails += run(e, 1, "n=37 (padded rows zero)");
  }
  // timing (MN-major, the layout the library uses)
  for (int na = 1; na <= 4; na *= 2) {
    Cfg a{64, 64, 1, 0, 0, 64, 8, 0, 4096, na};     run(a, 1, "time 1 CTA"); run(a, 148, "time 148 CTAs"); run(a, 296, "time 296 CTAs");
    Cfg b{128, 128, 1, 0, 0, 128, 8, 0, 4096, na};  run(b, 148, "time 148 CTAs");
    if (na <= 2) { Cfg d{128, 256, 1, 0, 0, 256, 8, 0, 4096, na};  run(d, 148, "time 148 CTAs"); }
  }
  printf(fails ? "PROBE: %d documented-layout checks FAILED\n" : "PROBE: all documented-layout checks OK\n", fails);
  return 0;
}
