// tmem_probe.cu -- can tensor memory serve as per-thread scratch for a non-MMA kernel?  (sm_100a)
// Each warp owns the 32 TMEM lanes of its sub-partition (warp id % 4); with the 32x32b shape thread t of the warp reads and
// writes lane t, i.e. a column of TMEM is one 32-bit word PER THREAD: 512 columns = 2 KB per thread, dynamically indexed,
// outside the register file and outside shared memory.  This probe checks data integrity and measures the round-trip
// latency and the throughput of double-precision loads against the same access pattern in shared memory.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tmem_probe tmem_probe.cu && ./tmem_probe
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void tmem_st2(uint32_t taddr, double v) {
  const unsigned long long b = (unsigned long long)__double_as_longlong(v);
  asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1, %2};" ::"r"(taddr), "r"((unsigned)b), "r"((unsigned)(b >> 32)) : "memory");
}
__device__ __forceinline__ double tmem_ld2(uint32_t taddr) {
  unsigned lo, hi;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];" : "=r"(lo), "=r"(hi) : "r"(taddr) : "memory");
  return __longlong_as_double((long long)(((unsigned long long)hi << 32) | lo));
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

constexpr int NT = 256;       // threads per block
constexpr int COLS = 256;     // TMEM columns per block (2 blocks per SM -> all 512)
constexpr int SLOTS = 45;     // doubles per thread

__global__ void __launch_bounds__(NT, 2) k_probe(double* out, long long* cyc, int iters, int mode) {
  __shared__ uint32_t s_base;
  extern __shared__ double sm[];  // [SLOTS][NT] for the shared-memory variant
  const int tid = threadIdx.x, w = tid >> 5;
  if (w == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(&s_base)), "n"(COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t base = s_base;
  // warps w and w + 4 share the lanes of sub-partition w % 4: each takes half of the block's columns
  const uint32_t t0 = base + ((uint32_t)((w & 3) * 32) << 16) + (uint32_t)((w >> 2) * (COLS / 2));
  // integrity: write SLOTS doubles, read them back in another order
  for (int i = 0; i < SLOTS; ++i) tmem_st2(t0 + 2 * i, 1000.0 * blockIdx.x + tid + 0.001 * i);
  tmem_wait_st();
  double acc = 0.0;
  int bad = 0;
  for (int i = SLOTS - 1; i >= 0; --i) {
    const double v = tmem_ld2(t0 + 2 * i);
    tmem_wait_ld();
    if (v != 1000.0 * blockIdx.x + tid + 0.001 * i) ++bad;
    acc += v;
  }
  for (int i = 0; i < SLOTS; ++i) sm[i * NT + tid] = 0.5 * i + tid;
  __syncthreads();
  // timing: mode 0 dependent TMEM loads (latency), 1 batches of 9 independent TMEM loads + one wait (throughput),
  //         2 dependent shared loads, 3 batches of 9 shared loads
  long long c0 = clock64();
  double s = 0.0;
  int idx = tid % SLOTS;
  if (mode == 0) {
    for (int it = 0; it < iters; ++it) { const double v = tmem_ld2(t0 + 2 * idx); tmem_wait_ld(); s += v; idx = (idx + 1 + ((int)v & 0)) % SLOTS; }
  } else if (mode == 1) {
    for (int it = 0; it < iters; ++it) {
      double v[9];
#pragma unroll
      for (int j = 0; j < 9; ++j) v[j] = tmem_ld2(t0 + 2 * ((idx + j) % 36));
      tmem_wait_ld();
#pragma unroll
      for (int j = 0; j < 9; ++j) s += v[j];
      idx = (idx + 9) % 36;
    }
  } else if (mode == 2) {
    for (int it = 0; it < iters; ++it) { const double v = sm[idx * NT + tid]; s += v; idx = (idx + 1 + ((int)v & 0)) % SLOTS; }
  } else {
    for (int it = 0; it < iters; ++it) {
      double v[9];
#pragma unroll
      for (int j = 0; j < 9; ++j) v[j] = *(volatile double*)&sm[((idx + j) % 36) * NT + tid];
#pragma unroll
      for (int j = 0; j < 9; ++j) s += v[j];
      idx = (idx + 9) % 36;
    }
  }
  long long c1 = clock64();
  if (tid == 0) cyc[blockIdx.x] = c1 - c0;
  out[blockIdx.x * NT + tid] = acc + s + bad * 1e9;
  if (bad) atomicAdd((unsigned long long*)&cyc[gridDim.x], 1ULL);
  __syncthreads();
  if (w == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "n"(COLS));
}

int main() {
  const int blocks = 296, iters = 2000;
  double* out; long long* cyc;
  cudaMalloc(&out, blocks * NT * sizeof(double));
  cudaMalloc(&cyc, (blocks + 1) * sizeof(long long));
  cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, SLOTS * NT * 8);
  const char* names[4] = {"TMEM dependent ld.x2 + wait", "TMEM 9 x ld.x2 + one wait", "smem dependent ld.f64", "smem 9 x ld.f64"};
  for (int mode = 0; mode < 4; ++mode) {
    cudaMemset(cyc, 0, (blocks + 1) * sizeof(long long));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k_probe<<<blocks, NT, SLOTS * NT * 8>>>(out, cyc, iters, mode);  // warm-up
    cudaEventRecord(e0);
    k_probe<<<blocks, NT, SLOTS * NT * 8>>>(out, cyc, iters, mode);
    cudaEventRecord(e1);
    cudaError_t err = cudaDeviceSynchronize();
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    long long h[blocks + 1]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    const double per = (double)h[0] / iters;
    printf("%-30s err=%s bad=%lld  cycles/iter (block 0, 8 warps x 2 blocks per SM) %.1f  kernel %.3f ms\n", names[mode], cudaGetErrorString(err), h[blocks], per, ms);
  }
  return 0;
}
