// fp64_lat.cu -- microbenchmark: DFMA dependent-issue latency and throughput vs (warps/SMSP, ILP) on sm_100a.
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/fp64_lat tools/fp64_lat.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP>
__global__ void k(double* out, int iters, long long* cyc) {
  double a[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) a[i] = 1.0 + threadIdx.x + i;
  const double m = 0.999999, c = 1e-9;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) a[i] = fma(a[i], m, c);
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += a[i];
  if (s == 1234.5) out[0] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int ILP>
void run(int warps_per_smsp) {
  double* d; long long* c; cudaMalloc(&d, 8); cudaMalloc(&c, 8);
  const int iters = 4096;
  k<ILP><<<148, 32 * 4 * warps_per_smsp>>>(d, iters, c);
  cudaDeviceSynchronize();
  k<ILP><<<148, 32 * 4 * warps_per_smsp>>>(d, iters, c);
  long long h; cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
  double cyc_per_fma_per_warp = (double)h / (iters * ILP);
  // pipe utilisation: warp-DFMA per cycle per SMSP vs 0.5 peak
  double rate = warps_per_smsp / cyc_per_fma_per_warp;
  printf("warps/SMSP=%d ILP=%d cycles/DFMA/warp=%.2f  SMSP rate=%.3f warp-DFMA/cycle (peak 0.5) util=%.2f\n", warps_per_smsp, ILP,
         cyc_per_fma_per_warp, rate, rate / 0.5);
  cudaFree(d); cudaFree(c);
}
int main() {
  for (int w : {1, 2, 3, 4, 6, 8}) { run<1>(w); run<2>(w); run<4>(w); run<8>(w); }
  return 0;
}
