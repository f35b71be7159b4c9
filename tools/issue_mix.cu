// issue_mix.cu -- microbenchmark: does a non-fp64 instruction cost an issue slot beside DFMA on sm_100a?
// Each warp runs ITERS trips of 8 independent DFMA chains, with NI integer (IMAD/LOP3) or NL shared-load instructions
// interleaved per trip.  If a DFMA holds the issue port for its two pipe cycles, time ~ 2*N64 + N_other; if other pipes
// issue in the second cycle, time ~ max(2*N64, N64 + N_other).
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/issue_mix tools/issue_mix.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int NI, int NL>
__global__ void __launch_bounds__(512) k(double* out, int* iout, int iters, long long* cyc) {
  __shared__ double sm[8][512];
  for (int i = 0; i < 8; ++i) sm[i][threadIdx.x] = 1e-9 * (i + 1);
  double a[8];
  int x[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { a[i] = 1.0 + threadIdx.x + i; x[i] = threadIdx.x * 3 + i; }
  const double m = 0.999999, c = 1e-9;
  double ls = 0.0;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      a[i] = fma(a[i], m, c);
      if (i < NI) x[i] = x[i] * 1664525 + it;  // one IMAD per slot, independent chains
      if (i < NL) {
        double v;
        asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"((unsigned)__cvta_generic_to_shared(&sm[i][threadIdx.x])));
        ls += v;  // one DADD more per load (counted as fp64 below)
      }
    }
  }
  long long t1 = clock64();
  double s = ls;
  int xs = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) { s += a[i]; xs ^= x[i]; }
  if (s == 1234.5) out[0] = s;
  if (xs == 12345) iout[0] = xs;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int NI, int NL>
void run(int warps_per_smsp) {
  double* d; int* di; long long* c; cudaMalloc(&d, 8); cudaMalloc(&di, 8); cudaMalloc(&c, 8);
  const int iters = 4096;
  for (int rep = 0; rep < 2; ++rep) k<NI, NL><<<148, 32 * 4 * warps_per_smsp>>>(d, di, iters, c);
  cudaDeviceSynchronize();
  long long h; cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
  const double per_trip = (double)h / iters / warps_per_smsp;  // SMSP cycles per warp-trip
  const int n64 = 8 + NL;
  printf("warps/SMSP=%d  per trip: %d fp64 + %d int + %d lds  -> %.2f cycles/warp-trip   (2*N64 = %d, 2*N64+other = %d, N64+other = %d)\n",
         warps_per_smsp, n64, NI, NL, per_trip, 2 * n64, 2 * n64 + NI + NL, n64 + NI + NL);
  cudaFree(d); cudaFree(di); cudaFree(c);
}
int main() {
  for (int w : {2, 4}) { run<0, 0>(w); run<4, 0>(w); run<8, 0>(w); run<0, 4>(w); run<0, 8>(w); run<8, 8>(w); }
  return 0;
}
