// Latency micro-benchmarks on sm_100a (B200): dependent-chain cycles per op for the FP64 operations the
// scalar optimiser and the likelihood passes are made of.  One warp, one block; clock64 around N dependent ops.
#include <cstdio>
#include <cuda_runtime.h>
#define N 4096
__global__ void k_lat(double *out, long long *cyc, double a, double b, int which) {
  double x = a + threadIdx.x * 1e-9, y = b;
  __shared__ double sh[1024];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) sh[i] = (double)((i * 7 + 1) & 1023);
  __syncthreads();
  long long t0 = clock64();
  if (which == 0) { for (int i = 0; i < N; ++i) x = fma(x, y, a); }
  else if (which == 1) { for (int i = 0; i < N; ++i) x = x * y; }
  else if (which == 2) { for (int i = 0; i < N; ++i) x = x + y; }
  else if (which == 3) { for (int i = 0; i < N; ++i) x = a / (x + 1.0); }
  else if (which == 4) { for (int i = 0; i < N; ++i) x = exp(-x * 1e-3 - 0.1); }
  else if (which == 5) { for (int i = 0; i < N; ++i) x = log(x + 2.0); }
  else if (which == 6) { int j = threadIdx.x; for (int i = 0; i < N; ++i) j = (int)sh[j & 1023]; x = j; }
  else if (which == 7) { for (int i = 0; i < N; ++i) { x = fma(x, y, a); __syncthreads(); } }
  else if (which == 8) { for (int i = 0; i < N; ++i) { x = x * y; int hi = __double2hiint(x); x = __hiloint2double((hi & 0x800fffff) | 0x3ff00000, __double2loint(x)); } }
  else if (which == 9) { for (int i = 0; i < N; ++i) x = __shfl_xor_sync(0xffffffffu, x, 1) + y; }
  else if (which == 10) { float f = (float)x; for (int i = 0; i < N; ++i) f = fmaf(f, (float)y, (float)a); x = f; }
  else if (which == 11) { for (int i = 0; i < N; ++i) x = fabs(x - y) > a ? x * y : x + y; }
  long long t1 = clock64();
  out[threadIdx.x + blockIdx.x * blockDim.x] = x;
  if (threadIdx.x == 0) cyc[which] = t1 - t0;
}
// throughput: many warps, independent chains
__global__ void k_tput(double *out, long long *cyc, double a, double b, int which) {
  double x0 = a + threadIdx.x * 1e-9, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
  long long t0 = clock64();
  for (int i = 0; i < N; ++i) {
    x0 = fma(x0, b, a); x1 = fma(x1, b, a); x2 = fma(x2, b, a); x3 = fma(x3, b, a);
    x4 = fma(x4, b, a); x5 = fma(x5, b, a); x6 = fma(x6, b, a); x7 = fma(x7, b, a);
  }
  long long t1 = clock64();
  out[threadIdx.x + blockIdx.x * blockDim.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[which] = t1 - t0;
}
int main() {
  double *out; long long *cyc;
  cudaMalloc(&out, 1 << 20); cudaMallocManaged(&cyc, 64 * sizeof(long long));
  const char *names[] = {"DFMA dependent", "DMUL dependent", "DADD dependent", "DDIV (a/(x+1)) dependent", "exp() dependent", "log() dependent",
                         "LDS.64 pointer chase", "DFMA + bar.sync (128 thr)", "WarpProd-like (DMUL + 2 int + rebuild)", "SHFL.64 + DADD", "FFMA dependent", "DSETP/select chain"};
  for (int w = 0; w < 12; ++w) {
    const int thr = w == 7 ? 128 : 32;
    k_lat<<<1, thr>>>(out, cyc, 0.999, 1.0000001, w); cudaDeviceSynchronize();
    k_lat<<<1, thr>>>(out, cyc, 0.999, 1.0000001, w); cudaDeviceSynchronize();
    printf("%-42s %8.1f cycles/op\n", names[w], (double)cyc[w] / N);
  }
  for (int warps = 1; warps <= 16; warps *= 2) {
    k_tput<<<1, 32 * warps>>>(out, cyc, 0.999, 1.0000001, 20); cudaDeviceSynchronize();
    printf("DFMA throughput, %2d warps x 8 chains on one SM: %6.2f cycles per warp-DFMA per SM (%.1f lanes/cycle)\n", warps,
           (double)cyc[20] / (N * 8.0 * warps), 32.0 * N * 8.0 * warps / (double)cyc[20]);
  }
  return 0;
}
