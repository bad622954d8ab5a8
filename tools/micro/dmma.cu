// FP64 tensor-core (DMMA) micro-benchmark on sm_100a: throughput and dependent latency of mma.sync m8n8k4 / m16n8k4 / m16n8k8 f64
// against plain DFMA, per SM and for the whole chip.
#include <cstdio>
#include <cuda_runtime.h>
#define N 2048
__device__ __forceinline__ void mma884(double &c0, double &c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void mma1684(double (&c)[4], double a0, double a1, double b) {
  asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
               : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3]) : "d"(a0), "d"(a1), "d"(b));
}
__device__ __forceinline__ void mma1688(double (&c)[4], const double (&a)[4], double b0, double b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3]) : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b0), "d"(b1));
}
template <int WHICH, int CH>
__global__ void k(double *out, long long *cyc, double a, double b) {
  double c[CH][4];
  for (int i = 0; i < CH; ++i) for (int j = 0; j < 4; ++j) c[i][j] = threadIdx.x * 1e-3 + i + j;
  double av[4] = {a, a * 1.01, a * 1.02, a * 1.03};
  long long t0 = clock64();
  for (int it = 0; it < N; ++it) {
#pragma unroll
    for (int i = 0; i < CH; ++i) {
      if (WHICH == 0) mma884(c[i][0], c[i][1], a, b);
      else if (WHICH == 1) mma1684(c[i], a, av[1], b);
      else if (WHICH == 2) mma1688(c[i], av, b, b * 0.99);
      else { c[i][0] = fma(c[i][0], a, b); }
    }
  }
  long long t1 = clock64();
  double s = 0; for (int i = 0; i < CH; ++i) for (int j = 0; j < 4; ++j) s += c[i][j];
  out[threadIdx.x + blockIdx.x * blockDim.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int WHICH, int CH>
void run(const char *name, double fma_per_inst, double *out, long long *cyc) {
  for (int warps : {1, 4, 8, 16}) {
    k<WHICH, CH><<<1, 32 * warps>>>(out, cyc, 0.999, 1.0001); cudaDeviceSynchronize();
    k<WHICH, CH><<<1, 32 * warps>>>(out, cyc, 0.999, 1.0001); cudaDeviceSynchronize();
    const double inst = (double)N * CH * warps;
    printf("%-22s chains/warp %d warps %2d: %7.2f cycles per warp-instr per SM, %7.1f FMA/cycle/SM\n", name, CH, warps, *cyc / inst, inst * fma_per_inst / *cyc);
  }
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<WHICH, CH><<<148 * 4, 256>>>(out, cyc, 0.999, 1.0001); cudaDeviceSynchronize();
  cudaEventRecord(e0); k<WHICH, CH><<<148 * 4, 256>>>(out, cyc, 0.999, 1.0001); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  printf("%-22s whole chip: %.2f TFLOP/s\n", name, 2.0 * fma_per_inst * N * CH * 148 * 4 * 8 / (ms * 1e-3) / 1e12);
}
int main() {
  double *out; long long *cyc;
  cudaMalloc(&out, 8 << 20); cudaMallocManaged(&cyc, 8);
  run<3, 8>("DFMA", 32, out, cyc);
  run<0, 1>("DMMA m8n8k4 (dep)", 256, out, cyc);
  run<0, 8>("DMMA m8n8k4", 256, out, cyc);
  run<1, 1>("DMMA m16n8k4 (dep)", 512, out, cyc);
  run<1, 8>("DMMA m16n8k4", 512, out, cyc);
  run<2, 8>("DMMA m16n8k8", 1024, out, cyc);
  return 0;
}
