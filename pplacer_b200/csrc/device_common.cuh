// Device-side common definitions for libpplacer_b200 (sm_100a).
//
// HBM layout (DESIGN.md "Data layout"): a GLV array holds, per node, K = Rg*S planes of Lp doubles
//   a[node][k][site],  k = rate*S + state,  Lp = L rounded up to 32
// and one plane of int32 base-two exponents e[node][site].  The reference keeps a[R][L][S]
// (state fastest, linear_c.c:5-12); here the SITE index is fastest so that a warp whose lanes
// walk consecutive sites issues fully coalesced 256-byte loads for every (rate,state) plane, in
// the pruning kernels, the prefilter and the three-taxon kernels alike.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace ppb {

constexpr int kMaxS = 20;
constexpr int kMaxRates = 64;
constexpr int kMaxCodesDev = 32;
constexpr int kSiteCodeShift = 24;           // qsite entries: site index | alphabet code << 24
constexpr int kSiteMask = (1 << kSiteCodeShift) - 1;
constexpr double kLn2 = 0.693147180559945309417232121458;

struct DevModel {
  int S, R, Rg, K, gcat, n_codes, any_code, _pad;
  double statd[kMaxS];
  double x[kMaxS * kMaxS];
  double lam[kMaxS];
  double xit[kMaxS * kMaxS];
  double rates[kMaxRates];
  uint32_t mask_of_code[kMaxCodesDev];
};

struct GlvArrays {
  double *a[3];   // darr, parr, snodes : [n_nodes][K][Lp]
  double *ea[2];  // DNA Gamma-4 second stage (threetax_eigen.cuh): darr, parr in the eigenbasis of the rate matrix, same layout
  int32_t *e[3];  // [n_nodes][Lp]
  int64_t a_stride;  // K * Lp
  int L, Lp, K, n_nodes;
};

__device__ __forceinline__ const double *glv_plane(const GlvArrays &g, int which, int node, int k) {
  return g.a[which] + (int64_t)node * g.a_stride + (int64_t)k * g.Lp;
}
__device__ __forceinline__ double *glv_plane_w(const GlvArrays &g, int which, int node, int k) {
  return g.a[which] + (int64_t)node * g.a_stride + (int64_t)k * g.Lp;
}

// Per-site likelihoods are multiplied into a (mantissa in [1,2), exponent) pair instead of
// summing one log per site: identical value to sum(log u) up to rounding (and more accurate),
// ~6 integer ops instead of a ~40-instruction FP64 log.  Zero / negative / non-finite / subnormal
// factors take the slow exact path through log() so -inf and NaN propagate as in the reference
// (linear_c.c:204,418,532: ll_tot += log(util)).
struct LogProd {
  double m;    // running mantissa product, in [1,2)
  double sp;   // sum of log(u) for special u
  int64_t e;   // running exponent
  __device__ __forceinline__ void init() { m = 1.0; sp = 0.0; e = 0; }
  __device__ __forceinline__ void mul(double u) {
    const int hi = __double2hiint(u);
    const int ef = (hi >> 20) & 0x7ff;
    if (hi < 0 || ef == 0 || ef == 0x7ff) { sp += log(u); return; }
    double p = m * u;  // m in [1,2), u normal  =>  p normal, no overflow
    const int phi = __double2hiint(p);
    e += ((phi >> 20) & 0x7ff) - 1023;
    m = __hiloint2double((phi & 0x800fffff) | 0x3ff00000, __double2loint(p));
  }
  __device__ __forceinline__ void add_exp(int64_t de) { e += de; }
  // natural log of the accumulated product
  __device__ __forceinline__ double value() const { return log(m) + (double)e * kLn2 + sp; }
};

// deterministic block-wide sum: every thread returns the same bits
template <int THREADS>
__device__ __forceinline__ double block_sum(double v, double *red /* >= 33 doubles smem */) {
  constexpr int NW = THREADS / 32;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();  // protect red[] from the previous use
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  double t = 0.0;
#pragma unroll
  for (int w = 0; w < NW; ++w) t += red[w];
  return t;
}

template <int THREADS>
__device__ __forceinline__ long long block_sum_ll(long long v, long long *red) {
  constexpr int NW = THREADS / 32;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  long long t = 0;
#pragma unroll
  for (int w = 0; w < NW; ++w) t += red[w];
  return t;
}

// Diagd.multi_exp (diagd.ml:79-98) + dediagonalize (linear_c.c:92-146) by one thread block:
// P[r][i][j] = sum_k (exp(rate_r*bl*lambda_k) * x[i][k]) * xit[j][k]   for r < Rt
// dst has a per-rate stride `rstride` (>= S*S; padded for Gcat to dodge bank conflicts).
// ex is Rt*S doubles of shared scratch.  Ends with a __syncthreads().
template <int S>
__device__ __forceinline__ void block_build_P(const DevModel *__restrict__ m, int Rt, double bl, double *dst,
                                              int rstride, double *ex) {
  __syncthreads();
  for (int t = threadIdx.x; t < Rt * S; t += blockDim.x) {
    const int r = t / S, k = t - r * S;
    ex[t] = exp(m->rates[r] * bl * m->lam[k]);
  }
  __syncthreads();
  for (int t = threadIdx.x; t < Rt * S * S; t += blockDim.x) {
    const int r = t / (S * S), ij = t - r * S * S, i = ij / S, j = ij - i * S;
    double acc = 0.0;
#pragma unroll
    for (int k = 0; k < S; ++k) acc = fma(ex[r * S + k] * m->x[i * S + k], m->xit[j * S + k], acc);
    dst[r * rstride + ij] = acc;
  }
  __syncthreads();
}

// Gcat rate-matrix stride in shared memory: S*S + 1 doubles, so lanes that hold different site
// categories hit different banks when they read the same (i,j).
template <int S, bool GCAT>
__host__ __device__ constexpr int p_rstride() { return GCAT ? S * S + 1 : S * S; }

}  // namespace ppb
