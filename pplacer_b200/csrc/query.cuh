// Query-side kernels: character coding / per-query site lists and the evolved-query table (the
// prefilter itself is in prefilter.cuh).
//   core.ml:121-158   query chars -> likelihood vectors, informative-site mask, mask_reindex
//   core.ml:187-194   score loc = Glv.masked_logdot full_query_evolv snodes.(loc) mask_vec
//   gmix_model.ml:194-198 / gcat_model.ml:220-224 ; linear_c.c:465-539 (ten) / :251-296 (mat)
#pragma once
#include "device_common.cuh"

namespace ppb {

// chars[Q][L] -> codes[Q][Lp] (alphabet code | 0x80 if uninformative), qsite[Q][Lp] = informative
// site indices in ascending order (Core.mask_reindex, core.ml:26-29) with the site's alphabet code
// packed into bits 24-30 (kSiteMask extracts the index; one load serves both), nq[Q], status[Q].
constexpr int kPrepThreads = 256;
__global__ void __launch_bounds__(kPrepThreads)
query_prep_kernel(const uint8_t *__restrict__ chars, const uint8_t *__restrict__ code_of_char, int L, int Lp,
                  uint8_t *__restrict__ codes, int32_t *__restrict__ qsite, int32_t *__restrict__ nq,
                  int32_t *__restrict__ status, int bad_status, int noinf_status) {
  __shared__ int warp_tot[kPrepThreads / 32];
  __shared__ int base_s;
  __shared__ int bad_s;
  const int64_t q = blockIdx.x;
  const uint8_t *row = chars + q * L;
  uint8_t *crow = codes + q * Lp;
  int32_t *srow = qsite + q * Lp;
  if (threadIdx.x == 0) { base_s = 0; bad_s = 0; }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int start = 0; start < Lp; start += kPrepThreads) {
    const int site = start + threadIdx.x;
    uint8_t code = 0xFF;
    bool inf = false;
    if (site < L) {
      code = code_of_char[row[site]];
      if (code == 0xFF) bad_s = 1;
      else inf = !(code & 0x80);
    }
    if (site < Lp) crow[site] = (site < L) ? code : 0xFF;
    const unsigned bal = __ballot_sync(0xffffffffu, inf);
    const int within = __popc(bal & ((1u << lane) - 1u));
    if (lane == 0) warp_tot[warp] = __popc(bal);
    __syncthreads();
    int off = base_s;
    for (int w = 0; w < warp; ++w) off += warp_tot[w];
    if (inf) srow[off + within] = site | ((int)code << kSiteCodeShift);
    __syncthreads();
    if (threadIdx.x == 0) {
      int t = 0;
      for (int w = 0; w < kPrepThreads / 32; ++w) t += warp_tot[w];
      base_s += t;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    nq[q] = base_s;
    status[q] = bad_s ? bad_status : (base_s == 0 ? noinf_status : 0);
  }
}

// Price-CAT models, second stage (three_tax_kernel): the query's informative sites ordered by (site category, site),
// each category's segment padded with -1 to a multiple of 32, so that the 32 lanes of a warp always share the
// category - the per-category rate matrices in shared memory are then read with broadcast loads (threetax.cuh).
// perm = the alignment's sites in (category, site) order, cat_start[R + 1] = the segment bounds in perm (host-built
// per category assignment).  Entries as in query_prep_kernel (site | code << 24); nqc[q] = padded length.
__global__ void __launch_bounds__(kPrepThreads)
qsite_by_category_kernel(const uint8_t *__restrict__ codes, const int32_t *__restrict__ perm, const int32_t *__restrict__ cat_start,
                         int R, int Lp, int64_t LpC, int32_t *__restrict__ qsite_cat, int32_t *__restrict__ nqc) {
  __shared__ int warp_tot[kPrepThreads / 32];
  __shared__ int base_s;
  const int64_t q = blockIdx.x;
  const uint8_t *crow = codes + q * Lp;
  int32_t *srow = qsite_cat + q * LpC;
  if (threadIdx.x == 0) base_s = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int c = 0; c < R; ++c) {
    const int j0 = cat_start[c], j1 = cat_start[c + 1];
    for (int start = j0; start < j1; start += kPrepThreads) {
      const int j = start + threadIdx.x;
      int site = 0;
      uint8_t code = 0xFF;
      if (j < j1) { site = __ldg(perm + j); code = crow[site]; }
      const bool inf = !(code & 0x80);
      const unsigned bal = __ballot_sync(0xffffffffu, inf);
      const int within = __popc(bal & ((1u << lane) - 1u));
      if (lane == 0) warp_tot[warp] = __popc(bal);
      __syncthreads();
      int off = base_s;
      for (int w = 0; w < warp; ++w) off += warp_tot[w];
      if (inf) srow[off + within] = site | ((int)code << kSiteCodeShift);
      __syncthreads();
      if (threadIdx.x == 0) {
        int t = 0;
        for (int w = 0; w < kPrepThreads / 32; ++w) t += warp_tot[w];
        base_s += t;
      }
      __syncthreads();
    }
    const int end = base_s, padded = (end + 31) & ~31;
    if (threadIdx.x < padded - end) srow[end + threadIdx.x] = -1;
    __syncthreads();
    if (threadIdx.x == 0) base_s = padded;
    __syncthreads();
  }
  if (threadIdx.x == 0) nqc[q] = base_s;
}

// Evolved-query table for pendant length t: the reference evolves the full-length query GLV
// (core.ml:137-141); an evolved indicator vector is a sum of columns of P_r(t), so it only depends
// on the character:  Tq[(cat,) code][r*S+s] = sum_{j in states(code)} P_r(t)[s][j]  (j ascending,
// the same additions gemmish_c performs on a 0/1 vector, linear_c.c:29-90).
// Gmix: Tq[n_codes][K] ; Gcat: Tq[R][n_codes][S].   One block.
template <int S>
__global__ void query_table_kernel(const DevModel *__restrict__ m, double t, double *__restrict__ Tq) {
  extern __shared__ double sm[];
  const int Rt = m->R;
  double *ex = sm;             // Rt*S
  double *P = sm + Rt * S;     // Rt*S*S
  block_build_P<S>(m, Rt, t, P, S * S, ex);
  const int nc = m->n_codes;
  for (int idx = threadIdx.x; idx < Rt * nc * S; idx += blockDim.x) {
    int r, code, s;
    if (m->gcat) { r = idx / (nc * S); code = (idx / S) % nc; s = idx % S; }
    else { code = idx / (Rt * S); r = (idx / S) % Rt; s = idx % S; }
    const uint32_t mask = m->mask_of_code[code];
    double acc = 0.0;
#pragma unroll
    for (int j = 0; j < S; ++j)
      if (mask >> j & 1u) acc += P[r * S * S + s * S + j];
    Tq[idx] = acc;
  }
}

}  // namespace ppb
