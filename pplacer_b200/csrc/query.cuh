// Query-side kernels: character coding / per-query site lists, the evolved-query table, and the
// baseball prefilter (first-stage scoring of every edge).
//   core.ml:121-158   query chars -> likelihood vectors, informative-site mask, mask_reindex
//   core.ml:187-194   score loc = Glv.masked_logdot full_query_evolv snodes.(loc) mask_vec
//   gmix_model.ml:194-198 / gcat_model.ml:220-224 ; linear_c.c:465-539 (ten) / :251-296 (mat)
#pragma once
#include "device_common.cuh"

namespace ppb {

// chars[Q][L] -> codes[Q][Lp] (alphabet code | 0x80 if uninformative), qsite[Q][Lp] = informative
// site indices in ascending order (Core.mask_reindex, core.ml:26-29), nq[Q], status[Q].
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
    if (inf) srow[off + within] = site;
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

// ---------------------------------------------------------------------------------------------
// Prefilter, direct form (ten_masked_logdot_c / mat_masked_logdot_c semantics):
//   score(q,e) = sum_{site in mask_q} log( sum_{r,s} Qev[r,site,s] * snode_e[r,site,s] )
//                - n_q log R  + ln2 * sum_{site in mask_q} e_snode[site]        (query exponents are 0)
// One block per (edge tile, query tile).  Lanes walk SITES, so every supernode plane is read with
// coalesced 256-byte warp loads straight into registers; each thread keeps TE edges x QT queries of
// running (mantissa, exponent) products; a warp-shuffle + shared-memory reduction finishes the sum
// over sites.  FP64 FMAs per (query, edge, site) unit: R*S (the algorithmic 2RS flops).
// ---------------------------------------------------------------------------------------------
template <int S, bool GCAT, int TE, int QT, int THREADS>
__global__ void __launch_bounds__(THREADS)
prefilter_direct_kernel(const DevModel *__restrict__ m, GlvArrays g, const double *__restrict__ Tq,
                        const uint8_t *__restrict__ codes, const int32_t *__restrict__ nq,
                        const int32_t *__restrict__ site_cat, int Q, int root_id, double *__restrict__ scores) {
  extern __shared__ double sm[];
  const int R = m->R, nc = m->n_codes;
  const int Rg = GCAT ? 1 : R;
  const int tq_len = R * nc * S;
  double *tq = sm;                       // evolved-query table
  double *red = sm + tq_len;             // THREADS/32 * TE*QT partial sums
  for (int i = threadIdx.x; i < tq_len; i += THREADS) tq[i] = Tq[i];
  __syncthreads();
  const int e0 = blockIdx.x * TE;
  const int q0 = blockIdx.y * QT;
  const int L = g.L;
  const int64_t Lp = g.Lp;
  LogProd acc[TE][QT];
#pragma unroll
  for (int a = 0; a < TE; ++a)
#pragma unroll
    for (int b = 0; b < QT; ++b) acc[a][b].init();

  for (int site = threadIdx.x; site < L; site += THREADS) {
    uint8_t code[QT];
    bool any = false;
#pragma unroll
    for (int b = 0; b < QT; ++b) {
      code[b] = (q0 + b < Q) ? codes[(int64_t)(q0 + b) * Lp + site] : (uint8_t)0xFF;
      any |= !(code[b] & 0x80);
    }
    if (!any) continue;
    double u[TE][QT];
#pragma unroll
    for (int a = 0; a < TE; ++a)
#pragma unroll
      for (int b = 0; b < QT; ++b) u[a][b] = 0.0;
    const int cat = GCAT ? site_cat[site] : 0;
    for (int r = 0; r < Rg; ++r) {
      double sn[TE][S];
#pragma unroll
      for (int a = 0; a < TE; ++a) {
        const int e = min(e0 + a, g.n_nodes - 1);
#pragma unroll
        for (int s = 0; s < S; ++s) sn[a][s] = glv_plane(g, 2, e, r * S + s)[site];
      }
#pragma unroll
      for (int b = 0; b < QT; ++b) {
        if (code[b] & 0x80) continue;
        const double *tv = GCAT ? tq + ((int64_t)cat * nc + code[b]) * S : tq + (int64_t)code[b] * (R * S) + r * S;
#pragma unroll
        for (int s = 0; s < S; ++s) {
          const double t = tv[s];
#pragma unroll
          for (int a = 0; a < TE; ++a) u[a][b] = fma(t, sn[a][s], u[a][b]);
        }
      }
    }
#pragma unroll
    for (int a = 0; a < TE; ++a) {
      const int e = min(e0 + a, g.n_nodes - 1);
      const int es = g.e[2][(int64_t)e * Lp + site];
#pragma unroll
      for (int b = 0; b < QT; ++b)
        if (!(code[b] & 0x80)) { acc[a][b].mul(u[a][b]); acc[a][b].add_exp(es); }
    }
  }
  // reduce over the block's threads
  constexpr int NW = THREADS / 32;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int a = 0; a < TE; ++a)
#pragma unroll
    for (int b = 0; b < QT; ++b) {
      double v = acc[a][b].value();
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane == 0) red[warp * (TE * QT) + a * QT + b] = v;
    }
  __syncthreads();
  for (int i = threadIdx.x; i < TE * QT; i += THREADS) {
    const int a = i / QT, b = i % QT;
    const int e = e0 + a, q = q0 + b;
    if (e >= g.n_nodes || q >= Q) continue;
    double v = 0.0;
    for (int w = 0; w < NW; ++w) v += red[w * (TE * QT) + i];
    if (!GCAT) v -= (double)((float)nq[q]) * log((double)((float)R));  // linear_c.c:536
    scores[(int64_t)q * g.n_nodes + e] = (e == root_id) ? __longlong_as_double(0x7ff8000000000000LL) : v;
  }
}

}  // namespace ppb
