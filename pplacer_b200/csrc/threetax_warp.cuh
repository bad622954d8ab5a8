// Second stage, warp-per-candidate engine for the four-state Gamma-mixture model (the DNA configs).
//   Three_tax.log_like / optimize / calc_marg_prob   pplacer_src/three_tax.ml:30-83,124-215
//   Model.log_like3    gmix_model.ml:298-304 -> ten_log_like3_c (linear_c.c:367-424)
//   Model.evolve_into  gmix_model.ml:308-315 ; Diagd.multi_exp diagd.ml:79-98
// The scalar optimiser / integrator is TTDriver (threetax.cuh); this file supplies the likelihood
// passes.  Design:
//   * ONE WARP owns one (query, edge) candidate: lanes stride the query's informative sites, the
//     scalar Brent / QNG state machine runs redundantly in all 32 lanes, reductions are xor
//     butterflies (every lane ends with the same bits), and there is no block-level barrier in the
//     candidate loop, so 16-24 candidates per SM hide each other's latency;
//   * objective values at points known IN ADVANCE are computed several per pass over the data:
//     Brent's first seven calls (f(left), f(right), f(start), then GSL's f(lower), f(upper),
//     f(minimum), f(lower + golden*(upper-lower))) touch four distinct points -> one pass instead of
//     seven; each QNG rule (21 / +22 / +44 abscissae) is evaluated eight points per pass.  The
//     values are kept in a small per-warp memo keyed by the exact argument bits, so the optimiser
//     sees precisely the numbers a call-by-call evaluation would give;
//   * the pass that caches w = (statd*prox_ev)*dist_ev is fused with the first pendant batch;
//   * candidates are pulled from a queue ORDERED BY EDGE, so concurrently running warps read the
//     same darr/parr planes from L2.
#pragma once
#include "threetax.cuh"

namespace ppb {

template <int R>
struct TTWarp : TTDriver<TTWarp<R>> {
  using Base = TTDriver<TTWarp<R>>;
  using Base::pend_bl; using Base::dist_bl; using Base::prox_bl; using Base::calls;
  static constexpr int S = 4, K = 4 * R, PM = 16 * R;
  static constexpr int NBP = 8;    // pendant points per pass
  static constexpr int NBD = 4;    // distal points per pass
  static constexpr int MEMO = 44;  // largest QNG rule extension
  // block-shared constants (doubles): statd[4] lam[4] x[16] xit[16] rates[R]
  static constexpr int C_STATD = 0, C_LAM = 4, C_X = 8, C_XIT = 24, C_RATES = 40, C_TOTAL = 40 + R;
  // per-warp shared layout (doubles)
  static constexpr int OFF_EX = 0;                // [8][K] exp scratch
  static constexpr int OFF_PQ = OFF_EX + 8 * K;   // current pendant P [R][4][4]
  static constexpr int OFF_PD = OFF_PQ + PM;      // current distal P, then proximal P
  static constexpr int OFF_PB = OFF_PD + 2 * PM;  // batch: pendant [NBP][PM] | distal [NBD][2][PM]
  static constexpr int OFF_MX = OFF_PB + 8 * PM;  // memo: argument
  static constexpr int OFF_MA = OFF_MX + MEMO;    //       proximal length (distal memo) or 0
  static constexpr int OFF_MF = OFF_MA + MEMO;    //       log-likelihood
  static constexpr int PER_WARP = OFF_MF + MEMO;
  static_assert(NBD * 2 <= 8 && NBP <= 8, "batch matrices share one region");

  const double *cst;  // block-shared constants
  double *sw;         // this warp's shared region
  int lane;
  // candidate
  const double *dbase, *pbase;
  const int32_t *ebase_d, *ebase_p;
  const int32_t *qsite;
  const uint8_t *qcode;
  double *w;
  int64_t Lp;
  int nq;
  double e_tot_ln2, rate_term;
  bool w_stale;
  int memo_n, memo_pending, memo_kind;  // memo_kind: -1 none, 0 pendant values, 1 distal values

  static __host__ __device__ constexpr size_t smem_bytes(int warps) { return sizeof(double) * (size_t)(C_TOTAL + 1 + warps * PER_WARP); }

  __device__ __forceinline__ static double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
  }

  // Diagd.multi_exp (diagd.ml:79-98; linear_c.c:92-146) for n branch lengths at once:
  // dst + b*stride receives P(bls[b]); the lanes split the exponentials and the matrix entries.
  template <int N>
  __device__ __forceinline__ void build_P_n(double *dst, int stride, const double (&bls)[N]) {
    double *ex = sw + OFF_EX;
    __syncwarp();
    for (int t = lane; t < N * K; t += 32) {
      const int b = t / K, kk = t - b * K;
      double blv = bls[0];
#pragma unroll
      for (int bb = 1; bb < N; ++bb) blv = (b == bb) ? bls[bb] : blv;
      ex[t] = exp(cst[C_RATES + (kk >> 2)] * blv * cst[C_LAM + (kk & 3)]);
    }
    __syncwarp();
    for (int t = lane; t < N * PM; t += 32) {
      const int b = t / PM, e = t - b * PM;
      const int r = e >> 4, i = (e >> 2) & 3, j = e & 3;
      double acc = 0.0;
#pragma unroll
      for (int k = 0; k < 4; ++k) acc = fma(ex[b * K + r * 4 + k] * cst[C_X + i * 4 + k], cst[C_XIT + j * 4 + k], acc);
      dst[b * stride + e] = acc;
    }
    __syncwarp();
  }
  __device__ __noinline__ void build_P(double *dst, double bl) { const double b1[1] = {bl}; build_P_n<1>(dst, 0, b1); }

  __device__ void memo_clear(int kind) { memo_n = 0; memo_pending = 0; memo_kind = kind; }
  __device__ void memo_put(double x, double aux, double f) {
    if (memo_n >= MEMO) return;
    __syncwarp();
    if (lane == 0) { sw[OFF_MX + memo_n] = x; sw[OFF_MA + memo_n] = aux; sw[OFF_MF + memo_n] = f; }
    __syncwarp();
    ++memo_n;
  }
  __device__ bool memo_get(int kind, double x, double aux, double *f) const {
    if (memo_kind != kind) return false;
    for (int i = 0; i < memo_n; ++i)
      if (sw[OFF_MX + i] == x && sw[OFF_MA + i] == aux) { *f = sw[OFF_MF + i]; return true; }
    return false;
  }

  // Glv_edge.set_bl (glv_edge.ml:33-35) on the three edges: only the rate matrices change here
  __device__ void set_pend(double bl) {
    pend_bl = bl;
    build_P(sw + OFF_PQ, bl);
    if (memo_kind == 1) memo_kind = -1;  // distal values were computed for another pendant length
  }
  __device__ void set_dist_prox(double d, double p) {
    dist_bl = d; prox_bl = p;
    const double b2[2] = {d, p};
    build_P_n<2>(sw + OFF_PD, PM, b2);
    w_stale = true;
    if (memo_kind == 0) memo_kind = -1;
  }

  // evolved pendant GLV entry for state row rs = r*4+s of matrix set Pq: sum of the columns of
  // P_r(pendant) selected by the query character (lv_of_nuc, nuc_models.ml:49-54; core.ml:137-141)
  __device__ __forceinline__ static double q_entry(const double *Pq, int rs, int col, uint32_t mask, bool single) {
    const double *row = Pq + rs * 4;
    if (single) return row[col];
    double acc = 0.0;
    while (mask) { const int j = __ffs(mask) - 1; mask &= mask - 1; acc += row[j]; }
    return acc;
  }

  __device__ __forceinline__ double finish(const LogProd &lp) const { return warp_sum(lp.value()) - rate_term + e_tot_ln2; }

  // Three_tax.log_like at NB (distal, proximal) pairs with the current pendant matrices: Pm holds
  // NB x {P(distal), P(proximal)}.  One pass over darr/parr planes at the query's sites.
  template <int NB>
  __device__ __forceinline__ void pass_dist(const double *Pm, double (&out)[NB]) {
    LogProd lp[NB];
#pragma unroll
    for (int b = 0; b < NB; ++b) lp[b].init();
    const double *Pq = sw + OFF_PQ;
    const double st0 = cst[C_STATD], st1 = cst[C_STATD + 1], st2 = cst[C_STATD + 2], st3 = cst[C_STATD + 3];
    for (int i = lane; i < nq; i += 32) {
      const int site = qsite[i];
      const uint32_t mask = qcode[site] & 0x7f;
      const bool single = (mask & (mask - 1)) == 0;
      const int col = __ffs(mask) - 1;
      double u[NB];
#pragma unroll
      for (int b = 0; b < NB; ++b) u[b] = 0.0;
#pragma unroll
      for (int r = 0; r < R; ++r) {
        double d[4], p[4], q[4];
#pragma unroll
        for (int s = 0; s < 4; ++s) {
          d[s] = __ldg(dbase + (int64_t)(r * 4 + s) * Lp + site);
          p[s] = __ldg(pbase + (int64_t)(r * 4 + s) * Lp + site);
          q[s] = q_entry(Pq, r * 4 + s, col, mask, single);
        }
#pragma unroll
        for (int b = 0; b < NB; ++b) {
          const double *Pd = Pm + b * 2 * PM + r * 16, *Pp = Pd + PM;
#pragma unroll
          for (int s = 0; s < 4; ++s) {
            double de = 0.0, pe = 0.0;
#pragma unroll
            for (int j = 0; j < 4; ++j) { de = fma(Pd[s * 4 + j], d[j], de); pe = fma(Pp[s * 4 + j], p[j], pe); }
            const double st = s == 0 ? st0 : s == 1 ? st1 : s == 2 ? st2 : st3;
            const double ww = (st * pe) * de;  // ((statd*x)*y), x = prox, y = dist (three_tax.ml:31-36)
            u[b] = fma(ww, q[s], u[b]);
          }
        }
      }
#pragma unroll
      for (int b = 0; b < NB; ++b) lp[b].mul(u[b]);
    }
#pragma unroll
    for (int b = 0; b < NB; ++b) out[b] = finish(lp[b]);
  }

  // log_like at NB pendant lengths (matrices Pqm[b]) with the current distal / proximal matrices.
  // FUSED: recompute w from darr/parr and store it; otherwise read the cached w.
  template <int NB, bool FUSED>
  __device__ __forceinline__ void pass_pend(const double *Pqm, double (&out)[NB]) {
    LogProd lp[NB];
#pragma unroll
    for (int b = 0; b < NB; ++b) lp[b].init();
    const double *Pd0 = sw + OFF_PD;
    const double st0 = cst[C_STATD], st1 = cst[C_STATD + 1], st2 = cst[C_STATD + 2], st3 = cst[C_STATD + 3];
    for (int i = lane; i < nq; i += 32) {
      const int site = qsite[i];
      const uint32_t mask = qcode[site] & 0x7f;
      const bool single = (mask & (mask - 1)) == 0;
      const int col = __ffs(mask) - 1;
      double u[NB];
#pragma unroll
      for (int b = 0; b < NB; ++b) u[b] = 0.0;
#pragma unroll
      for (int r = 0; r < R; ++r) {
        double ww[4];
        if (FUSED) {
          double d[4], p[4];
#pragma unroll
          for (int s = 0; s < 4; ++s) {
            d[s] = __ldg(dbase + (int64_t)(r * 4 + s) * Lp + site);
            p[s] = __ldg(pbase + (int64_t)(r * 4 + s) * Lp + site);
          }
          const double *Pd = Pd0 + r * 16, *Pp = Pd + PM;
#pragma unroll
          for (int s = 0; s < 4; ++s) {
            double de = 0.0, pe = 0.0;
#pragma unroll
            for (int j = 0; j < 4; ++j) { de = fma(Pd[s * 4 + j], d[j], de); pe = fma(Pp[s * 4 + j], p[j], pe); }
            const double st = s == 0 ? st0 : s == 1 ? st1 : s == 2 ? st2 : st3;
            ww[s] = (st * pe) * de;
            w[(int64_t)(r * 4 + s) * Lp + i] = ww[s];
          }
        } else {
#pragma unroll
          for (int s = 0; s < 4; ++s) ww[s] = w[(int64_t)(r * 4 + s) * Lp + i];
        }
#pragma unroll
        for (int b = 0; b < NB; ++b)
#pragma unroll
          for (int s = 0; s < 4; ++s) u[b] = fma(ww[s], q_entry(Pqm + b * PM, r * 4 + s, col, mask, single), u[b]);
      }
#pragma unroll
      for (int b = 0; b < NB; ++b) lp[b].mul(u[b]);
    }
#pragma unroll
    for (int b = 0; b < NB; ++b) out[b] = finish(lp[b]);
  }

  __device__ __noinline__ double pend_single_cached() { double o[1]; pass_pend<1, false>(sw + OFF_PQ, o); return o[0]; }
  __device__ __noinline__ double pend_single_fused() { double o[1]; pass_pend<1, true>(sw + OFF_PQ, o); w_stale = false; return o[0]; }
  __device__ __noinline__ double dist_single() { double o[1]; pass_dist<1>(sw + OFF_PD, o); return o[0]; }
  __device__ __noinline__ void pend_batch4(double (&o)[4]) {
    if (w_stale) { pass_pend<4, true>(sw + OFF_PB, o); w_stale = false; } else pass_pend<4, false>(sw + OFF_PB, o);
  }
  __device__ __noinline__ void pend_batch8(double (&o)[8]) {
    if (w_stale) { pass_pend<8, true>(sw + OFF_PB, o); w_stale = false; } else pass_pend<8, false>(sw + OFF_PB, o);
  }
  __device__ __noinline__ void dist_batch4(double (&o)[4]) { pass_dist<4>(sw + OFF_PB, o); }

  // ---- TTDriver engine interface ----
  __device__ double f_pend(double v) {
    set_pend(v);
    double f;
    if (memo_get(0, v, 0.0, &f)) return f;
    return w_stale ? pend_single_fused() : pend_single_cached();
  }
  __device__ double f_dist(double d, double p) {
    double f;
    const bool hit = memo_get(1, d, p, &f);  // distal values stay valid while the pendant length is fixed
    set_dist_prox(d, p);
    return hit ? f : dist_single();
  }
  __device__ void cache_w() {}  // lazy: set_dist_prox marks w stale, the next pendant pass recomputes it (fused)
  __device__ double eval_full_cache() { return w_stale ? pend_single_fused() : pend_single_cached(); }
  __device__ double eval_full_plain() { return dist_single(); }

  __device__ __noinline__ void hint_brent(int kind, double left, double right, double start, double v) {
    if (kind == OBJ_DIST) {
      // replay the drift of cut = dist + prox over the seven calls (set_dist_bl, three_tax.ml:41-45)
      memo_clear(1);
      const double xs[7] = {left, right, start, left, right, start, v};
      const int first[4] = {0, 1, 2, 6};
      double cut = dist_bl + prox_bl, px[7];
#pragma unroll
      for (int i = 0; i < 7; ++i) {
        const double prox = cut - xs[i];
        px[i] = prox;
        if (prox >= 0.0) cut = xs[i] + prox;
      }
      double bls[8], o[4];
#pragma unroll
      for (int b = 0; b < 4; ++b) { bls[2 * b] = xs[first[b]]; bls[2 * b + 1] = px[first[b]]; }
      build_P_n<8>(sw + OFF_PB, PM, bls);
      dist_batch4(o);
#pragma unroll
      for (int b = 0; b < 4; ++b)
        if (px[first[b]] >= 0.0) memo_put(xs[first[b]], px[first[b]], o[b]);
    } else {
      memo_clear(0);
      const double bls[4] = {left, right, start, v};
      double o[4];
      build_P_n<4>(sw + OFF_PB, PM, bls);
      pend_batch4(o);
#pragma unroll
      for (int b = 0; b < 4; ++b) memo_put(bls[b], 0.0, o[b]);
    }
  }
  __device__ void hint_begin() { memo_clear(0); }
  __device__ void hint_add(double x) {
    const int at = memo_n + memo_pending;
    if (at >= MEMO) return;
    __syncwarp();
    if (lane == 0) { sw[OFF_MX + at] = x; sw[OFF_MA + at] = 0.0; }
    __syncwarp();
    ++memo_pending;
  }
  __device__ __noinline__ void hint_flush() {
    while (memo_pending > 0) {
      const int n = memo_pending < NBP ? memo_pending : NBP;
      double bls[NBP], o[NBP];
#pragma unroll
      for (int b = 0; b < NBP; ++b) bls[b] = sw[OFF_MX + memo_n + (b < n ? b : n - 1)];
      build_P_n<NBP>(sw + OFF_PB, PM, bls);
      pend_batch8(o);
      __syncwarp();
#pragma unroll
      for (int b = 0; b < NBP; ++b)
        if (b < n && lane == 0) sw[OFF_MF + memo_n + b] = o[b];
      __syncwarp();
      memo_n += n;
      memo_pending -= n;
    }
  }

  // core.ml:225-239 set_tt_edges (no gather): fix the candidate and its three branch lengths
  __device__ void set_tt_edges(double cut_bl, double pendant, double distal) {
    long long es = 0;
    for (int i = lane; i < nq; i += 32) {
      const int site = qsite[i];
      es += (long long)ebase_d[site] + (long long)ebase_p[site];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) es += __shfl_xor_sync(0xffffffffu, es, o);
    e_tot_ln2 = kLn2 * (double)es;  // gstar_support.ml:43-49
    rate_term = (double)((float)nq) * log((double)((float)R));  // linear_c.c:421
    memo_clear(-1);
    set_pend(pendant);
    set_dist_prox(distal, cut_bl - distal);
  }
};

// Persistent kernel: every warp pulls candidates (in `order`, sorted by edge) until n are done.
template <int R, int WARPS, int MINB>
__global__ void __launch_bounds__(WARPS * 32, MINB)
tt_warp_kernel(const DevModel *__restrict__ m, GlvArrays g, const double *__restrict__ bl, const uint8_t *__restrict__ codes,
               const int32_t *__restrict__ qsite, const int32_t *__restrict__ nq, const CandIn *__restrict__ in,
               const int32_t *__restrict__ order, CandOut *__restrict__ out, int n, TTParams prm, double *__restrict__ wbuf,
               int *counter) {
  using TT = TTWarp<R>;
  extern __shared__ double sm[];
  for (int i = threadIdx.x; i < 4; i += blockDim.x) { sm[TT::C_STATD + i] = m->statd[i]; sm[TT::C_LAM + i] = m->lam[i]; }
  for (int i = threadIdx.x; i < 16; i += blockDim.x) { sm[TT::C_X + i] = m->x[i]; sm[TT::C_XIT + i] = m->xit[i]; }
  for (int i = threadIdx.x; i < R; i += blockDim.x) sm[TT::C_RATES + i] = m->rates[i];
  __syncthreads();
  const int warp = threadIdx.x >> 5;
  TT tt;
  tt.cst = sm;
  tt.sw = sm + ((TT::C_TOTAL + 1) & ~1) + warp * TT::PER_WARP;
  tt.lane = threadIdx.x & 31;
  tt.Lp = g.Lp;
  tt.w = wbuf + (int64_t)(blockIdx.x * WARPS + warp) * TT::K * g.Lp;
  for (;;) {
    int idx = 0;
    if (tt.lane == 0) idx = atomicAdd(counter, 1);
    idx = __shfl_sync(0xffffffffu, idx, 0);
    if (idx >= n) break;
    const int ci = order ? order[idx] : idx;
    const CandIn c = in[ci];
    tt.qsite = qsite + (int64_t)c.q * g.Lp;
    tt.qcode = codes + (int64_t)c.q * g.Lp;
    tt.nq = nq[c.q];
    tt.dbase = g.a[0] + (int64_t)c.loc * g.a_stride;
    tt.pbase = g.a[1] + (int64_t)c.loc * g.a_stride;
    tt.ebase_d = g.e[0] + (int64_t)c.loc * g.Lp;
    tt.ebase_p = g.e[1] + (int64_t)c.loc * g.Lp;
    tt.calls = 0;
    const double cut_bl = bl[c.loc];
    CandOut o;
    o.marg = __longlong_as_double(0x7ff8000000000000LL);
    const double p0 = prm.use_defaults ? prm.start_pend : c.pend0;
    const double d0 = prm.use_defaults ? cut_bl / 2.0 : c.dist0;
    tt.set_tt_edges(cut_bl, p0, d0);
    int st = ST_OK;
    if (prm.mode == TT_MODE_OPTIMIZE) {
      // core.ml:254-279: ExceededMaxIter keeps the current state; a GSL error retries once from the
      // default branch lengths (tt_edges_default, core.ml:240-245).
      st = tt.optimize(prm.tol, prm.max_pend, prm.max_iter);
      if (st_is_gsl(st)) {
        tt.set_tt_edges(cut_bl, prm.start_pend, cut_bl / 2.0);
        st = tt.optimize(prm.tol, prm.max_pend, prm.max_iter);
      }
      if (st == ST_EXCEEDED_MAX_ITER) st = ST_OK;
      o.ll = tt.eval_full_plain();  // Three_tax.get_results (three_tax.ml:85)
    } else if (prm.mode == TT_MODE_LOGLIKE) {
      o.ll = tt.eval_full_plain();
    } else {
      tt.prior.kind = prm.prior_kind;
      tt.prior.mean = prm.prior_kind == 2 ? prm.prior_means[c.loc] : prm.prior_mean;
      double mp = __longlong_as_double(0x7ff8000000000000LL);
      st = tt.calc_marg_prob(prm.pp_rel_err, prm.max_pend, &mp);
      o.marg = mp;
      o.ll = mp;
    }
    o.pend = tt.pend_bl;
    o.dist = tt.dist_bl;
    o.status = st;
    o.calls = tt.calls;
    if (tt.lane == 0) out[ci] = o;
  }
}

// queue order: candidate indices sorted by edge (stable), so neighbours share GLV planes
__global__ void cand_keys_kernel(const CandIn *__restrict__ in, int n, int32_t *__restrict__ keys, int32_t *__restrict__ vals) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { keys[i] = in[i].loc; vals[i] = i; }
}

}  // namespace ppb
