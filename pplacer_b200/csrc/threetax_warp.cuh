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

// running product of per-site likelihoods as (mantissa in [1,2), exponent); non-finite logs go to
// a per-point shared slot (see LogProd in device_common.cuh for the idea).  Subnormal factors are
// rescaled into the normal path, so the slot only ever receives -inf / +inf / NaN, whose sum does
// not depend on the order of the additions.
struct WarpProd {
  double m;
  int e;
  __device__ __forceinline__ void init() { m = 1.0; e = 0; }
  // zero / negative / NaN / inf / subnormal factor (rare): subnormals are rescaled exactly, the
  // rest add log(u) = -inf / NaN / +inf to the point's shared slot and leave the product alone
  __device__ __noinline__ static double special(double u, double *bad_slot, int *eadj) {
    if (u > 0.0 && u < 2.2250738585072014e-308) { *eadj = -600; return u * 0x1p600; }
    const double inf = __longlong_as_double(0x7ff0000000000000LL);
    atomicAdd(bad_slot, u == 0.0 ? -inf : (u == inf ? inf : __longlong_as_double(0x7ff8000000000000LL)));  // linear_c.c:418 sums log(util)
    *eadj = 0;
    return 1.0;
  }
  __device__ __forceinline__ void mul(double u, double *bad_slot) {
    const int hi = __double2hiint(u);
    const int ef = (hi >> 20) & 0x7ff;
    if (hi < 0 || ef == 0 || ef == 0x7ff) {
      int adj;
      u = special(u, bad_slot, &adj);
      e += adj;
    }
    const double p = m * u;  // m in [1,2), u normal  =>  p normal, no overflow
    const int phi = __double2hiint(p);
    e += ((phi >> 20) & 0x7ff) - 1023;
    m = __hiloint2double((phi & 0x800fffff) | 0x3ff00000, __double2loint(p));
  }
};

extern __shared__ double2 ppb_ttw_smem[];

// TS = sites per lane per iteration in the one-point passes of the optimiser
template <int R, int TS>
struct TTWarp : TTDriver<TTWarp<R, TS>>, TTMachine<TTWarp<R, TS>> {
  using Base = TTDriver<TTWarp<R, TS>>;
  using Mach = TTMachine<TTWarp<R, TS>>;
  using Base::pend_bl; using Base::dist_bl; using Base::prox_bl; using Base::calls;
  using Mach::opt_init; using Mach::opt_advance; using Mach::b_kind; using Mach::req_x; using Mach::o_status;
  static constexpr int S = 4, K = 4 * R, PM = 16 * R;
  static constexpr int PQS = 5 * K;   // a pendant matrix set: rows [R*4][5] = P[rs][0..3] then QN[rs] (row sum, for 'N')
  static constexpr int NBP = 6;    // pendant points per pass
  static constexpr int NBD = 4;    // distal points per pass
  static constexpr int MEMO = 24;  // QNG's 21- and 22-point rule extensions (the 44-point one is batched in part)
  static constexpr unsigned FULL = 0xffffffffu;
  // block-shared constants (doubles): statd[4] lam[4] x[16] xit[16] rates[R]
  static constexpr int C_STATD = 0, C_LAM = 4, C_X = 8, C_XIT = 24, C_RATES = 40, C_TOTAL = (40 + R + 1) & ~1;
  // per-warp shared layout (doubles; every offset even so that double2 loads are 16-byte aligned)
  static constexpr int OFF_EX = 0;                 // [8][K] exp scratch
  static constexpr int OFF_BL = OFF_EX + 8 * K;    // [8] branch lengths handed to build_P
  static constexpr int OFF_OUT = OFF_BL + 8;       // [8] pass results
  static constexpr int OFF_BAD = OFF_OUT + 8;      // [8] non-finite log terms per point
  static constexpr int OFF_PQ = OFF_BAD + 8;       // current pendant set (PQS)
  static constexpr int OFF_PD = OFF_PQ + PQS;      // current distal P, then proximal P
  static constexpr int OFF_PB = OFF_PD + 2 * PM;   // batch: pendant [NBP][PQS] | distal [NBD][2][PM]
  static constexpr int PB_SIZE = NBP * PQS > NBD * 2 * PM ? NBP * PQS : NBD * 2 * PM;
  static constexpr int OFF_MX = OFF_PB + PB_SIZE;  // memo: argument
  static constexpr int OFF_MA = OFF_MX + MEMO;     //       proximal length (distal memo) or 0
  static constexpr int OFF_MF = OFF_MA + MEMO;     //       log-likelihood
  static constexpr int OFF_OBJ = OFF_MF + MEMO;     // the TTWarp object itself
  static constexpr int OBJ_DOUBLES = 64;
  static constexpr int PER_WARP = OFF_OBJ + OBJ_DOUBLES;
  static_assert(NBD * 2 <= 8 && NBP <= 8, "scratch holds eight branch lengths");
  static_assert((PQS % 2) == 0 && (PM % 2) == 0 && (MEMO % 2) == 0 && (K % 2) == 0, "keep double2 alignment");

  // The object itself lives in this warp's shared region (every member is warp-uniform; all lanes
  // store identical values), so neither `this` nor the optimiser state ever touches local memory.
  int sw_off;  // this warp's region, in doubles from the start of dynamic shared memory
  // candidate
  const double *dbase, *pbase;
  const int32_t *ebase_d, *ebase_p;
  const int32_t *qsite;
  double *w;
  int64_t Lp;
  int nq;
  double e_tot_ln2, rate_term;
  bool w_stale;
  int flags;  // TTParams.flags
  int calls_dist;  // distal-objective calls among `calls`
  int memo_n, memo_pending, memo_kind;  // memo_kind: -1 none, 0 pendant values, 1 distal values

  static __host__ __device__ constexpr size_t smem_bytes(int warps) { return sizeof(double) * (size_t)(C_TOTAL + warps * PER_WARP); }

  // shared-memory views (rebuilt from the extern symbol so that the compiler emits LDS/STS)
  __device__ __forceinline__ static const double *CST() { return reinterpret_cast<const double *>(ppb_ttw_smem); }
  __device__ __forceinline__ double *SW() const { return reinterpret_cast<double *>(ppb_ttw_smem) + sw_off; }
  __device__ __forceinline__ static int lane_id() { return threadIdx.x & 31; }
  __device__ __forceinline__ static void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

  __device__ __forceinline__ static double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
  }

  // Diagd.multi_exp (diagd.ml:79-98; linear_c.c:92-146) for the n branch lengths in SW()[OFF_BL..]:
  // dst + b*stride receives P(bl_b) and, when qn, the row sums ((P0+P1)+P2)+P3 used for the all-ones
  // character 'N' (the evolved indicator vector of an ambiguity code is the sum of the selected
  // columns in ascending state order: lv_of_nuc, nuc_models.ml:49-54 through gemmish_c, linear_c.c:29-90).
  __device__ __noinline__ void build_P(int dst_off, int stride, int n, bool qn) {
    double *s = SW();
    const double *c = CST();
    double *ex = s + OFF_EX, *dst = s + dst_off;
    __syncwarp();
    for (int t = lane_id(); t < n * K; t += 32) {
      const int b = t / K, kk = t - b * K;
      ex[t] = exp(c[C_RATES + (kk >> 2)] * s[OFF_BL + b] * c[C_LAM + (kk & 3)]);
    }
    __syncwarp();
    for (int t = lane_id(); t < n * PM; t += 32) {
      const int b = t / PM, e = t - b * PM;
      const int r = e >> 4, i = (e >> 2) & 3, j = e & 3;
      double acc = 0.0;
#pragma unroll
      for (int k = 0; k < 4; ++k) acc = fma(ex[b * K + r * 4 + k] * c[C_X + i * 4 + k], c[C_XIT + j * 4 + k], acc);
      dst[b * stride + (qn ? (e >> 2) * 5 + (e & 3) : e)] = acc;
    }
    __syncwarp();
    if (qn) {
      for (int t = lane_id(); t < n * K; t += 32) {
        const int b = t / K, rs = t - b * K;
        double *row = dst + b * stride + rs * 5;
        row[4] = (((0.0 + row[0]) + row[1]) + row[2]) + row[3];
      }
      __syncwarp();
    }
  }
  __device__ __forceinline__ void put_bl(int b, double v) { if (lane_id() == 0) SW()[OFF_BL + b] = v; }

  __device__ void memo_clear(int kind) { memo_n = 0; memo_pending = 0; memo_kind = kind; }
  __device__ void memo_put(double x, double aux, double f) {
    if (memo_n >= MEMO) return;
    double *s = SW();
    __syncwarp();
    if (lane_id() == 0) { s[OFF_MX + memo_n] = x; s[OFF_MA + memo_n] = aux; s[OFF_MF + memo_n] = f; }
    __syncwarp();
    ++memo_n;
  }
  __device__ bool memo_get(int kind, double x, double aux, double *f) const {
    if (memo_kind != kind) return false;
    const double *s = SW();
    for (int i = 0; i < memo_n; ++i)
      if (s[OFF_MX + i] == x && s[OFF_MA + i] == aux) { *f = s[OFF_MF + i]; return true; }
    return false;
  }

  // Glv_edge.set_bl (glv_edge.ml:33-35) on the three edges: only the rate matrices change here
  __device__ void set_pend(double bl) {
    pend_bl = bl;
    __syncwarp();
    put_bl(0, bl);
    build_P(OFF_PQ, PQS, 1, true);
    if (memo_kind == 1) memo_kind = -1;  // distal values were computed for another pendant length
  }
  __device__ void set_dist_prox(double d, double p) {
    dist_bl = d; prox_bl = p;
    __syncwarp();
    put_bl(0, d); put_bl(1, p);
    build_P(OFF_PD, PM, 2, false);
    w_stale = true;
    if (memo_kind == 0) memo_kind = -1;
  }

  // Per-row (32 sites) query-character decode, uniform across the warp.  The evolved pendant GLV
  // entry for state row rs is P[rs][col] for a plain base, QN[rs] for 'N' (fast: one shared load at
  // set[rs*5 + qcol]); any other ambiguity code in the row sends the whole row down the slow path
  // (sum of the selected columns in ascending order - exact for plain bases too, since 0 + x = x).
  struct RowCodes {
    uint32_t mask;
    int qcol;
    bool fast;
  };
  __device__ __forceinline__ static RowCodes decode(int packed) {
    RowCodes c;
    c.mask = ((uint32_t)packed >> kSiteCodeShift) & 0x7f;
    const bool single = (c.mask & (c.mask - 1)) == 0;
    c.qcol = single ? __ffs(c.mask) - 1 : 4;
    c.fast = __all_sync(FULL, single || c.mask == 15u);
    return c;
  }
  __device__ __noinline__ static double q_slow(const double *set, int rs, uint32_t mask) {
    const double *row = set + rs * 5;
    double acc = 0.0;
#pragma unroll
    for (int j = 0; j < 4; ++j) acc = ((mask >> j) & 1u) ? acc + row[j] : acc;
    return acc;
  }

  // natural log of the accumulated product, summed over the warp, plus the constant terms
  __device__ __noinline__ double finish(double m, int e, int b) {
    const double v = warp_sum(log(m) + (double)e * kLn2);
    return (v + SW()[OFF_BAD + b]) - rate_term + e_tot_ln2;
  }

  // de = P_d d, pe = P_p p, ww[s] = (statd_s * pe_s) * de_s   (log_like3's ((statd*x)*y), x = prox,
  // y = dist: three_tax.ml:31-36, linear_c.c:367-424) for one rate category and T sites per lane:
  // every matrix entry fetched from shared memory serves T sites.
  template <int T>
  __device__ __forceinline__ static void site_w(const double *Pd, const double *Pp, const double (&d)[T][4], const double (&p)[T][4],
                                                const double *st, double (&ww)[T][4]) {
    const double2 *Pd2 = reinterpret_cast<const double2 *>(Pd), *Pp2 = reinterpret_cast<const double2 *>(Pp);
#pragma unroll
    for (int s = 0; s < 4; ++s) {
      const double2 a0 = Pd2[s * 2], a1 = Pd2[s * 2 + 1], b0 = Pp2[s * 2], b1 = Pp2[s * 2 + 1];
      const double sts = st[s];
#pragma unroll
      for (int t = 0; t < T; ++t) {
        double de = fma(a0.x, d[t][0], 0.0), pe = fma(b0.x, p[t][0], 0.0);
        de = fma(a0.y, d[t][1], de); pe = fma(b0.y, p[t][1], pe);
        de = fma(a1.x, d[t][2], de); pe = fma(b1.x, p[t][2], pe);
        de = fma(a1.y, d[t][3], de); pe = fma(b1.y, p[t][3], pe);
        ww[t][s] = (sts * pe) * de;
      }
    }
  }

  __device__ __forceinline__ void pass_begin() {
    double *s = SW();
    __syncwarp();
    if (lane_id() < 8) s[OFF_BAD + lane_id()] = 0.0;
    __syncwarp();
  }
  template <int NB>
  __device__ __forceinline__ void pass_end(const WarpProd (&lp)[NB]) {
    double *s = SW();
    __syncwarp();
#pragma unroll
    for (int b = 0; b < NB; ++b) {
      const double v = finish(lp[b].m, lp[b].e, b);
      if (lane_id() == 0) s[OFF_OUT + b] = v;
    }
    __syncwarp();
  }

  // Row bookkeeping shared by the passes: each iteration handles T rows of 32 sites (lane l owns
  // sites base + t*32 + l); the packed (site, code) words of the next iteration are loaded one
  // iteration ahead, and the planes two iterations ahead are pulled towards L2 (flags bit 1).
  // Three_tax.log_like at NB (distal, proximal) pairs with the current pendant matrices: the set at
  // pm_off holds NB x {P(distal), P(proximal)}.  One pass over darr/parr planes at the query's sites;
  // results go to SW()[OFF_OUT + b].
  template <int NB, int T>
  __device__ __forceinline__ void pass_dist(int pm_off) {
    pass_begin();
    double *s = SW();
    const double *Pm = s + pm_off, *Pq = s + OFF_PQ, *st = CST() + C_STATD;
    WarpProd lp[NB];
#pragma unroll
    for (int b = 0; b < NB; ++b) lp[b].init();
    const int lane = lane_id();
    int next[T];
#pragma unroll
    for (int t = 0; t < T; ++t) next[t] = qsite[min(t * 32 + lane, nq - 1)];
    const bool pf = (flags & 2) != 0;
    const double *pfp = (lane < 16 ? dbase : pbase) + (int64_t)(lane & 15) * Lp;  // lane -> plane, for prefetch
    const int pf_max = (int)Lp - 16;
    for (int base = 0; base < nq; base += 32 * T) {
      int packed[T];
      bool act[T], fast = true;
      const double *dp[T], *pp[T], *Pqc[T];
      uint32_t mask[T];
#pragma unroll
      for (int t = 0; t < T; ++t) {
        const int i = base + t * 32 + lane;
        act[t] = i < nq;
        packed[t] = next[t];
        next[t] = qsite[min(i + 32 * T, nq - 1)];
        const RowCodes c = decode(packed[t]);
        fast = fast && c.fast;
        mask[t] = c.mask;
        Pqc[t] = Pq + c.qcol;
        const int site = packed[t] & kSiteMask;
        dp[t] = dbase + site; pp[t] = pbase + site;
      }
      if (pf) {
        const int s0 = __shfl_sync(FULL, packed[0] & kSiteMask, 0) + 64 * T;
#pragma unroll
        for (int h = 0; h < 2 * T; ++h) prefetch_l2(pfp + min(s0 + 16 * h, pf_max));
      }
      double u[T][NB];
#pragma unroll
      for (int t = 0; t < T; ++t)
#pragma unroll
        for (int b = 0; b < NB; ++b) u[t][b] = 0.0;
#pragma unroll 1
      for (int r = 0; r < R; ++r) {
        double d[T][4], p[T][4], q[T][4];
#pragma unroll
        for (int t = 0; t < T; ++t)
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            d[t][k] = __ldg(dp[t] + (int64_t)(r * 4 + k) * Lp);
            p[t][k] = __ldg(pp[t] + (int64_t)(r * 4 + k) * Lp);
          }
        if (fast) {
#pragma unroll
          for (int t = 0; t < T; ++t)
#pragma unroll
            for (int k = 0; k < 4; ++k) q[t][k] = Pqc[t][(r * 4 + k) * 5];
        } else {
#pragma unroll
          for (int t = 0; t < T; ++t)
#pragma unroll
            for (int k = 0; k < 4; ++k) q[t][k] = q_slow(Pq, r * 4 + k, mask[t]);
        }
#pragma unroll
        for (int b = 0; b < NB; ++b) {
          double ww[T][4];
          site_w<T>(Pm + b * 2 * PM + r * 16, Pm + b * 2 * PM + PM + r * 16, d, p, st, ww);
#pragma unroll
          for (int t = 0; t < T; ++t)
#pragma unroll
            for (int k = 0; k < 4; ++k) u[t][b] = fma(ww[t][k], q[t][k], u[t][b]);
        }
      }
#pragma unroll
      for (int t = 0; t < T; ++t)
        if (act[t]) {
#pragma unroll
          for (int b = 0; b < NB; ++b) lp[b].mul(u[t][b], s + OFF_BAD + b);
        }
    }
    pass_end<NB>(lp);
  }

  // log_like at NB pendant lengths (sets at pq_off + b*PQS) with the current distal / proximal
  // matrices.  MODE 0: read the cached w; 1: recompute w from darr/parr and store it; 2: recompute,
  // no store.  w is streamed (evict-first) so that it does not push the shared darr/parr planes out of L2.
  template <int NB, int MODE, int T>
  __device__ __forceinline__ void pass_pend(int pq_off) {
    constexpr bool FUSED = MODE != 0;
    pass_begin();
    double *s = SW();
    const double *Pqm = s + pq_off, *Pd0 = s + OFF_PD, *st = CST() + C_STATD;
    WarpProd lp[NB];
#pragma unroll
    for (int b = 0; b < NB; ++b) lp[b].init();
    const int lane = lane_id();
    int next[T];
#pragma unroll
    for (int t = 0; t < T; ++t) next[t] = qsite[min(t * 32 + lane, nq - 1)];
    const bool pf = (flags & 2) != 0;
    const double *pfp = FUSED ? (lane < 16 ? dbase : pbase) + (int64_t)(lane & 15) * Lp
                              : w + (int64_t)(lane & 15) * Lp + (lane >> 4) * 16;
    const int pf_max = (int)Lp - 16;
    for (int base = 0; base < nq; base += 32 * T) {
      int packed[T];
      bool act[T], fast = true;
      const double *dp[T], *pp[T], *Pqc[T];
      double *wp[T];
      uint32_t mask[T];
#pragma unroll
      for (int t = 0; t < T; ++t) {
        const int i = base + t * 32 + lane;
        act[t] = i < nq;
        packed[t] = next[t];
        next[t] = qsite[min(i + 32 * T, nq - 1)];
        const RowCodes c = decode(packed[t]);
        fast = fast && c.fast;
        mask[t] = c.mask;
        Pqc[t] = Pqm + c.qcol;
        const int site = packed[t] & kSiteMask;
        dp[t] = dbase + site; pp[t] = pbase + site;
        wp[t] = w + (act[t] ? i : nq - 1);
      }
      if (pf) {
        if (FUSED) {
          const int s0 = __shfl_sync(FULL, packed[0] & kSiteMask, 0) + 64 * T;
#pragma unroll
          for (int h = 0; h < 2 * T; ++h) prefetch_l2(pfp + min(s0 + 16 * h, pf_max));
        } else {
#pragma unroll
          for (int h = 0; h < T; ++h) prefetch_l2(pfp + min(base + 64 * T + 32 * h, pf_max));
        }
      }
      double u[T][NB];
#pragma unroll
      for (int t = 0; t < T; ++t)
#pragma unroll
        for (int b = 0; b < NB; ++b) u[t][b] = 0.0;
#pragma unroll 1
      for (int r = 0; r < R; ++r) {
        double ww[T][4];
        if (FUSED) {
          double d[T][4], p[T][4];
#pragma unroll
          for (int t = 0; t < T; ++t)
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              d[t][k] = __ldg(dp[t] + (int64_t)(r * 4 + k) * Lp);
              p[t][k] = __ldg(pp[t] + (int64_t)(r * 4 + k) * Lp);
            }
          site_w<T>(Pd0 + r * 16, Pd0 + PM + r * 16, d, p, st, ww);
          if (MODE == 1) {
#pragma unroll
            for (int t = 0; t < T; ++t)
              if (act[t]) {
#pragma unroll
                for (int k = 0; k < 4; ++k) __stcs(wp[t] + (int64_t)(r * 4 + k) * Lp, ww[t][k]);
              }
          }
        } else {
#pragma unroll
          for (int t = 0; t < T; ++t)
#pragma unroll
            for (int k = 0; k < 4; ++k) ww[t][k] = __ldcs(wp[t] + (int64_t)(r * 4 + k) * Lp);
        }
        if (fast) {
#pragma unroll
          for (int b = 0; b < NB; ++b)
#pragma unroll
            for (int t = 0; t < T; ++t)
#pragma unroll
              for (int k = 0; k < 4; ++k) u[t][b] = fma(ww[t][k], Pqc[t][b * PQS + (r * 4 + k) * 5], u[t][b]);
        } else {
#pragma unroll
          for (int b = 0; b < NB; ++b)
#pragma unroll
            for (int t = 0; t < T; ++t)
#pragma unroll
              for (int k = 0; k < 4; ++k) u[t][b] = fma(ww[t][k], q_slow(Pqm + b * PQS, r * 4 + k, mask[t]), u[t][b]);
        }
      }
#pragma unroll
      for (int t = 0; t < T; ++t)
        if (act[t]) {
#pragma unroll
          for (int b = 0; b < NB; ++b) lp[b].mul(u[t][b], s + OFF_BAD + b);
        }
    }
    pass_end<NB>(lp);
  }

  __device__ __noinline__ double pend_single_cached() { pass_pend<1, 0, 1>(OFF_PQ); return SW()[OFF_OUT]; }
  __device__ __noinline__ double pend_single_fused() { pass_pend<1, 1, 1>(OFF_PQ); w_stale = false; return SW()[OFF_OUT]; }
  __device__ __noinline__ double dist_single() { pass_dist<1, 1>(OFF_PD); return SW()[OFF_OUT]; }
  __device__ __noinline__ void pend_batch_cached() { pass_pend<NBP, 0, 1>(OFF_PB); }
  __device__ __noinline__ void pend_batch_fused() { pass_pend<NBP, 1, 1>(OFF_PB); w_stale = false; }
  __device__ void pend_batch() { if (w_stale) pend_batch_fused(); else pend_batch_cached(); }
  __device__ __noinline__ void dist_batch() { pass_dist<NBD, 1>(OFF_PB); }

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

  // Brent's first calls are known in advance: f(left), f(right), f(start), GSL's repeats of the same
  // three, f(lower + golden*(upper - lower)) and - for the pendant objective - the first golden-section
  // step from (z = start, d = e = 0).  hint_prepare builds their matrices and memo keys, the caller
  // runs ONE batched pass (run_batch), hint_commit stores the values.
  __device__ void hint_prepare(int kind, double left, double right, double start, double v) {
    double *s = SW();
    if (kind == OBJ_DIST) {
      // replay the drift of cut = dist + prox over the seven calls (set_dist_bl, three_tax.ml:41-45);
      // a point whose proximal length would be negative keeps a negative key and never matches
      memo_clear(1);
      const double xs[7] = {left, right, start, left, right, start, v};
      double cut = dist_bl + prox_bl, px[7];
#pragma unroll
      for (int i = 0; i < 7; ++i) {
        const double prox = cut - xs[i];
        px[i] = prox;
        if (prox >= 0.0) cut = xs[i] + prox;
      }
      const double bx[4] = {xs[0], xs[1], xs[2], xs[6]}, bp[4] = {px[0], px[1], px[2], px[6]};
      __syncwarp();
      if (lane_id() == 0) {
#pragma unroll
        for (int b = 0; b < 4; ++b) { s[OFF_BL + 2 * b] = bx[b]; s[OFF_BL + 2 * b + 1] = bp[b]; s[OFF_MX + b] = bx[b]; s[OFF_MA + b] = bp[b]; }
      }
      memo_pending = 4;
      build_P(OFF_PB, PM, 8, false);
    } else {
      memo_clear(0);
      const double mid = 0.5 * (left + right);
      const double ee = (start < mid) ? right - start : -(start - left);
      const double u1 = start + 0.3819660 * ee;
      const double bls[5] = {left, right, start, v, u1};
      __syncwarp();
      if (lane_id() == 0) {
#pragma unroll
        for (int b = 0; b < NBP; ++b) { s[OFF_BL + b] = bls[b < 5 ? b : 4]; if (b < 5) { s[OFF_MX + b] = bls[b]; s[OFF_MA + b] = 0.0; } }
      }
      memo_pending = 5;
      build_P(OFF_PB, PQS, NBP, true);
    }
  }
  __device__ void hint_commit() {
    double *s = SW();
    __syncwarp();
    if (lane_id() < memo_pending) s[OFF_MF + lane_id()] = s[OFF_OUT + lane_id()];
    __syncwarp();
    memo_n = memo_pending;
    memo_pending = 0;
  }
  __device__ void hint_brent(int kind, double left, double right, double start, double v) {  // TTDriver path (--pp)
    hint_prepare(kind, left, right, start, v);
    if (kind == OBJ_DIST) dist_batch(); else pend_batch();
    hint_commit();
  }
  __device__ void hint_begin() { memo_clear(0); }
  __device__ void hint_add(double x) {
    const int at = memo_n + memo_pending;
    if (at >= MEMO) return;
    double *s = SW();
    __syncwarp();
    if (lane_id() == 0) { s[OFF_MX + at] = x; s[OFF_MA + at] = 0.0; }
    __syncwarp();
    ++memo_pending;
  }
  __device__ __noinline__ void hint_flush() {
    double *s = SW();
    while (memo_pending > 0) {
      const int n = memo_pending < NBP ? memo_pending : NBP;
      __syncwarp();
      if (lane_id() < NBP) s[OFF_BL + lane_id()] = s[OFF_MX + memo_n + (lane_id() < n ? lane_id() : n - 1)];
      build_P(OFF_PB, PQS, NBP, true);
      pend_batch();
      if (lane_id() < n) s[OFF_MF + memo_n + lane_id()] = s[OFF_OUT + lane_id()];
      __syncwarp();
      memo_n += n;
      memo_pending -= n;
    }
  }


  // One (query, edge) candidate in TT_MODE_OPTIMIZE / TT_MODE_LOGLIKE: Three_tax.optimize through the
  // state machine, the retry from the default branch lengths after a GSL error (core.ml:254-279,
  // tt_edges_default :240-245), then Three_tax.get_results (three_tax.ml:85).  Every likelihood pass
  // is inlined exactly once, here.
  enum : int { OBJ_FINAL = 3 };
  __device__ __forceinline__ int run_candidate(bool optimise, double cut_bl, const TTParams &prm, double *ll_out) {
    int st = ST_OK, feed = 0, attempt = 0;
    double f = 0.0;
    if (optimise) opt_init(prm.tol, prm.max_pend, prm.max_iter);
    for (;;) {
      const int more = optimise ? opt_advance(feed, f) : 0;
      if (more == 2) {  // Brent's predictable calls, several per pass
        if (memo_kind == 1) pass_dist<NBD, 1>(OFF_PB);
        else if (w_stale) { pass_pend<NBP, 1, 1>(OFF_PB); w_stale = false; }
        else pass_pend<NBP, 0, 1>(OFF_PB);
        feed = 2;
        continue;
      }
      int kind = OBJ_FINAL;
      if (more == 1) kind = b_kind;
      else if (optimise) {
        st = o_status;
        if (st_is_gsl(st) && attempt == 0) {
          attempt = 1;
          set_tt_edges(cut_bl, prm.start_pend, cut_bl / 2.0);
          opt_init(prm.tol, prm.max_pend, prm.max_iter);
          feed = 0;
          continue;
        }
        if (st == ST_EXCEEDED_MAX_ITER) st = ST_OK;
      }
      // ---- the evaluation site: TTDriver::objective with f_pend / f_dist inlined ----
      const double x = req_x;
      double val = 0.0;
      bool have = false, nan_out = false;
      if (kind == OBJ_PEND) {
        calls++;
        set_pend(x);
        have = memo_get(0, x, 0.0, &val);
        if (!have) {
          if (w_stale) { pass_pend<1, 1, TS>(OFF_PQ); w_stale = false; }
          else pass_pend<1, 0, TS>(OFF_PQ);
        }
      } else {
        if (kind == OBJ_DIST) {
          calls++;
          calls_dist++;
          const double prox = (dist_bl + prox_bl) - x;
          if (!(prox >= 0.0)) nan_out = true;
          else {
            have = memo_get(1, x, prox, &val);
            set_dist_prox(x, prox);
          }
        }
        if (!have && !nan_out) pass_dist<1, TS>(OFF_PD);
      }
      if (!have) val = SW()[OFF_OUT];
      if (more == 0) { *ll_out = val; return st; }
      f = nan_out ? __longlong_as_double(0x7ff8000000000000LL) : -val;
      feed = 1;
    }
  }

  // core.ml:225-239 set_tt_edges (no gather): fix the candidate and its three branch lengths
  __device__ __noinline__ void set_tt_edges(double cut_bl, double pendant, double distal) {
    long long es = 0;
    for (int i = lane_id(); i < nq; i += 32) {
      const int site = qsite[i] & kSiteMask;
      es += (long long)ebase_d[site] + (long long)ebase_p[site];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) es += __shfl_xor_sync(FULL, es, o);
    e_tot_ln2 = kLn2 * (double)es;  // gstar_support.ml:43-49
    rate_term = (double)((float)nq) * log((double)((float)R));  // linear_c.c:421
    memo_clear(-1);
    set_pend(pendant);
    set_dist_prox(distal, cut_bl - distal);
  }
};

// Persistent kernel: every warp pulls candidates (in `order`, sorted by edge) until n are done.
template <int R, int WARPS, int MINB, int TS>
__global__ void __launch_bounds__(WARPS * 32, MINB)
tt_warp_kernel(const DevModel *__restrict__ m, GlvArrays g, const double *__restrict__ bl, const uint8_t *__restrict__ codes,
               const int32_t *__restrict__ qsite, const int32_t *__restrict__ nq, const CandIn *__restrict__ in,
               const int32_t *__restrict__ order, CandOut *__restrict__ out, int n, TTParams prm, double *__restrict__ wbuf,
               int *counter, unsigned long long *__restrict__ work) {
  using TT = TTWarp<R, TS>;
  double *sm = reinterpret_cast<double *>(ppb_ttw_smem);
  for (int i = threadIdx.x; i < 4; i += blockDim.x) { sm[TT::C_STATD + i] = m->statd[i]; sm[TT::C_LAM + i] = m->lam[i]; }
  for (int i = threadIdx.x; i < 16; i += blockDim.x) { sm[TT::C_X + i] = m->x[i]; sm[TT::C_XIT + i] = m->xit[i]; }
  for (int i = threadIdx.x; i < R; i += blockDim.x) sm[TT::C_RATES + i] = m->rates[i];
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  static_assert(sizeof(TT) <= TT::OBJ_DOUBLES * sizeof(double), "TTWarp object does not fit its shared slot");
  TT &tt = *reinterpret_cast<TT *>(sm + TT::C_TOTAL + warp * TT::PER_WARP + TT::OFF_OBJ);
  tt.sw_off = TT::C_TOTAL + warp * TT::PER_WARP;
  tt.Lp = g.Lp;
  tt.flags = prm.flags;
  tt.w = wbuf + (int64_t)(blockIdx.x * WARPS + warp) * TT::K * g.Lp;
  for (;;) {
    int idx = 0;
    if (lane == 0) idx = atomicAdd(counter, 1);
    idx = __shfl_sync(0xffffffffu, idx, 0);
    if (idx >= n) break;
    const int ci = order ? order[idx] : idx;
    const CandIn c = in[ci];
    tt.qsite = qsite + (int64_t)c.q * g.Lp;
    tt.nq = nq[c.q];
    tt.dbase = g.a[0] + (int64_t)c.loc * g.a_stride;
    tt.pbase = g.a[1] + (int64_t)c.loc * g.a_stride;
    tt.ebase_d = g.e[0] + (int64_t)c.loc * g.Lp;
    tt.ebase_p = g.e[1] + (int64_t)c.loc * g.Lp;
    tt.calls = 0;
    tt.calls_dist = 0;
    const double cut_bl = bl[c.loc];
    CandOut o;
    o.marg = __longlong_as_double(0x7ff8000000000000LL);
    const double p0 = prm.use_defaults ? prm.start_pend : c.pend0;
    const double d0 = prm.use_defaults ? cut_bl / 2.0 : c.dist0;
    tt.set_tt_edges(cut_bl, p0, d0);
    int st = ST_OK;
    if (prm.mode != TT_MODE_MARGINAL) {
      double ll;
      st = tt.run_candidate(prm.mode == TT_MODE_OPTIMIZE, cut_bl, prm, &ll);
      o.ll = ll;
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
    if (lane == 0) {
      out[ci] = o;
      const unsigned long long nd = tt.calls_dist, np = tt.calls - tt.calls_dist, ns = tt.nq;
      atomicAdd(work + 0, np); atomicAdd(work + 1, nd); atomicAdd(work + 2, np * ns); atomicAdd(work + 3, nd * ns);
    }
  }
}

// queue order: longest queries first (four length classes - a persistent warp that draws a long
// candidate last would otherwise set the launch time), and by edge within a class (stable), so
// that candidates running at the same time share darr/parr planes in L2
__global__ void cand_keys_kernel(const CandIn *__restrict__ in, const int32_t *__restrict__ nq, int n, int L, int loc_bits,
                                 int32_t *__restrict__ keys, int32_t *__restrict__ vals) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const CandIn c = in[i];
  const int cls = 3 - min(3, (4 * nq[c.q]) / (L + 1));
  keys[i] = (cls << loc_bits) | c.loc;
  vals[i] = i;
}

}  // namespace ppb
