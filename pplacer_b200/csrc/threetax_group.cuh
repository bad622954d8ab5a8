// Second stage, block-per-candidate engine with the cached product IN SHARED MEMORY, for the
// four-state Gamma-mixture model (Three_tax.optimize / log_like only; --pp stays on the warp engine).
//   Three_tax.log_like / optimize   pplacer_src/three_tax.ml:30-83
//   Model.log_like3    gmix_model.ml:298-304 -> ten_log_like3_c (linear_c.c:367-424)
//   Model.evolve_into  gmix_model.ml:308-315 ; Diagd.multi_exp diagd.ml:79-98
//
// Why: three out of four objective calls move only the pendant length and need nothing but
// w = (statd*prox_ev)*dist_ev (16 doubles per site).  With one warp per candidate (threetax_warp.cuh)
// thousands of candidates are in flight, their w copies (250 MB) miss L2 and every pendant call
// streams 128 B/site from HBM - measured 2.9 TB/s, the bound of that engine.  Here ONE THREAD BLOCK
// (8 warps) owns a candidate: a thread owns the sites tid, tid+256, ..., w for those sites sits in
// the block's shared memory (each thread reads back only what it wrote itself), and a pendant call
// touches no global memory at all.  Only ~300 candidates are in flight, so the darr/parr planes they
// gather (queue ordered by edge) stay in L2.
//
// Division of labour inside the block, per objective call:
//   warp 0      runs the scalar optimiser (TTMachine, threetax.cuh) up to the next call that needs
//               a pass over the sites - calls answered by the memo of batched values cost no barrier
//               at all - and posts a COMMAND: which pass, and which rate-matrix sets to (re)build
//   all warps   build the posted matrix sets cooperatively (one thread per entry), run the pass,
//               and leave one partial sum per warp; warp 0 adds them in a fixed order.
// Four block barriers per pass: command, exponentials, matrices, partial sums.
//
// Redundant execution: warp 0's scalar state lives in ONE shared-memory object that all 32 lanes read
// and write with identical values while executing the same instruction converged (no lane-dependent
// branch exists in that code; __syncwarp() at the seams).  compute-sanitizer racecheck reports exactly
// these same-value write/write pairs (profiles/README.md, r2n); memcheck is clean.
#pragma once
#include "threetax_warp.cuh"

namespace ppb {

struct GCand {  // per-candidate constants every thread keeps in registers
  const double *dbase, *pbase;
  const int32_t *qsite;
  double *wg;  // global overflow of w for sites >= WS
  int64_t Lp;
  int nq;
};

// G warps per candidate, TS sites per thread per iteration in the one-point passes, WS sites of w in shared memory
template <int R, int G, int TS, int WS, bool RU = false>
struct TTGroup : TTDriver<TTGroup<R, G, TS, WS, RU>>, TTMachine<TTGroup<R, G, TS, WS, RU>> {
  using Drv = TTDriver<TTGroup<R, G, TS, WS, RU>>;
  using Mach = TTMachine<TTGroup<R, G, TS, WS, RU>>;
  using Drv::pend_bl; using Drv::dist_bl; using Drv::prox_bl; using Drv::calls;
  using Mach::opt_init; using Mach::opt_advance; using Mach::b_kind; using Mach::req_x; using Mach::o_status;
  static constexpr int S = 4, K = 4 * R, PM = 16 * R, NT = 32 * G;
  static constexpr int PQS = 5 * K;  // pendant matrix set: rows [R*4][5] = P[rs][0..3], QN[rs] (row sum, for 'N')
  static constexpr int NBP = 6, NBD = 4, MEMO = 88, MAXJOBS = 12;
  static constexpr int NBE = 44, CHE = 8;  // --pp: pendant points per eigenbasis pass (a whole QNG rule), points per register chunk
  static constexpr unsigned FULL = 0xffffffffu;
  enum : int { CMD_DONE = 0, CMD_DIST1, CMD_PEND1_FUSED, CMD_PEND1_CACHED, CMD_DIST_BATCH, CMD_PEND_BATCH_FUSED, CMD_PEND_BATCH_CACHED, CMD_PEND_EIG };
  // block-shared layout (doubles)
  static constexpr int C_STATD = 0, C_LAM = 4, C_X = 8, C_XIT = 24, C_RATES = 40, C_TOTAL = (40 + R + 1) & ~1;
  static constexpr int OFF_EX = C_TOTAL;                 // [MAXJOBS][K] exponentials of the posted builds
  static constexpr int OFF_JBL = OFF_EX + MAXJOBS * K;   // [MAXJOBS] branch length of each build job
  static constexpr int OFF_JOB = OFF_JBL + MAXJOBS;      // ints: [0] pass, [1] n_jobs, [2] queue slot, [4..] dst[MAXJOBS], qn[MAXJOBS]
  static constexpr int OFF_OUT = OFF_JOB + 2 + MAXJOBS;  // [8] pass results ( JOB()[3]: points of an eigenbasis pass )
  static constexpr int OFF_BAD = OFF_OUT + 8;            // [8] non-finite log terms per point
  static constexpr int OFF_PART = OFF_BAD + 8;           // [G][8] per-warp partial sums
  static constexpr int OFF_ESUM = OFF_PART + G * 8;      // [G] per-warp exponent sums
  static constexpr int OFF_PQ = OFF_ESUM + G;            // current pendant set
  static constexpr int OFF_PD = OFF_PQ + PQS;            // current distal P, then proximal P
  static constexpr int OFF_PB = OFF_PD + 2 * PM;         // batch: pendant [NBP][PQS] | distal [NBD][2][PM]
  static constexpr int PB_SIZE = NBP * PQS > NBD * 2 * PM ? NBP * PQS : NBD * 2 * PM;
  static constexpr int OFF_MX = OFF_PB + PB_SIZE;        // memo: argument, proximal length, value
  static constexpr int OFF_MA = OFF_MX + MEMO;
  static constexpr int OFF_MF = OFF_MA + MEMO;
  static constexpr int OFF_MI = OFF_MF + MEMO;           // --pp: integrand value exp(f - base_ll) * prior(x) of a memo entry
  static constexpr int OFF_OBJ = OFF_MI + MEMO;          // warp 0's scalar state (this object)
  static constexpr int OBJ_DOUBLES = 96;
  // --pp, eigenbasis pendant batch (pass_pend_eig): exponentials [NBE][K], results [NBE], per-warp partials [G][NBE],
  // change of basis of the query column xq[mask][k] = sum_{c in mask} xit[c][k]
  static constexpr int OFF_EE = OFF_OBJ + OBJ_DOUBLES;
  static constexpr int OFF_OUTE = OFF_EE + NBE * K;
  static constexpr int OFF_PARTE = OFF_OUTE + NBE;
  static constexpr int OFF_XQ = OFF_PARTE + G * NBE;
  static constexpr int OFF_W = OFF_XQ + 64;              // w[K][WS] for the first WS sites of the query
  static_assert(WS % NT == 0, "an iteration's sites are all in shared memory or all in the global overflow");
  static_assert((OFF_EE % 2) == 0 && (K % 2) == 0, "double2 alignment of the exponentials");
  static_assert((PQS % 2) == 0 && (PM % 2) == 0 && (MEMO % 2) == 0 && (G % 2) == 0 && (MAXJOBS % 2) == 0 && (OFF_PD % 2) == 0 &&
                (OFF_PB % 2) == 0, "double2 alignment of the matrix sets");

  static __host__ __device__ constexpr size_t smem_bytes() { return sizeof(double) * ((size_t)OFF_W + (size_t)K * WS); }

  // ---- scalar state: read and written by warp 0 only (all lanes store identical values) ----
  double e_tot_ln2, rate_term, fval, cut_bl, start_pend, tol, max_pend;
  int calls_dist, max_iter;
  GCand cand;  // warp 0's copy, for the nested --pp driver
  int feed, attempt, status, optimise, final_pending;
  bool w_stale, pq_dirty, pd_dirty;
  int memo_n, memo_pending, memo_kind;  // memo_kind: -1 none, 0 pendant values, 1 distal values
  int memo_cur;     // QNG reads its abscissae back in the order it announced them: the expected slot is tried first
  bool memo_integrand;  // OFF_MI holds the integrand values of the memo entries
  int n_jobs;

  __device__ __forceinline__ static double *SM() { return reinterpret_cast<double *>(ppb_ttw_smem); }
  __device__ __forceinline__ static int *JOB() { return reinterpret_cast<int *>(SM() + OFF_JOB); }
  __device__ __forceinline__ static int lane_id() { return threadIdx.x & 31; }
  __device__ __forceinline__ static int warp_id() { return threadIdx.x >> 5; }
  // darr/parr gathers: read-only path; RU = keep them in L2 ahead of the streamed w overflow (evict_last)
  __device__ __forceinline__ static unsigned long long glv_policy() {
    unsigned long long pol = 0;
    if (RU) asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
  }
  __device__ __forceinline__ static double ld_glv(const double *p, unsigned long long pol) {
    if (!RU) return __ldg(p);
    double v;
    asm volatile("ld.global.nc.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(p), "l"(pol));
    return v;
  }

  __device__ __forceinline__ static double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
  }

  // ---- warp 0: build jobs and the memo ----
  __device__ void add_job(double bl, int dst_off, int qn) {
    __syncwarp();
    if (lane_id() == 0) { SM()[OFF_JBL + n_jobs] = bl; JOB()[4 + n_jobs] = dst_off; JOB()[4 + MAXJOBS + n_jobs] = qn; }
    __syncwarp();
    ++n_jobs;
  }
  __device__ void memo_clear(int kind) { memo_n = 0; memo_pending = 0; memo_kind = kind; memo_cur = 0; memo_integrand = false; }
  __device__ int memo_find(int kind, double x, double aux) {
    if (memo_kind != kind) return -1;
    const double *s = SM();
    if (memo_cur < memo_n && s[OFF_MX + memo_cur] == x && s[OFF_MA + memo_cur] == aux) return memo_cur++;
    for (int i = 0; i < memo_n; ++i)
      if (s[OFF_MX + i] == x && s[OFF_MA + i] == aux) { memo_cur = i + 1; return i; }
    return -1;
  }
  __device__ bool memo_get(int kind, double x, double aux, double *f) {
    const int i = memo_find(kind, x, aux);
    if (i < 0) return false;
    *f = SM()[OFF_MF + i];
    return true;
  }
  // TTDriver::integrand_inner: exp(f_pend(x) - base_ll) * prior(x), computed for a whole rule by the lanes of warp 0
  __device__ bool integrand_memo(double x, double *v) {
    if (!memo_integrand) return false;
    const int i = memo_find(0, x, 0.0);
    if (i < 0) return false;
    pend_bl = x; pq_dirty = true;  // what f_pend leaves behind
    *v = SM()[OFF_MI + i];
    return true;
  }
  // Brent's predictable first calls, several per pass (TTWarp::hint_prepare): posts the batch's build jobs
  __device__ void hint_prepare(int kind, double left, double right, double start, double v) {
    double *s = SM();
    if (kind == OBJ_DIST) {
      // replay the drift of cut = dist + prox over the seven calls (set_dist_bl, three_tax.ml:41-45)
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
        for (int b = 0; b < 4; ++b) { s[OFF_MX + b] = bx[b]; s[OFF_MA + b] = bp[b]; }
      }
      memo_pending = 4;
#pragma unroll
      for (int b = 0; b < 4; ++b) { add_job(bx[b], OFF_PB + 2 * b * PM, 0); add_job(bp[b], OFF_PB + (2 * b + 1) * PM, 0); }
    } else {
      memo_clear(0);
      const double mid = 0.5 * (left + right);
      const double ee = (start < mid) ? right - start : -(start - left);
      const double u1 = start + 0.3819660 * ee;  // Brent's first golden-section step (GSL min/brent.c)
      const double bls[5] = {left, right, start, v, u1};
      __syncwarp();
      if (lane_id() == 0) {
#pragma unroll
        for (int b = 0; b < 5; ++b) { s[OFF_MX + b] = bls[b]; s[OFF_MA + b] = 0.0; }
      }
      memo_pending = 5;
#pragma unroll
      for (int b = 0; b < NBP; ++b) add_job(bls[b < 5 ? b : 4], OFF_PB + b * PQS, 1);
    }
  }
  __device__ void hint_commit() {
    double *s = SM();
    __syncwarp();
    if (lane_id() < memo_pending) s[OFF_MF + lane_id()] = s[OFF_OUT + lane_id()];
    __syncwarp();
    memo_n = memo_pending;
    memo_pending = 0;
  }

  // core.ml:225-239 set_tt_edges / :240-245 tt_edges_default: branch lengths only (the sites are
  // never gathered); the matrices are rebuilt lazily by the next pass that needs them
  __device__ void reset_edges(double pendant, double distal) {
    pend_bl = pendant; dist_bl = distal; prox_bl = cut_bl - distal;
    pq_dirty = true; pd_dirty = true; w_stale = true;
    memo_clear(-1);
  }
  __device__ void post(int pass, bool need_pq, bool need_pd) {
    if (need_pq && pq_dirty) { add_job(pend_bl, OFF_PQ, 1); pq_dirty = false; }
    if (need_pd && pd_dirty) { add_job(dist_bl, OFF_PD, 0); add_job(prox_bl, OFF_PD + PM, 0); pd_dirty = false; }
    __syncwarp();
    if (lane_id() == 0) { JOB()[0] = pass; JOB()[1] = n_jobs; }
    __syncwarp();
  }

  // warp 0: advance the optimiser to the next call that needs a pass; returns false when the candidate is done
  __device__ __forceinline__ bool scalar_step(bool first, double *ll_out) {
    double *s = SM();
    __syncwarp();  // the 32 lanes run this uniform code converged (see the note on redundant execution above)
    if (!first) {
      // consume the pass just finished
      if (final_pending) { *ll_out = s[OFF_OUT]; return false; }
      if (feed != 2) { fval = -s[OFF_OUT]; feed = 1; }
    }
    n_jobs = 0;
    for (;;) {
      const int more = optimise ? opt_advance(feed, fval) : 0;
      __syncwarp();
      if (more == 2) {  // Brent's predictable calls: hint_prepare posted the batch's matrices
        feed = 2;
        if (memo_kind == 1) post(CMD_DIST_BATCH, true, false);
        else if (w_stale) { post(CMD_PEND_BATCH_FUSED, false, true); w_stale = false; }
        else post(CMD_PEND_BATCH_CACHED, false, false);
        return true;
      }
      if (more == 0) {
        if (optimise) {
          status = o_status;
          if (st_is_gsl(status) && attempt == 0) {  // core.ml:254-279: retry once from the default branch lengths
            attempt = 1;
            reset_edges(start_pend, cut_bl / 2.0);
            opt_init(tol, max_pend, max_iter);
            feed = 0;
            n_jobs = 0;
            continue;
          }
          if (status == ST_EXCEEDED_MAX_ITER) status = ST_OK;
        }
        final_pending = 1;  // Three_tax.get_results (three_tax.ml:85): log_like at the state left by the last call
        post(CMD_DIST1, true, true);
        return true;
      }
      // ---- TTDriver::objective, with calls answered by the memo costing no pass ----
      const double x = req_x;
      double val;
      if (b_kind == OBJ_PEND) {
        calls++;
        pend_bl = x; pq_dirty = true;
        if (memo_kind == 1) memo_kind = -1;
        if (memo_get(0, x, 0.0, &val)) { fval = -val; feed = 1; continue; }
        if (w_stale) { post(CMD_PEND1_FUSED, true, true); w_stale = false; } else post(CMD_PEND1_CACHED, true, false);
        feed = 1;
        return true;
      }
      calls++;
      calls_dist++;
      const double prox = (dist_bl + prox_bl) - x;
      if (!(prox >= 0.0)) { fval = __longlong_as_double(0x7ff8000000000000LL); feed = 1; continue; }
      const bool hit = memo_get(1, x, prox, &val);
      dist_bl = x; prox_bl = prox; pd_dirty = true; w_stale = true;
      if (memo_kind == 0) memo_kind = -1;
      if (hit) { fval = -val; feed = 1; continue; }
      post(CMD_DIST1, true, true);
      feed = 1;
      return true;
    }
  }

  // ---- all threads: Diagd.multi_exp (diagd.ml:79-98; linear_c.c:92-146) for the posted jobs ----
  __device__ __forceinline__ static void build_jobs(int n) {
    double *s = SM();
    const int *job = JOB();
    const int tid = threadIdx.x;
    for (int t = tid; t < n * K; t += NT) {
      const int j = t / K, kk = t - j * K;
      s[OFF_EX + t] = exp(s[C_RATES + (kk >> 2)] * s[OFF_JBL + j] * s[C_LAM + (kk & 3)]);
    }
    __syncthreads();
    for (int t0 = 0; t0 < n * PM; t0 += NT) {  // uniform trip count: the shuffles below need whole warps
      const int t = t0 + tid;
      const bool on = t < n * PM;
      const int j = on ? t / PM : 0, e = on ? t - j * PM : 0;
      const int r = e >> 4, i = (e >> 2) & 3, c = e & 3;
      double acc = 0.0;
#pragma unroll
      for (int k = 0; k < 4; ++k) acc = fma(s[OFF_EX + j * K + r * 4 + k] * s[C_X + i * 4 + k], s[C_XIT + c * 4 + k], acc);
      const int qn = job[4 + MAXJOBS + j];
      double *dst = s + job[4 + j];
      if (on) dst[qn ? (e >> 2) * 5 + c : e] = acc;
      // row sum ((0+P0)+P1)+P2)+P3 for the all-ones character (lanes 4m..4m+3 hold row m)
      const double p1 = __shfl_down_sync(FULL, acc, 1), p2 = __shfl_down_sync(FULL, acc, 2), p3 = __shfl_down_sync(FULL, acc, 3);
      if (on && qn && c == 0) dst[(e >> 2) * 5 + 4] = (((0.0 + acc) + p1) + p2) + p3;
    }
    __syncthreads();
  }

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
        ww[t][s] = (sts * pe) * de;  // ((statd*x)*y), x = prox, y = dist (three_tax.ml:31-36)
      }
    }
  }

  // per-warp partial sums of the NB running products; warp 0 adds them (fixed order) after the barrier
  template <int NB>
  __device__ __forceinline__ void pass_end(const WarpProd (&lp)[NB]) {
    double *s = SM();
#pragma unroll
    for (int b = 0; b < NB; ++b) {
      const double v = warp_sum(log(lp[b].m) + (double)lp[b].e * kLn2);
      if (lane_id() == 0) s[OFF_PART + warp_id() * 8 + b] = v;
    }
    __syncthreads();
    if (warp_id() == 0) {
#pragma unroll
      for (int b = 0; b < NB; ++b) {
        double tot = 0.0;
#pragma unroll
        for (int g = 0; g < G; ++g) tot += s[OFF_PART + g * 8 + b];
        const double v = (tot + s[OFF_BAD + b]) - rate_term + e_tot_ln2;
        __syncwarp();
        if (lane_id() == 0) { s[OFF_OUT + b] = v; s[OFF_BAD + b] = 0.0; }
      }
      __syncwarp();
    }
  }

  // Three_tax.log_like at NB (distal, proximal) pairs with the current pendant matrices
  template <int NB, int T>
  __device__ __forceinline__ void pass_dist(const GCand &c, int pm_off) {
    double *s = SM();
    const double *Pm = s + pm_off, *Pq = s + OFF_PQ, *st = s + C_STATD;
    double *bad = s + OFF_BAD;
    WarpProd lp[NB];
#pragma unroll
    for (int b = 0; b < NB; ++b) lp[b].init();
    const int tid = threadIdx.x;
    const unsigned long long pol = glv_policy();
    for (int base = 0; base < c.nq; base += NT * T) {
      bool act[T], fast = true;
      const double *dp[T], *pp[T], *Pqc[T];
      uint32_t mask[T];
#pragma unroll
      for (int t = 0; t < T; ++t) {
        const int i = base + t * NT + tid;
        act[t] = i < c.nq;
        const int packed = c.qsite[min(i, c.nq - 1)];
        const RowCodes rc = decode(packed);
        fast = fast && rc.fast;
        mask[t] = rc.mask;
        Pqc[t] = Pq + rc.qcol;
        const int site = packed & kSiteMask;
        dp[t] = c.dbase + site; pp[t] = c.pbase + site;
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
            d[t][k] = ld_glv(dp[t] + (int64_t)(r * 4 + k) * c.Lp, pol);
            p[t][k] = ld_glv(pp[t] + (int64_t)(r * 4 + k) * c.Lp, pol);
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
          for (int b = 0; b < NB; ++b) lp[b].mul(u[t][b], bad + b);
        }
    }
    pass_end<NB>(lp);
  }

  // log_like at NB pendant lengths.  FUSED: recompute w from darr/parr and keep it (shared memory,
  // global beyond ws sites); otherwise read it back.
  template <int NB, bool FUSED, int T>
  __device__ __forceinline__ void pass_pend(const GCand &c, int pq_off) {
    double *s = SM();
    const double *Pqm = s + pq_off, *Pd0 = s + OFF_PD, *st = s + C_STATD;
    double *bad = s + OFF_BAD;
    WarpProd lp[NB];
#pragma unroll
    for (int b = 0; b < NB; ++b) lp[b].init();
    const int tid = threadIdx.x;
    const unsigned long long pol = glv_policy();
    for (int base = 0; base < c.nq; base += NT * T) {
      bool act[T], fast = true;
      int iw[T];
      const double *dp[T], *pp[T], *Pqc[T];
      uint32_t mask[T];
#pragma unroll
      for (int t = 0; t < T; ++t) {
        const int i = base + t * NT + tid;
        act[t] = i < c.nq;
        iw[t] = min(i, c.nq - 1);
        const int packed = c.qsite[iw[t]];
        const RowCodes rc = decode(packed);
        fast = fast && rc.fast;
        mask[t] = rc.mask;
        Pqc[t] = Pqm + rc.qcol;
        const int site = packed & kSiteMask;
        dp[t] = c.dbase + site; pp[t] = c.pbase + site;
      }
      double u[T][NB];
#pragma unroll
      for (int t = 0; t < T; ++t)
#pragma unroll
        for (int b = 0; b < NB; ++b) u[t][b] = 0.0;
#pragma unroll(!FUSED && NB == 1 ? R : 1)
      for (int r = 0; r < R; ++r) {
        double ww[T][4];
        if (FUSED) {
          double d[T][4], p[T][4];
#pragma unroll
          for (int t = 0; t < T; ++t)
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              d[t][k] = ld_glv(dp[t] + (int64_t)(r * 4 + k) * c.Lp, pol);
              p[t][k] = ld_glv(pp[t] + (int64_t)(r * 4 + k) * c.Lp, pol);
            }
          site_w<T>(Pd0 + r * 16, Pd0 + PM + r * 16, d, p, st, ww);
#pragma unroll
          for (int t = 0; t < T; ++t)
            if (act[t]) {
              if (base + t * NT < WS) {  // block-uniform: WS is a multiple of the block size
                double *wp = SM() + OFF_W + (r * 4) * WS + iw[t];
#pragma unroll
                for (int k = 0; k < 4; ++k) wp[k * WS] = ww[t][k];
              } else {
                double *wp = c.wg + (int64_t)(r * 4) * c.Lp + iw[t];
#pragma unroll
                for (int k = 0; k < 4; ++k) __stcs(wp + (int64_t)k * c.Lp, ww[t][k]);
              }
            }
        } else {
#pragma unroll
          for (int t = 0; t < T; ++t) {
            if (base + t * NT < WS) {
              const double *wp = SM() + OFF_W + (r * 4) * WS + iw[t];
#pragma unroll
              for (int k = 0; k < 4; ++k) ww[t][k] = wp[k * WS];
            } else {
              const double *wp = c.wg + (int64_t)(r * 4) * c.Lp + iw[t];
#pragma unroll
              for (int k = 0; k < 4; ++k) ww[t][k] = __ldcs(wp + (int64_t)k * c.Lp);
            }
          }
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
          for (int b = 0; b < NB; ++b) lp[b].mul(u[t][b], bad + b);
        }
    }
    pass_end<NB>(lp);
  }

  // Three_tax.log_like at n_pts pendant lengths from the cached product w, in the eigenbasis of the rate matrix: with
  // P_r(t) = X exp(rho_r t Lambda) Xit^T the site likelihood is  u(t) = sum_{r,k} exp(rho_r t lambda_k) * y[r,k],
  // y[r,k] = (sum_s w[r,s] X[s,k]) * (sum_{c in code} Xit[c,k])  - no pendant table per point, 16 multiply-adds per
  // point and site against exponentials shared by the block (OFF_EE, computed by warp 0 before the pass is posted).
  // A whole QNG rule (21 / 22 / 44 abscissae, integration.ml:27-31) is one pass instead of four to eight.
  // running product of site likelihoods as (mantissa product, exponent): the factor's exponent is split off BEFORE the
  // multiplication, so the chain from site to site is one DMUL; special factors (zero, negative, NaN, inf, subnormal)
  // go through log() into `sp` so that -inf / NaN propagate as in the reference's sum of logs (linear_c.c:418)
  struct EigProd {
    double m, sp;
    int e;
    __device__ __forceinline__ void init() { m = 1.0; sp = 0.0; e = 0; }
    __device__ __forceinline__ void mul(double u) {
      const int hi = __double2hiint(u);
      const int ef = (hi >> 20) & 0x7ff;
      if (hi < 0 || ef == 0 || ef == 0x7ff) { sp += log(u); return; }
      e += ef - 1023;
      m *= __hiloint2double((hi & 0x800fffff) | 0x3ff00000, __double2loint(u));
    }
    __device__ __forceinline__ void renorm() {
      const int hi = __double2hiint(m);
      e += ((hi >> 20) & 0x7ff) - 1023;
      m = __hiloint2double((hi & 0x800fffff) | 0x3ff00000, __double2loint(m));
    }
    __device__ __forceinline__ double value() { renorm(); return log(m) + (double)e * kLn2 + sp; }
  };
  __device__ __forceinline__ void eig_load_w(const GCand &c, int i, double (&w)[K]) const {
    const double *s = SM();
    if (i < WS) {
#pragma unroll
      for (int k = 0; k < K; ++k) w[k] = s[OFF_W + k * WS + i];
    } else {
#pragma unroll
      for (int k = 0; k < K; ++k) w[k] = __ldcs(c.wg + (int64_t)k * c.Lp + i);
    }
  }
  __device__ __noinline__ void pass_pend_eig(const GCand &c, int n_pts) {
    double *s = SM();
    const double *X = s + C_X, *XQ = s + OFF_XQ, *E = s + OFF_EE;
    const int tid = threadIdx.x;
    for (int c0 = 0; c0 < n_pts; c0 += CHE) {
      EigProd lp[CHE];
#pragma unroll
      for (int b = 0; b < CHE; ++b) lp[b].init();
      int iter = 0;
      for (int i = tid; i < c.nq; i += NT, ++iter) {
        double w[K];
        eig_load_w(c, i, w);
        const uint32_t mask = ((uint32_t)c.qsite[i] >> kSiteCodeShift) & 15u;
        double y[K];
#pragma unroll
        for (int r = 0; r < R; ++r)
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            double t = w[r * 4] * X[k];
#pragma unroll
            for (int q = 1; q < 4; ++q) t = fma(w[r * 4 + q], X[q * 4 + k], t);
            y[r * 4 + k] = t * XQ[mask * 4 + k];
          }
#pragma unroll
        for (int b = 0; b < CHE; ++b) {
          const double2 *Eb = reinterpret_cast<const double2 *>(E + min(c0 + b, n_pts - 1) * K);  // block-uniform: broadcast loads
          double u0 = 0.0, u1 = 0.0;
#pragma unroll
          for (int j = 0; j < K / 2; ++j) { const double2 e = Eb[j]; u0 = fma(e.x, y[2 * j], u0); u1 = fma(e.y, y[2 * j + 1], u1); }
          lp[b].mul(u0 + u1);
        }
        if ((iter & 255) == 255) {
#pragma unroll
          for (int b = 0; b < CHE; ++b) lp[b].renorm();
        }
      }
#pragma unroll
      for (int b = 0; b < CHE; ++b) {
        const double v = warp_sum(lp[b].value());
        if (lane_id() == 0 && c0 + b < n_pts) s[OFF_PARTE + warp_id() * NBE + c0 + b] = v;
      }
    }
    __syncthreads();
    if (warp_id() == 0) {
      for (int b = lane_id(); b < n_pts; b += 32) {
        double tot = 0.0;
#pragma unroll
        for (int g = 0; g < G; ++g) tot += s[OFF_PARTE + g * NBE + b];
        s[OFF_OUTE + b] = tot - rate_term + e_tot_ln2;
      }
      __syncwarp();
    }
  }

  // ---------------------------------------------------------------------------------------------
  // --pp on this engine: warp 0 runs TTDriver::calc_marg_prob (nested Brent / QNG, ordinary calls)
  // and every objective value it needs is a BLOCKING request: post the command, then execute it
  // together with the other warps, which sit in a loop on run_command() until CMD_DONE.
  // ---------------------------------------------------------------------------------------------
  __device__ __noinline__ bool run_command(const GCand &c) {
    __syncthreads();  // the command is posted
    const int cmd = JOB()[0], nj = JOB()[1];
    if (cmd == CMD_DONE) return false;
    if (nj > 0) build_jobs(nj);
    switch (cmd) {
      case CMD_DIST1: pass_dist<1, 1>(c, OFF_PD); break;
      case CMD_PEND1_FUSED: pass_pend<1, true, 1>(c, OFF_PQ); break;
      case CMD_PEND1_CACHED: pass_pend<1, false, 1>(c, OFF_PQ); break;
      case CMD_DIST_BATCH: pass_dist<NBD, 1>(c, OFF_PB); break;
      case CMD_PEND_BATCH_FUSED: pass_pend<NBP, true, 1>(c, OFF_PB); break;
      case CMD_PEND_EIG: pass_pend_eig(c, JOB()[3]); break;
      default: pass_pend<NBP, false, 1>(c, OFF_PB); break;
    }
    return true;
  }
  __device__ void post_done() {
    __syncwarp();
    if (lane_id() == 0) { JOB()[0] = CMD_DONE; JOB()[1] = 0; }
    __syncwarp();
  }
  __device__ double pend_now() {  // log_like at the current branch lengths through a pendant pass
    n_jobs = 0;
    if (w_stale) { post(CMD_PEND1_FUSED, true, true); w_stale = false; } else post(CMD_PEND1_CACHED, true, false);
    run_command(cand);
    return SM()[OFF_OUT];
  }
  // TTDriver engine interface (warp 0 only)
  __device__ double f_pend(double v) {
    pend_bl = v; pq_dirty = true;
    if (memo_kind == 1) memo_kind = -1;
    double f;
    if (memo_get(0, v, 0.0, &f)) return f;
    return pend_now();
  }
  __device__ double f_dist(double d, double p) {
    double f;
    const bool hit = memo_get(1, d, p, &f);
    set_dist_prox(d, p);
    if (hit) return f;
    n_jobs = 0;
    post(CMD_DIST1, true, true);
    run_command(cand);
    return SM()[OFF_OUT];
  }
  __device__ void set_dist_prox(double d, double p) {
    dist_bl = d; prox_bl = p; pd_dirty = true; w_stale = true;
    if (memo_kind == 0) memo_kind = -1;
  }
  __device__ void cache_w() {}  // lazy: the next pendant pass after a distal move recomputes w (fused)
  __device__ double eval_full_cache() { return pend_now(); }
  __device__ void hint_brent(int kind, double left, double right, double start, double v) {
    n_jobs = 0;
    hint_prepare(kind, left, right, start, v);
    if (memo_kind == 1) post(CMD_DIST_BATCH, true, false);
    else if (w_stale) { post(CMD_PEND_BATCH_FUSED, false, true); w_stale = false; }
    else post(CMD_PEND_BATCH_CACHED, false, false);
    run_command(cand);
    hint_commit();
  }
  // gsl_integration_qng (GSL integration/qng.c; TTDriver::qng) for the INNER integral, with every rule's integrand values
  // produced by one batched evaluation (hint_flush: lanes write the abscissae, one or two passes, lanes take exp * prior)
  // and the rule's sums then read from shared memory - the same values combined in the same order as TTDriver::qng,
  // without a function call, a memo search and two exponentials on one thread per abscissa.
  __device__ __noinline__ int qng_inner(double a, double b, double epsabs, double epsrel, double *result) {
    double *s = SM();
    const double half_length = 0.5 * (b - a), abs_half_length = fabs(half_length), center = 0.5 * (b + a);
    auto announce = [&](const double *xs, int n_pairs, int at, bool with_center) {  // evaluation order: (center,) +x_k, -x_k ...
      __syncwarp();
      const int n = 2 * n_pairs + (with_center ? 1 : 0);
      for (int t = lane_id(); t < n; t += 32) {
        double x;
        if (with_center && t == 0) x = center;
        else { const int u = t - (with_center ? 1 : 0); const double ab = half_length * xs[u >> 1]; x = (u & 1) ? center - ab : center + ab; }
        s[OFF_MX + at + t] = x; s[OFF_MA + at + t] = 0.0;
      }
      __syncwarp();
      memo_pending = n;
      hint_flush();
      pend_bl = s[OFF_MX + at + n - 1]; pq_dirty = true;  // the last call's set_pend_bl
    };
    memo_clear(0);
    {  // the 21-point rule: center, x1 pairs, x2 pairs (the first 21 slots)
      __syncwarp();
      for (int t = lane_id(); t < 21; t += 32) {
        double x;
        if (t == 0) x = center;
        else { const int u = t - 1, k = u >> 1; const double ab = half_length * (k < 5 ? c_qng_x1[k] : c_qng_x2[k - 5]); x = (u & 1) ? center - ab : center + ab; }
        s[OFF_MX + t] = x; s[OFF_MA + t] = 0.0;
      }
      __syncwarp();
      memo_pending = 21;
      hint_flush();
      pend_bl = s[OFF_MX + 20]; pq_dirty = true;
    }
    const double *v = s + OFF_MI;
    double fv1[5], fv2[5], fv3[5], fv4[5], savfun[21];
    const double f_center = v[0];
    if (epsabs <= 0.0 && (epsrel < 50.0 * kDblEps || epsrel < 0.5e-28)) { *result = 0.0; return ST_GSL_EBADTOL; }
    double res10 = 0.0, res21 = c_qng_w21b[5] * f_center, resabs = c_qng_w21b[5] * fabs(f_center);
    for (int k = 0; k < 5; ++k) {
      const double f1 = v[1 + 2 * k], f2 = v[2 + 2 * k], fv = f1 + f2;
      res10 += c_qng_w10[k] * fv;
      res21 += c_qng_w21a[k] * fv;
      resabs += c_qng_w21a[k] * (fabs(f1) + fabs(f2));
      savfun[k] = fv; fv1[k] = f1; fv2[k] = f2;
    }
    for (int k = 0; k < 5; ++k) {
      const double f1 = v[11 + 2 * k], f2 = v[12 + 2 * k], fv = f1 + f2;
      res21 += c_qng_w21b[k] * fv;
      resabs += c_qng_w21b[k] * (fabs(f1) + fabs(f2));
      savfun[k + 5] = fv; fv3[k] = f1; fv4[k] = f2;
    }
    resabs *= abs_half_length;
    const double mean = 0.5 * res21;
    double resasc = c_qng_w21b[5] * fabs(f_center - mean);
    for (int k = 0; k < 5; ++k)
      resasc += (c_qng_w21a[k] * (fabs(fv1[k] - mean) + fabs(fv2[k] - mean)) +
                 c_qng_w21b[k] * (fabs(fv3[k] - mean) + fabs(fv4[k] - mean)));
    resasc *= abs_half_length;
    double result_kronrod = res21 * half_length;
    double err = Drv::rescale_error((res21 - res10) * half_length, resabs, resasc);
    if (err < epsabs || err < epsrel * fabs(result_kronrod)) { *result = result_kronrod; return ST_OK; }
    announce(c_qng_x3, 11, 21, false);
    double res43 = c_qng_w43b[11] * f_center;
    for (int k = 0; k < 10; ++k) res43 += savfun[k] * c_qng_w43a[k];
    for (int k = 0; k < 11; ++k) {
      const double fv = v[21 + 2 * k] + v[22 + 2 * k];
      res43 += fv * c_qng_w43b[k];
      savfun[k + 10] = fv;
    }
    result_kronrod = res43 * half_length;
    err = Drv::rescale_error((res43 - res21) * half_length, resabs, resasc);
    if (err < epsabs || err < epsrel * fabs(result_kronrod)) { *result = result_kronrod; return ST_OK; }
    announce(c_qng_x4, 22, 43, false);
    double res87 = c_qng_w87b[22] * f_center;
    for (int k = 0; k < 21; ++k) res87 += savfun[k] * c_qng_w87a[k];
    for (int k = 0; k < 22; ++k) res87 += c_qng_w87b[k] * (v[43 + 2 * k] + v[44 + 2 * k]);
    result_kronrod = res87 * half_length;
    err = Drv::rescale_error((res87 - res43) * half_length, resabs, resasc);
    *result = result_kronrod;
    if (err < epsabs || err < epsrel * fabs(result_kronrod)) return ST_OK;
    return ST_GSL_ETOL;
  }
  // QNG announces each rule's abscissae before evaluating them: six per pass
  __device__ void hint_begin() { memo_clear(0); }
  __device__ void hint_add(double x) {
    const int at = memo_n + memo_pending;
    if (at >= MEMO) return;
    __syncwarp();
    if (lane_id() == 0) { SM()[OFF_MX + at] = x; SM()[OFF_MA + at] = 0.0; }
    __syncwarp();
    ++memo_pending;
  }
  __device__ void hint_flush() {
    double *s = SM();
    while (memo_pending > 0) {
      if (!w_stale) {  // the cached product is current: the rest of the rule in ONE eigenbasis pass
        const int n = memo_pending < NBE ? memo_pending : NBE;
        __syncwarp();
        for (int t = lane_id(); t < n * K; t += 32) {
          const int b = t / K, rk = t - b * K;
          s[OFF_EE + t] = exp(s[C_RATES + (rk >> 2)] * s[OFF_MX + memo_n + b] * s[C_LAM + (rk & 3)]);
        }
        __syncwarp();
        n_jobs = 0;
        if (lane_id() == 0) { JOB()[0] = CMD_PEND_EIG; JOB()[1] = 0; JOB()[3] = n; }
        __syncwarp();
        run_command(cand);
        __syncwarp();
        for (int b = lane_id(); b < n; b += 32) {
          const double f = s[OFF_OUTE + b];
          s[OFF_MF + memo_n + b] = f;
          s[OFF_MI + memo_n + b] = exp(f - this->base_ll) * this->prior(s[OFF_MX + memo_n + b]);
        }
        __syncwarp();
        pend_bl = s[OFF_MX + memo_n + n - 1]; pq_dirty = true;  // as after the last of these calls (set_pend_bl)
        memo_n += n;
        memo_pending -= n;
        continue;
      }
      const int n = memo_pending < NBP ? memo_pending : NBP;
      n_jobs = 0;
      for (int b = 0; b < NBP; ++b) add_job(s[OFF_MX + memo_n + (b < n ? b : n - 1)], OFF_PB + b * PQS, 1);
      if (w_stale) { post(CMD_PEND_BATCH_FUSED, false, true); w_stale = false; } else post(CMD_PEND_BATCH_CACHED, false, false);
      run_command(cand);
      __syncwarp();
      if (lane_id() < n) {
        const double f = s[OFF_OUT + lane_id()];
        s[OFF_MF + memo_n + lane_id()] = f;
        s[OFF_MI + memo_n + lane_id()] = exp(f - this->base_ll) * this->prior(s[OFF_MX + memo_n + lane_id()]);
      }
      __syncwarp();
      memo_n += n;
      memo_pending -= n;
    }
    memo_integrand = true;
  }
};

// Persistent kernel: every BLOCK pulls candidates (in `order`: length class, then edge) until n are done.
template <int R, int G, int MINB, int TS, int WS, bool RU = false>
__global__ void __launch_bounds__(G * 32, MINB)
tt_group_kernel(const DevModel *__restrict__ m, GlvArrays g, const double *__restrict__ bl, const int32_t *__restrict__ qsite,
                const int32_t *__restrict__ nq, const CandIn *__restrict__ in, const int32_t *__restrict__ order,
                CandOut *__restrict__ out, int n, TTParams prm, double *__restrict__ wbuf, int *counter,
                unsigned long long *__restrict__ work) {
  using TT = TTGroup<R, G, TS, WS, RU>;
  double *sm = reinterpret_cast<double *>(ppb_ttw_smem);
  for (int i = threadIdx.x; i < 4; i += blockDim.x) { sm[TT::C_STATD + i] = m->statd[i]; sm[TT::C_LAM + i] = m->lam[i]; }
  for (int i = threadIdx.x; i < 16; i += blockDim.x) { sm[TT::C_X + i] = m->x[i]; sm[TT::C_XIT + i] = m->xit[i]; }
  for (int i = threadIdx.x; i < R; i += blockDim.x) sm[TT::C_RATES + i] = m->rates[i];
  for (int i = threadIdx.x; i < 8; i += blockDim.x) sm[TT::OFF_BAD + i] = 0.0;
  static_assert(sizeof(TT) <= TT::OBJ_DOUBLES * sizeof(double), "TTGroup object does not fit its shared slot");
  TT &tt = *reinterpret_cast<TT *>(sm + TT::OFF_OBJ);  // warp 0 only
  int *job = TT::JOB();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (;;) {
    __syncthreads();
    if (threadIdx.x == 0) job[2] = atomicAdd(counter, 1);
    __syncthreads();
    const int idx = job[2];
    if (idx >= n) break;
#ifdef PPB_TT_PROFILE
    long long pf_t0 = clock64(), pf_scalar = 0, pf_build = 0, pf_pass[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    int pf_n[8] = {0, 0, 0, 0, 0, 0, 0, 0}, pf_nb = 0, pf_jobs = 0;
#endif
    const int ci = order ? order[idx] : idx;
    const CandIn cin = in[ci];
    GCand c;
    c.qsite = qsite + (int64_t)cin.q * g.Lp;
    c.nq = nq[cin.q];
    c.dbase = g.a[0] + (int64_t)cin.loc * g.a_stride;
    c.pbase = g.a[1] + (int64_t)cin.loc * g.a_stride;
    c.Lp = g.Lp;
    c.wg = wbuf + (int64_t)blockIdx.x * TT::K * g.Lp;
    {  // total_twoexp of the distal and proximal GLVs over the query's sites (gstar_support.ml:43-49)
      const int32_t *ed = g.e[0] + (int64_t)cin.loc * g.Lp, *ep = g.e[1] + (int64_t)cin.loc * g.Lp;
      long long es = 0;
      for (int i = threadIdx.x; i < c.nq; i += TT::NT) {
        const int site = c.qsite[i] & kSiteMask;
        es += (long long)ed[site] + (long long)ep[site];
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) es += __shfl_xor_sync(0xffffffffu, es, o);
      if (lane == 0) sm[TT::OFF_ESUM + warp] = (double)es;  // exact: |es| < 2^53
    }
    __syncthreads();
    double ll = 0.0;
    if (warp == 0) {
      double tot = 0.0;
#pragma unroll
      for (int w = 0; w < G; ++w) tot += sm[TT::OFF_ESUM + w];
      tt.e_tot_ln2 = kLn2 * tot;
      tt.rate_term = (double)((float)c.nq) * log((double)((float)R));  // linear_c.c:421
      tt.cut_bl = bl[cin.loc];
      tt.start_pend = prm.start_pend; tt.tol = prm.tol; tt.max_pend = prm.max_pend; tt.max_iter = prm.max_iter;
      tt.calls = 0; tt.calls_dist = 0; tt.feed = 0; tt.attempt = 0; tt.status = ST_OK; tt.final_pending = 0; tt.fval = 0.0;
      tt.optimise = prm.mode == TT_MODE_OPTIMIZE;
      tt.reset_edges(prm.use_defaults ? prm.start_pend : cin.pend0, prm.use_defaults ? tt.cut_bl / 2.0 : cin.dist0);
      if (tt.optimise) tt.opt_init(prm.tol, prm.max_pend, prm.max_iter);
    }
    for (bool first = true;; first = false) {
#ifdef PPB_TT_PROFILE
      const long long pf_a = clock64();
#endif
      if (warp == 0) {
        if (!tt.scalar_step(first, &ll)) {
          __syncwarp();
          if (lane == 0) { job[0] = TT::CMD_DONE; job[1] = 0; }
        }
      }
#ifdef PPB_TT_PROFILE
      const long long pf_b = clock64();
      pf_scalar += pf_b - pf_a;
#endif
      __syncthreads();  // the command is posted
      const int cmd = job[0], nj = job[1];
      if (cmd == TT::CMD_DONE) break;
      if (nj > 0) TT::build_jobs(nj);
#ifdef PPB_TT_PROFILE
      const long long pf_c = clock64();
      pf_build += pf_c - pf_b; pf_nb += nj > 0; pf_jobs += nj;
#endif
      switch (cmd) {
        case TT::CMD_DIST1: tt.template pass_dist<1, TS>(c, TT::OFF_PD); break;
        case TT::CMD_PEND1_FUSED: tt.template pass_pend<1, true, TS>(c, TT::OFF_PQ); break;
        case TT::CMD_PEND1_CACHED: tt.template pass_pend<1, false, TS>(c, TT::OFF_PQ); break;
        case TT::CMD_DIST_BATCH: tt.template pass_dist<TT::NBD, 1>(c, TT::OFF_PB); break;
        case TT::CMD_PEND_BATCH_FUSED: tt.template pass_pend<TT::NBP, true, 1>(c, TT::OFF_PB); break;
        default: tt.template pass_pend<TT::NBP, false, 1>(c, TT::OFF_PB); break;
      }
#ifdef PPB_TT_PROFILE
      pf_pass[cmd & 7] += clock64() - pf_c; pf_n[cmd & 7]++;
#endif
    }
#ifdef PPB_TT_PROFILE
    if (threadIdx.x == 0) {  // work[4..]: cycles of thread 0 by phase, pass counts by command
      atomicAdd(work + 4, (unsigned long long)pf_scalar); atomicAdd(work + 5, (unsigned long long)pf_build);
      atomicAdd(work + 6, (unsigned long long)pf_nb); atomicAdd(work + 7, (unsigned long long)pf_jobs);
      for (int k = 1; k < 7; ++k) { atomicAdd(work + 8 + k, (unsigned long long)pf_pass[k]); atomicAdd(work + 16 + k, (unsigned long long)pf_n[k]); }
      atomicAdd(work + 24, 1ull); atomicAdd(work + 25, (unsigned long long)(clock64() - pf_t0));
    }
#endif
    if (threadIdx.x == 0) {
      CandOut o;
      o.marg = __longlong_as_double(0x7ff8000000000000LL);
      o.ll = ll; o.pend = tt.pend_bl; o.dist = tt.dist_bl; o.status = tt.status; o.calls = tt.calls;
      out[ci] = o;
      const unsigned long long nd = tt.calls_dist, np = tt.calls - tt.calls_dist, ns = c.nq;
      atomicAdd(work + 0, np); atomicAdd(work + 1, nd); atomicAdd(work + 2, np * ns); atomicAdd(work + 3, nd * ns);
    }
  }
}


// --pp (TT_MODE_MARGINAL): Three_tax.calc_marg_prob per kept placement, block per candidate
template <int R, int G, int MINB, int WS>
__global__ void __launch_bounds__(G * 32, MINB)
tt_group_pp_kernel(const DevModel *__restrict__ m, GlvArrays g, const double *__restrict__ bl, const int32_t *__restrict__ qsite,
                   const int32_t *__restrict__ nq, const CandIn *__restrict__ in, const int32_t *__restrict__ order,
                   CandOut *__restrict__ out, int n, TTParams prm, double *__restrict__ wbuf, int *counter) {
  using TT = TTGroup<R, G, 1, WS>;
  double *sm = reinterpret_cast<double *>(ppb_ttw_smem);
  for (int i = threadIdx.x; i < 4; i += blockDim.x) { sm[TT::C_STATD + i] = m->statd[i]; sm[TT::C_LAM + i] = m->lam[i]; }
  for (int i = threadIdx.x; i < 16; i += blockDim.x) { sm[TT::C_X + i] = m->x[i]; sm[TT::C_XIT + i] = m->xit[i]; }
  for (int i = threadIdx.x; i < R; i += blockDim.x) sm[TT::C_RATES + i] = m->rates[i];
  for (int i = threadIdx.x; i < 8; i += blockDim.x) sm[TT::OFF_BAD + i] = 0.0;
  for (int i = threadIdx.x; i < 64; i += blockDim.x) {  // xq[mask][k] = sum_{c in mask} xit[c][k]
    double acc = 0.0;
    for (int cc = 0; cc < 4; ++cc) if ((i >> 2) >> cc & 1) acc += m->xit[cc * 4 + (i & 3)];
    sm[TT::OFF_XQ + i] = acc;
  }
  TT &tt = *reinterpret_cast<TT *>(sm + TT::OFF_OBJ);  // warp 0 only
  int *job = TT::JOB();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (;;) {
    __syncthreads();
    if (threadIdx.x == 0) job[2] = atomicAdd(counter, 1);
    __syncthreads();
    const int idx = job[2];
    if (idx >= n) break;
    const int ci = order ? order[idx] : idx;
    const CandIn cin = in[ci];
    GCand c;
    c.qsite = qsite + (int64_t)cin.q * g.Lp;
    c.nq = nq[cin.q];
    c.dbase = g.a[0] + (int64_t)cin.loc * g.a_stride;
    c.pbase = g.a[1] + (int64_t)cin.loc * g.a_stride;
    c.Lp = g.Lp;
    c.wg = wbuf + (int64_t)blockIdx.x * TT::K * g.Lp;
    {
      const int32_t *ed = g.e[0] + (int64_t)cin.loc * g.Lp, *ep = g.e[1] + (int64_t)cin.loc * g.Lp;
      long long es = 0;
      for (int i = threadIdx.x; i < c.nq; i += TT::NT) {
        const int site = c.qsite[i] & kSiteMask;
        es += (long long)ed[site] + (long long)ep[site];
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) es += __shfl_xor_sync(0xffffffffu, es, o);
      if (lane == 0) sm[TT::OFF_ESUM + warp] = (double)es;
    }
    __syncthreads();
    if (warp == 0) {
      double tot = 0.0;
#pragma unroll
      for (int w = 0; w < G; ++w) tot += sm[TT::OFF_ESUM + w];
      tt.e_tot_ln2 = kLn2 * tot;
      tt.rate_term = (double)((float)c.nq) * log((double)((float)R));
      tt.cut_bl = bl[cin.loc];
      tt.cand = c;
      tt.calls = 0; tt.calls_dist = 0;
      tt.reset_edges(prm.use_defaults ? prm.start_pend : cin.pend0, prm.use_defaults ? tt.cut_bl / 2.0 : cin.dist0);
      tt.prior.kind = prm.prior_kind;
      tt.prior.mean = prm.prior_kind == 2 ? prm.prior_means[cin.loc] : prm.prior_mean;
      double mp = __longlong_as_double(0x7ff8000000000000LL);
      const int st = tt.calc_marg_prob(prm.pp_rel_err, prm.max_pend, &mp);
      tt.post_done();
      tt.run_command(c);
      if (lane == 0) {
        CandOut o;
        o.marg = mp; o.ll = mp; o.pend = tt.pend_bl; o.dist = tt.dist_bl; o.status = st; o.calls = tt.calls;
        out[ci] = o;
      }
    } else {
      while (tt.run_command(c)) {}
    }
  }
}

}  // namespace ppb
