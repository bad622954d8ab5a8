// Second stage, engine 4: block-per-candidate with the PENDANT objective in the eigenbasis of the rate
// matrix, for the four-state Gamma-mixture models (Three_tax.optimize / log_like).
//   Three_tax.log_like / optimize   pplacer_src/three_tax.ml:30-83
//   Model.log_like3    gmix_model.ml:298-304 -> ten_log_like3_c (linear_c.c:367-424)
//   Model.evolve_into  gmix_model.ml:308-315 ; Diagd.multi_exp diagd.ml:79-98 ; dediagonalize linear_c.c:92-146
//
// Why (profiles/README.md r3a: 20 of a candidate's 30 passes move only the pendant length, and each took
// 20k cycles in engine 3 - a matrix build, two barriers and 16 global + 16 shared loads per site):
// with P_r(t) = X exp(rho_r t Lambda) Xit^T the site likelihood at pendant length t is
//     u_i(t) = sum_{r,s} w_i[r,s] * sum_{c in code_i} P_r(t)[s,c]
//            = sum_{r,k} exp(rho_r t lambda_k) * y_i[r,k],   y_i[r,k] = (sum_s w_i[r,s] X[s,k]) * (sum_{c in code_i} Xit[c,k])
// where w = (statd*prox_ev)*dist_ev is the product the reference caches implicitly by mutating only the
// pendant edge (three_tax.ml:51-54).  lambda_k0 = 0 (rows of Q sum to zero), so the R terms of k0 collapse
// into one constant: a pendant pass is 3R fused multiply-adds per site against 3R WARP-UNIFORM exponentials -
// no rate-matrix build, no dediagonalize, no per-lane table lookups, no global memory: y (3R+1 doubles per
// site) lives in the block's shared memory for the first WS sites of the query (global overflow beyond).
//
// The passes that gather the candidate's distal / proximal vectors read them IN THE EIGENBASIS too (r5m): pp_precompute
// keeps alpha = Xit^T d and beta = Xit^T p per (node, rate, site) beside darr / parr, so that
//     P_r(t) d = X (exp(rho_r t Lambda) .* alpha)
// needs four exponentials per rate instead of a 4x4 matrix from shared memory (the X entries are immediates from the
// constant bank), and a distal evaluation contracts directly in the eigenbasis:
//     u_i = sum_r a_r^T M_r[code_i] b_r,   a = E(dist) .* alpha,  b = E(prox) .* beta,
//     M_r[code][k][k'] = sum_s statd_s q_r,s(code) X[s,k] X[s,k']          (q = the evolved pendant vector of the code)
// - 28 FP64 operations per (site, rate, point) instead of 44, and M (built once per pendant length) is shared by
// the four points of a batched pass.  M and the batch's exponentials live in the head of y, which is dead while
// the distal length moves.
//
// Division of labour, per objective call:
//   thread 0    copies the scalar optimiser state (TTMachine, threetax.cuh) from shared memory into registers,
//               advances it to the next call that needs a pass (memo hits cost nothing), posts a COMMAND;
//   warp 0      computes the 3R exponentials of each posted pendant point (one lane each);
//   all warps   build the posted rate matrices (distal / proximal moves only), run the pass and leave one
//               (mantissa product, exponent) partial per warp and point; lane b of warp 0 folds point b.
// Two block barriers per pendant pass (command, partials), four when matrices are built.
// The likelihood Three_tax.get_results reports (three_tax.ml:85) is the value of the LAST objective call -
// the state it left is the state reported - so no extra pass is spent on it.
#pragma once
#include "threetax_warp.cuh"

namespace ppb {

// model constants of engine 4, passed BY VALUE (kernel parameter = constant bank: uniform operands)
template <int R>
struct EigModel {
  double statd[4], lam[4], x[16], xit[16], rates[R];  // as built (reference order): used for the P(t) builds
  double lamn[3];                                     // the three non-zero eigenvalues (k0 = index of the zero one removed)
  double xz[4];                                       // x[s][k0]
  double xn[12];                                      // x[s][kk[j]], [s*3+j]
  double xq[64];                                      // [mask][0] = sum_{c in mask} xit[c][k0], [mask][1+j] = sum xit[c][kk[j]]
};

struct ECand {
  const double *dbase, *pbase;
  const int32_t *qsite;
  double *wg;  // global overflow of y for sites >= WS: plane stride Lp
  int64_t Lp;
  int nq;
};

template <int R, int NW, int WS>
struct TTEig {
  static constexpr int NT = 32 * NW, K = 4 * R, PM = 16 * R, NE = 3 * R, NY = NE + 1, NEP = (NE + 1) & ~1;
  static constexpr int NBP = 5, NBD = 4, MEMO = 24, MAXJOBS = 10, NPT = 8;
  static constexpr int MSTR = 18;  // doubles per (rate, code) block of M: 16 + 2, so that the four base codes fall into different banks
  static constexpr unsigned FULL = 0xffffffffu;
  enum : int { CMD_DONE = 0, CMD_PEND, CMD_BUILD_PEND, CMD_DIST1, CMD_DIST4 };
  enum : int { JOB_P = 0, JOB_Q = 1 };
  // block-shared layout (doubles)
  static constexpr int OFF_XQ = 0;                                    // [16][4]
  static constexpr int OFF_EX = OFF_XQ + 64;                          // [MAXJOBS][K] exponentials of the posted builds
  static constexpr int OFF_EXQ = OFF_EX + MAXJOBS * K;                // [NW][K] each warp's copy of the pendant-table job's exponentials
  static constexpr int OFF_JBL = OFF_EXQ + NW * K;                    // [MAXJOBS] branch length of each build job
  static constexpr int OFF_JOB = OFF_JBL + MAXJOBS;                   // ints: [0] cmd, [1] n_jobs, [2] queue slot, [3] points, [4..] dst, kind
  static constexpr int OFF_PT = OFF_JOB + (4 + 2 * MAXJOBS + 1) / 2;  // [NPT] pendant points of the posted pass
  static constexpr int OFF_E = (OFF_PT + NPT + 1) & ~1;               // [NPT][NEP] exp(rate_r * t_b * lamn_j)
  static constexpr int OFF_OUT = OFF_E + NPT * NEP;                   // [NPT] pass results
  static constexpr int OFF_BAD = OFF_OUT + NPT;                       // [NPT] non-finite log terms per point
  static constexpr int OFF_PART = OFF_BAD + NPT;                      // [NW][NPT] per-warp partial log-likelihoods
  static constexpr int OFF_ESUM = OFF_PART + NW * NPT;                // [NW]
  static constexpr int OFF_PD = (OFF_ESUM + NW + 1) & ~1;             // current distal P, then proximal P
  static constexpr int OFF_PQ = OFF_PD + 2 * PM;                      // pendant table q[rs][mask] = sum_{c in mask} P_r(pend)[s][c]
  static constexpr int OFF_MX = OFF_PQ + K * 16;                      // memo: argument, proximal length, value
  static constexpr int OFF_MA = OFF_MX + MEMO;
  static constexpr int OFF_MF = OFF_MA + MEMO;
  static constexpr int OFF_OBJ = OFF_MF + MEMO;                       // thread 0's scalar state between steps
  static constexpr int OBJ_DOUBLES = 64;
  static constexpr int OFF_QS = OFF_OBJ + OBJ_DOUBLES;                // int32 [WS]: site | code << 24 of the candidate's query
  static constexpr int OFF_Y = OFF_QS + WS / 2;                       // y[NY][WS]
  static constexpr int OFF_PB = OFF_Y;                                // distal batch exponentials [NBD][2][K]: y is dead while the distal length moves
  static constexpr int OFF_M = OFF_PB + NBD * 2 * K;                  // M[R][16 codes][MSTR]: also in the dead head of y
  static_assert(WS % 32 == 0, "a warp's sites are all in shared memory or all in the global overflow");
  static_assert(NBD * 2 * K + R * 16 * MSTR <= NY * WS, "the distal batch exponentials and M alias the head of y");
  static_assert((OFF_M % 2) == 0 && (MSTR % 2) == 0, "16-byte alignment of the M rows");
  static_assert((OFF_PD % 2) == 0 && (OFF_E % 2) == 0 && (OFF_Y % 2) == 0 && (PM % 2) == 0, "16-byte alignment of the matrix sets");
  static_assert(NBP <= NPT && NBD <= NPT, "points per pass");

  static __host__ __device__ constexpr size_t smem_bytes() { return sizeof(double) * ((size_t)OFF_Y + (size_t)NY * WS); }

  __device__ __forceinline__ static double *SM() { return reinterpret_cast<double *>(ppb_ttw_smem); }
  __device__ __forceinline__ static int *JOB() { return reinterpret_cast<int *>(SM() + OFF_JOB); }
  __device__ __forceinline__ static int *QS() { return reinterpret_cast<int *>(SM() + OFF_QS); }
  __device__ __forceinline__ static int lane_id() { return threadIdx.x & 31; }
  __device__ __forceinline__ static int warp_id() { return threadIdx.x >> 5; }

  // ---------------------------------------------------------------------------------------------
  // thread 0: the scalar optimiser.  Lives in shared memory between steps; step() runs on a register copy.
  // ---------------------------------------------------------------------------------------------
  struct Scalar : TTMachine<Scalar> {
    using Mach = TTMachine<Scalar>;
    double pend_bl, dist_bl, prox_bl;
    double e_tot_ln2, rate_term, fval, last_ll, cut_bl, start_pend, tol, max_pend;
    int calls, calls_dist, max_iter, feed, attempt, status, optimise, final_pending, last_ok;
    int w_stale, pq_dirty, pd_dirty, memo_n, memo_pending, memo_kind, n_jobs;

    __device__ __forceinline__ void add_job(int kind, double bl, int dst_off) {
      SM()[OFF_JBL + n_jobs] = bl;
      JOB()[4 + n_jobs] = dst_off;
      JOB()[4 + MAXJOBS + n_jobs] = kind;
      ++n_jobs;
    }
    __device__ __forceinline__ void memo_clear(int kind) { memo_n = 0; memo_pending = 0; memo_kind = kind; }
    __device__ __forceinline__ bool memo_get(int kind, double x, double aux, double *f) const {
      if (memo_kind != kind) return false;
      const double *s = SM();
      for (int i = 0; i < memo_n; ++i)
        if (s[OFF_MX + i] == x && s[OFF_MA + i] == aux) { *f = s[OFF_MF + i]; return true; }
      return false;
    }
    // Brent's predictable first calls, several per pass (see threetax_warp.cuh): f(left), f(right), f(start), GSL's
    // repeats of the same three, f(lower + golden*(upper-lower)) and - pendant only - the first golden-section step
    __device__ __forceinline__ void hint_prepare(int kind, double left, double right, double start, double v) {
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
#pragma unroll
        for (int b = 0; b < 4; ++b) { s[OFF_MX + b] = bx[b]; s[OFF_MA + b] = bp[b]; }
        memo_pending = 4;
        w_stale = 1;  // the batch matrices overwrite the head of y
#pragma unroll
        for (int b = 0; b < 4; ++b) { add_job(JOB_P, bx[b], OFF_PB + 2 * b * K); add_job(JOB_P, bp[b], OFF_PB + (2 * b + 1) * K); }
      } else {
        memo_clear(0);
        const double mid = 0.5 * (left + right);
        const double ee = (start < mid) ? right - start : -(start - left);
        const double u1 = start + 0.3819660 * ee;  // Brent's first golden-section step (GSL min/brent.c)
        const double bls[5] = {left, right, start, v, u1};
#pragma unroll
        for (int b = 0; b < 5; ++b) { s[OFF_MX + b] = bls[b]; s[OFF_MA + b] = 0.0; s[OFF_PT + b] = bls[b]; }
        memo_pending = 5;
      }
    }
    __device__ __forceinline__ void hint_commit() {
      double *s = SM();
      for (int i = 0; i < memo_pending; ++i) s[OFF_MF + i] = s[OFF_OUT + i];
      memo_n = memo_pending;
      memo_pending = 0;
    }
    // core.ml:225-239 set_tt_edges / :240-245 tt_edges_default: branch lengths only (sites are never gathered)
    __device__ __forceinline__ void reset_edges(double pendant, double distal) {
      pend_bl = pendant; dist_bl = distal; prox_bl = cut_bl - distal;
      pq_dirty = 1; pd_dirty = 1; w_stale = 1; last_ok = 0;
      memo_clear(-1);
    }
    __device__ __forceinline__ void post(int cmd, int points) { JOB()[0] = cmd; JOB()[1] = n_jobs; JOB()[3] = points; }
    __device__ __forceinline__ void need_pd() {
      if (pd_dirty) { add_job(JOB_P, dist_bl, OFF_PD); add_job(JOB_P, prox_bl, OFF_PD + K); pd_dirty = 0; }
    }
    __device__ __forceinline__ void post_pend(int points) {  // the points are in OFF_PT
      if (w_stale) { need_pd(); w_stale = 0; pq_dirty = 1; post(CMD_BUILD_PEND, points); } else post(CMD_PEND, points);  // writing y overwrites M
    }
    __device__ __forceinline__ void post_dist(bool batch) {
      if (pq_dirty) { add_job(JOB_Q, pend_bl, OFF_PQ); pq_dirty = 0; }
      if (batch) post(CMD_DIST4, NBD);
      else { need_pd(); post(CMD_DIST1, 1); }
    }

    // advance to the next call that needs a pass; false when the candidate is done (*ll_out = its log-likelihood)
    __device__ __forceinline__ bool step(bool first, double *ll_out) {
      double *s = SM();
      if (!first) {  // consume the pass just finished
        if (final_pending) { *ll_out = s[OFF_OUT]; return false; }
        if (feed != 2) { last_ll = s[OFF_OUT]; fval = -last_ll; last_ok = 1; feed = 1; }
      }
      n_jobs = 0;
      for (;;) {
        const int more = optimise ? this->opt_advance(feed, fval) : 0;
        if (more == 2) {  // Brent's predictable calls: hint_prepare posted the batch
          feed = 2;
          if (memo_kind == 1) post_dist(true); else post_pend(NBP);
          return true;
        }
        if (more == 0) {
          if (optimise) {
            status = this->o_status;
            if (st_is_gsl(status) && attempt == 0) {  // core.ml:254-279: retry once from the default branch lengths
              attempt = 1;
              reset_edges(start_pend, cut_bl / 2.0);
              this->opt_init(tol, max_pend, max_iter);
              feed = 0;
              n_jobs = 0;
              continue;
            }
            if (status == ST_EXCEEDED_MAX_ITER) status = ST_OK;
            // Three_tax.get_results (three_tax.ml:85): log_like at the state the last objective call left
            if (last_ok) { *ll_out = last_ll; return false; }
          }
          final_pending = 1;
          post_dist(false);
          return true;
        }
        // ---- TTDriver::objective (three_tax.ml:51-54), calls answered by the memo cost no pass ----
        const double x = this->req_x;
        double val;
        if (this->b_kind == OBJ_PEND) {
          calls++;
          pend_bl = x; pq_dirty = 1;
          if (memo_kind == 1) memo_kind = -1;
          const bool hit_p = memo_get(0, x, 0.0, &val);
          if (this->b_ph == Mach::B_ITER) memo_kind = -1;  // the batch held Brent's start-up calls and its first iterate only
          if (hit_p) { last_ll = val; fval = -val; last_ok = 1; feed = 1; continue; }
          s[OFF_PT] = x;
          post_pend(1);
          feed = 1;
          return true;
        }
        calls++;
        calls_dist++;
        const double prox = (dist_bl + prox_bl) - x;
        if (!(prox >= 0.0)) { fval = __longlong_as_double(0x7ff8000000000000LL); last_ok = 0; feed = 1; continue; }
        const bool hit = memo_get(1, x, prox, &val);
        dist_bl = x; prox_bl = prox; pd_dirty = 1; w_stale = 1;
        if (memo_kind == 0 || this->b_ph == Mach::B_ITER) memo_kind = -1;
        if (hit) { last_ll = val; fval = -val; last_ok = 1; feed = 1; continue; }
        post_dist(false);
        feed = 1;
        return true;
      }
    }
  };
  static_assert(sizeof(Scalar) <= OBJ_DOUBLES * sizeof(double), "scalar state does not fit its shared slot");

  // ---- all threads: the posted jobs.  JOB_P: the K exponentials exp(rho_r bl lambda_k) of a branch length (what is
  // left of Diagd.multi_exp, diagd.ml:79-98, when the vectors are already in the eigenbasis), one warp per job.
  // JOB_Q: the evolved pendant vectors q[rs][mask] of the sixteen codes (lv_of_nuc, nuc_models.ml:49-54 through
  // gemmish_c, linear_c.c:29-90), then M[r][mask][k][k'] = sum_s statd_s q[rs][mask] X[s,k] X[s,k'] ----
  __device__ __forceinline__ static void build_jobs(const EigModel<R> &em, int n) {
    double *s = SM();
    const int *job = JOB();
    const int w = warp_id(), l = lane_id();
    bool has_q = false;
    for (int j = w; j < n; j += NW) {
      if (job[4 + MAXJOBS + j] != JOB_P) continue;
      if (l < K) s[job[4 + j] + l] = exp(em.rates[l >> 2] * s[OFF_JBL + j] * em.lam[l & 3]);
    }
    for (int j = 0; j < n; ++j) {
      if (job[4 + MAXJOBS + j] != JOB_Q) continue;
      has_q = true;
      double *ex = s + OFF_EXQ + w * K;
      for (int t = l; t < K; t += 32) ex[t] = exp(em.rates[t >> 2] * s[OFF_JBL + j] * em.lam[t & 3]);
      __syncwarp();
      double *dst = s + job[4 + j];
      for (int e = threadIdx.x; e < K * 16; e += NT) {
        const int rs = e >> 4, mask = e & 15, r = rs >> 2, i = rs & 3;
        double tot = 0.0;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          double acc = 0.0;
#pragma unroll
          for (int k = 0; k < 4; ++k) acc = fma(ex[r * 4 + k] * em.x[i * 4 + k], em.xit[c * 4 + k], acc);
          tot = ((mask >> c) & 1) ? tot + acc : tot;
        }
        dst[e] = tot;
      }
    }
    __syncthreads();
    if (has_q) {  // block-uniform
      const double *q = s + OFF_PQ;
      for (int e = threadIdx.x; e < R * 16 * 16; e += NT) {
        const int r = e >> 8, mask = (e >> 4) & 15, k = (e >> 2) & 3, k2 = e & 3;
        double acc = 0.0;
#pragma unroll
        for (int i = 0; i < 4; ++i) acc = fma((em.statd[i] * q[(r * 4 + i) * 16 + mask]) * em.x[i * 4 + k], em.x[i * 4 + k2], acc);
        s[OFF_M + (r * 16 + mask) * MSTR + k * 4 + k2] = acc;
      }
      __syncthreads();
    }
  }

  // Running product of per-site likelihoods as (product, exponent sum).  A factor inside [2^-120, 2^120) - every
  // factor of a well-conditioned problem - is multiplied in as it is: one integer range test and ONE DMUL per factor
  // (the chain from site to site is 8 cycles, tools/micro/lat.cu); the callers renorm() at least every 8 factors, so
  // the product stays inside the double range.  Anything else takes the exact path below.
  struct SiteProd {
    double m;
    int e;
    __device__ __forceinline__ void init() { m = 1.0; e = 0; }
    __device__ __forceinline__ void mul(double u, double *bad_slot) {
      int hi = __double2hiint(u);
      if ((unsigned)(hi - ((1023 - 120) << 20)) >= (unsigned)(240 << 20)) {
        // zero / negative / NaN / inf / subnormal / far-out-of-range factor (rare; inline - a call here would spill every
        // value that is live across the site loop): subnormals are rescaled exactly, zero / negative / non-finite values
        // add log(u) = -inf / NaN / +inf to the point's shared slot (linear_c.c:418 sums log(util)) and leave the product
        // alone; what remains is a normal number, whose exponent is split off exactly
        int ef = (hi >> 20) & 0x7ff;
        if (hi < 0 || ef == 0 || ef == 0x7ff) {
          if (u > 0.0 && u < 2.2250738585072014e-308) { u *= 0x1p600; e -= 600; }
          else {
            const double inf = __longlong_as_double(0x7ff0000000000000LL);
            atomicAdd(bad_slot, u == 0.0 ? -inf : (u == inf ? inf : __longlong_as_double(0x7ff8000000000000LL)));
            u = 1.0;
          }
          hi = __double2hiint(u);
          ef = (hi >> 20) & 0x7ff;
        }
        e += ef - 1023;
        u = __hiloint2double((hi & 0x800fffff) | 0x3ff00000, __double2loint(u));
      }
      m *= u;
    }
    __device__ __forceinline__ void renorm() {
      const int hi = __double2hiint(m);
      e += ((hi >> 20) & 0x7ff) - 1023;
      m = __hiloint2double((hi & 0x800fffff) | 0x3ff00000, __double2loint(m));
    }
  };

  // one partial log-likelihood per warp and point: after the butterflies lane b keeps point b and takes its log, so
  // the NB logs of a warp cost one log latency and are off thread 0's serial path; lane b of warp 0 adds the warps'
  // partials in a fixed order (finish)
  template <int NB>
  __device__ __forceinline__ static void pass_end(SiteProd (&lp)[NB]) {
    double *s = SM();
    double mm = 1.0;
    int ee = 0;
#pragma unroll
    for (int b = 0; b < NB; ++b) {
      lp[b].renorm();
      double m = lp[b].m;  // in [1,2): the product of 32 of them stays below 2^32
      int e = lp[b].e;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) m *= __shfl_xor_sync(FULL, m, o);
      e = __reduce_add_sync(FULL, e);
      if (lane_id() == b) { mm = m; ee = e; }
    }
    if (lane_id() < NB) s[OFF_PART + warp_id() * NPT + lane_id()] = log(mm) + (double)ee * kLn2;
    __syncthreads();
  }
  // warp 0, after the barrier that ends a pass: lane b < points adds up log-likelihood b
  __device__ __forceinline__ static void finish(int points) {
    double *s = SM();
    const Scalar *sp = reinterpret_cast<const Scalar *>(s + OFF_OBJ);
    const int b = lane_id();
    if (b < points) {
      double tot = 0.0;
#pragma unroll
      for (int w = 0; w < NW; ++w) tot += s[OFF_PART + w * NPT + b];
      s[OFF_OUT + b] = (tot + s[OFF_BAD + b]) - sp->rate_term + sp->e_tot_ln2;
      s[OFF_BAD + b] = 0.0;
    }
    __syncwarp();
  }

  // site | code << 24 of the query's i-th informative site (clamped: a tail thread re-reads the last site)
  __device__ __forceinline__ static int site_packed(const ECand &c, int i) {
    const int ii = min(i, c.nq - 1);
    return ii < WS ? QS()[ii] : __ldg(c.qsite + ii);
  }
  template <int T>
  __device__ __forceinline__ static void load_dp(const ECand &c, const int (&pk)[T], int r, double (&d)[T][4], double (&p)[T][4]) {
#pragma unroll
    for (int t = 0; t < T; ++t) {
      const int64_t off = (int64_t)(r * 4) * c.Lp + (pk[t] & kSiteMask);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        d[t][k] = __ldg(c.dbase + off + (int64_t)k * c.Lp);
        p[t][k] = __ldg(c.pbase + off + (int64_t)k * c.Lp);
      }
    }
  }

  // The passes that gather darr / parr at the query's sites (lanes = consecutive informative sites: coalesced).
  //   BUILD = false: Three_tax.log_like at NB (distal, proximal) pairs (matrices at pm_off) with the current pendant table
  //   BUILD = true : y of every site from the current distal / proximal matrices (header) -> shared memory, global beyond WS
  // T sites per thread per iteration.  The loads of step (iteration, rate) + 1 are issued before step (iteration, rate) is
  // computed and the site list runs two iterations ahead, so one L2 latency is exposed per pass instead of one per step
  // (r3d: long_scoreboard on the first FMA of the site body was the top stall of the straightforward loop).
  template <bool BUILD, int NB, int T>
  __device__ __forceinline__ static void pass_gather(const EigModel<R> &em, const ECand &c, int pm_off) {
    double *s = SM();
    const double *EV = s + pm_off, *MT = s + OFF_M;  // EV: per point the K exponentials of the distal, then of the proximal length
    double *bad = s + OFF_BAD;
    SiteProd lp[NB];
#pragma unroll
    for (int b = 0; b < NB; ++b) lp[b].init();
    const int tid = threadIdx.x;
    int pk[T], pk1[T];  // this iteration's sites, the next iteration's
    double dn[T][4], pn[T][4];
#pragma unroll
    for (int t = 0; t < T; ++t) { pk[t] = site_packed(c, t * NT + tid); pk1[t] = site_packed(c, NT * T + t * NT + tid); }
    load_dp<T>(c, pk, 0, dn, pn);
    int iter = 0;
    for (int base = 0; base < c.nq; base += NT * T, ++iter) {
      int pk2[T];
#pragma unroll
      for (int t = 0; t < T; ++t) pk2[t] = site_packed(c, base + 2 * NT * T + t * NT + tid);
      double u[T][NB], xq[T][4], cacc[T];
#pragma unroll
      for (int t = 0; t < T; ++t) {
#pragma unroll
        for (int b = 0; b < NB; ++b) u[t][b] = 0.0;
        cacc[t] = 0.0;
        if (BUILD) {
          const int mask = (pk[t] >> kSiteCodeShift) & 15;
          const double2 a = *reinterpret_cast<const double2 *>(s + OFF_XQ + mask * 4), b2 = *reinterpret_cast<const double2 *>(s + OFF_XQ + mask * 4 + 2);
          xq[t][0] = a.x; xq[t][1] = a.y; xq[t][2] = b2.x; xq[t][3] = b2.y;
        }
      }
#pragma unroll 1  // one rate per trip: the loop body stays within the instruction cache (r3g: unrolled, no_instruction stalls 4.3 warps per issue)
      for (int r = 0; r < R; ++r) {
        double d[T][4], p[T][4];
#pragma unroll
        for (int t = 0; t < T; ++t)
#pragma unroll
          for (int k = 0; k < 4; ++k) { d[t][k] = dn[t][k]; p[t][k] = pn[t][k]; }
        {
          const bool last = r + 1 == R;
          int pks[T];
#pragma unroll
          for (int t = 0; t < T; ++t) pks[t] = last ? pk1[t] : pk[t];
          load_dp<T>(c, pks, last ? 0 : r + 1, dn, pn);
        }
        if (!BUILD) {
          double m[T][16];
#pragma unroll
          for (int t = 0; t < T; ++t) {
            const double2 *mp = reinterpret_cast<const double2 *>(MT + (r * 16 + ((pk[t] >> kSiteCodeShift) & 15)) * MSTR);
#pragma unroll
            for (int i = 0; i < 8; ++i) { const double2 v = mp[i]; m[t][2 * i] = v.x; m[t][2 * i + 1] = v.y; }
          }
#pragma unroll
          for (int b = 0; b < NB; ++b) {
            const double2 *ed = reinterpret_cast<const double2 *>(EV + b * 2 * K + r * 4), *ep = reinterpret_cast<const double2 *>(EV + b * 2 * K + K + r * 4);
            const double2 e0 = ed[0], e1 = ed[1], f0 = ep[0], f1 = ep[1];
#pragma unroll
            for (int t = 0; t < T; ++t) {
              const double a0 = e0.x * d[t][0], a1 = e0.y * d[t][1], a2 = e1.x * d[t][2], a3 = e1.y * d[t][3];
              const double b0 = f0.x * p[t][0], b1 = f0.y * p[t][1], b2 = f1.x * p[t][2], b3 = f1.y * p[t][3];
              const double t0 = fma(m[t][3], b3, fma(m[t][2], b2, fma(m[t][1], b1, m[t][0] * b0)));
              const double t1 = fma(m[t][7], b3, fma(m[t][6], b2, fma(m[t][5], b1, m[t][4] * b0)));
              const double t2 = fma(m[t][11], b3, fma(m[t][10], b2, fma(m[t][9], b1, m[t][8] * b0)));
              const double t3 = fma(m[t][15], b3, fma(m[t][14], b2, fma(m[t][13], b1, m[t][12] * b0)));
              u[t][b] = fma(a3, t3, fma(a2, t2, fma(a1, t1, fma(a0, t0, u[t][b]))));
            }
          }
        } else {
          double ww[T][4];
          {
            const double2 *ed = reinterpret_cast<const double2 *>(EV + r * 4), *ep = reinterpret_cast<const double2 *>(EV + K + r * 4);
            const double2 e0 = ed[0], e1 = ed[1], f0 = ep[0], f1 = ep[1];
#pragma unroll
            for (int t = 0; t < T; ++t) {
              const double a0 = e0.x * d[t][0], a1 = e0.y * d[t][1], a2 = e1.x * d[t][2], a3 = e1.y * d[t][3];
              const double b0 = f0.x * p[t][0], b1 = f0.y * p[t][1], b2 = f1.x * p[t][2], b3 = f1.y * p[t][3];
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const double de = fma(em.x[i * 4 + 3], a3, fma(em.x[i * 4 + 2], a2, fma(em.x[i * 4 + 1], a1, em.x[i * 4] * a0)));
                const double pe = fma(em.x[i * 4 + 3], b3, fma(em.x[i * 4 + 2], b2, fma(em.x[i * 4 + 1], b1, em.x[i * 4] * b0)));
                ww[t][i] = (em.statd[i] * pe) * de;  // ((statd*x)*y), x = prox, y = dist (three_tax.ml:31-36)
              }
            }
          }
#pragma unroll
          for (int t = 0; t < T; ++t) {
            double t0 = ww[t][0] * em.xz[0];
#pragma unroll
            for (int k = 1; k < 4; ++k) t0 = fma(ww[t][k], em.xz[k], t0);
            cacc[t] = fma(t0, xq[t][0], cacc[t]);
            double yv[3];
#pragma unroll
            for (int j = 0; j < 3; ++j) {
              double tj = ww[t][0] * em.xn[j];
#pragma unroll
              for (int k = 1; k < 4; ++k) tj = fma(ww[t][k], em.xn[k * 3 + j], tj);
              yv[j] = tj * xq[t][1 + j];
            }
            const int i = base + t * NT + tid;
            if (i < c.nq) {
              if (i < WS) {  // warp-uniform: WS is a multiple of 32
                double *yp = s + OFF_Y + (1 + r * 3) * WS + i;
#pragma unroll
                for (int j = 0; j < 3; ++j) yp[j * WS] = yv[j];
              } else {
                double *yp = c.wg + (int64_t)(1 + r * 3) * c.Lp + i;
#pragma unroll
                for (int j = 0; j < 3; ++j) yp[(int64_t)j * c.Lp] = yv[j];
              }
            }
          }
        }
      }
#pragma unroll
      for (int t = 0; t < T; ++t) {
        const int i = base + t * NT + tid;
        if (i < c.nq) {
          if (BUILD) {
            if (i < WS) s[OFF_Y + i] = cacc[t]; else c.wg[i] = cacc[t];
          } else {
#pragma unroll
            for (int b = 0; b < NB; ++b) lp[b].mul(u[t][b], bad + b);
          }
        }
        pk[t] = pk1[t]; pk1[t] = pk2[t];
      }
      if (!BUILD && (iter & (8 / T - 1)) == 8 / T - 1) {  // every 8 factors (SiteProd)
#pragma unroll
        for (int b = 0; b < NB; ++b) lp[b].renorm();
      }
    }
    if (!BUILD) pass_end<NB>(lp);
    // BUILD: every thread reads back only what it wrote itself - no barrier before the pendant pass that follows
  }

  // Three_tax.log_like at NB pendant lengths from the cached y: u = y0 + sum_j E_j y_j
  template <int NB>
  __device__ __forceinline__ static void pend_site(const double (&y)[NY], const double (&E1)[NE], SiteProd (&lp)[NB], double *bad) {
    const double *s = SM();
    if (NB == 1) {
      double u0 = y[0], u1 = 0.0;  // two chains
#pragma unroll
      for (int j = 0; j < NE; ++j) {
        if (j & 1) u1 = fma(E1[j], y[1 + j], u1); else u0 = fma(E1[j], y[1 + j], u0);
      }
      lp[0].mul(u0 + u1, bad);
    } else {
#pragma unroll
      for (int b = 0; b < NB; ++b) {
        const double2 *Eb = reinterpret_cast<const double2 *>(s + OFF_E + b * NEP);  // warp-uniform: broadcast loads
        double u0 = y[0], u1 = 0.0;
#pragma unroll
        for (int j2 = 0; j2 < NE / 2; ++j2) {
          const double2 e = Eb[j2];
          u0 = fma(e.x, y[1 + 2 * j2], u0);
          u1 = fma(e.y, y[2 + 2 * j2], u1);
        }
        if (NE & 1) u0 = fma(s[OFF_E + b * NEP + NE - 1], y[NE], u0);
        lp[b].mul(u0 + u1, bad + b);
      }
    }
  }
  template <int NB>
  __device__ __forceinline__ static void pass_pend(const ECand &c) {
    double *s = SM();
    const double *Y = s + OFF_Y;
    double *bad = s + OFF_BAD;
    SiteProd lp[NB];
#pragma unroll
    for (int b = 0; b < NB; ++b) lp[b].init();
    const int tid = threadIdx.x;
    // this thread's first site beyond the shared-memory part: its y is requested first and used last
    const int io = tid + NT * ((WS - tid + NT - 1) / NT);
    constexpr int PF = NB == 1 ? 1 : 0;
    double yo[PF ? PF : 1][NY];
#pragma unroll
    for (int o = 0; o < PF; ++o)
      if (io + o * NT < c.nq) {
        const double *yp = c.wg + io + o * NT;
#pragma unroll
        for (int j = 0; j < NY; ++j) yo[o][j] = __ldcg(yp + (int64_t)j * c.Lp);
      }
    double E1[NE];
#pragma unroll
    for (int j = 0; j < NE; ++j) E1[j] = NB == 1 ? s[OFF_E + j] : 0.0;
    const int n_sm = min(c.nq, WS);
    int i = tid, nf = 0;  // nf: factors since the last renorm (SiteProd: at most 8)
    auto tick = [&](int k) {
      nf += k;
      if (nf >= 7) {
        nf = 0;
#pragma unroll
        for (int b = 0; b < NB; ++b) lp[b].renorm();
      }
    };
    if (NB == 1) {
      for (; i + NT < n_sm; i += 2 * NT) {  // two sites per trip: independent chains
        double ya[NY], yb[NY];
#pragma unroll
        for (int j = 0; j < NY; ++j) { ya[j] = Y[j * WS + i]; yb[j] = Y[j * WS + i + NT]; }
        pend_site<NB>(ya, E1, lp, bad);
        pend_site<NB>(yb, E1, lp, bad);
        tick(2);
      }
    }
    for (; i < n_sm; i += NT) {
      double y[NY];
#pragma unroll
      for (int j = 0; j < NY; ++j) y[j] = Y[j * WS + i];
      pend_site<NB>(y, E1, lp, bad);
      tick(1);
    }
#pragma unroll
    for (int o = 0; o < PF; ++o)
      if (io + o * NT < c.nq) { pend_site<NB>(yo[o], E1, lp, bad); tick(1); }
    for (int i2 = io + PF * NT; i2 < c.nq; i2 += NT) {
      double y[NY];
      const double *yp = c.wg + i2;
#pragma unroll
      for (int j = 0; j < NY; ++j) y[j] = __ldcg(yp + (int64_t)j * c.Lp);
      pend_site<NB>(y, E1, lp, bad);
      tick(1);
    }
    pass_end<NB>(lp);
  }
};

// alpha[node][r*4 + k][site] = sum_c Xit[c][k] * glv[node][r*4 + c][site]: darr / parr in the eigenbasis of the rate matrix
// (exponents unchanged - the transform is linear per site and rate).  One thread per (node, rate, site).
struct XitMat { double v[16]; };
__global__ void __launch_bounds__(256)
eigen_glv_kernel(const double *__restrict__ src, double *__restrict__ dst, int64_t n_planes4 /* nodes * R */, int64_t Lp, const XitMat xit) {
  const int64_t site = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (site >= Lp) return;
  for (int64_t pl = blockIdx.y; pl < n_planes4; pl += gridDim.y) {
    const double *a = src + pl * 4 * Lp + site;
    const double d0 = __ldg(a), d1 = __ldg(a + Lp), d2 = __ldg(a + 2 * Lp), d3 = __ldg(a + 3 * Lp);
    double *o = dst + pl * 4 * Lp + site;
#pragma unroll
    for (int k = 0; k < 4; ++k) o[(int64_t)k * Lp] = fma(xit.v[12 + k], d3, fma(xit.v[8 + k], d2, fma(xit.v[4 + k], d1, xit.v[k] * d0)));
  }
}

// Persistent kernel: every BLOCK pulls candidates (in `order`: length class, then edge) until n are done.
template <int R, int NW, int MINB, int WS>
__global__ void __launch_bounds__(NW * 32, MINB)
tt_eigen_kernel(const __grid_constant__ EigModel<R> em, GlvArrays g, const double *__restrict__ bl, const int32_t *__restrict__ qsite,
                const int32_t *__restrict__ nq, const CandIn *__restrict__ in, const int32_t *__restrict__ order,
                CandOut *__restrict__ out, int n, TTParams prm, double *__restrict__ wbuf, int *counter,
                unsigned long long *__restrict__ work) {
  using TT = TTEig<R, NW, WS>;
  using Scalar = typename TT::Scalar;
  constexpr int TG = MINB * NW <= 12 ? 2 : 1;  // sites per thread per step of the one-point gather passes: 2 when the register budget allows (<= 384 threads/SM)
  double *sm = TT::SM();
  for (int i = threadIdx.x; i < 64; i += TT::NT) sm[TT::OFF_XQ + i] = em.xq[i];
  for (int i = threadIdx.x; i < TT::NPT; i += TT::NT) sm[TT::OFF_BAD + i] = 0.0;
  Scalar *sp = reinterpret_cast<Scalar *>(sm + TT::OFF_OBJ);
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
    ECand c;
    c.qsite = qsite + (int64_t)cin.q * g.Lp;
    c.nq = nq[cin.q];
    c.dbase = g.ea[0] + (int64_t)cin.loc * g.a_stride;  // the eigenbasis copies of darr / parr (eigen_glv_kernel)
    c.pbase = g.ea[1] + (int64_t)cin.loc * g.a_stride;
    c.Lp = g.Lp;
    c.wg = wbuf + (int64_t)blockIdx.x * TT::K * g.Lp;
    {  // the query's site list into shared memory; total_twoexp of the distal and proximal GLVs over it (gstar_support.ml:43-49)
      const int32_t *ed = g.e[0] + (int64_t)cin.loc * g.Lp, *ep = g.e[1] + (int64_t)cin.loc * g.Lp;
      int *qs = TT::QS();
      long long es = 0;
      for (int base = threadIdx.x; base < c.nq; base += 4 * TT::NT) {  // four sites per thread per trip: the loads go out together
        int pk[4], ev[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) pk[u] = base + u * TT::NT < c.nq ? __ldg(c.qsite + base + u * TT::NT) : -1;
#pragma unroll
        for (int u = 0; u < 4; ++u) ev[u] = pk[u] >= 0 ? __ldg(ed + (pk[u] & kSiteMask)) + __ldg(ep + (pk[u] & kSiteMask)) : 0;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          if (pk[u] >= 0 && base + u * TT::NT < WS) qs[base + u * TT::NT] = pk[u];
          es += ev[u];
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) es += __shfl_xor_sync(0xffffffffu, es, o);
      if (lane == 0) sm[TT::OFF_ESUM + warp] = (double)es;  // exact: |es| < 2^53
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      Scalar sc;
      double tot = 0.0;
#pragma unroll
      for (int w = 0; w < NW; ++w) tot += sm[TT::OFF_ESUM + w];
      sc.e_tot_ln2 = kLn2 * tot;
      sc.rate_term = (double)((float)c.nq) * log((double)((float)R));  // linear_c.c:421
      sc.cut_bl = bl[cin.loc];
      sc.start_pend = prm.start_pend; sc.tol = prm.tol; sc.max_pend = prm.max_pend; sc.max_iter = prm.max_iter;
      sc.calls = 0; sc.calls_dist = 0; sc.feed = 0; sc.attempt = 0; sc.status = ST_OK; sc.final_pending = 0; sc.fval = 0.0;
      sc.last_ll = 0.0; sc.n_jobs = 0;
      sc.optimise = prm.mode == TT_MODE_OPTIMIZE;
      sc.reset_edges(prm.use_defaults ? prm.start_pend : cin.pend0, prm.use_defaults ? sc.cut_bl / 2.0 : cin.dist0);
      if (sc.optimise) sc.opt_init(prm.tol, prm.max_pend, prm.max_iter);
      *sp = sc;
    }
    __syncthreads();
    double ll = 0.0;
    for (bool first = true;; first = false) {
#ifdef PPB_TT_PROFILE
      const long long pf_a = clock64();
#endif
      if (warp == 0) {
        if (!first) TT::finish(job[3]);
        if (lane == 0) {  // the only thread that touches *sp between barriers
          if (!sp->step(first, &ll)) { job[0] = TT::CMD_DONE; job[1] = 0; }
        }
        __syncwarp();
        const int cmd0 = job[0];
        if (cmd0 == TT::CMD_PEND || cmd0 == TT::CMD_BUILD_PEND) {  // the posted points' exponentials, one lane each
          const int ne = job[3] * TT::NE;
          for (int t = lane; t < ne; t += 32) {
            const int b = t / TT::NE, e = t - b * TT::NE, r = e / 3, j = e - 3 * r;
            sm[TT::OFF_E + b * TT::NEP + e] = exp(em.rates[r] * sm[TT::OFF_PT + b] * em.lamn[j]);
          }
        }
      }
#ifdef PPB_TT_PROFILE
      const long long pf_b = clock64();
      pf_scalar += pf_b - pf_a;
#endif
      __syncthreads();  // the command is posted
      const int cmd = job[0], nj = job[1];
      if (cmd == TT::CMD_DONE) break;
      if (nj > 0) TT::build_jobs(em, nj);
#ifdef PPB_TT_PROFILE
      const long long pf_c = clock64();
      pf_build += pf_c - pf_b; pf_nb += nj > 0; pf_jobs += nj;
#endif
      switch (cmd) {
        case TT::CMD_DIST1: TT::template pass_gather<false, 1, TG>(em, c, TT::OFF_PD); break;
        case TT::CMD_DIST4: TT::template pass_gather<false, TT::NBD, TG>(em, c, TT::OFF_PB); break;
        case TT::CMD_BUILD_PEND:
          TT::template pass_gather<true, 1, TG>(em, c, TT::OFF_PD);
          // fall through: every thread evaluates the sites it has just written
        default:
          if (job[3] == 1) TT::template pass_pend<1>(c); else TT::template pass_pend<TT::NBP>(c);
          break;
      }
#ifdef PPB_TT_PROFILE
      pf_pass[cmd & 7] += clock64() - pf_c; pf_n[cmd & 7]++;
#endif
    }
    if (threadIdx.x == 0) {
      CandOut o;
      o.marg = __longlong_as_double(0x7ff8000000000000LL);
      o.ll = ll; o.pend = sp->pend_bl; o.dist = sp->dist_bl; o.status = sp->status; o.calls = sp->calls;
      out[ci] = o;
      const unsigned long long nd = sp->calls_dist, np = sp->calls - sp->calls_dist, ns = c.nq;
      atomicAdd(work + 0, np); atomicAdd(work + 1, nd); atomicAdd(work + 2, np * ns); atomicAdd(work + 3, nd * ns);
#ifdef PPB_TT_PROFILE
      atomicAdd(work + 4, (unsigned long long)pf_scalar); atomicAdd(work + 5, (unsigned long long)pf_build);
      atomicAdd(work + 6, (unsigned long long)pf_nb); atomicAdd(work + 7, (unsigned long long)pf_jobs);
      for (int k = 1; k < 7; ++k) { atomicAdd(work + 8 + k, (unsigned long long)pf_pass[k]); atomicAdd(work + 16 + k, (unsigned long long)pf_n[k]); }
      atomicAdd(work + 24, 1ull); atomicAdd(work + 25, (unsigned long long)(clock64() - pf_t0));
#endif
    }
  }
}

}  // namespace ppb
