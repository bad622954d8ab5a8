// Second stage on device: the three-taxon tree, its likelihood, alternating Brent optimisation
// of pendant and distal branch lengths, and the --pp marginal likelihood.
//   Three_tax          pplacer_src/three_tax.ml:30-83 (log_like, set_*_bl, optimize), :124-215 (pp)
//   Glv_edge.set_bl    pplacer_src/glv_edge.ml:29-35   (re-evolve on every branch-length change)
//   Model.log_like3    gmix_model.ml:298-304 -> ten_log_like3_c linear_c.c:367-424
//                      gcat_model.ml:317-323 -> mat_log_like3_c linear_c.c:170-216
//   Model.evolve_into  gmix_model.ml:308-315 / gcat_model.ml:327-345 ; Diagd.multi_exp diagd.ml:79-98
//   Minimization.brent pplacer_src/minimization.ml:11-41,83-101 over GSL's gsl_min_fminimizer_brent
//   Integration        pplacer_src/integration.ml:27-31 over GSL's gsl_integration_qng
//   core.ml:225-279    set_tt_edges / tt_edges_default / (safe_)ml_optimize_location
//
// One thread block owns one (query, edge) candidate at a time (persistent blocks pull candidates
// from an atomic queue).  Threads stride the query's informative sites; the scalar Brent / QNG
// state machines are executed redundantly and identically by every thread (all inputs are
// block-uniform and the block reductions are deterministic), so no broadcast is needed.
//
// Restructurings relative to the reference (same arithmetic per site):
//   * the distal / proximal GLVs are never gathered into masked copies (Glv.mask_into,
//     core.ml:229-233): the kernel reads darr/parr planes at the query's site list directly;
//   * the evolved pendant GLV is a per-character table (sum of columns of P(pendant));
//   * while only the pendant length moves, w = (statd*prox_ev)*dist_ev is cached per site.
#pragma once
#include "device_common.cuh"
#include "qng_tables.h"

namespace ppb {

enum : int {
  ST_OK = 0,
  ST_GSL_EINVAL = 4,
  ST_GSL_EBADFUNC = 9,
  ST_GSL_EBADTOL = 13,
  ST_GSL_ETOL = 14,
  ST_EXCEEDED_MAX_ITER = 100,
  ST_FIND_START_FAILURE = 101,
  ST_INVALID_START = 102,
  ST_ASSERT = 106
};
__host__ __device__ inline bool st_is_gsl(int st) {
  return st == ST_GSL_EINVAL || st == ST_GSL_EBADFUNC || st == ST_GSL_ETOL || st == ST_GSL_EBADTOL;
}

struct CandIn {
  int32_t q, loc;
  double pend0, dist0;
};
struct CandOut {
  double ll, pend, dist, marg;
  int32_t status, calls;
};

enum : int { TT_MODE_OPTIMIZE = 0, TT_MODE_LOGLIKE = 1, TT_MODE_MARGINAL = 2 };

struct TTParams {
  double start_pend, max_pend, tol, pp_rel_err, prior_mean;
  const double *prior_means;  // informative prior (per edge) or null
  int max_iter, mode, prior_kind, use_defaults;
  int flags;  // warp engine: bit 0 = recompute w in every pendant pass instead of caching it, bit 1 = L2 prefetch
};

__constant__ double c_qng_x1[5], c_qng_w10[5], c_qng_x2[5], c_qng_w21a[5], c_qng_w21b[6], c_qng_x3[11],
    c_qng_w43a[10], c_qng_w43b[12], c_qng_x4[22], c_qng_w87a[21], c_qng_w87b[23];

constexpr double kSqrtDblEps = 1.4901161193847656e-08;
constexpr double kDblEps = 2.2204460492503131e-16;


// ---------------------------------------------------------------------------------------------
// The scalar part of the second stage, shared by the block-per-candidate engine below and the
// warp-per-candidate engine of threetax_warp.cuh (CRTP: D supplies the likelihood passes).
// D must provide:
//   double f_pend(double pend)               set the pendant length, return Three_tax.log_like
//   double f_dist(double dist, double prox)  set distal / proximal lengths, return log_like
//   void   set_dist_prox(double, double), void cache_w(), double eval_full_cache()
//   void   hint_brent(kind, left, right, start, v)   points Brent is about to evaluate (may batch)
//   void   hint_begin() / hint_add(x) / hint_flush() pendant points QNG is about to evaluate
// Every thread of the owning warp / block executes this code redundantly on identical values.
// ---------------------------------------------------------------------------------------------
struct Prior {
  int kind;
  double mean;
  __device__ double operator()(double x) const {  // core.ml:87-92; gsl_ran_exponential_pdf
    if (kind == 0) return 1.0;
    if (x < 0.0) return 0.0;
    return exp(-x / mean) / mean;
  }
};
enum : int { OBJ_PEND = 0, OBJ_DIST = 1, OBJ_UPPER = 2 };

template <class D>
struct TTDriver {
  double pend_bl, dist_bl, prox_bl;
  int calls;
  Prior prior;
  double target_ll, base_ll, upper_limit, pp_rel_err;
  int pp_status;

  __device__ __forceinline__ D &self() { return *static_cast<D *>(this); }

  __device__ __noinline__ double objective(int kind, double v) {
    if (kind == OBJ_PEND) {  // three_tax.ml:51-54 with set_pend_bl
      calls++;
      return -self().f_pend(v);
    }
    if (kind == OBJ_DIST) {  // three_tax.ml:41-45,51-54 with set_dist_bl
      calls++;
      const double prox = (dist_bl + prox_bl) - v;
      if (!(prox >= 0.0)) return __longlong_as_double(0x7ff8000000000000LL);
      return -self().f_dist(v, prox);
    }
    // OBJ_UPPER: three_tax.ml:144-150
    const double like = self().f_pend(v) + log(prior(v));
    const double x = fabs(target_ll - like);
    return (x == __longlong_as_double(0x7ff0000000000000LL)) ? 100000.0 : x;
  }

  // Minimization.brent (minimization.ml:83-101) with bisection_start_finder (:24-41) and GSL's
  // gsl_min_fminimizer_set + brent (GSL 2.7 min/fsolver.c, min/brent.c).
  __device__ __noinline__ int brent(int kind, double raw_start, double left, double right, double tolerance,
                                    int max_iters, double *xmin_out) {
    auto f = [this, kind](double v) { return objective(kind, v); };
    if (left >= raw_start || raw_start >= right) return ST_INVALID_START;
    const double golden = 0.3819660;
    // f(left), f(right), f(raw_start) and GSL's first interior point are known in advance
    self().hint_brent(kind, left, right, raw_start, left + golden * (right - left));
    const double lefty = f(left), righty = f(right);
    const double smaller = (lefty < righty) ? left : right;
    double miny = fmin(lefty, righty);
    if (lefty != lefty || righty != righty) miny = __longlong_as_double(0x7ff8000000000000LL);
    double start, prev = raw_start;
    for (int iter = 1;; ++iter) {
      const double pv = f(prev);
      if (pv < miny) { start = prev; break; }
      if (fabs(prev - smaller) < tolerance) { *xmin_out = smaller; return ST_OK; }  // FoundMin
      if (iter > max_iters) return ST_FIND_START_FAILURE;
      prev = (prev + smaller) / 2.0;
    }
    double x_minimum = start, x_lower = left, x_upper = right;
    double f_lower = f(x_lower);
    if (!isfinite(f_lower)) return ST_GSL_EBADFUNC;
    double f_upper = f(x_upper);
    if (!isfinite(f_upper)) return ST_GSL_EBADFUNC;
    double f_minimum = f(x_minimum);
    if (!isfinite(f_minimum)) return ST_GSL_EBADFUNC;
    if (x_lower > x_upper) return ST_GSL_EINVAL;
    if (x_minimum >= x_upper || x_minimum <= x_lower) return ST_GSL_EINVAL;
    if (f_minimum >= f_lower || f_minimum >= f_upper) return ST_GSL_EINVAL;
    double v = x_lower + golden * (x_upper - x_lower), wv = v, d = 0.0, e = 0.0;
    const double f_vw = f(v);
    if (!isfinite(f_vw)) return ST_GSL_EBADFUNC;
    double f_v = f_vw, f_w = f_vw;
    for (int step = 1;; ++step) {
      if (step > max_iters) return ST_EXCEEDED_MAX_ITER;
      const double x_left = x_lower, x_right = x_upper, z = x_minimum, f_z = f_minimum;
      double dd = e, ee = d;  // GSL: d = state->e; e = state->d
      double u;
      const double w_lower = z - x_left, w_upper = x_right - z;
      const double tol = kSqrtDblEps * fabs(z);
      double p = 0.0, q = 0.0, r = 0.0;
      const double midpoint = 0.5 * (x_left + x_right);
      if (fabs(ee) > tol) {
        r = (z - wv) * (f_z - f_v);
        q = (z - v) * (f_z - f_w);
        p = (z - v) * q - (z - wv) * r;
        q = 2.0 * (q - r);
        if (q > 0.0) p = -p; else q = -q;
        r = ee;
        ee = dd;
      }
      if (fabs(p) < fabs(0.5 * q * r) && p < q * w_lower && p < q * w_upper) {
        const double t2 = 2.0 * tol;
        dd = p / q;
        u = z + dd;
        if ((u - x_left) < t2 || (x_right - u) < t2) dd = (z < midpoint) ? tol : -tol;
      } else {
        ee = (z < midpoint) ? x_right - z : -(z - x_left);
        dd = golden * ee;
      }
      if (fabs(dd) >= tol) u = z + dd; else u = z + ((dd > 0.0) ? tol : -tol);
      e = ee;
      d = dd;
      const double f_u = f(u);
      if (!isfinite(f_u)) return ST_GSL_EBADFUNC;
      if (f_u <= f_z) {
        if (u < z) { x_upper = z; f_upper = f_z; } else { x_lower = z; f_lower = f_z; }
        v = wv; f_v = f_w; wv = z; f_w = f_z; x_minimum = u; f_minimum = f_u;
      } else {
        if (u < z) { x_lower = u; f_lower = f_u; } else { x_upper = u; f_upper = f_u; }
        if (f_u <= f_w || wv == z) { v = wv; f_v = f_w; wv = u; f_w = f_u; }
        else if (f_u <= f_v || v == z || v == wv) { v = u; f_v = f_u; }
      }
      if (x_upper - x_lower < tolerance) { *xmin_out = x_minimum; return ST_OK; }
    }
  }

  // Three_tax.optimize (three_tax.ml:70-83)
  __device__ __noinline__ int optimize(double tolerance, double max_pend, int max_iter) {
    double prev_pend = pend_bl, prev_dist = dist_bl;
    for (int step = 1;; ++step) {
      if (step > max_iter) return ST_EXCEEDED_MAX_ITER;
      double curr_pend = pend_bl, curr_dist = dist_bl;
      if (1e-8 != max_pend) {
        self().cache_w();  // w = (statd*prox_ev)*dist_ev per site, for the pendant-only evaluations
        const int st = brent(OBJ_PEND, pend_bl, 1e-8, max_pend, tolerance, 100, &curr_pend);
        if (st != ST_OK) return st;
      }
      const double cut = dist_bl + prox_bl;
      if (0.0 != cut) {
        const int st = brent(OBJ_DIST, dist_bl, 0.0, cut, tolerance, 100, &curr_dist);
        if (st != ST_OK) return st;
      }
      if (fabs(prev_pend - curr_pend) > tolerance || fabs(prev_dist - curr_dist) > tolerance) {
        prev_pend = curr_pend;
        prev_dist = curr_dist;
      } else
        return ST_OK;
    }
  }

  // ---- --pp : Three_tax.calc_marg_prob (three_tax.ml:169-215) ----
  __device__ int find_upper_limit(double max_pend, double orig_ll, double *out) {  // :124-159
    const double orig_pend = pend_bl;
    if (prior(orig_pend) == 0.0) {
      double curr = orig_pend;
      for (;;) {
        const double next = curr / 2.0;
        if (prior(next) > exp(-10.0)) { *out = curr; return ST_OK; }
        curr = next;
      }
    }
    target_ll = -10.0 + orig_ll + log(prior(orig_pend));
    const double start = orig_pend + ((max_pend - orig_pend) / 2.0);
    const int st = brent(OBJ_UPPER, start, orig_pend, max_pend, 0.02, 10, out);
    if (st == ST_GSL_EINVAL) { *out = max_pend; return ST_OK; }
    return st;
  }

  static __device__ double rescale_error(double err, double result_abs, double result_asc) {
    err = fabs(err);
    if (result_asc != 0.0 && err != 0.0) {
      const double scale = pow((200.0 * err / result_asc), 1.5);
      err = (scale < 1.0) ? result_asc * scale : result_asc;
    }
    if (result_abs > 2.2250738585072014e-308 / (50.0 * kDblEps)) {
      const double min_err = 50.0 * kDblEps * result_abs;
      if (min_err > err) err = min_err;
    }
    return err;
  }

  // integrands: inner (three_tax.ml:187-189) and outer (:192-197)
  // the inner (pendant) integral: an engine may replace it by a form that evaluates a whole rule per pass
  __device__ int qng_inner(double a, double b, double epsabs, double epsrel, double *result) { return qng<false>(a, b, epsabs, epsrel, result); }
  // an engine that evaluates a whole rule's abscissae in one pass may also keep the integrand values (same arithmetic)
  __device__ bool integrand_memo(double, double *) { return false; }
  __device__ __noinline__ double integrand_inner(double pend) {
    double v;
    if (self().integrand_memo(pend, &v)) return v;
    return exp(self().f_pend(pend) - base_ll) * prior(pend);
  }
  __device__ __noinline__ double integrand_outer(double dist) {
    if (pp_status) return 0.0;  // the reference's exception is already unwinding
    const double prox = (dist_bl + prox_bl) - dist;
    if (!(prox >= 0.0)) { pp_status = ST_ASSERT; return 0.0; }
    self().set_dist_prox(dist, prox);
    self().cache_w();
    double res;
    const int st = self().qng_inner(0.0, upper_limit, 0.0, pp_rel_err, &res);
    if (st) pp_status = st;
    return res;
  }

  // gsl_integration_qng (GSL 2.7 integration/qng.c).  The abscissae of each
  // rule are announced to the engine (hint_*) before the rule is evaluated, in evaluation order.
  template <bool OUTER>
  __device__ __noinline__ int qng(double a, double b, double epsabs, double epsrel, double *result) {
    auto f = [this](double v) { return OUTER ? integrand_outer(v) : integrand_inner(v); };
    double fv1[5], fv2[5], fv3[5], fv4[5], savfun[21];
    const double half_length = 0.5 * (b - a), abs_half_length = fabs(half_length), center = 0.5 * (b + a);
    if (!OUTER) {
      self().hint_begin();
      self().hint_add(center);
      for (int k = 0; k < 5; ++k) { const double ab = half_length * c_qng_x1[k]; self().hint_add(center + ab); self().hint_add(center - ab); }
      for (int k = 0; k < 5; ++k) { const double ab = half_length * c_qng_x2[k]; self().hint_add(center + ab); self().hint_add(center - ab); }
      self().hint_flush();
    }
    const double f_center = f(center);
    if (epsabs <= 0.0 && (epsrel < 50.0 * kDblEps || epsrel < 0.5e-28)) { *result = 0.0; return ST_GSL_EBADTOL; }
    double res10 = 0.0, res21 = c_qng_w21b[5] * f_center, resabs = c_qng_w21b[5] * fabs(f_center);
    for (int k = 0; k < 5; ++k) {
      const double ab = half_length * c_qng_x1[k];
      const double f1 = f(center + ab), f2 = f(center - ab), fv = f1 + f2;
      res10 += c_qng_w10[k] * fv;
      res21 += c_qng_w21a[k] * fv;
      resabs += c_qng_w21a[k] * (fabs(f1) + fabs(f2));
      savfun[k] = fv; fv1[k] = f1; fv2[k] = f2;
    }
    for (int k = 0; k < 5; ++k) {
      const double ab = half_length * c_qng_x2[k];
      const double f1 = f(center + ab), f2 = f(center - ab), fv = f1 + f2;
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
    double err = rescale_error((res21 - res10) * half_length, resabs, resasc);
    if (err < epsabs || err < epsrel * fabs(result_kronrod)) { *result = result_kronrod; return ST_OK; }
    if (!OUTER) {
      self().hint_begin();
      for (int k = 0; k < 11; ++k) { const double ab = half_length * c_qng_x3[k]; self().hint_add(center + ab); self().hint_add(center - ab); }
      self().hint_flush();
    }
    double res43 = c_qng_w43b[11] * f_center;
    for (int k = 0; k < 10; ++k) res43 += savfun[k] * c_qng_w43a[k];
    for (int k = 0; k < 11; ++k) {
      const double ab = half_length * c_qng_x3[k];
      const double fv = f(center + ab) + f(center - ab);
      res43 += fv * c_qng_w43b[k];
      savfun[k + 10] = fv;
    }
    result_kronrod = res43 * half_length;
    err = rescale_error((res43 - res21) * half_length, resabs, resasc);
    if (err < epsabs || err < epsrel * fabs(result_kronrod)) { *result = result_kronrod; return ST_OK; }
    if (!OUTER) {
      self().hint_begin();
      for (int k = 0; k < 22; ++k) { const double ab = half_length * c_qng_x4[k]; self().hint_add(center + ab); self().hint_add(center - ab); }
      self().hint_flush();
    }
    double res87 = c_qng_w87b[22] * f_center;
    for (int k = 0; k < 21; ++k) res87 += savfun[k] * c_qng_w87a[k];
    for (int k = 0; k < 22; ++k) {
      const double ab = half_length * c_qng_x4[k];
      res87 += c_qng_w87b[k] * (f(center + ab) + f(center - ab));
    }
    result_kronrod = res87 * half_length;
    err = rescale_error((res87 - res43) * half_length, resabs, resasc);
    *result = result_kronrod;
    if (err < epsabs || err < epsrel * fabs(result_kronrod)) return ST_OK;
    return ST_GSL_ETOL;
  }

  __device__ int calc_marg_prob(double rel_err, double max_pend, double *out) {
    pp_rel_err = rel_err;
    base_ll = self().eval_full_cache();
    const double cut_bl = dist_bl + prox_bl;
    int st = find_upper_limit(max_pend, base_ll, &upper_limit);
    if (st) return st;
    for (int n_exceptions = 0;; ++n_exceptions) {
      if (n_exceptions >= 10) { *out = base_ll; return ST_OK; }
      pp_status = 0;
      double val;
      if (fabs(cut_bl - 0.0) < 1e-5) {  // cut_bl =~ 0.  (ppatteries.ml:96)
        self().set_dist_prox(cut_bl / 2.0, (dist_bl + prox_bl) - cut_bl / 2.0);
        self().cache_w();
        pp_status = self().qng_inner(0.0, upper_limit, 0.0, rel_err, &val);
      } else {
        double res;
        const int so = qng<true>(0.0, cut_bl, 0.0, rel_err, &res);
        if (!pp_status && so) pp_status = so;
        val = res / cut_bl;
      }
      if (pp_status == ST_GSL_ETOL) { upper_limit = upper_limit / 2.0; continue; }
      if (pp_status) return pp_status;
      *out = base_ll + log(val);
      return ST_OK;
    }
  }
};

// ---------------------------------------------------------------------------------------------
// Three_tax.optimize as a RESUMABLE state machine: the same statements as TTDriver::brent /
// ::optimize (threetax.cuh; minimization.ml:83-101, three_tax.ml:70-83, GSL min/brent.c), cut at
// every objective call, so that a kernel has ONE evaluation site and the optimiser's variables
// sit in the engine's (shared-memory) object instead of being spilled around nine call sites.
// D supplies pend_bl / dist_bl / prox_bl, hint_prepare(kind, left, right, start, v) and hint_commit().
//   opt_advance(0, -)  starts;  opt_advance(1, f) feeds the value of the requested call.
//   returns 1: evaluate objective(b_kind, req_x) next;  0: finished, status in o_status.
// ---------------------------------------------------------------------------------------------
template <class D>
struct TTMachine {
  __device__ __forceinline__ D &self() { return *static_cast<D *>(this); }
  enum : int { B_LEFT, B_RIGHT, B_START, B_LOWER, B_UPPER, B_MIN, B_V, B_ITER };
  enum : int { O_PEND, O_DIST };
  double b_left, b_right, b_tol, b_lefty, b_righty, b_smaller, b_miny, b_prev;
  double b_xmin, b_xlo, b_xup, b_flo, b_fup, b_fmin, b_v, b_w, b_d, b_e, b_fv, b_fw, b_out, req_x;
  int b_ph, b_kind, b_iter, b_step, b_max, b_status;
  double o_prev_pend, o_prev_dist, o_curr_pend, o_curr_dist, o_tol, o_max_pend;
  int o_ph, o_step, o_max_iter, o_status;

  __device__ int brent_begin(int kind, double raw_start, double left, double right, double tolerance, int max_iters) {
    b_kind = kind; b_left = left; b_right = right; b_tol = tolerance; b_max = max_iters;
    if (left >= raw_start || raw_start >= right) { b_status = ST_INVALID_START; return 0; }
    self().hint_prepare(kind, left, right, raw_start, left + 0.3819660 * (right - left));
    b_prev = raw_start;
    return 2;  // run the batched pass, then opt_advance(2, -)
  }
  __device__ int brent_done(int st) { b_status = st; return 0; }
  __device__ int brent_feed(double fv) {
    const double golden = 0.3819660;
    switch (b_ph) {
      case B_LEFT: b_lefty = fv; b_ph = B_RIGHT; req_x = b_right; return 1;
      case B_RIGHT: {
        b_righty = fv;
        b_smaller = (b_lefty < b_righty) ? b_left : b_right;
        double miny = fmin(b_lefty, b_righty);
        if (b_lefty != b_lefty || b_righty != b_righty) miny = __longlong_as_double(0x7ff8000000000000LL);
        b_miny = miny; b_iter = 1; b_ph = B_START; req_x = b_prev;
        return 1;
      }
      case B_START: {
        if (fv < b_miny) { b_xmin = b_prev; b_xlo = b_left; b_xup = b_right; b_ph = B_LOWER; req_x = b_xlo; return 1; }
        if (fabs(b_prev - b_smaller) < b_tol) { b_out = b_smaller; return brent_done(ST_OK); }  // FoundMin
        if (b_iter > b_max) return brent_done(ST_FIND_START_FAILURE);
        b_prev = (b_prev + b_smaller) / 2.0; ++b_iter; req_x = b_prev;
        return 1;
      }
      case B_LOWER: b_flo = fv; if (!isfinite(fv)) return brent_done(ST_GSL_EBADFUNC); b_ph = B_UPPER; req_x = b_xup; return 1;
      case B_UPPER: b_fup = fv; if (!isfinite(fv)) return brent_done(ST_GSL_EBADFUNC); b_ph = B_MIN; req_x = b_xmin; return 1;
      case B_MIN: {
        b_fmin = fv;
        if (!isfinite(fv)) return brent_done(ST_GSL_EBADFUNC);
        if (b_xlo > b_xup) return brent_done(ST_GSL_EINVAL);
        if (b_xmin >= b_xup || b_xmin <= b_xlo) return brent_done(ST_GSL_EINVAL);
        if (b_fmin >= b_flo || b_fmin >= b_fup) return brent_done(ST_GSL_EINVAL);
        b_v = b_xlo + golden * (b_xup - b_xlo); b_w = b_v; b_d = 0.0; b_e = 0.0;
        b_ph = B_V; req_x = b_v;
        return 1;
      }
      case B_V: {
        if (!isfinite(fv)) return brent_done(ST_GSL_EBADFUNC);
        b_fv = fv; b_fw = fv; b_step = 1;
        break;
      }
      default: {  // B_ITER: the value at u
        const double u = req_x, f_u = fv, z = b_xmin, f_z = b_fmin;
        if (!isfinite(f_u)) return brent_done(ST_GSL_EBADFUNC);
        if (f_u <= f_z) {
          if (u < z) { b_xup = z; b_fup = f_z; } else { b_xlo = z; b_flo = f_z; }
          b_v = b_w; b_fv = b_fw; b_w = z; b_fw = f_z; b_xmin = u; b_fmin = f_u;
        } else {
          if (u < z) { b_xlo = u; b_flo = f_u; } else { b_xup = u; b_fup = f_u; }
          if (f_u <= b_fw || b_w == z) { b_v = b_w; b_fv = b_fw; b_w = u; b_fw = f_u; }
          else if (f_u <= b_fv || b_v == z || b_v == b_w) { b_v = u; b_fv = f_u; }
        }
        if (b_xup - b_xlo < b_tol) { b_out = b_xmin; return brent_done(ST_OK); }
        ++b_step;
      }
    }
    // next iterate (GSL brent_iterate up to the function evaluation)
    if (b_step > b_max) return brent_done(ST_EXCEEDED_MAX_ITER);
    {
      const double x_left = b_xlo, x_right = b_xup, z = b_xmin, f_z = b_fmin, v = b_v, wv = b_w, f_v = b_fv, f_w = b_fw;
      double dd = b_e, ee = b_d;  // GSL: d = state->e; e = state->d
      double u;
      const double w_lower = z - x_left, w_upper = x_right - z;
      const double tol = kSqrtDblEps * fabs(z);
      double p = 0.0, q = 0.0, r = 0.0;
      const double midpoint = 0.5 * (x_left + x_right);
      if (fabs(ee) > tol) {
        r = (z - wv) * (f_z - f_v);
        q = (z - v) * (f_z - f_w);
        p = (z - v) * q - (z - wv) * r;
        q = 2.0 * (q - r);
        if (q > 0.0) p = -p; else q = -q;
        r = ee;
        ee = dd;
      }
      if (fabs(p) < fabs(0.5 * q * r) && p < q * w_lower && p < q * w_upper) {
        const double t2 = 2.0 * tol;
        dd = p / q;
        u = z + dd;
        if ((u - x_left) < t2 || (x_right - u) < t2) dd = (z < midpoint) ? tol : -tol;
      } else {
        ee = (z < midpoint) ? x_right - z : -(z - x_left);
        dd = golden * ee;
      }
      if (fabs(dd) >= tol) u = z + dd; else u = z + ((dd > 0.0) ? tol : -tol);
      b_e = ee;
      b_d = dd;
      b_ph = B_ITER; req_x = u;
    }
    return 1;
  }

  __device__ void opt_init(double tolerance, double max_pend, int max_iter) {
    o_tol = tolerance; o_max_pend = max_pend; o_max_iter = max_iter;
    o_prev_pend = self().pend_bl; o_prev_dist = self().dist_bl; o_step = 1;
  }
  __device__ __forceinline__ int opt_advance(int feed, double fv) {
    if (feed == 1) {
      if (brent_feed(fv)) return 1;
      goto after_brent;
    }
    if (feed == 2) { self().hint_commit(); b_ph = B_LEFT; req_x = b_left; return 1; }
  round:
    if (o_step > o_max_iter) { o_status = ST_EXCEEDED_MAX_ITER; return 0; }
    o_curr_pend = self().pend_bl; o_curr_dist = self().dist_bl;
    if (1e-8 != o_max_pend) {
      o_ph = O_PEND;
      if (brent_begin(OBJ_PEND, self().pend_bl, 1e-8, o_max_pend, o_tol, 100)) return 2;
      goto after_brent;
    }
  start_dist:
    {
      const double cut = self().dist_bl + self().prox_bl;
      if (0.0 != cut) {
        o_ph = O_DIST;
        if (brent_begin(OBJ_DIST, self().dist_bl, 0.0, cut, o_tol, 100)) return 2;
        goto after_brent;
      }
    }
    goto check;
  after_brent:
    if (b_status != ST_OK) { o_status = b_status; return 0; }
    if (o_ph == O_PEND) { o_curr_pend = b_out; goto start_dist; }
    o_curr_dist = b_out;
  check:
    if (fabs(o_prev_pend - o_curr_pend) > o_tol || fabs(o_prev_dist - o_curr_dist) > o_tol) {
      o_prev_pend = o_curr_pend;
      o_prev_dist = o_curr_dist;
      ++o_step;
      goto round;
    }
    o_status = ST_OK;
    return 0;
  }

};

extern __shared__ double ppb_tt1_smem[];  // the general block engine's dynamic shared memory (ThreeTax)

template <int S, bool GCAT, int THREADS>
struct ThreeTax : TTDriver<ThreeTax<S, GCAT, THREADS>> {
  using Base = TTDriver<ThreeTax<S, GCAT, THREADS>>;
  using Base::pend_bl; using Base::dist_bl; using Base::prox_bl; using Base::calls;
  static constexpr int RS = GCAT ? S * S : p_rstride<S, GCAT>();  // Gcat: a warp reads ONE category's matrix (broadcast, 16-byte rows): no padding
  // ---- shared memory views ----
  // Offsets into the block's dynamic shared memory, turned into pointers by the accessors below: the object itself
  // lives in local memory (its address goes to the non-inlined driver), and pointers loaded from there are GENERIC -
  // every read of a rate matrix was an LD through the local/global path (ncu r5e: long_scoreboard / lg_throttle on
  // them).  Derived from the __shared__ symbol they are LDS.
  int o_statd, o_lam, o_x, o_xit, o_rates, o_ex, o_Pd, o_Pp, o_Pq, o_Tp, o_red, o_moc;
  __device__ __forceinline__ double *statd() const { return ppb_tt1_smem + o_statd; }
  __device__ __forceinline__ double *lam() const { return ppb_tt1_smem + o_lam; }
  __device__ __forceinline__ double *x() const { return ppb_tt1_smem + o_x; }
  __device__ __forceinline__ double *xit() const { return ppb_tt1_smem + o_xit; }
  __device__ __forceinline__ double *rates() const { return ppb_tt1_smem + o_rates; }
  __device__ __forceinline__ double *ex() const { return ppb_tt1_smem + o_ex; }
  __device__ __forceinline__ double *Pd() const { return ppb_tt1_smem + o_Pd; }
  __device__ __forceinline__ double *Pp() const { return ppb_tt1_smem + o_Pp; }
  __device__ __forceinline__ double *Pq() const { return ppb_tt1_smem + o_Pq; }
  __device__ __forceinline__ double *Tp() const { return ppb_tt1_smem + o_Tp; }
  __device__ __forceinline__ double *red() const { return ppb_tt1_smem + o_red; }
  __device__ __forceinline__ uint32_t *mask_of_code() const { return reinterpret_cast<uint32_t *>(ppb_tt1_smem + o_moc); }
  // ---- block-uniform state ----
  GlvArrays g;
  int R, Rt, Rg, K, nc;
  const int32_t *qsite;
  const uint8_t *qcode;
  const int32_t *site_cat;
  double *w;
  int64_t wstride;
  int nq, loc;
  // Gcat: the candidate's distal / proximal vectors gathered ONCE at the query's sites in (category, site) order,
  // each category's segment padded to a multiple of 32 (query.cuh: qsite_by_category_kernel), cd/cp[s][i]
  double *cd, *cp;
  int nqc;
  double e_tot_ln2, rate_term;

  static __host__ __device__ size_t smem_doubles(int Rt, int nc) {
    size_t n = 0;
    n += S + S + 2 * S * S + Rt;              // statd lam x xit rates
    n += (size_t)Rt * S;                      // ex
    n += 1;                                   // Pd starts on a 16-byte boundary (Gcat rows are read as double2)
    n += 2 * (size_t)Rt * RS;                 // Pd Pp
    n += (size_t)Rt * RS;                     // Pq (Gcat pendant tensor / Gmix scratch)
    n += GCAT ? 0 : (size_t)nc * Rt * S;      // Tp
    n += 40;                                  // red
    n += kMaxCodesDev / 2;                    // mask_of_code (uint32)
    return n;
  }

  __device__ void carve(double *sm, const DevModel *m) {
    R = m->R; Rt = m->R; Rg = GCAT ? 1 : m->R; K = Rg * S; nc = m->n_codes;
    int o = 0;
    o_statd = o; o += S; o_lam = o; o += S; o_x = o; o += S * S; o_xit = o; o += S * S; o_rates = o; o += Rt;
    o_ex = o; o += Rt * S;
    if (o & 1) ++o;
    o_Pd = o; o += Rt * RS; o_Pp = o; o += Rt * RS; o_Pq = o; o += Rt * RS;
    o_Tp = o; o += GCAT ? 0 : nc * Rt * S;
    o_red = o; o += 40;
    o_moc = o;
    (void)sm;
    for (int i = threadIdx.x; i < S; i += THREADS) { statd()[i] = m->statd[i]; lam()[i] = m->lam[i]; }
    // xit is kept TRANSPOSED (xitT[k][j] = xit[j][k]): build_P's lanes walk j, so they read consecutive words
    for (int i = threadIdx.x; i < S * S; i += THREADS) { x()[i] = m->x[i]; xit()[(i % S) * S + i / S] = m->xit[i]; }
    for (int i = threadIdx.x; i < Rt; i += THREADS) rates()[i] = m->rates[i];
    for (int i = threadIdx.x; i < kMaxCodesDev; i += THREADS) mask_of_code()[i] = m->mask_of_code[i];
    __syncthreads();
  }

  // Diagd.multi_exp for all Rt rates into dst (diagd.ml:79-98; linear_c.c:92-146)
  __device__ __noinline__ void build_P(int o_dst, double bl) {
    double *dst = ppb_tt1_smem + o_dst;
    __syncthreads();
    for (int t = threadIdx.x; t < Rt * S; t += THREADS) {
      const int r = t / S, k = t - r * S;
      ex()[t] = exp(rates()[r] * bl * lam()[k]);
    }
    __syncthreads();
    // four adjacent entries of a row per thread: ex * x is formed once per k for the four and xitT[k][j .. j+3] comes as
    // two 16-byte loads - a third of the shared-memory loads per FMA of the one-entry loop (r5k: 14 % of the C3 samples);
    // each entry still sums its k terms in ascending order, as dediagonalize does (linear_c.c:92-146)
    constexpr int JB = 4;
    static_assert(S % JB == 0, "row segments");
    const double *xt = xit();
    for (int t = threadIdx.x; t < Rt * S * (S / JB); t += THREADS) {
      const int r = t / (S * (S / JB)), rem = t - r * S * (S / JB), i = rem / (S / JB), j0 = (rem - i * (S / JB)) * JB;
      double acc[JB];
#pragma unroll
      for (int u = 0; u < JB; ++u) acc[u] = 0.0;
#pragma unroll
      for (int k = 0; k < S; ++k) {
        const double exx = ex()[r * S + k] * x()[i * S + k];
        const double2 v0 = *reinterpret_cast<const double2 *>(xt + k * S + j0), v1 = *reinterpret_cast<const double2 *>(xt + k * S + j0 + 2);
        acc[0] = fma(exx, v0.x, acc[0]); acc[1] = fma(exx, v0.y, acc[1]); acc[2] = fma(exx, v1.x, acc[2]); acc[3] = fma(exx, v1.y, acc[3]);
      }
#pragma unroll
      for (int u = 0; u < JB; ++u) dst[r * RS + i * S + j0 + u] = acc[u];
    }
    __syncthreads();
  }

  // evolved pendant GLV as a table (Gmix) or tensor (Gcat): Glv_edge.set_bl on pend (core.ml:234)
  __device__ __noinline__ void set_pend(double bl) {
    pend_bl = bl;
    build_P(o_Pq, bl);
    if (!GCAT) {
      for (int idx = threadIdx.x; idx < nc * Rt * S; idx += THREADS) {
        const int code = idx / (Rt * S), rs = idx - code * (Rt * S), r = rs / S, s = rs - r * S;
        uint32_t mask = mask_of_code()[code];
        double acc = 0.0;
        while (mask) { const int j = __ffs(mask) - 1; mask &= mask - 1; acc += Pq()[r * RS + s * S + j]; }
        Tp()[idx] = acc;
      }
      __syncthreads();
    }
  }
  __device__ void set_dist_prox(double d, double p) {
    dist_bl = d; prox_bl = p;
    build_P(o_Pd, d);
    build_P(o_Pp, p);
  }

  __device__ __forceinline__ double pend_val(int cat, int code, int r, int s) const {
    if (!GCAT) return Tp()[code * (Rt * S) + r * S + s];
    uint32_t mask = mask_of_code()[code];
    const double *row = Pq() + cat * RS + s * S;
    if ((mask & (mask - 1)) == 0) return row[__ffs(mask) - 1];
    double acc = 0.0;
    while (mask) { const int j = __ffs(mask) - 1; mask &= mask - 1; acc += row[j]; }
    return acc;
  }

  // Three_tax.log_like over all of the query's sites with the current P matrices; optionally
  // caches w[k][i] = (statd_s * prox_ev) * dist_ev for the pendant-only evaluations.
  // Gcat: lanes of a warp share the site category, so the category's matrix rows are broadcast LDS.128 loads (the
  // general loop below reads a different matrix per lane: one 8-byte shared-memory word per lane and FMA, bound by
  // shared-memory bandwidth - r5a: 304 ms for 256 ar53 genomes); d / p come coalesced from the compact copies.
  template <bool WRITE_W>
  __device__ __forceinline__ double eval_full_gcat() {
    LogProd lp;
    lp.init();
    for (int i = threadIdx.x; i < nqc; i += THREADS) {
      const int pk = __ldg(qsite + i);
      const bool valid = pk >= 0;
      const int pk0 = __shfl_sync(0xffffffffu, pk, 0);  // a segment starts at a multiple of 32 with a real site
      const int cat = site_cat[pk0 & kSiteMask];
      const int code = (pk >> kSiteCodeShift) & 0x7f;
      double d[S], p[S];
#pragma unroll
      for (int j = 0; j < S; ++j) { d[j] = __ldcg(cd + (int64_t)j * wstride + i); p[j] = __ldcg(cp + (int64_t)j * wstride + i); }
      const double *Prd = Pd() + cat * RS, *Prp = Pp() + cat * RS;
      const uint32_t mask = valid ? mask_of_code()[code] : 1u;
      const bool single = __all_sync(0xffffffffu, (mask & (mask - 1)) == 0);
      const double *col = Pq() + cat * RS + (__ffs(mask) - 1);
      double u = 0.0;
#pragma unroll 2
      for (int s = 0; s < S; ++s) {
        const double2 *rd = reinterpret_cast<const double2 *>(Prd + s * S), *rp = reinterpret_cast<const double2 *>(Prp + s * S);
        double de0 = 0.0, de1 = 0.0, pe0 = 0.0, pe1 = 0.0;
#pragma unroll
        for (int j2 = 0; j2 < S / 2; ++j2) {
          const double2 a = rd[j2], b = rp[j2];
          de0 = fma(a.x, d[2 * j2], de0); de1 = fma(a.y, d[2 * j2 + 1], de1);
          pe0 = fma(b.x, p[2 * j2], pe0); pe1 = fma(b.y, p[2 * j2 + 1], pe1);
        }
        const double ww = (statd()[s] * (pe0 + pe1)) * (de0 + de1);  // ((statd*x)*y) with x = prox, y = dist (three_tax.ml:31-36)
        if (WRITE_W) w[(int64_t)s * wstride + i] = ww;
        u = fma(ww, single ? col[s * S] : (valid ? pend_val(cat, code, 0, s) : 0.0), u);
      }
      if (valid) lp.mul(u);
    }
    return block_sum<THREADS>(lp.value(), red()) - rate_term + e_tot_ln2;
  }

  template <bool WRITE_W>
  __device__ __forceinline__ double eval_full_impl() {
    if (GCAT) return eval_full_gcat<WRITE_W>();
    LogProd lp;
    lp.init();
    const int64_t Lp = g.Lp;
    const double *dbase = g.a[0] + (int64_t)loc * g.a_stride;
    const double *pbase = g.a[1] + (int64_t)loc * g.a_stride;
    for (int i = threadIdx.x; i < nq; i += THREADS) {
      const int site = qsite[i] & kSiteMask;
      const int code = qcode[site] & 0x7f;
      const int cat = GCAT ? site_cat[site] : 0;
      double u = 0.0;
      for (int r = 0; r < Rg; ++r) {
        const double *Prd = Pd() + (GCAT ? cat : r) * RS, *Prp = Pp() + (GCAT ? cat : r) * RS;
        double d[S], p[S];
#pragma unroll
        for (int s = 0; s < S; ++s) {
          d[s] = dbase[(int64_t)(r * S + s) * Lp + site];
          p[s] = pbase[(int64_t)(r * S + s) * Lp + site];
        }
#pragma unroll(S == 4 ? 4 : 1)
        for (int s = 0; s < S; ++s) {
          double de = 0.0, pe = 0.0;
#pragma unroll
          for (int j = 0; j < S; ++j) { de = fma(Prd[s * S + j], d[j], de); pe = fma(Prp[s * S + j], p[j], pe); }
          const double ww = (statd()[s] * pe) * de;  // ((statd*x)*y) with x = prox, y = dist (three_tax.ml:31-36)
          if (WRITE_W) w[(int64_t)(r * S + s) * wstride + i] = ww;
          u = fma(ww, pend_val(cat, code, r, s), u);
        }
      }
      lp.mul(u);
    }
    return block_sum<THREADS>(lp.value(), red()) - rate_term + e_tot_ln2;
  }
  __device__ __noinline__ double eval_full_cache() { return eval_full_impl<true>(); }
  __device__ __noinline__ double eval_full_plain() { return eval_full_impl<false>(); }
  template <bool WRITE_W>
  __device__ __forceinline__ double eval_full() { return WRITE_W ? eval_full_cache() : eval_full_plain(); }

  __device__ __noinline__ double eval_pend_cached() {
    LogProd lp;
    lp.init();
    if (GCAT) {
      // A thread takes PAIRS of adjacent list entries (16-byte loads: twice the bytes per load instruction in flight),
      // TP pairs per trip, and requests all its S * TP loads of w before anything depends on them.  ncu r5e / r5g: this
      // pass is bound by memory latency (63 % of the kernel's samples wait here at 8 % of the DRAM bandwidth).
      // Lanes 0-15 of a warp cover one 32-entry group of the list, lanes 16-31 the next: the category is uniform per half-warp.
      constexpr int TP = 2;
      const int lane = threadIdx.x & 31;
      const int npair = nqc >> 1;
      for (int base = 0; base < npair; base += TP * THREADS) {  // block-uniform trip count: every lane reaches the warp-wide shuffles and votes
        double2 wv[TP][S];
        int2 pk[TP];
#pragma unroll
        for (int t = 0; t < TP; ++t) {
          const int j = base + t * THREADS + threadIdx.x;
          const bool in = j < npair;  // uniform per HALF-warp only (npair is a multiple of 16)
          pk[t] = in ? __ldg(reinterpret_cast<const int2 *>(qsite) + j) : make_int2(-1, -1);
#pragma unroll
          for (int s = 0; s < S; ++s)
            wv[t][s] = in ? __ldcg(reinterpret_cast<const double2 *>(w + (int64_t)s * wstride) + j) : make_double2(0.0, 0.0);
        }
#pragma unroll
        for (int t = 0; t < TP; ++t) {
          const int pk0 = __shfl_sync(0xffffffffu, pk[t].x, lane & 16);  // a group starts with a real site
          const int cat = pk0 >= 0 ? site_cat[pk0 & kSiteMask] : 0;
          const bool va = pk[t].x >= 0, vb = pk[t].y >= 0;
          const int ca = (pk[t].x >> kSiteCodeShift) & 0x7f, cb = (pk[t].y >> kSiteCodeShift) & 0x7f;
          const uint32_t ma = va ? mask_of_code()[ca] : 1u, mb = vb ? mask_of_code()[cb] : 1u;
          double ua0 = 0.0, ua1 = 0.0, ub0 = 0.0, ub1 = 0.0;
          if (__all_sync(0xffffffffu, ((ma & (ma - 1)) | (mb & (mb - 1))) == 0)) {  // one state per site (the rule): columns of P_cat(pend)
            const double *cola = Pq() + cat * RS + (__ffs(ma) - 1), *colb = Pq() + cat * RS + (__ffs(mb) - 1);
#pragma unroll
            for (int s = 0; s < S; s += 2) {
              ua0 = fma(wv[t][s].x, cola[s * S], ua0); ua1 = fma(wv[t][s + 1].x, cola[(s + 1) * S], ua1);
              ub0 = fma(wv[t][s].y, colb[s * S], ub0); ub1 = fma(wv[t][s + 1].y, colb[(s + 1) * S], ub1);
            }
          } else {
#pragma unroll  // static indices: wv must stay in registers
            for (int s = 0; s < S; ++s) {
              ua0 = fma(wv[t][s].x, va ? pend_val(cat, ca, 0, s) : 0.0, ua0);
              ub0 = fma(wv[t][s].y, vb ? pend_val(cat, cb, 0, s) : 0.0, ub0);
            }
          }
          if (va) lp.mul(ua0 + ua1);
          if (vb) lp.mul(ub0 + ub1);
        }
      }
      return block_sum<THREADS>(lp.value(), red()) - rate_term + e_tot_ln2;
    }
    for (int i = threadIdx.x; i < nq; i += THREADS) {
      const int site = qsite[i] & kSiteMask;
      const int code = qcode[site] & 0x7f;
      const int cat = GCAT ? site_cat[site] : 0;
      double u = 0.0;
      for (int r = 0; r < Rg; ++r)
#pragma unroll(S == 4 ? 4 : 1)
        for (int s = 0; s < S; ++s) u = fma(w[(int64_t)(r * S + s) * wstride + i], pend_val(cat, code, r, s), u);
      lp.mul(u);
    }
    return block_sum<THREADS>(lp.value(), red()) - rate_term + e_tot_ln2;
  }

  // core.ml:225-239 set_tt_edges (without the gather): fix the candidate and its branch lengths
  __device__ void set_tt_edges(int loc_, double cut_bl, double pendant, double distal) {
    loc = loc_;
    long long es = 0;
    if (GCAT) {
      // one scattered pass over the candidate's planes; every later pass reads the compact copies coalesced
      const double *dbase = g.a[0] + (int64_t)loc * g.a_stride, *pbase = g.a[1] + (int64_t)loc * g.a_stride;
      __syncthreads();  // the previous candidate's passes are done with cd / cp
      for (int i = threadIdx.x; i < nqc; i += THREADS) {
        const int pk = __ldg(qsite + i);
        const int site = pk & kSiteMask;
        if (pk >= 0) es += (long long)g.e[0][(int64_t)loc * g.Lp + site] + (long long)g.e[1][(int64_t)loc * g.Lp + site];
#pragma unroll 4
        for (int j = 0; j < S; ++j) {
          cd[(int64_t)j * wstride + i] = pk >= 0 ? __ldg(dbase + (int64_t)j * g.Lp + site) : 0.0;
          cp[(int64_t)j * wstride + i] = pk >= 0 ? __ldg(pbase + (int64_t)j * g.Lp + site) : 0.0;
        }
      }
    } else
    for (int i = threadIdx.x; i < nq; i += THREADS) {
      const int site = qsite[i] & kSiteMask;
      es += (long long)g.e[0][(int64_t)loc * g.Lp + site] + (long long)g.e[1][(int64_t)loc * g.Lp + site];
    }
    es = block_sum_ll<THREADS>(es, reinterpret_cast<long long *>(red()));
    e_tot_ln2 = kLn2 * (double)es;  // ln2 * (total_twoexp prox + total_twoexp dist), gstar_support.ml:43-49
    rate_term = GCAT ? 0.0 : (double)((float)nq) * log((double)((float)R));  // linear_c.c:421
    set_pend(pendant);
    set_dist_prox(distal, cut_bl - distal);
  }

  // ---- TTDriver engine interface: one likelihood pass per objective call ----
  __device__ __forceinline__ double f_pend(double v) { set_pend(v); return eval_pend_cached(); }
  __device__ __forceinline__ double f_dist(double d, double p) { set_dist_prox(d, p); return eval_full_plain(); }
  __device__ __forceinline__ void cache_w() { (void)eval_full_cache(); }
  __device__ __forceinline__ void hint_brent(int, double, double, double, double) {}
  __device__ __forceinline__ void hint_begin() {}
  __device__ __forceinline__ void hint_add(double) {}
  __device__ __forceinline__ void hint_flush() {}
};

// Persistent kernel: blocks pull candidates from `counter` until n are done.
template <int S, bool GCAT, int THREADS>
__global__ void __launch_bounds__(THREADS)
three_tax_kernel(const DevModel *__restrict__ m, GlvArrays g, const double *__restrict__ bl,
                 const int32_t *__restrict__ site_cat, const uint8_t *__restrict__ codes,
                 const int32_t *__restrict__ qsite, const int32_t *__restrict__ nq, const CandIn *__restrict__ in,
                 CandOut *__restrict__ out, int n, TTParams prm, double *__restrict__ wbuf, int *counter,
                 const int32_t *__restrict__ nqc, int64_t LpC) {
  // Gcat: qsite = the (category, site)-ordered padded lists [Q][LpC], nqc = their padded lengths; wbuf holds w, cd, cp per block
  double *sm = ppb_tt1_smem;
  __shared__ int s_idx;
  ThreeTax<S, GCAT, THREADS> tt;
  tt.g = g;
  tt.carve(sm, m);
  tt.site_cat = site_cat;
  tt.wstride = GCAT ? LpC : g.Lp;
  tt.w = wbuf + (int64_t)blockIdx.x * tt.K * tt.wstride * (GCAT ? 3 : 1);
  tt.cd = tt.w + (int64_t)tt.K * tt.wstride;
  tt.cp = tt.cd + (int64_t)tt.K * tt.wstride;
  for (;;) {
    __syncthreads();
    if (threadIdx.x == 0) s_idx = atomicAdd(counter, 1);
    __syncthreads();
    const int idx = s_idx;
    if (idx >= n) break;
    const CandIn c = in[idx];
    tt.qsite = qsite + (int64_t)c.q * (GCAT ? LpC : g.Lp);
    tt.qcode = codes + (int64_t)c.q * g.Lp;
    tt.nq = nq[c.q];
    tt.nqc = GCAT ? nqc[c.q] : 0;
    tt.calls = 0;
    const double cut_bl = bl[c.loc];
    CandOut o;
    o.marg = __longlong_as_double(0x7ff8000000000000LL);
    const double p0 = prm.use_defaults ? prm.start_pend : c.pend0;
    const double d0 = prm.use_defaults ? cut_bl / 2.0 : c.dist0;
    tt.set_tt_edges(c.loc, cut_bl, p0, d0);
    int st = ST_OK;
    if (prm.mode == TT_MODE_OPTIMIZE) {
      // core.ml:254-279: ExceededMaxIter keeps the current state; a GSL error retries once from the
      // default branch lengths (tt_edges_default, core.ml:240-245).
      st = tt.optimize(prm.tol, prm.max_pend, prm.max_iter);
      if (st_is_gsl(st)) {
        tt.set_tt_edges(c.loc, cut_bl, prm.start_pend, cut_bl / 2.0);
        st = tt.optimize(prm.tol, prm.max_pend, prm.max_iter);
      }
      if (st == ST_EXCEEDED_MAX_ITER) st = ST_OK;
      o.ll = tt.template eval_full<false>();  // Three_tax.get_results (three_tax.ml:85)
    } else if (prm.mode == TT_MODE_LOGLIKE) {
      o.ll = tt.template eval_full<false>();
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
    if (threadIdx.x == 0) out[idx] = o;
  }
}

}  // namespace ppb
