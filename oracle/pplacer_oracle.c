/* TEST INFRASTRUCTURE ONLY — see pplacer_oracle.h.  CPU FP64 restatement of the reference's
 * per-query placement path.  Every function cites the reference file:line it follows (paths are
 * relative to /root/reference).  Arithmetic order follows the reference loops;
 * compile with -ffp-contract=off so sums/products round as in the reference's x86-64 build.
 */
#define _GNU_SOURCE
#include "pplacer_oracle.h"
#include "qng_tables.h"

#include <errno.h>
#include <fcntl.h>
#include <float.h>
#include <math.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <time.h>
#include <unistd.h>

/* ------------------------------------------------------------------------------------------ */
/* discrete Gamma (gamma.ml) on top of restated GSL special functions                          */
/* ------------------------------------------------------------------------------------------ */

/* gsl_sf_gamma_inc_P(a,x): regularised lower incomplete gamma.  GSL is not vendored in the
 * reference; restated with the classical series / Lentz continued fraction in long double. */
static long double gamma_inc_P_l(long double a, long double x) {
  if (x <= 0.0L) return 0.0L;
  long double lg = lgammal(a);
  if (x < a + 1.0L) {
    long double ap = a, del = 1.0L / a, sum = del;
    for (int n = 0; n < 100000; n++) {
      ap += 1.0L;
      del *= x / ap;
      sum += del;
      if (fabsl(del) < fabsl(sum) * 1e-21L) break;
    }
    return sum * expl(-x + a * logl(x) - lg);
  } else {
    const long double tiny = 1e-4000L;
    long double b = x + 1.0L - a, c = 1.0L / tiny, d = 1.0L / b, h = d;
    for (int i = 1; i < 100000; i++) {
      long double an = -(long double)i * ((long double)i - a);
      b += 2.0L;
      d = an * d + b; if (fabsl(d) < tiny) d = tiny;
      c = b + an / c; if (fabsl(c) < tiny) c = tiny;
      d = 1.0L / d;
      long double del = d * c;
      h *= del;
      if (fabsl(del - 1.0L) < 1e-21L) break;
    }
    long double q = expl(-x + a * logl(x) - lg) * h;
    return 1.0L - q;
  }
}
double orc_gamma_inc_P(double a, double x) { return (double)gamma_inc_P_l(a, x); }

/* gsl_cdf_gamma_Pinv(p, a, 1): safeguarded Newton on P(a,x) = p to full double precision. */
double orc_gamma_Pinv(double p, double a) {
  if (p <= 0.0) return 0.0;
  if (p >= 1.0) return INFINITY;
  long double lo = 0.0L, hi = 1.0L;
  while ((double)gamma_inc_P_l(a, hi) < p) { lo = hi; hi *= 2.0L; }
  long double x = 0.5L * (lo + hi), lg = lgammal((long double)a);
  for (int it = 0; it < 500; it++) {
    long double f = gamma_inc_P_l(a, x) - (long double)p;
    if (f > 0) hi = x; else lo = x;
    long double pdf = expl(((long double)a - 1.0L) * logl(x) - x - lg);
    long double xn = (pdf > 0) ? x - f / pdf : -1.0L;
    if (!(xn > lo && xn < hi)) xn = 0.5L * (lo + hi);
    if (fabsl(xn - x) <= 1e-19L * fabsl(x)) { x = xn; break; }
    x = xn;
  }
  return (double)x;
}

/* gamma.ml:14-25 inverse_gamma_cdf ~alpha ~beta:1. (epsilon 1e-7 guard) */
static double inverse_gamma_cdf(double alpha, double y) {
  const double epsilon = 1e-7;
  if (y < orc_gamma_inc_P(alpha, epsilon)) return 0.0; /* gsl_cdf_gamma_P(x,a,b=1) = P(a,x) */
  return orc_gamma_Pinv(y, alpha);
}

/* gamma.ml:52-70 discrete_gamma; make_mean_one gamma.ml:27-30 */
void orc_discrete_gamma(int n_cat, double alpha, double *out) {
  double *freq = (double *)malloc(sizeof(double) * (size_t)(n_cat + 1));
  for (int i = 0; i <= n_cat; i++) {
    if (i == 0) freq[i] = 0.0;
    else if (i == n_cat) freq[i] = 1.0;
    else {
      double cutpoint = inverse_gamma_cdf(alpha, (double)i / (double)n_cat); /* Yang 1994 eq 9 */
      freq[i] = orc_gamma_inc_P(alpha + 1.0, cutpoint);                      /* eq 10 */
    }
  }
  double tot = 0.0;
  for (int i = 0; i < n_cat; i++) { out[i] = freq[i + 1] - freq[i]; tot += out[i]; }
  double f_n = (double)n_cat;
  for (int i = 0; i < n_cat; i++) out[i] = f_n * out[i] / tot;
  free(freq);
}

/* ------------------------------------------------------------------------------------------ */
/* Diagd (diagd.ml) + dediagonalize (linear_c.c:92-146)                                        */
/* ------------------------------------------------------------------------------------------ */

/* Gsl.Eigen.symmv (linear_utils.ml:206-208): any symmetric eigensolver gives the same P(t) to
 * rounding; cyclic Jacobi here.  Eigenvectors are the COLUMNS of evec (test_matrix.ml:9-16). */
void orc_symm_eigs(int n, const double *m, double *eval, double *evec) {
  double *a = (double *)malloc(sizeof(double) * (size_t)n * n);
  memcpy(a, m, sizeof(double) * (size_t)n * n);
  for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) evec[i * n + j] = (i == j);
  for (int sweep = 0; sweep < 100; sweep++) {
    double off = 0.0;
    for (int i = 0; i < n; i++) for (int j = i + 1; j < n; j++) off += a[i * n + j] * a[i * n + j];
    if (off < 1e-300) break;
    for (int p = 0; p < n; p++) for (int q = p + 1; q < n; q++) {
      double apq = a[p * n + q];
      if (apq == 0.0) continue;
      double theta = (a[q * n + q] - a[p * n + p]) / (2.0 * apq);
      double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
      double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
      for (int k = 0; k < n; k++) {
        double akp = a[k * n + p], akq = a[k * n + q];
        a[k * n + p] = c * akp - s * akq; a[k * n + q] = s * akp + c * akq;
      }
      for (int k = 0; k < n; k++) {
        double apk = a[p * n + k], aqk = a[q * n + k];
        a[p * n + k] = c * apk - s * aqk; a[q * n + k] = s * apk + c * aqk;
      }
      for (int k = 0; k < n; k++) {
        double vkp = evec[k * n + p], vkq = evec[k * n + q];
        evec[k * n + p] = c * vkp - s * vkq; evec[k * n + q] = s * vkp + c * vkq;
      }
    }
  }
  for (int i = 0; i < n; i++) eval[i] = a[i * n + i];
  free(a);
}

/* linear_c.c:92-146: dst[i][j] = sum_k (lambda[k]*u[i][k])*uit[j][k], k ascending */
static void dediagonalize(double *dst, const double *u, const double *lambda, const double *uit, int n) {
  for (int i = 0; i < n; i++)
    for (int j = 0; j < n; j++) {
      double acc = 0;
      for (int k = 0; k < n; k++) acc += lambda[k] * u[i * n + k] * uit[j * n + k];
      dst[i * n + j] = acc;
    }
}

static double safe_invert(double x) { return x == 0.0 ? 1.0 : 1.0 / x; } /* diagd.ml:118 */

/* diagd.ml:169-172 normed_of_exchangeable_pair = of_exchangeable_pair (:154) + normalize_rate (:166) */
void orc_diagd_normed_of_exchangeable_pair(int S, const double *r, const double *pi, double *x,
                                           double *lambda, double *xit) {
  size_t nn = (size_t)S * S;
  double *b = (double *)malloc(sizeof(double) * nn), *sym = (double *)malloc(sizeof(double) * nn);
  double *u = (double *)malloc(sizeof(double) * nn), *q = (double *)malloc(sizeof(double) * nn);
  /* b_of_exchangeable_pair diagd.ml:140-152 */
  for (int i = 0; i < S; i++)
    for (int j = 0; j < S; j++) {
      if (i != j) b[i * S + j] = r[i * S + j];
      else {
        double total = 0.0;
        for (int k = 0; k < S; k++) if (k != i) total += r[k * S + i] * pi[k];
        b[i * S + j] = -(total * safe_invert(pi[i]));
      }
    }
  /* of_d_b diagd.ml:123-132: symm_eigs (D^1/2 B D^1/2); x = D^-1/2 U ; xit = D^1/2 U */
  double d_root[ORC_MAX_STATES];
  for (int i = 0; i < S; i++) d_root[i] = sqrt(pi[i]);
  for (int i = 0; i < S; i++) for (int j = 0; j < S; j++) sym[i * S + j] = d_root[i] * (b[i * S + j] * d_root[j]);
  /* enforce exact symmetry against rounding (assert_symmetric in the reference tolerates it) */
  for (int i = 0; i < S; i++) for (int j = i + 1; j < S; j++) {
    double m = 0.5 * (sym[i * S + j] + sym[j * S + i]); sym[i * S + j] = sym[j * S + i] = m; }
  orc_symm_eigs(S, sym, lambda, u);
  for (int i = 0; i < S; i++) for (int j = 0; j < S; j++) {
    x[i * S + j] = safe_invert(d_root[i]) * u[i * S + j];
    xit[i * S + j] = d_root[i] * u[i * S + j];
  }
  /* normalize_rate diagd.ml:158-167: rate = -sum_i q_ii pi_i ; lambda *= 1/rate */
  dediagonalize(q, x, lambda, xit, S);
  double rate = 0.0;
  for (int i = 0; i < S; i++) rate = rate - q[i * S + i] * pi[i];
  double inv = 1.0 / rate;
  for (int i = 0; i < S; i++) lambda[i] *= inv;
  free(b); free(sym); free(u); free(q);
}

/* diagd.ml:79-98 multi_exp: tensor[r] = X diag(exp(rate_r * bl * lambda)) Xit^T for unmasked r */
void orc_multi_exp(int S, int R, const double *x, const double *lambda, const double *xit,
                   const double *rates, const uint8_t *mask, double bl, double *tensor) {
  double util[ORC_MAX_STATES];
  memset(tensor, 0, sizeof(double) * (size_t)R * S * S);
  for (int r = 0; r < R; r++) {
    if (mask && !mask[r]) continue;
    for (int i = 0; i < S; i++) util[i] = exp(rates[r] * bl * lambda[i]);
    dediagonalize(tensor + (size_t)r * S * S, x, util, xit, S);
  }
}

/* ------------------------------------------------------------------------------------------ */
/* alphabets (nuc_models.ml:24-54, prot_models.ml:5-51)                                        */
/* ------------------------------------------------------------------------------------------ */
static const char PROT_CODE[21] = "ARNDCQEGHILKMFPSTWYV";

int orc_lv_of_char(int S, uint8_t c, double *lv) {
  for (int i = 0; i < S; i++) lv[i] = 0.0;
  if (S == 4) {
    const char *set;
    switch (c) {
      case 'A': set = "A"; break; case 'C': set = "C"; break; case 'G': set = "G"; break;
      case 'T': set = "T"; break; case 'R': set = "AG"; break; case 'Y': set = "CT"; break;
      case 'M': set = "AC"; break; case 'K': set = "GT"; break; case 'S': set = "CG"; break;
      case 'W': set = "AT"; break; case 'H': set = "ACT"; break; case 'B': set = "CGT"; break;
      case 'V': set = "ACG"; break; case 'D': set = "AGT"; break;
      case 'N': case '-': case '?': case 'X': set = "ACGT"; break;
      default: return 0;
    }
    for (const char *p = set; *p; p++) lv[*p == 'A' ? 0 : *p == 'C' ? 1 : *p == 'G' ? 2 : 3] = 1.0;
    return 1;
  }
  if (S == 20) {
    const char *p = c ? strchr(PROT_CODE, c) : NULL;
    if (p) { lv[p - PROT_CODE] = 1.0; return 1; }
    switch (c) {
      case '-': case '?': case 'X': for (int i = 0; i < 20; i++) lv[i] = 1.0; return 1;
      case 'B': lv[3] = 1.0; lv[2] = 1.0; return 1;  /* D or N */
      case 'J': lv[9] = 1.0; lv[10] = 1.0; return 1; /* I or L */
      case 'Z': lv[6] = 1.0; lv[5] = 1.0; return 1;  /* E or Q */
      default: return 0;
    }
  }
  return 0;
}

static int informative(uint8_t c) { return !(c == '?' || c == '-'); } /* alignment.ml:235 */

/* alignment.ml:186-217 emper_freq: per character add the l1-normalised likelihood vector
 * ('-' and '?' skipped, 'X' counted), then l1-normalise the total. */
int orc_emper_freq(int S, const uint8_t *chars, int64_t n, double *out) {
  double normed[256][ORC_MAX_STATES];
  uint8_t known[256];
  for (int c = 0; c < 256; c++) {
    double lv[ORC_MAX_STATES];
    known[c] = (uint8_t)orc_lv_of_char(S, (uint8_t)c, lv);
    if (known[c]) {
      double nrm = 0.0;
      for (int i = 0; i < S; i++) nrm += pow(lv[i], 1.0);
      nrm = pow(nrm, 1.0);
      double sc = 1.0 / nrm;
      for (int i = 0; i < S; i++) normed[c][i] = lv[i] * sc;
    }
  }
  for (int i = 0; i < S; i++) out[i] = 0.0;
  for (int64_t k = 0; k < n; k++) {
    uint8_t c = chars[k];
    if (c == '-' || c == '?') continue;
    if (!known[c]) return ORC_BAD_CHAR;
    for (int i = 0; i < S; i++) out[i] += normed[c][i];
  }
  double nrm = 0.0;
  for (int i = 0; i < S; i++) nrm += out[i];
  double sc = 1.0 / nrm;
  for (int i = 0; i < S; i++) out[i] *= sc;
  return ORC_OK;
}

/* ------------------------------------------------------------------------------------------ */
/* GSL one-dimensional Brent minimiser + Minimization.brent (minimization.ml)                  */
/* ------------------------------------------------------------------------------------------ */
#define GSL_SQRT_DBL_EPSILON 1.4901161193847656e-08
#define GSL_DBL_EPSILON 2.2204460492503131e-16

typedef struct {
  orc_fn f; void *ud; int *n_calls;
} fwrap;
static double callf(fwrap *w, double x) { if (w->n_calls) (*w->n_calls)++; return w->f(x, w->ud); }

/* Minimization.brent minimization.ml:83-101 with bisection_start_finder :24-41 and
 * start_finder_prep :11-17; GSL gsl_min_fminimizer_set / brent_init / brent_iterate restated from
 * GSL 2.7 min/fsolver.c + min/brent.c. */
int orc_brent(orc_fn f, void *ud, double raw_start, double left, double right, double tolerance,
              int max_iters, double *xmin_out, int *n_calls) {
  fwrap w = {f, ud, n_calls};
  /* start_finder_prep */
  if (left >= raw_start || raw_start >= right) return ORC_INVALID_START;
  double lefty = callf(&w, left), righty = callf(&w, right);
  double smaller = (lefty < righty) ? left : right;
  double miny = fmin(lefty, righty); /* OCaml min: polymorphic; NaN cases are moot for finite LL */
  if (lefty != lefty || righty != righty) miny = NAN;
  /* bisection_start_finder */
  double start, prev = raw_start;
  int iter = 1;
  for (;;) {
    double prev_val = callf(&w, prev);
    if (prev_val < miny) { start = prev; break; }
    else if (fabs(prev - smaller) < tolerance) { *xmin_out = smaller; return ORC_OK; } /* FoundMin */
    else if (iter > max_iters) return ORC_FIND_START_FAILURE;
    prev = (prev + smaller) / 2.0; iter++;
  }
  /* Gsl.Min.make BRENT f ~min:start ~lo:left ~up:right == gsl_min_fminimizer_set */
  double x_minimum = start, x_lower = left, x_upper = right;
  double f_lower = callf(&w, x_lower); if (!isfinite(f_lower)) return ORC_GSL_EBADFUNC;
  double f_upper = callf(&w, x_upper); if (!isfinite(f_upper)) return ORC_GSL_EBADFUNC;
  double f_minimum = callf(&w, x_minimum); if (!isfinite(f_minimum)) return ORC_GSL_EBADFUNC;
  if (x_lower > x_upper) return ORC_GSL_EINVAL;
  if (x_minimum >= x_upper || x_minimum <= x_lower) return ORC_GSL_EINVAL;
  if (f_minimum >= f_lower || f_minimum >= f_upper) return ORC_GSL_EINVAL;
  /* brent_init */
  const double golden = 0.3819660;
  double v = x_lower + golden * (x_upper - x_lower), wv = v, d = 0, e = 0;
  double f_vw = callf(&w, v); if (!isfinite(f_vw)) return ORC_GSL_EBADFUNC;
  double f_v = f_vw, f_w = f_vw;
  /* run minimization.ml:88-99 */
  for (int step = 1;; step++) {
    if (step > max_iters) return ORC_EXCEEDED_MAX_ITER;
    /* brent_iterate */
    {
      const double x_left = x_lower, x_right = x_upper, z = x_minimum;
      double dd = e, ee = d; /* sic: d = state->e; e = state->d */
      double u, f_u;
      const double f_z = f_minimum;
      const double w_lower = (z - x_left), w_upper = (x_right - z);
      const double tol = GSL_SQRT_DBL_EPSILON * fabs(z);
      double p = 0, q = 0, r = 0;
      const double midpoint = 0.5 * (x_left + x_right);
      if (fabs(ee) > tol) {
        r = (z - wv) * (f_z - f_v);
        q = (z - v) * (f_z - f_w);
        p = (z - v) * q - (z - wv) * r;
        q = 2 * (q - r);
        if (q > 0) p = -p; else q = -q;
        r = ee; ee = dd;
      }
      if (fabs(p) < fabs(0.5 * q * r) && p < q * w_lower && p < q * w_upper) {
        double t2 = 2 * tol;
        dd = p / q; u = z + dd;
        if ((u - x_left) < t2 || (x_right - u) < t2) dd = (z < midpoint) ? tol : -tol;
      } else {
        ee = (z < midpoint) ? x_right - z : -(z - x_left);
        dd = golden * ee;
      }
      if (fabs(dd) >= tol) u = z + dd; else u = z + ((dd > 0) ? tol : -tol);
      e = ee; d = dd;
      f_u = callf(&w, u); if (!isfinite(f_u)) return ORC_GSL_EBADFUNC;
      if (f_u <= f_z) {
        if (u < z) { x_upper = z; f_upper = f_z; } else { x_lower = z; f_lower = f_z; }
        v = wv; f_v = f_w; wv = z; f_w = f_z; x_minimum = u; f_minimum = f_u;
      } else {
        if (u < z) { x_lower = u; f_lower = f_u; } else { x_upper = u; f_upper = f_u; }
        if (f_u <= f_w || wv == z) { v = wv; f_v = f_w; wv = u; f_w = f_u; }
        else if (f_u <= f_v || v == z || v == wv) { v = u; f_v = f_u; }
      }
    }
    if (x_upper - x_lower < tolerance) { *xmin_out = x_minimum; return ORC_OK; }
  }
}

/* ------------------------------------------------------------------------------------------ */
/* GSL gsl_integration_qng restated (GSL 2.7 integration/qng.c)        */
/* ------------------------------------------------------------------------------------------ */
static double rescale_error(double err, const double result_abs, const double result_asc) {
  err = fabs(err);
  if (result_asc != 0 && err != 0) {
    double scale = pow((200 * err / result_asc), 1.5);
    if (scale < 1) err = result_asc * scale; else err = result_asc;
  }
  if (result_abs > DBL_MIN / (50 * GSL_DBL_EPSILON)) {
    double min_err = 50 * GSL_DBL_EPSILON * result_abs;
    if (min_err > err) err = min_err;
  }
  return err;
}

int orc_qng(orc_fn f, void *ud, double a, double b, double epsabs, double epsrel, double *result,
            double *abserr, int *neval) {
  const double *x1 = ppb_qng_x1, *w10 = ppb_qng_w10, *x2 = ppb_qng_x2, *w21a = ppb_qng_w21a,
               *w21b = ppb_qng_w21b, *x3 = ppb_qng_x3, *w43a = ppb_qng_w43a, *w43b = ppb_qng_w43b,
               *x4 = ppb_qng_x4, *w87a = ppb_qng_w87a, *w87b = ppb_qng_w87b;
  double fv1[5], fv2[5], fv3[5], fv4[5], savfun[21];
  double res10, res21, res43, res87, result_kronrod, err, resabs, resasc;
  const double half_length = 0.5 * (b - a), abs_half_length = fabs(half_length);
  const double center = 0.5 * (b + a);
  const double f_center = f(center, ud);
  int k;
  if (epsabs <= 0 && (epsrel < 50 * GSL_DBL_EPSILON || epsrel < 0.5e-28)) {
    *result = 0; *abserr = 0; *neval = 0; return ORC_GSL_EBADTOL;
  }
  res10 = 0; res21 = w21b[5] * f_center; resabs = w21b[5] * fabs(f_center);
  for (k = 0; k < 5; k++) {
    const double abscissa = half_length * x1[k];
    const double fval1 = f(center + abscissa, ud), fval2 = f(center - abscissa, ud);
    const double fval = fval1 + fval2;
    res10 += w10[k] * fval; res21 += w21a[k] * fval;
    resabs += w21a[k] * (fabs(fval1) + fabs(fval2));
    savfun[k] = fval; fv1[k] = fval1; fv2[k] = fval2;
  }
  for (k = 0; k < 5; k++) {
    const double abscissa = half_length * x2[k];
    const double fval1 = f(center + abscissa, ud), fval2 = f(center - abscissa, ud);
    const double fval = fval1 + fval2;
    res21 += w21b[k] * fval;
    resabs += w21b[k] * (fabs(fval1) + fabs(fval2));
    savfun[k + 5] = fval; fv3[k] = fval1; fv4[k] = fval2;
  }
  resabs *= abs_half_length;
  {
    const double mean = 0.5 * res21;
    resasc = w21b[5] * fabs(f_center - mean);
    for (k = 0; k < 5; k++)
      resasc += (w21a[k] * (fabs(fv1[k] - mean) + fabs(fv2[k] - mean)) +
                 w21b[k] * (fabs(fv3[k] - mean) + fabs(fv4[k] - mean)));
    resasc *= abs_half_length;
  }
  result_kronrod = res21 * half_length;
  err = rescale_error((res21 - res10) * half_length, resabs, resasc);
  if (err < epsabs || err < epsrel * fabs(result_kronrod)) {
    *result = result_kronrod; *abserr = err; *neval = 21; return ORC_OK;
  }
  res43 = w43b[11] * f_center;
  for (k = 0; k < 10; k++) res43 += savfun[k] * w43a[k];
  for (k = 0; k < 11; k++) {
    const double abscissa = half_length * x3[k];
    const double fval = (f(center + abscissa, ud) + f(center - abscissa, ud));
    res43 += fval * w43b[k]; savfun[k + 10] = fval;
  }
  result_kronrod = res43 * half_length;
  err = rescale_error((res43 - res21) * half_length, resabs, resasc);
  if (err < epsabs || err < epsrel * fabs(result_kronrod)) {
    *result = result_kronrod; *abserr = err; *neval = 43; return ORC_OK;
  }
  res87 = w87b[22] * f_center;
  for (k = 0; k < 21; k++) res87 += savfun[k] * w87a[k];
  for (k = 0; k < 22; k++) {
    const double abscissa = half_length * x4[k];
    res87 += w87b[k] * (f(center + abscissa, ud) + f(center - abscissa, ud));
  }
  result_kronrod = res87 * half_length;
  err = rescale_error((res87 - res43) * half_length, resabs, resasc);
  *result = result_kronrod; *abserr = err; *neval = 87;
  if (err < epsabs || err < epsrel * fabs(result_kronrod)) return ORC_OK;
  return ORC_GSL_ETOL;
}

/* ------------------------------------------------------------------------------------------ */
/* GLVs and the model operations (glvm.ml Model signature; gmix_model.ml, gcat_model.ml)        */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
  int n;       /* sites */
  int64_t *e;  /* [n] base-two exponents (OCaml native int) */
  double *a;   /* Gmix: [R][n][S] (rate outermost, linear_c.c:5-12); Gcat: [n][S] */
} glv;

struct orc_ctx {
  int S, R, gcat, Rg; /* Rg = rate dimension of a GLV: R for Gmix, 1 for Gcat */
  double statd[ORC_MAX_STATES], x[ORC_MAX_STATES * ORC_MAX_STATES], l[ORC_MAX_STATES],
      xit[ORC_MAX_STATES * ORC_MAX_STATES];
  double *rates;
  int32_t *site_cat;       /* gcat: [L] */
  uint8_t *occupied_rates; /* gcat: [R] gcat_model.ml:47-49 */
  int n_nodes, root_id;
  int32_t *child_ptr, *child_idx, *leaf_row;
  double *bl;
  uint8_t *ref_chars;
  int n_leaves, L;
  size_t glv_len; /* Rg*L*S doubles */
  double *A[3];
  int64_t *E[3];
  size_t map_bytes[3];
  int mapped;
  int precomputed;
  /* Fig.t (fig.ml:4-12): n_figs == 0 is Dummy */
  int n_figs;
  int32_t *fig_rep, *fig_ptr, *fig_idx;
};

/* per-thread scratch: the reference mutates model.tensor and util vectors in place */
typedef struct {
  double *tensor;  /* [R][S][S] model.tensor gmix_model.ml:15 */
  uint8_t *rmask;  /* [R] */
  double *util;    /* [L] utilv_nsites core.ml:93 */
  glv t1, t2;      /* full-length temporaries */
} scratch;

static void *xcalloc(size_t n, size_t sz) {
  void *p = calloc(n ? n : 1, sz);
  if (!p) { fprintf(stderr, "oracle: out of memory (%zu x %zu)\n", n, sz); abort(); }
  return p;
}
static glv glv_make(const orc_ctx *c, int n) {
  glv g; g.n = n;
  g.e = (int64_t *)xcalloc((size_t)n, sizeof(int64_t));
  g.a = (double *)xcalloc((size_t)c->Rg * n * c->S, sizeof(double));
  return g;
}
static void glv_free(glv *g) { free(g->e); free(g->a); g->e = NULL; g->a = NULL; }
static glv glv_at(const orc_ctx *c, int which, int node) {
  glv g; g.n = c->L;
  g.e = c->E[which] + (size_t)node * c->L;
  g.a = c->A[which] + (size_t)node * c->glv_len;
  return g;
}
static void scratch_init(const orc_ctx *c, scratch *s) {
  s->tensor = (double *)xcalloc((size_t)c->R * c->S * c->S, sizeof(double));
  s->rmask = (uint8_t *)xcalloc((size_t)c->R, 1);
  s->util = (double *)xcalloc((size_t)c->L, sizeof(double));
  s->t1 = glv_make(c, c->L); s->t2 = glv_make(c, c->L);
}
static void scratch_free(scratch *s) {
  free(s->tensor); free(s->rmask); free(s->util); glv_free(&s->t1); glv_free(&s->t2);
}

/* linear_c.c:29-90 gemmish: dst[site][i] = sum_j a[i][j]*b[site][j], j ascending, acc from 0 */
static void gemmish(double *dst, const double *a, const double *b, int S, int n_sites) {
  for (int site = 0; site < n_sites; site++) {
    for (int i = 0; i < S; i++) {
      double acc = 0;
      for (int j = 0; j < S; j++) acc += a[i * S + j] * b[j];
      dst[i] = acc;
    }
    dst += S; b += S;
  }
}

/* Model.evolve_into: gmix_model.ml:308-315 (+ gstar_support.ml:70-79 tensor_mul);
 * gcat_model.ml:327-345 (+ :72-82 prep_tensor_for_bl, GSL gemv row-dot linear_utils.ml:179-182).
 * reind (Gcat second stage): masked site i uses the category of original site reind[i]. */
static void evolve_into(const orc_ctx *c, scratch *s, glv *dst, const glv *src, double bl,
                        const int32_t *reind) {
  const int S = c->S, n = src->n;
  memcpy(dst->e, src->e, sizeof(int64_t) * (size_t)n);
  if (!c->gcat) {
    orc_multi_exp(S, c->R, c->x, c->l, c->xit, c->rates, NULL, bl, s->tensor);
    for (int r = 0; r < c->R; r++)
      gemmish(dst->a + (size_t)r * n * S, s->tensor + (size_t)r * S * S, src->a + (size_t)r * n * S, S, n);
  } else {
    const uint8_t *mask = c->occupied_rates;
    if (reind) {
      memset(s->rmask, 0, (size_t)c->R);
      for (int i = 0; i < n; i++) s->rmask[c->site_cat[reind[i]]] = 1;
      mask = s->rmask;
    }
    orc_multi_exp(S, c->R, c->x, c->l, c->xit, c->rates, mask, bl, s->tensor);
    for (int i = 0; i < n; i++) {
      const double *evo = s->tensor + (size_t)c->site_cat[reind ? reind[i] : i] * S * S;
      const double *v = src->a + (size_t)i * S;
      double *d = dst->a + (size_t)i * S;
      for (int r = 0; r < S; r++) { /* cblas_dgemv NoTrans: temp = sum_j x[j]*A[r][j]; y = 0 + 1*temp */
        double temp = 0.0;
        for (int j = 0; j < S; j++) temp += v[j] * evo[r * S + j];
        d[r] = 0.0 + 1.0 * temp;
      }
    }
  }
}

/* Glv.perhaps_pull_exponent gmix_model.ml:163-187 / gcat_model.ml:195-212 */
static void perhaps_pull_exponent(const orc_ctx *c, int min_allowed_twoexp, glv *g) {
  const int S = c->S, n = g->n;
  for (int site = 0; site < n; site++) {
    long max_twoexp = -0x3fffffffffffffffL; /* -max_int */
    for (int r = 0; r < c->Rg; r++) {
      const double *sl = g->a + ((size_t)r * n + site) * S;
      double mx = sl[0];
      for (int k = 1; k < S; k++) if (sl[k] > mx) mx = sl[k]; /* Gsl.Vector.max */
      int twoexp; (void)frexp(mx, &twoexp);
      if (twoexp > max_twoexp) max_twoexp = twoexp;
    }
    if (max_twoexp < min_allowed_twoexp) {
      double sc = ldexp(0.5, (int)(-max_twoexp) + 1); /* of_twoexp gstar_support.ml:40 */
      for (int r = 0; r < c->Rg; r++) {
        double *sl = g->a + ((size_t)r * n + site) * S;
        for (int k = 0; k < S; k++) sl[k] *= sc;
      }
      g->e[site] += max_twoexp;
    }
  }
}

/* Glv.pairwise_prod gmix_model.ml:215-218 (linear_c.c:426-441 / :218-230; exponents add
 * gstar_support.ml:24-30) */
static void pairwise_prod(const orc_ctx *c, glv *dst, const glv *g1, const glv *g2) {
  size_t len = (size_t)c->Rg * g1->n * c->S;
  for (int i = 0; i < g1->n; i++) dst->e[i] = g1->e[i] + g2->e[i];
  for (size_t i = 0; i < len; i++) dst->a[i] = g1->a[i] * g2->a[i];
}
static void glv_memcpy(const orc_ctx *c, glv *dst, const glv *src) {
  memcpy(dst->e, src->e, sizeof(int64_t) * (size_t)src->n);
  memcpy(dst->a, src->a, sizeof(double) * (size_t)c->Rg * src->n * c->S);
}
/* Glv.listwise_prod gmix_model.ml:224-233 */
static void listwise_prod(const orc_ctx *c, glv *dst, glv **list, int n) {
  if (n >= 2) {
    pairwise_prod(c, dst, list[0], list[1]);
    for (int i = 2; i < n; i++) pairwise_prod(c, dst, dst, list[i]);
  } else if (n == 1) glv_memcpy(c, dst, list[0]);
  else abort();
}
/* Model.statd_pairwise_prod gmix_model.ml:319-322 (linear_c.c:443-463 / :232-249) */
static void statd_pairwise_prod(const orc_ctx *c, glv *dst, const glv *g1, const glv *g2) {
  const int S = c->S, n = g1->n;
  for (int i = 0; i < n; i++) dst->e[i] = g1->e[i] + g2->e[i];
  size_t k = 0;
  for (int r = 0; r < c->Rg; r++)
    for (int site = 0; site < n; site++)
      for (int st = 0; st < S; st++, k++) dst->a[k] = c->statd[st] * g1->a[k] * g2->a[k];
}
/* Glv.prep_constant_rate_glv_from_lv_arr gmix_model.ml:149-156 / gcat_model.ml:186-191 */
static int prep_constant_rate_glv_from_chars(const orc_ctx *c, glv *g, const uint8_t *chars,
                                             const int32_t *pick /* or NULL */, int n) {
  const int S = c->S;
  double lv[ORC_MAX_STATES];
  for (int site = 0; site < n; site++) {
    if (!orc_lv_of_char(S, chars[pick ? pick[site] : site], lv)) return ORC_BAD_CHAR;
    for (int r = 0; r < c->Rg; r++) memcpy(g->a + ((size_t)r * n + site) * S, lv, sizeof(double) * S);
    g->e[site] = 0;
  }
  return ORC_OK;
}
static void glv_set_unit(const orc_ctx *c, glv *g) {
  memset(g->e, 0, sizeof(int64_t) * (size_t)g->n);
  size_t len = (size_t)c->Rg * g->n * c->S;
  for (size_t i = 0; i < len; i++) g->a[i] = 1.0;
}
/* gstar_support.ml:43-49 total_twoexp (float accumulation) */
static double total_twoexp(const int64_t *e, int n) {
  double tot = 0.0;
  for (int i = 0; i < n; i++) tot += (double)e[i];
  return tot;
}
static double masked_total_twoexp(const int64_t *e, const uint16_t *mask, int n) { /* :52-59 */
  double tot = 0.0;
  for (int i = 0; i < n; i++) if (mask[i]) tot += (double)e[i];
  return tot;
}
static const double LOG_OF_2 = 0.693147180559945309417232121458; /* log 2. gstar_support.ml:9 */

/* Model.log_like3: gmix_model.ml:298-304 -> ten_log_like3_c linear_c.c:367-424;
 * gcat_model.ml:317-323 -> mat_log_like3_c linear_c.c:170-216 */
static double log_like3(const orc_ctx *c, double *util, const glv *x, const glv *y, const glv *z) {
  const int S = c->S, n = x->n;
  double ll_tot = 0;
  if (!c->gcat) {
    for (int site = 0; site < n; site++) util[site] = 0.0;
    const double *xp = x->a, *yp = y->a, *zp = z->a;
    for (int r = 0; r < c->R; r++)
      for (int site = 0; site < n; site++) {
        for (int st = 0; st < S; st++) util[site] += c->statd[st] * xp[st] * yp[st] * zp[st];
        xp += S; yp += S; zp += S;
      }
    for (int site = 0; site < n; site++) ll_tot += log(util[site]);
    ll_tot -= ((float)n) * log((float)c->R);
  } else {
    const double *xp = x->a, *yp = y->a, *zp = z->a;
    for (int site = 0; site < n; site++) {
      double u = 0;
      for (int st = 0; st < S; st++) u += c->statd[st] * xp[st] * yp[st] * zp[st];
      ll_tot += log(u);
      xp += S; yp += S; zp += S;
    }
  }
  return ll_tot + (LOG_OF_2 * ((total_twoexp(x->e, n) + total_twoexp(y->e, n)) + total_twoexp(z->e, n)));
}

/* Glv.masked_logdot: gmix_model.ml:194-198 -> ten_masked_logdot_c linear_c.c:465-539;
 * gcat_model.ml:220-224 -> mat_masked_logdot_c linear_c.c:251-296 */
static double masked_logdot(const orc_ctx *c, double *util, const glv *x, const glv *y,
                            const uint16_t *mask) {
  const int S = c->S, n = x->n;
  double ll_tot = 0;
  if (!c->gcat) {
    for (int site = 0; site < n; site++) util[site] = 0.0;
    for (int r = 0; r < c->R; r++) {
      const double *xp = x->a + (size_t)r * n * S, *yp = y->a + (size_t)r * n * S;
      for (int site = 0; site < n; site++) {
        if (mask[site]) for (int st = 0; st < S; st++) util[site] += xp[st] * yp[st];
        xp += S; yp += S;
      }
    }
    int n_used = 0;
    for (int site = 0; site < n; site++) if (mask[site]) { ll_tot += log(util[site]); n_used++; }
    ll_tot -= ((float)n_used) * log((float)c->R);
  } else {
    const double *xp = x->a, *yp = y->a;
    for (int site = 0; site < n; site++) {
      if (mask[site]) {
        double u = 0;
        for (int st = 0; st < S; st++) u += xp[st] * yp[st];
        ll_tot += log(u);
      }
      xp += S; yp += S;
    }
  }
  return ll_tot + (LOG_OF_2 * (masked_total_twoexp(x->e, mask, n) + masked_total_twoexp(y->e, mask, n)));
}

/* Glv.logdot = bounded_logdot 0..n-1 (gmix_model.ml:202-213; linear_c.c:541-605 / :298-341) */
static double logdot_all(const orc_ctx *c, double *util, const glv *x, const glv *y) {
  const int S = c->S, n = x->n;
  double ll_tot = 0;
  if (!c->gcat) {
    for (int site = 0; site < n; site++) util[site] = 0.0;
    for (int r = 0; r < c->R; r++) {
      const double *xp = x->a + (size_t)r * n * S, *yp = y->a + (size_t)r * n * S;
      for (int site = 0; site < n; site++) {
        for (int st = 0; st < S; st++) util[site] += xp[st] * yp[st];
        xp += S; yp += S;
      }
    }
    for (int site = 0; site < n; site++) ll_tot += log(util[site]);
    ll_tot -= ((float)n) * log((float)c->R);
  } else {
    const double *xp = x->a, *yp = y->a;
    for (int site = 0; site < n; site++) {
      double u = 0;
      for (int st = 0; st < S; st++) u += xp[st] * yp[st];
      xp += S; yp += S;
      ll_tot += log(u);
    }
  }
  return ll_tot + (LOG_OF_2 * (total_twoexp(x->e, n) + total_twoexp(y->e, n)));
}

/* Model.slow_log_like3 gmix_model.ml:325-356 / gcat_model.ml:357-382 */
static double slow_log_like3(const orc_ctx *c, const glv *x, const glv *y, const glv *z) {
  const int S = c->S, n = x->n;
  double tot = 0.0;
  for (int site = 0; site < n; site++) {
    double site_like = 0.0;
    for (int r = 0; r < c->Rg; r++)
      for (int st = 0; st < S; st++) {
        size_t k = ((size_t)r * n + site) * S + st;
        site_like = site_like + c->statd[st] * x->a[k] * y->a[k] * z->a[k];
      }
    double v = c->gcat ? log(site_like) : log(site_like / (double)c->R);
    tot += v + LOG_OF_2 * (double)(x->e[site] + y->e[site] + z->e[site]);
  }
  return tot;
}

/* ------------------------------------------------------------------------------------------ */
/* context                                                                                      */
/* ------------------------------------------------------------------------------------------ */
orc_ctx *orc_create(int S, int R, int gcat, const double *statd, const double *x,
                    const double *lambda, const double *xit, const double *rates,
                    const int32_t *site_cat, int n_nodes, int root_id, const int32_t *child_ptr,
                    const int32_t *child_idx, const double *bl, const int32_t *leaf_row,
                    const uint8_t *ref_chars, int n_leaves, int L, const char *mmap_dir) {
  if (S > ORC_MAX_STATES || S < 2 || R < 1) return NULL;
  orc_ctx *c = (orc_ctx *)xcalloc(1, sizeof(orc_ctx));
  c->S = S; c->R = R; c->gcat = gcat; c->Rg = gcat ? 1 : R;
  memcpy(c->statd, statd, sizeof(double) * S);
  memcpy(c->x, x, sizeof(double) * S * S);
  memcpy(c->l, lambda, sizeof(double) * S);
  memcpy(c->xit, xit, sizeof(double) * S * S);
  c->rates = (double *)xcalloc((size_t)R, sizeof(double)); memcpy(c->rates, rates, sizeof(double) * R);
  c->n_nodes = n_nodes; c->root_id = root_id; c->L = L; c->n_leaves = n_leaves;
  if (gcat) {
    c->site_cat = (int32_t *)xcalloc((size_t)L, sizeof(int32_t));
    memcpy(c->site_cat, site_cat, sizeof(int32_t) * L);
    c->occupied_rates = (uint8_t *)xcalloc((size_t)R, 1);
    for (int i = 0; i < L; i++) c->occupied_rates[site_cat[i]] = 1;
  }
  c->child_ptr = (int32_t *)xcalloc((size_t)n_nodes + 1, sizeof(int32_t));
  memcpy(c->child_ptr, child_ptr, sizeof(int32_t) * ((size_t)n_nodes + 1));
  int nch = child_ptr[n_nodes];
  c->child_idx = (int32_t *)xcalloc((size_t)nch, sizeof(int32_t));
  memcpy(c->child_idx, child_idx, sizeof(int32_t) * (size_t)nch);
  c->leaf_row = (int32_t *)xcalloc((size_t)n_nodes, sizeof(int32_t));
  memcpy(c->leaf_row, leaf_row, sizeof(int32_t) * (size_t)n_nodes);
  c->bl = (double *)xcalloc((size_t)n_nodes, sizeof(double));
  memcpy(c->bl, bl, sizeof(double) * (size_t)n_nodes);
  c->ref_chars = (uint8_t *)xcalloc((size_t)n_leaves * L, 1);
  memcpy(c->ref_chars, ref_chars, (size_t)n_leaves * L);
  c->glv_len = (size_t)c->Rg * L * S;
  /* 3 x (E+1) GLVs: darr, parr, snodes (pplacer_run.ml:289-309); optionally file-backed like
   * --mmap-file (gstar_support.ml:100-115) when mmap_dir is given. */
  for (int w = 0; w < 3; w++) {
    size_t abytes = sizeof(double) * c->glv_len * (size_t)n_nodes;
    size_t ebytes = sizeof(int64_t) * (size_t)L * (size_t)n_nodes;
    if (mmap_dir) {
      char path[4096];
      snprintf(path, sizeof path, "%s/orc_glv_%d_%d.mmap", mmap_dir, (int)getpid(), w);
      int fd = open(path, O_RDWR | O_CREAT | O_TRUNC, 0600);
      if (fd < 0 || ftruncate(fd, (off_t)(abytes + ebytes)) != 0) { perror("oracle mmap file"); abort(); }
      unlink(path);
      void *p = mmap(NULL, abytes + ebytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
      close(fd);
      if (p == MAP_FAILED) { perror("oracle mmap"); abort(); }
      c->A[w] = (double *)p; c->E[w] = (int64_t *)((char *)p + abytes);
      c->map_bytes[w] = abytes + ebytes; c->mapped = 1;
    } else {
      c->A[w] = (double *)xcalloc(c->glv_len * (size_t)n_nodes, sizeof(double));
      c->E[w] = (int64_t *)xcalloc((size_t)L * (size_t)n_nodes, sizeof(int64_t));
    }
  }
  return c;
}

void orc_destroy(orc_ctx *c) {
  if (!c) return;
  for (int w = 0; w < 3; w++) {
    if (c->mapped) munmap(c->A[w], c->map_bytes[w]);
    else { free(c->A[w]); free(c->E[w]); }
  }
  free(c->rates); free(c->site_cat); free(c->occupied_rates); free(c->child_ptr); free(c->child_idx);
  free(c->fig_rep); free(c->fig_ptr); free(c->fig_idx);
  free(c->leaf_row); free(c->bl); free(c->ref_chars); free(c);
}

const double *orc_glv_a(orc_ctx *c, int which, int node) { return c->A[which] + (size_t)node * c->glv_len; }
const int64_t *orc_glv_e(orc_ctx *c, int which, int node) { return c->E[which] + (size_t)node * c->L; }
int64_t orc_glv_a_len(orc_ctx *c) { return (int64_t)c->glv_len; }

/* ------------------------------------------------------------------------------------------ */
/* Like_stree.calc_distal_and_proximal (like_stree.ml:43-110) + pplacer_run.ml:322-340          */
/* ------------------------------------------------------------------------------------------ */
#define MIN_ALLOWED_TWOEXP (-50) /* like_stree.ml:8 */

/* like_stree.ml:43-69; evolv_dist goes to the snodes array used as scratch (:107-108) */
static void calc_distal(orc_ctx *c, scratch *s, int id) {
  glv distal = glv_at(c, 0, id), evolv_dist = glv_at(c, 2, id);
  int c0 = c->child_ptr[id], c1 = c->child_ptr[id + 1], nch = c1 - c0;
  if (nch > 0) {
    glv *kids = (glv *)xcalloc((size_t)nch, sizeof(glv));
    glv **ptrs = (glv **)xcalloc((size_t)nch, sizeof(glv *));
    for (int k = 0; k < nch; k++) { /* List.map calc tL: children in list order */
      calc_distal(c, s, c->child_idx[c0 + k]);
      kids[k] = glv_at(c, 2, c->child_idx[c0 + k]); ptrs[k] = &kids[k];
    }
    listwise_prod(c, &distal, ptrs, nch);
    perhaps_pull_exponent(c, MIN_ALLOWED_TWOEXP, &distal);
    free(kids); free(ptrs);
  } else {
    prep_constant_rate_glv_from_chars(c, &distal, c->ref_chars + (size_t)c->leaf_row[id] * c->L, NULL, c->L);
  }
  evolve_into(c, s, &evolv_dist, &distal, c->bl[id], NULL);
}

/* like_stree.ml:74-100 */
static void calc_proximal(orc_ctx *c, scratch *s, glv *evolved_prox, int id) {
  int c0 = c->child_ptr[id], c1 = c->child_ptr[id + 1], nch = c1 - c0;
  if (nch == 0) return;
  glv proximal = glv_at(c, 1, id);
  evolve_into(c, s, evolved_prox, &proximal, c->bl[id], NULL);
  glv *others = (glv *)xcalloc((size_t)nch, sizeof(glv));
  glv **ptrs = (glv **)xcalloc((size_t)nch + 1, sizeof(glv *));
  for (int k = 0; k < nch; k++) { /* pull_each_out: (chosen, rest in original order) */
    glv prox_below = glv_at(c, 1, c->child_idx[c0 + k]);
    int m = 0;
    ptrs[m++] = evolved_prox;
    for (int j = 0; j < nch; j++) if (j != k) { others[j] = glv_at(c, 2, c->child_idx[c0 + j]); ptrs[m++] = &others[j]; }
    listwise_prod(c, &prox_below, ptrs, m);
    perhaps_pull_exponent(c, MIN_ALLOWED_TWOEXP, &prox_below);
  }
  free(others); free(ptrs);
  for (int k = 0; k < nch; k++) calc_proximal(c, s, evolved_prox, c->child_idx[c0 + k]);
}

int orc_precompute(orc_ctx *c) {
  scratch s; scratch_init(c, &s);
  calc_distal(c, &s, c->root_id);
  glv top_prox = glv_at(c, 1, c->root_id);
  glv_set_unit(c, &top_prox); /* like_stree.ml:99 */
  calc_proximal(c, &s, &s.t1, c->root_id);
  /* pplacer_run.ml:333-335 pull exponents at -10 on darr and parr */
  for (int w = 0; w < 2; w++)
    for (int i = 0; i < c->n_nodes; i++) { glv g = glv_at(c, w, i); perhaps_pull_exponent(c, -10, &g); }
  /* Glv_arr.prep_supernodes glv_arr.ml:46-56 with half_bl_fun pplacer_run.ml:338 */
  for (int i = 0; i < c->n_nodes; i++) {
    glv d = glv_at(c, 0, i), p = glv_at(c, 1, i), sn = glv_at(c, 2, i);
    evolve_into(c, &s, &s.t1, &d, c->bl[i] / 2.0, NULL);
    evolve_into(c, &s, &s.t2, &p, c->bl[i] / 2.0, NULL);
    statd_pairwise_prod(c, &sn, &s.t1, &s.t2);
  }
  scratch_free(&s);
  c->precomputed = 1;
  return ORC_OK;
}

/* Gcat_model.refine (gcat_model.ml:387-419) over Like_stree.site_log_like_arr (like_stree.ml:113-131) and
 * site_log_like_arr3 (gcat_model.ml:355-375).  Replaces c->site_cat; returns ORC_OK or 1 when a site has zero
 * likelihood (the reference raises Failure).  site_ll_out [R x L] may be NULL. */
int orc_refine_site_categories(orc_ctx *c, int32_t *site_cat_out, double *site_ll_out) {
  if (!c->gcat) return ORC_OK; /* gmix_model.ml: refine is a no-op */
  const int L = c->L, S = c->S, R = c->R;
  scratch s; scratch_init(c, &s);
  double *best = (double *)xcalloc((size_t)L, sizeof(double));
  int32_t *best_cat = (int32_t *)xcalloc((size_t)L, sizeof(int32_t));
  for (int i = 0; i < L; i++) { best[i] = -INFINITY; best_cat[i] = -1; }
  const int c0 = c->child_ptr[c->root_id], nch = c->child_ptr[c->root_id + 1] - c0;
  int bad = 0;
  glv util = glv_make(c, L);
  for (int cat = 0; cat < R && !bad; cat++) {
    for (int i = 0; i < L; i++) c->site_cat[i] = cat;              /* set_site_categories (:380-384) */
    memset(c->occupied_rates, 0, (size_t)R); c->occupied_rates[cat] = 1;
    for (int k = 0; k < nch; k++) calc_distal(c, &s, c->child_idx[c0 + k]); /* calc_distal_and_evolv_dist below the top */
    glv x = glv_at(c, 2, c->child_idx[c0]), y = glv_at(c, 2, c->child_idx[c0 + 1]);
    glv *z = &util;
    if (nch == 2) glv_set_unit(c, &util);
    else if (nch == 3) { static glv zz; zz = glv_at(c, 2, c->child_idx[c0 + 2]); z = &zz; }
    else {
      glv *rest = (glv *)xcalloc((size_t)nch, sizeof(glv));
      glv **ptrs = (glv **)xcalloc((size_t)nch, sizeof(glv *));
      for (int k = 2; k < nch; k++) { rest[k] = glv_at(c, 2, c->child_idx[c0 + k]); ptrs[k - 2] = &rest[k]; }
      listwise_prod(c, &util, ptrs, nch - 2);
      free(rest); free(ptrs);
    }
    const double lr = log(c->rates[cat]);
    for (int site = 0; site < L; site++) {
      double like = 0.0;
      for (int st = 0; st < S; st++)
        like = like + c->statd[st] * x.a[(size_t)site * S + st] * y.a[(size_t)site * S + st] * z->a[(size_t)site * S + st];
      if (0.0 >= like) { bad = 1; break; }
      const double ll = log(like) + LOG_OF_2 * (double)(x.e[site] + y.e[site] + z->e[site]);
      if (site_ll_out) site_ll_out[(size_t)cat * L + site] = ll;
      const double v = ll + 2.0 * lr - 3.0 * c->rates[cat];
      if (v > best[site]) { best[site] = v; best_cat[site] = cat; }
    }
  }
  if (!bad) {
    memset(c->occupied_rates, 0, (size_t)R);
    for (int i = 0; i < L; i++) { c->site_cat[i] = best_cat[i]; c->occupied_rates[best_cat[i]] = 1; if (site_cat_out) site_cat_out[i] = best_cat[i]; }
  }
  c->precomputed = 0;
  glv_free(&util); free(best); free(best_cat); scratch_free(&s);
  return bad ? 1 : ORC_OK;
}

/* Fig.t for the ranking (fig.ml): m figs as (representative, CSR edge lists), built by the host (Fig._figs_of_gtree);
 * m == 0 restores the Dummy ranking of every non-root edge */
int orc_set_figs(orc_ctx *c, int m, const int32_t *rep, const int32_t *ptr, const int32_t *idx) {
  free(c->fig_rep); free(c->fig_ptr); free(c->fig_idx);
  c->fig_rep = c->fig_ptr = c->fig_idx = NULL;
  c->n_figs = 0;
  if (m <= 0) return ORC_OK;
  c->fig_rep = (int32_t *)xcalloc((size_t)m, sizeof(int32_t));
  c->fig_ptr = (int32_t *)xcalloc((size_t)m + 1, sizeof(int32_t));
  c->fig_idx = (int32_t *)xcalloc((size_t)ptr[m] + 1, sizeof(int32_t));
  memcpy(c->fig_rep, rep, sizeof(int32_t) * (size_t)m);
  memcpy(c->fig_ptr, ptr, sizeof(int32_t) * ((size_t)m + 1));
  memcpy(c->fig_idx, idx, sizeof(int32_t) * (size_t)ptr[m]);
  c->n_figs = m;
  return ORC_OK;
}

/* tests/pplacer/test_like.ml:57-66 */
void orc_node_tree_likes(orc_ctx *c, int node, double *slow_ll3, double *sn_logdot) {
  scratch s; scratch_init(c, &s);
  glv one = glv_make(c, c->L); glv_set_unit(c, &one);
  glv d = glv_at(c, 0, node), p = glv_at(c, 1, node), sn = glv_at(c, 2, node);
  evolve_into(c, &s, &s.t1, &d, c->bl[node] / 2.0, NULL);
  evolve_into(c, &s, &s.t2, &p, c->bl[node] / 2.0, NULL);
  *slow_ll3 = slow_log_like3(c, &s.t1, &s.t2, &one);
  *sn_logdot = logdot_all(c, s.util, &sn, &one);
  glv_free(&one); scratch_free(&s);
}

/* ------------------------------------------------------------------------------------------ */
/* Three_tax (three_tax.ml) over Glv_edge (glv_edge.ml)                                         */
/* ------------------------------------------------------------------------------------------ */
typedef struct { glv orig, evolv; double bl; } glv_edge;

typedef struct {
  const orc_ctx *c; scratch *s;
  glv_edge prox, dist, pend;
  const int32_t *reind; /* Gcat only (core.ml:176-179) */
  int nq;
  int64_t n_like_calls;
} three_tax;

static void edge_set_bl(three_tax *tt, glv_edge *e, double bl) { /* glv_edge.ml:29-35 */
  e->bl = bl;
  evolve_into(tt->c, tt->s, &e->evolv, &e->orig, bl, tt->c->gcat ? tt->reind : NULL);
}
static double tt_cut_bl(const three_tax *tt) { return tt->dist.bl + tt->prox.bl; } /* three_tax.ml:26 */
static double tt_log_like(three_tax *tt) { /* three_tax.ml:30-36 */
  return log_like3(tt->c, tt->s->util, &tt->prox.evolv, &tt->dist.evolv, &tt->pend.evolv);
}
static void tt_set_pend_bl(three_tax *tt, double v) { edge_set_bl(tt, &tt->pend, v); }
static int tt_set_dist_bl(three_tax *tt, double dist_bl) { /* three_tax.ml:41-45 */
  double prox_bl = tt_cut_bl(tt) - dist_bl;
  if (!(prox_bl >= 0.)) return ORC_ASSERT;
  edge_set_bl(tt, &tt->dist, dist_bl);
  edge_set_bl(tt, &tt->prox, prox_bl);
  return ORC_OK;
}
static double opt_pend_fun(double v, void *ud) { /* three_tax.ml:51-54 */
  three_tax *tt = (three_tax *)ud; tt->n_like_calls++;
  tt_set_pend_bl(tt, v); return -tt_log_like(tt);
}
static double opt_dist_fun(double v, void *ud) {
  three_tax *tt = (three_tax *)ud; tt->n_like_calls++;
  if (tt_set_dist_bl(tt, v) != ORC_OK) return NAN;
  return -tt_log_like(tt);
}
/* three_tax.ml:60-68 */
static int optimize_pend_bl(three_tax *tt, double tolerance, double max_value, double *out) {
  double start_v = tt->pend.bl, min_v = 1e-8, max_v = max_value;
  if (min_v == max_v) { *out = start_v; return ORC_OK; }
  return orc_brent(opt_pend_fun, tt, start_v, min_v, max_v, tolerance, 100, out, NULL);
}
static int optimize_dist_bl(three_tax *tt, double tolerance, double *out) {
  double start_v = tt->dist.bl, min_v = 0., max_v = tt_cut_bl(tt);
  if (min_v == max_v) { *out = start_v; return ORC_OK; }
  return orc_brent(opt_dist_fun, tt, start_v, min_v, max_v, tolerance, 100, out, NULL);
}
/* three_tax.ml:70-83 */
static int tt_optimize(three_tax *tt, double tolerance, double max_pend_bl, int max_iter) {
  double prev_pend = tt->pend.bl, prev_dist = tt->dist.bl;
  tt->n_like_calls = 0;
  for (int which_step = 1;; which_step++) {
    if (which_step > max_iter) return ORC_EXCEEDED_MAX_ITER;
    double curr_pend, curr_dist;
    int st = optimize_pend_bl(tt, tolerance, max_pend_bl, &curr_pend); if (st) return st;
    st = optimize_dist_bl(tt, tolerance, &curr_dist); if (st) return st;
    if (fabs(prev_pend - curr_pend) > tolerance || fabs(prev_dist - curr_dist) > tolerance) {
      prev_pend = curr_pend; prev_dist = curr_dist;
    } else return ORC_OK;
  }
}

/* prior (core.ml:87-92; gsl_ran_exponential_pdf) */
typedef struct { int kind; double mean; } prior_t;
static double prior_eval(const prior_t *p, double x) {
  if (p->kind == 0) return 1.0;
  if (x < 0) return 0.0;
  return exp(-x / p->mean) / p->mean;
}

/* three_tax.ml:124-159 find_upper_limit */
typedef struct { three_tax *tt; const prior_t *prior; double target_ll; } ful_ud;
static double ful_fun(double pend_bl, void *ud_) {
  ful_ud *ud = (ful_ud *)ud_;
  tt_set_pend_bl(ud->tt, pend_bl);
  double like = tt_log_like(ud->tt) + log(prior_eval(ud->prior, pend_bl));
  double x = fabs(ud->target_ll - like);
  if (x == INFINITY) x = 100000.0;
  return x;
}
static int find_upper_limit(three_tax *tt, double max_pend, const prior_t *prior, double orig_ll, double *out) {
  double orig_pend_bl = tt->pend.bl;
  if (prior_eval(prior, orig_pend_bl) == 0.) {
    double curr = orig_pend_bl;
    for (;;) { double next = curr / 2.; if (prior_eval(prior, next) > exp(-10.)) { *out = curr; return ORC_OK; } curr = next; }
  }
  ful_ud ud = {tt, prior, -10. + orig_ll + log(prior_eval(prior, orig_pend_bl))};
  double start = orig_pend_bl + ((max_pend - orig_pend_bl) / 2.);
  int st = orc_brent(ful_fun, &ud, start, orig_pend_bl, max_pend, 0.02, 10, out, NULL);
  if (st == ORC_GSL_EINVAL) { *out = max_pend; return ORC_OK; }
  return st; /* any other failure is uncaught in the reference (fatal) */
}

/* three_tax.ml:169-215 calc_marg_prob */
/* diagnostic counters (tools/pp_stats.py): placements, outer integrand calls, inner integrand calls, ETOL retries */
static long long g_pp_stats[4];
typedef struct { three_tax *tt; const prior_t *prior; double base_ll, upper_limit, rel_err; int status; } marg_ud;
static double marg_inner_fun(double pend_bl, void *ud_) {
  marg_ud *ud = (marg_ud *)ud_;
  tt_set_pend_bl(ud->tt, pend_bl);
  return exp(tt_log_like(ud->tt) - ud->base_ll) * prior_eval(ud->prior, pend_bl);
}
static double marg_inner_integration(marg_ud *ud) {
  double res, err; int neval;
  int st = orc_qng(marg_inner_fun, ud, 0., ud->upper_limit, 0., ud->rel_err, &res, &err, &neval);
  __atomic_fetch_add(&g_pp_stats[2], neval, __ATOMIC_RELAXED);
  if (st && !ud->status) ud->status = st;
  return res;
}
void orc_pp_stats(long long *out4, int reset) {
  for (int i = 0; i < 4; i++) { out4[i] = g_pp_stats[i]; if (reset) g_pp_stats[i] = 0; }
}
static double marg_outer_fun(double dist_bl, void *ud_) {
  marg_ud *ud = (marg_ud *)ud_;
  __atomic_fetch_add(&g_pp_stats[1], 1, __ATOMIC_RELAXED);
  if (ud->status) return 0.0; /* exception already unwinding */
  if (tt_set_dist_bl(ud->tt, dist_bl) != ORC_OK) { ud->status = ORC_ASSERT; return 0.0; }
  return marg_inner_integration(ud);
}
static int calc_marg_prob(three_tax *tt, const prior_t *prior, double rel_err, double max_pend, double *out) {
  const int max_n_exceptions = 10;
  double base_ll = tt_log_like(tt), cut_bl = tt_cut_bl(tt);
  int n_exceptions = 0;
  __atomic_fetch_add(&g_pp_stats[0], 1, __ATOMIC_RELAXED);
  double upper_limit;
  int st = find_upper_limit(tt, max_pend, prior, base_ll, &upper_limit);
  if (st) return st;
  for (;;) {
    if (n_exceptions >= max_n_exceptions) { *out = base_ll; return ORC_OK; }
    marg_ud ud = {tt, prior, base_ll, upper_limit, rel_err, 0};
    double val;
    if (fabs(cut_bl - 0.) < 1e-5) { /* =~ ppatteries.ml:96 */
      if (tt_set_dist_bl(tt, cut_bl / 2.) != ORC_OK) return ORC_ASSERT;
      val = marg_inner_integration(&ud);
    } else {
      double res, err; int neval;
      int so = orc_qng(marg_outer_fun, &ud, 0., cut_bl, 0., rel_err, &res, &err, &neval);
      if (!ud.status && so) ud.status = so;
      val = res / cut_bl;
    }
    if (ud.status == ORC_GSL_ETOL) { n_exceptions++; upper_limit = upper_limit / 2.; __atomic_fetch_add(&g_pp_stats[3], 1, __ATOMIC_RELAXED); continue; }
    if (ud.status) return ud.status;
    *out = base_ll + log(val);
    return ORC_OK;
  }
}

/* ------------------------------------------------------------------------------------------ */
/* Core.pplacer_core (core.ml:76-409)                                                           */
/* ------------------------------------------------------------------------------------------ */
typedef struct { double score; int loc; } scored;
static int ocaml_float_compare(double a, double b) { /* Stdlib.compare on floats: nan < everything, nan = nan */
  if (a < b) return -1; if (a > b) return 1; if (a == b) return 0;
  if (a != a) return (b != b) ? 0 : -1;
  return 1;
}
/* List.sort (comparing fst |> flip): stable merge sort, descending */
static void stable_sort_desc(scored *v, scored *tmp, int n) {
  if (n < 2) return;
  int h = n / 2;
  stable_sort_desc(v, tmp, h); stable_sort_desc(v + h, tmp, n - h);
  int i = 0, j = h, k = 0;
  while (i < h && j < n) { if (ocaml_float_compare(v[j].score, v[i].score) > 0) tmp[k++] = v[j++]; else tmp[k++] = v[i++]; }
  while (i < h) tmp[k++] = v[i++];
  while (j < n) tmp[k++] = v[j++];
  memcpy(v, tmp, sizeof(scored) * (size_t)n);
}

typedef struct { int loc; double like, pend, dist; } ml_res;

typedef struct {
  orc_ctx *c; scratch s;
  glv full_query_orig, full_query_evolv;
  uint16_t *mask_vec; int32_t *reind;
  scored *rank, *rank_tmp;
  /* three-taxon storage sized L */
  glv_edge dist, prox, pend;
} worker;

static void worker_init(orc_ctx *c, worker *w) {
  w->c = c; scratch_init(c, &w->s);
  w->full_query_orig = glv_make(c, c->L); w->full_query_evolv = glv_make(c, c->L);
  w->mask_vec = (uint16_t *)xcalloc((size_t)c->L, sizeof(uint16_t));
  w->reind = (int32_t *)xcalloc((size_t)c->L, sizeof(int32_t));
  w->rank = (scored *)xcalloc((size_t)c->n_nodes, sizeof(scored));
  w->rank_tmp = (scored *)xcalloc((size_t)c->n_nodes, sizeof(scored));
  glv_edge *es[3] = {&w->dist, &w->prox, &w->pend};
  for (int i = 0; i < 3; i++) { es[i]->orig = glv_make(c, c->L); es[i]->evolv = glv_make(c, c->L); es[i]->bl = 0; }
}
static void worker_free(worker *w) {
  scratch_free(&w->s); glv_free(&w->full_query_orig); glv_free(&w->full_query_evolv);
  free(w->mask_vec); free(w->reind); free(w->rank); free(w->rank_tmp);
  glv_edge *es[3] = {&w->dist, &w->prox, &w->pend};
  for (int i = 0; i < 3; i++) { glv_free(&es[i]->orig); glv_free(&es[i]->evolv); }
}

/* Glv.mask_into gmix_model.ml:130-141 / gcat_model.ml:167-180 via reind (copy_site per informative site) */
static void mask_into(const orc_ctx *c, const int32_t *reind, int nq, const glv *src, glv *dst) {
  const int S = c->S;
  for (int i = 0; i < nq; i++) {
    int si = reind[i];
    dst->e[i] = src->e[si];
    for (int r = 0; r < c->Rg; r++)
      memcpy(dst->a + ((size_t)r * nq + i) * S, src->a + ((size_t)r * src->n + si) * S, sizeof(double) * S);
  }
}

/* core.ml:225-239 set_tt_edges */
static void set_tt_edges(three_tax *tt, orc_ctx *c, int loc, double pendant, double distal) {
  double cut_bl = c->bl[loc];
  tt_set_pend_bl(tt, pendant);
  glv d = glv_at(c, 0, loc), p = glv_at(c, 1, loc);
  mask_into(c, tt->reind, tt->nq, &d, &tt->dist.orig); edge_set_bl(tt, &tt->dist, distal);
  mask_into(c, tt->reind, tt->nq, &p, &tt->prox.orig); edge_set_bl(tt, &tt->prox, cut_bl - distal);
}

static int is_gsl_err(int st) { return st == ORC_GSL_EINVAL || st == ORC_GSL_EBADFUNC || st == ORC_GSL_ETOL || st == ORC_GSL_EBADTOL; }

/* core.ml:254-270 ml_optimize_location: ExceededMaxIter is caught and the current state kept */
static int ml_optimize_location(three_tax *tt, const orc_prefs *pf, double tol, ml_res *out, int loc, int64_t *calls) {
  int st = tt_optimize(tt, tol, pf->max_pend, pf->max_iter);
  if (calls) *calls += tt->n_like_calls;
  if (st != ORC_OK && st != ORC_EXCEEDED_MAX_ITER) return st;
  out->loc = loc; out->like = tt_log_like(tt); out->pend = tt->pend.bl; out->dist = tt->dist.bl; /* :85 */
  return ORC_OK;
}
/* core.ml:272-279 safe_ml_optimize_location */
static int safe_ml_optimize_location(three_tax *tt, orc_ctx *c, const orc_prefs *pf, double tol, int loc, ml_res *out, int64_t *calls) {
  int st = ml_optimize_location(tt, pf, tol, out, loc, calls);
  if (is_gsl_err(st)) {
    set_tt_edges(tt, c, loc, pf->start_pend, c->bl[loc] / 2.); /* tt_edges_default :240-245 */
    st = ml_optimize_location(tt, pf, tol, out, loc, calls);
  }
  return st;
}

static double now_s(void) { struct timespec ts; clock_gettime(CLOCK_THREAD_CPUTIME_ID, &ts); return ts.tv_sec + 1e-9 * ts.tv_nsec; }

static int process_query(worker *w, const orc_prefs *pf, const uint8_t *query, orc_result *out,
                         int32_t *n_out, orc_debug *dbg, int64_t qi) {
  orc_ctx *c = w->c;
  const int L = c->L;
  double t0 = now_s();
  *n_out = 0;
  int t_max_pitches, t_max_strikes; /* core.ml:95-106 (fantasy not supported) */
  if (pf->max_strikes == 0) { t_max_pitches = 0x7fffffff; t_max_strikes = 0x7fffffff; }
  else { t_max_pitches = pf->max_pitches; t_max_strikes = pf->max_strikes; }
  /* core.ml:121-141 */
  int st = prep_constant_rate_glv_from_chars(c, &w->full_query_orig, query, NULL, L);
  if (st) return st;
  evolve_into(c, &w->s, &w->full_query_evolv, &w->full_query_orig, pf->start_pend, NULL);
  /* core.ml:146-158 */
  int nq = 0;
  for (int i = 0; i < L; i++) { int inf = informative(query[i]); w->mask_vec[i] = inf ? 1 : 0; if (inf) w->reind[nq++] = i; }
  if (nq == 0) return ORC_NO_INFORMATIVE;
  /* core.ml:166-184: query_glv and the three edges, all with nq sites */
  three_tax tt; memset(&tt, 0, sizeof tt);
  tt.c = c; tt.s = &w->s; tt.reind = w->reind; tt.nq = nq;
  tt.dist = w->dist; tt.prox = w->prox; tt.pend = w->pend;
  tt.dist.orig.n = tt.dist.evolv.n = tt.prox.orig.n = tt.prox.evolv.n = tt.pend.orig.n = tt.pend.evolv.n = nq;
  st = prep_constant_rate_glv_from_chars(c, &tt.pend.orig, query, w->reind, nq);
  if (st) return st;
  tt.dist.bl = tt.prox.bl = pf->start_pend;
  edge_set_bl(&tt, &tt.pend, pf->start_pend);
  double t1 = now_s();
  int n_rank = 0;
  /* Fig._enum_by_score (fig.ml:120-138): representatives scored and sorted up front, groups expanded lazily */
  scored *fig_sc = NULL, *fig_tmp = NULL;
  uint8_t *yielded = NULL;
  int fig_next = 0;
  if (c->n_figs > 0) {
    fig_sc = (scored *)xcalloc((size_t)c->n_figs, sizeof(scored)); fig_tmp = (scored *)xcalloc((size_t)c->n_figs, sizeof(scored));
    yielded = (uint8_t *)xcalloc((size_t)c->n_nodes, 1);
    if (dbg && dbg->scores) for (int loc = 0; loc < c->n_nodes; loc++) dbg->scores[qi * c->n_nodes + loc] = NAN;
    for (int f = 0; f < c->n_figs; f++) {
      glv sn = glv_at(c, 2, c->fig_rep[f]);
      fig_sc[f].score = masked_logdot(c, w->s.util, &w->full_query_evolv, &sn, w->mask_vec);
      fig_sc[f].loc = f;
    }
    stable_sort_desc(fig_sc, fig_tmp, c->n_figs);
  } else {
    /* core.ml:187-197 + fig.ml:143-147 (Dummy: every non-root id ascending, stable sort descending) */
    for (int loc = 0; loc < c->n_nodes; loc++) {
      if (loc == c->root_id) { if (dbg && dbg->scores) dbg->scores[qi * c->n_nodes + loc] = NAN; continue; }
      glv sn = glv_at(c, 2, loc);
      double sc = masked_logdot(c, w->s.util, &w->full_query_evolv, &sn, w->mask_vec);
      if (dbg && dbg->scores) dbg->scores[qi * c->n_nodes + loc] = sc;
      w->rank[n_rank].score = sc; w->rank[n_rank].loc = loc; n_rank++;
    }
    stable_sort_desc(w->rank, w->rank_tmp, n_rank);
  }
  double t2 = now_s();
  /* core.ml:283-318 play_ball */
  int cap = (pf->max_strikes == 0) ? n_rank : (pf->max_pitches + 1);
  if (cap > n_rank) cap = n_rank;
  ml_res *results = (ml_res *)xcalloc((size_t)c->n_nodes, sizeof(ml_res));
  int n_results = 0, n_strikes = 0, next = 0;
  double like_record = -INFINITY;
  int64_t calls = 0;
  int fatal = 0;
  for (;;) {
    if (next >= n_rank) {
      if (fig_next >= c->n_figs) break;
      /* the next group: this fig plus the following ones whose representative scores lie within strike_box of its own
       * (approx_equal, ppatteries.ml:96: strict <), minus everything yielded before, members scored and sorted */
      const double s0 = fig_sc[fig_next].score;
      int g0 = fig_next++;
      while (fig_next < c->n_figs && fabs(s0 - fig_sc[fig_next].score) < pf->strike_box) fig_next++;
      int n_new = 0;
      uint8_t *in_group = (uint8_t *)xcalloc((size_t)c->n_nodes, 1);
      for (int g = g0; g < fig_next; g++) {
        const int f = fig_sc[g].loc;
        for (int k = c->fig_ptr[f]; k < c->fig_ptr[f + 1]; k++) in_group[c->fig_idx[k]] = 1;
      }
      for (int loc = 0; loc < c->n_nodes; loc++) { /* IntSet.elements: ascending */
        if (!in_group[loc] || yielded[loc]) continue;
        yielded[loc] = 1;
        glv sn = glv_at(c, 2, loc);
        double sc = masked_logdot(c, w->s.util, &w->full_query_evolv, &sn, w->mask_vec);
        if (dbg && dbg->scores) dbg->scores[qi * c->n_nodes + loc] = sc;
        w->rank[n_rank + n_new].score = sc; w->rank[n_rank + n_new].loc = loc; n_new++;
      }
      free(in_group);
      stable_sort_desc(w->rank + n_rank, w->rank_tmp, n_new);
      n_rank += n_new;
      continue;
    }
    int loc = w->rank[next++].loc;
    set_tt_edges(&tt, c, loc, pf->start_pend, c->bl[loc] / 2.); /* prepare_tt */
    ml_res r;
    st = safe_ml_optimize_location(&tt, c, pf, pf->initial_tolerance, loc, &r, &calls);
    if (is_gsl_err(st)) continue; /* skipped with warning core.ml:291-297 */
    if (st) { fatal = st; break; }
    int len_before = n_results;
    results[n_results++] = r;
    if (len_before >= t_max_pitches) break;
    else if (r.like > like_record) like_record = r.like;
    else if (r.like < like_record - pf->strike_box) {
      if (n_strikes + 1 >= t_max_strikes) break;
      n_strikes++;
    }
  }
  (void)cap;
  free(fig_sc); free(fig_tmp); free(yielded);
  if (fatal) { free(results); return fatal; }
  if (n_results == 0) { free(results); return ORC_EMPTY_RESULTS; }
  if (dbg) {
    int K = pf->max_pitches + 1;
    if (dbg->n_pitches) dbg->n_pitches[qi] = n_results;
    for (int i = 0; i < n_results && i < K; i++) {
      if (dbg->pitch_edges) dbg->pitch_edges[qi * K + i] = results[i].loc;
      if (dbg->pitch_ll) dbg->pitch_ll[qi * K + i] = results[i].like;
      if (dbg->pitch_pend) dbg->pitch_pend[qi * K + i] = results[i].pend;
      if (dbg->pitch_dist) dbg->pitch_dist[qi * K + i] = results[i].dist;
    }
  }
  /* core.ml:336-352: stable sort by like descending, keep filter */
  {
    scored *sv = (scored *)xcalloc((size_t)n_results, sizeof(scored)), *tmp = (scored *)xcalloc((size_t)n_results, sizeof(scored));
    ml_res *cp = (ml_res *)xcalloc((size_t)n_results, sizeof(ml_res));
    for (int i = 0; i < n_results; i++) { sv[i].score = results[i].like; sv[i].loc = i; }
    stable_sort_desc(sv, tmp, n_results);
    for (int i = 0; i < n_results; i++) cp[i] = results[sv[i].loc];
    memcpy(results, cp, sizeof(ml_res) * (size_t)n_results);
    free(sv); free(tmp); free(cp);
  }
  double best_like = results[0].like;
  double log_keep_factor = log(pf->keep_factor);
  int n_keep = 0;
  ml_res *keep = (ml_res *)xcalloc((size_t)n_results, sizeof(ml_res));
  for (int i = 0; i < n_results; i++)
    if ((i < pf->keep_at_most) && (results[i].like >= log_keep_factor + best_like)) keep[n_keep++] = results[i];
  /* core.ml:354-369 refinement */
  for (int i = 0; i < n_keep; i++) {
    ml_res r;
    set_tt_edges(&tt, c, keep[i].loc, keep[i].pend, keep[i].dist);
    st = safe_ml_optimize_location(&tt, c, pf, pf->final_tolerance, keep[i].loc, &r, &calls);
    if (is_gsl_err(st)) continue; /* keeps `initial` */
    if (st) { free(results); free(keep); return st; }
    keep[i] = r;
  }
  /* core.ml:370-377 ll_normalized_prob ppatteries.ml:175-183 */
  for (int i = 0; i < n_keep; i++) {
    double tot = 0.0;
    for (int j = 0; j < n_keep; j++) tot += exp(keep[j].like - keep[i].like);
    out[i].edge = keep[i].loc; out[i].log_like = keep[i].like; out[i].pendant = keep[i].pend;
    out[i].distal = keep[i].dist; out[i].ml_ratio = 1. / tot; out[i].marginal_like = NAN; out[i].post_prob = NAN;
  }
  double t3 = now_s(), t4 = t3;
  /* core.ml:379-404 */
  if (pf->calc_pp) {
    for (int i = 0; i < n_keep; i++) {
      prior_t prior = {pf->prior_kind, pf->prior_kind == 2 ? pf->prior_means[out[i].edge] : pf->prior_mean};
      set_tt_edges(&tt, c, out[i].edge, out[i].pendant, out[i].distal); /* tt_edges_from_placement */
      double mp;
      st = calc_marg_prob(&tt, &prior, pf->pp_rel_err, pf->max_pend, &mp);
      if (st) { free(results); free(keep); return st; }
      out[i].marginal_like = mp;
    }
    for (int i = 0; i < n_keep; i++) {
      double tot = 0.0;
      for (int j = 0; j < n_keep; j++) tot += exp(out[j].marginal_like - out[i].marginal_like);
      out[i].post_prob = 1. / tot;
    }
    t4 = now_s();
  }
  /* Pquery.make_ml_sorted pquery.ml:86-92: stable sort by ml_ratio descending */
  {
    scored *sv = (scored *)xcalloc((size_t)n_keep, sizeof(scored)), *tmp = (scored *)xcalloc((size_t)n_keep, sizeof(scored));
    orc_result *cp = (orc_result *)xcalloc((size_t)n_keep, sizeof(orc_result));
    for (int i = 0; i < n_keep; i++) { sv[i].score = out[i].ml_ratio; sv[i].loc = i; }
    stable_sort_desc(sv, tmp, n_keep);
    for (int i = 0; i < n_keep; i++) cp[i] = out[sv[i].loc];
    memcpy(out, cp, sizeof(orc_result) * (size_t)n_keep);
    free(sv); free(tmp); free(cp);
  }
  *n_out = n_keep;
  if (dbg && dbg->n_like_calls) dbg->n_like_calls[qi] = calls;
  if (dbg && dbg->timing) {
    /* not atomic across threads on purpose: only read when n_threads == 1 */
    dbg->timing[0] += t2 - t1; dbg->timing[1] += t3 - t2; dbg->timing[2] += t4 - t3; dbg->timing[3] += t1 - t0;
  }
  free(results); free(keep);
  return ORC_OK;
}

void orc_default_prefs(orc_prefs *p) { /* prefs.ml:65-120, core.ml:10,14 */
  memset(p, 0, sizeof *p);
  p->start_pend = 0.1; p->max_pend = 2.; p->initial_tolerance = 0.01; p->final_tolerance = 1e-5;
  p->max_iter = 200; p->strike_box = 3.; p->max_strikes = 6; p->max_pitches = 40; p->keep_at_most = 7;
  p->keep_factor = 0.01; p->calc_pp = 0; p->pp_rel_err = 0.01; p->prior_kind = 1; p->prior_mean = 1.0;
}

typedef struct {
  orc_ctx *c; const orc_prefs *pf; int64_t Q; const uint8_t *chars; orc_result *out; int32_t *n_out, *status;
  orc_debug *dbg; int64_t *next; pthread_mutex_t *mu;
} job;

static void *worker_main(void *arg) {
  job *j = (job *)arg;
  worker w; worker_init(j->c, &w);
  for (;;) {
    pthread_mutex_lock(j->mu); int64_t q = (*j->next)++; pthread_mutex_unlock(j->mu);
    if (q >= j->Q) break;
    j->status[q] = process_query(&w, j->pf, j->chars + (size_t)q * j->c->L, j->out + (size_t)q * j->pf->keep_at_most,
                                 &j->n_out[q], j->dbg, q);
  }
  worker_free(&w);
  return NULL;
}

int orc_place(orc_ctx *c, const orc_prefs *pf, int64_t Q, const uint8_t *query_chars, orc_result *out,
              int32_t *n_out, int32_t *status, orc_debug *dbg, int n_threads) {
  if (!c->precomputed) return ORC_ASSERT;
  if (n_threads < 1) n_threads = 1;
  int64_t next = 0; pthread_mutex_t mu = PTHREAD_MUTEX_INITIALIZER;
  job j = {c, pf, Q, query_chars, out, n_out, status, dbg, &next, &mu};
  if (n_threads == 1) { worker_main(&j); return ORC_OK; }
  pthread_t *th = (pthread_t *)xcalloc((size_t)n_threads, sizeof(pthread_t));
  for (int t = 0; t < n_threads; t++) pthread_create(&th[t], NULL, worker_main, &j);
  for (int t = 0; t < n_threads; t++) pthread_join(th[t], NULL);
  free(th);
  return ORC_OK;
}

int orc_three_tax_ll(orc_ctx *c, const uint8_t *query, int edge, double pendant, double distal, double *ll_out) {
  worker w; worker_init(c, &w);
  int nq = 0;
  for (int i = 0; i < c->L; i++) if (informative(query[i])) w.reind[nq++] = i;
  if (!nq) { worker_free(&w); return ORC_NO_INFORMATIVE; }
  three_tax tt; memset(&tt, 0, sizeof tt);
  tt.c = c; tt.s = &w.s; tt.reind = w.reind; tt.nq = nq; tt.dist = w.dist; tt.prox = w.prox; tt.pend = w.pend;
  tt.dist.orig.n = tt.dist.evolv.n = tt.prox.orig.n = tt.prox.evolv.n = tt.pend.orig.n = tt.pend.evolv.n = nq;
  int st = prep_constant_rate_glv_from_chars(c, &tt.pend.orig, query, w.reind, nq);
  if (!st) { set_tt_edges(&tt, c, edge, pendant, distal); *ll_out = tt_log_like(&tt); }
  worker_free(&w);
  return st;
}

/* ------------------------------------------------------------------------------------------ */
/* Kernel-level hooks: the static inner loops above on caller-owned arrays, so that             */
/* tests/test_oracle_vs_ref.py can hold them BIT FOR BIT against the reference's own            */
/* linear_c.c (oracle/_ref).  Exponent vectors are zero, so the ln2 terms add exactly 0.        */
/* ------------------------------------------------------------------------------------------ */
static void k_ctx(orc_ctx *c, int S, int R, int gcat, const double *statd) {
  memset(c, 0, sizeof *c);
  c->S = S; c->R = R; c->gcat = gcat; c->Rg = gcat ? 1 : R;
  if (statd) memcpy(c->statd, statd, sizeof(double) * (size_t)S);
}
static glv k_view(const double *a, int64_t *zeros, int n) { glv g; g.n = n; g.e = zeros; g.a = (double *)a; return g; }

void orc_k_gemmish(double *dst, const double *a, const double *b, int S, int n_sites) { gemmish(dst, a, b, S, n_sites); }
void orc_k_dediagonalize(double *dst, const double *u, const double *lambda, const double *uit, int n) {
  dediagonalize(dst, u, lambda, uit, n);
}
void orc_k_pairwise_prod(double *dst, const double *x, const double *y, int R, int gcat, int n, int S) {
  orc_ctx c; k_ctx(&c, S, R, gcat, NULL);
  int64_t *z = (int64_t *)xcalloc((size_t)n, sizeof(int64_t)), *ze = (int64_t *)xcalloc((size_t)n, sizeof(int64_t));
  glv gx = k_view(x, z, n), gy = k_view(y, z, n), gd = k_view(dst, ze, n);
  pairwise_prod(&c, &gd, &gx, &gy);
  free(z); free(ze);
}
void orc_k_statd_pairwise_prod(const double *statd, double *dst, const double *a, const double *b, int R, int gcat,
                               int n, int S) {
  orc_ctx c; k_ctx(&c, S, R, gcat, statd);
  int64_t *z = (int64_t *)xcalloc((size_t)n, sizeof(int64_t)), *ze = (int64_t *)xcalloc((size_t)n, sizeof(int64_t));
  glv ga = k_view(a, z, n), gb = k_view(b, z, n), gd = k_view(dst, ze, n);
  statd_pairwise_prod(&c, &gd, &ga, &gb);
  free(z); free(ze);
}
double orc_k_log_like3(const double *statd, const double *x, const double *y, const double *z_, int R, int gcat, int n,
                       int S) {
  orc_ctx c; k_ctx(&c, S, R, gcat, statd);
  int64_t *z = (int64_t *)xcalloc((size_t)n, sizeof(int64_t));
  double *util = (double *)xcalloc((size_t)n, sizeof(double));
  glv gx = k_view(x, z, n), gy = k_view(y, z, n), gz = k_view(z_, z, n);
  const double r = log_like3(&c, util, &gx, &gy, &gz);
  free(z); free(util);
  return r;
}
double orc_k_masked_logdot(const double *x, const double *y, const uint16_t *mask, int R, int gcat, int n, int S) {
  orc_ctx c; k_ctx(&c, S, R, gcat, NULL);
  int64_t *z = (int64_t *)xcalloc((size_t)n, sizeof(int64_t));
  double *util = (double *)xcalloc((size_t)n, sizeof(double));
  glv gx = k_view(x, z, n), gy = k_view(y, z, n);
  const double r = masked_logdot(&c, util, &gx, &gy, mask);
  free(z); free(util);
  return r;
}
double orc_k_logdot(const double *x, const double *y, int R, int gcat, int n, int S) {
  orc_ctx c; k_ctx(&c, S, R, gcat, NULL);
  int64_t *z = (int64_t *)xcalloc((size_t)n, sizeof(int64_t));
  double *util = (double *)xcalloc((size_t)n, sizeof(double));
  glv gx = k_view(x, z, n), gy = k_view(y, z, n);
  const double r = logdot_all(&c, util, &gx, &gy);
  free(z); free(util);
  return r;
}
