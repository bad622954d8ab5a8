/* TEST INFRASTRUCTURE ONLY — never linked into or called by the product path.
 *
 * Plain-C entry points over the reference's own pplacer_src/linear_c.c (compiled UNMODIFIED from
 * /root/reference by oracle/Makefile against oracle/caml_shim/ into oracle/_ref/liblinear_ref.so).
 * Each wrapper builds the shim's bigarray descriptors around caller buffers and calls the
 * reference CAMLprim function of the same name (linear_c.c line cited per wrapper).
 */
#include <stddef.h>
#include <stdint.h>
#include <setjmp.h>
#include <caml/mlvalues.h>
#include <caml/bigarray.h>
#include <caml/alloc.h>
#include <caml/fail.h>

__thread jmp_buf ppb_shim_fail_jmp;
__thread const char *ppb_shim_fail_msg;

/* the reference symbols */
value gemmish_c(value, value, value);
value dediagonalize(value, value, value, value);
value mat_log_like3_c(value, value, value, value);
value mat_pairwise_prod_c(value, value, value);
value mat_statd_pairwise_prod_c(value, value, value, value);
value mat_masked_logdot_c(value, value, value);
value mat_bounded_logdot_c(value, value, value, value);
value ten_log_like3_c(value, value, value, value, value);
value ten_pairwise_prod_c(value, value, value);
value ten_statd_pairwise_prod_c(value, value, value, value);
value ten_masked_logdot_c(value, value, value, value);
value ten_bounded_logdot_c(value, value, value, value, value);
value vec_pairwise_prod_c(value, value, value);
value int_vec_tot_c(value);
value int_vec_pairwise_prod_c(value, value, value);
value float_mat_int_vec_mul_c(value, value, value);

static struct caml_ba_array ba1(const void *p, intnat d0) {
  struct caml_ba_array b = {(void *)p, {d0, 0, 0, 0}}; return b; }
static struct caml_ba_array ba2(const void *p, intnat d0, intnat d1) {
  struct caml_ba_array b = {(void *)p, {d0, d1, 0, 0}}; return b; }
static struct caml_ba_array ba3(const void *p, intnat d0, intnat d1, intnat d2) {
  struct caml_ba_array b = {(void *)p, {d0, d1, d2, 0}}; return b; }
#define V(x) ((value)&(x))

/* linear_c.c:29 — dst[n_sites][S] = b[n_sites][S] * a^T, a is S x S */
void ref_gemmish(double *dst, const double *a, const double *b, int n_states, int n_sites) {
  struct caml_ba_array D = ba2(dst, n_sites, n_states), A = ba2(a, n_states, n_states),
                       B = ba2(b, n_sites, n_states);
  gemmish_c(V(D), V(A), V(B));
}
/* linear_c.c:92 — dst_ij = sum_k lambda_k u_ik uit_jk */
void ref_dediagonalize(double *dst, const double *u, const double *lambda, const double *uit, int n) {
  struct caml_ba_array D = ba2(dst, n, n), U = ba2(u, n, n), L = ba1(lambda, n), T = ba2(uit, n, n);
  dediagonalize(V(D), V(U), V(L), V(T));
}
/* linear_c.c:170 */
double ref_mat_log_like3(const double *statd, const double *x, const double *y, const double *z,
                         int n_sites, int n_states) {
  struct caml_ba_array S = ba1(statd, n_states), X = ba2(x, n_sites, n_states),
                       Y = ba2(y, n_sites, n_states), Z = ba2(z, n_sites, n_states);
  return Double_val(mat_log_like3_c(V(S), V(X), V(Y), V(Z)));
}
/* linear_c.c:218 */
void ref_mat_pairwise_prod(double *dst, const double *x, const double *y, int n_sites, int n_states) {
  struct caml_ba_array D = ba2(dst, n_sites, n_states), X = ba2(x, n_sites, n_states),
                       Y = ba2(y, n_sites, n_states);
  mat_pairwise_prod_c(V(D), V(X), V(Y));
}
/* linear_c.c:232 */
void ref_mat_statd_pairwise_prod(const double *statd, double *dst, const double *a, const double *b,
                                 int n_sites, int n_states) {
  struct caml_ba_array S = ba1(statd, n_states), D = ba2(dst, n_sites, n_states),
                       A = ba2(a, n_sites, n_states), B = ba2(b, n_sites, n_states);
  mat_statd_pairwise_prod_c(V(S), V(D), V(A), V(B));
}
/* linear_c.c:251 */
double ref_mat_masked_logdot(const double *x, const double *y, const uint16_t *mask, int n_sites,
                             int n_states) {
  struct caml_ba_array X = ba2(x, n_sites, n_states), Y = ba2(y, n_sites, n_states),
                       M = ba1(mask, n_sites);
  return Double_val(mat_masked_logdot_c(V(X), V(Y), V(M)));
}
/* linear_c.c:298 */
double ref_mat_bounded_logdot(const double *x, const double *y, int first, int last, int n_sites,
                              int n_states) {
  struct caml_ba_array X = ba2(x, n_sites, n_states), Y = ba2(y, n_sites, n_states);
  return Double_val(mat_bounded_logdot_c(V(X), V(Y), Val_int(first), Val_int(last)));
}
/* linear_c.c:367 */
double ref_ten_log_like3(const double *statd, const double *x, const double *y, const double *z,
                         double *util, int n_rates, int n_sites, int n_states) {
  struct caml_ba_array S = ba1(statd, n_states), X = ba3(x, n_rates, n_sites, n_states),
                       Y = ba3(y, n_rates, n_sites, n_states), Z = ba3(z, n_rates, n_sites, n_states),
                       U = ba1(util, n_sites);
  return Double_val(ten_log_like3_c(V(S), V(X), V(Y), V(Z), V(U)));
}
/* linear_c.c:426 */
void ref_ten_pairwise_prod(double *dst, const double *x, const double *y, int n_rates, int n_sites,
                           int n_states) {
  struct caml_ba_array D = ba3(dst, n_rates, n_sites, n_states), X = ba3(x, n_rates, n_sites, n_states),
                       Y = ba3(y, n_rates, n_sites, n_states);
  ten_pairwise_prod_c(V(D), V(X), V(Y));
}
/* linear_c.c:443 */
void ref_ten_statd_pairwise_prod(const double *statd, double *dst, const double *a, const double *b,
                                 int n_rates, int n_sites, int n_states) {
  struct caml_ba_array S = ba1(statd, n_states), D = ba3(dst, n_rates, n_sites, n_states),
                       A = ba3(a, n_rates, n_sites, n_states), B = ba3(b, n_rates, n_sites, n_states);
  ten_statd_pairwise_prod_c(V(S), V(D), V(A), V(B));
}
/* linear_c.c:465 */
double ref_ten_masked_logdot(const double *x, const double *y, const uint16_t *mask, double *util,
                             int n_rates, int n_sites, int n_states) {
  struct caml_ba_array X = ba3(x, n_rates, n_sites, n_states), Y = ba3(y, n_rates, n_sites, n_states),
                       M = ba1(mask, n_sites), U = ba1(util, n_sites);
  return Double_val(ten_masked_logdot_c(V(X), V(Y), V(M), V(U)));
}
/* linear_c.c:541 */
double ref_ten_bounded_logdot(const double *x, const double *y, int first, int last, double *util,
                              int n_rates, int n_sites, int n_states) {
  struct caml_ba_array X = ba3(x, n_rates, n_sites, n_states), Y = ba3(y, n_rates, n_sites, n_states),
                       U = ba1(util, n_sites);
  return Double_val(ten_bounded_logdot_c(V(X), V(Y), Val_int(first), Val_int(last), V(U)));
}
/* linear_c.c:607 */
void ref_vec_pairwise_prod(double *dst, const double *x, const double *y, int n) {
  struct caml_ba_array D = ba1(dst, n), X = ba1(x, n), Y = ba1(y, n);
  vec_pairwise_prod_c(V(D), V(X), V(Y));
}
/* linear_c.c:620 */
int ref_int_vec_tot(const uint16_t *x, int n) {
  struct caml_ba_array X = ba1(x, n);
  return Int_val(int_vec_tot_c(V(X)));
}
/* linear_c.c:632 */
void ref_int_vec_pairwise_prod(uint16_t *dst, const uint16_t *x, const uint16_t *y, int n) {
  struct caml_ba_array D = ba1(dst, n), X = ba1(x, n), Y = ba1(y, n);
  int_vec_pairwise_prod_c(V(D), V(X), V(Y));
}
/* linear_c.c:644 — returns 0, or 1 if the reference raised Failure (dimension mismatch) */
int ref_float_mat_int_vec_mul(double *dst, int dst_n, const double *mat, int n, int k,
                              const uint16_t *vec, int vec_n) {
  struct caml_ba_array D = ba1(dst, dst_n), M = ba2(mat, n, k), W = ba1(vec, vec_n);
  if (setjmp(ppb_shim_fail_jmp)) return 1;
  float_mat_int_vec_mul_c(V(D), V(M), V(W));
  return 0;
}
