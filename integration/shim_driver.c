/* TEST INFRASTRUCTURE (used by tests/test_ocaml_stub.py only): plain-C entry points that play the OCaml runtime for
 * integration/pplacer_b200_stubs.c under the mock caml/ headers of oracle/caml_shim - build the "values" (bigarray
 * descriptors, options, boxed floats), catch Failure (longjmp), run the custom block's finaliser. */
#include <setjmp.h>
#include <stddef.h>
#include <stdint.h>
#include <caml/mlvalues.h>
#include <caml/bigarray.h>
#include <caml/alloc.h>
#include <caml/custom.h>
#include <caml/fail.h>

__thread jmp_buf ppb_shim_fail_jmp;
__thread const char *ppb_shim_fail_msg;

value ppb_create(value, value, value, value, value, value, value, value, value, value, value, value, value, value, value, value);
value ppb_precompute(value);
value ppb_place(value, value, value, value, value, value, value, value);
value ppb_export_glv(value, value, value, value, value);

static struct caml_ba_array ba(const void *p, intnat d0, intnat d1) { struct caml_ba_array b = {(void *)p, {d0, d1, 0, 0}}; return b; }
#define V(x) ((value)&(x))
static value some(value *cell, value v) { cell[0] = v; return (value)cell; }
#define NONE Val_int(0)

const char *drv_last_failure(void) { return ppb_shim_fail_msg ? ppb_shim_fail_msg : ""; }

/* returns the ctx value (0 on Failure) */
intptr_t drv_create(int n_states, int model_class, const double *exch, const double *statd, const uint8_t *statd_chars,
                    int64_t n_statd_chars, const double *rates, double alpha, int n_rates, const int32_t *site_cat, int L,
                    int root, int n_nodes, const int32_t *child_ptr, const int32_t *child_idx, int n_child,
                    const double *bl, const int32_t *leaf_row, const uint8_t *ref_chars, int n_leaves,
                    const int32_t *devices, int n_devices) {
  struct caml_ba_array b_exch = ba(exch, n_states, n_states), b_statd = ba(statd, n_states, 0),
                       b_sc = ba(statd_chars, n_statd_chars, 0), b_rates = ba(rates, n_rates, 0), b_cat = ba(site_cat, L, 0),
                       b_ptr = ba(child_ptr, n_nodes + 1, 0), b_idx = ba(child_idx, n_child, 0), b_bl = ba(bl, n_nodes, 0),
                       b_leaf = ba(leaf_row, n_nodes, 0), b_ref = ba(ref_chars, n_leaves, L), b_dev = ba(devices, n_devices, 0);
  value c1[1], c2[1], c3[1], c4[1];
  value a = caml_copy_double(alpha);
  ppb_shim_fail_msg = NULL;
  if (setjmp(ppb_shim_fail_jmp)) return 0;
  return ppb_create(Val_int(n_states), Val_int(model_class), V(b_exch), statd ? some(c1, V(b_statd)) : NONE,
                    statd_chars ? some(c2, V(b_sc)) : NONE, rates ? some(c3, V(b_rates)) : NONE, a, Val_int(n_rates),
                    site_cat ? some(c4, V(b_cat)) : NONE, Val_int(root), V(b_ptr), V(b_idx), V(b_bl), V(b_leaf), V(b_ref), V(b_dev));
}
int drv_precompute(intptr_t ctx) {
  ppb_shim_fail_msg = NULL;
  if (setjmp(ppb_shim_fail_jmp)) return 1;
  ppb_precompute((value)ctx);
  return 0;
}
int drv_place(intptr_t ctx, const double *prefs14, const double *prior_means, int n_nodes, const uint8_t *queries, int64_t Q,
              int L, int keep_at_most, int32_t *edges, double *out, int32_t *n_out, int32_t *status) {
  struct caml_ba_array b_p = ba(prefs14, 14, 0), b_pm = ba(prior_means, n_nodes, 0), b_q = ba(queries, Q, L),
                       b_e = ba(edges, Q * keep_at_most, 0), b_o = ba(out, Q * keep_at_most * 6, 0), b_n = ba(n_out, Q, 0),
                       b_s = ba(status, Q, 0);
  value c1[1];
  ppb_shim_fail_msg = NULL;
  if (setjmp(ppb_shim_fail_jmp)) return 1;
  ppb_place((value)ctx, V(b_p), prior_means ? some(c1, V(b_pm)) : NONE, V(b_q), V(b_e), V(b_o), V(b_n), V(b_s));
  return 0;
}
int drv_export_glv(intptr_t ctx, int which, int node, double *a, int64_t n_a, int64_t *e, int L) {
  struct caml_ba_array b_a = ba(a, n_a, 0), b_e = ba(e, L, 0);
  ppb_shim_fail_msg = NULL;
  if (setjmp(ppb_shim_fail_jmp)) return 1;
  ppb_export_glv((value)ctx, Val_int(which), Val_int(node), V(b_a), V(b_e));
  return 0;
}
void drv_release(intptr_t ctx) { if (ctx) ppb_shim_finalize_custom((value)ctx); }
