/* OCaml C stubs binding libpplacer_b200.so for the reference's host code (the file a maintainer adds as
 * pplacer_src/pplacer_b200_stubs.c; same conventions as pplacer_src/linear_c.c: arguments are OCaml values wrapping
 * Bigarrays, every buffer is owned by OCaml, errors surface as Failure like float_mat_int_vec_mul_c, linear_c.c:654-657).
 *
 *   Gpu_place.create      <-  Model.build + Like_stree inputs            pplacer_run.ml:200-205,289-309
 *   Gpu_place.precompute  <-  calc_distal_and_proximal, pulls, snodes    pplacer_run.ml:322-340
 *   Gpu_place.place       <-  Core.pplacer_core mapped through Multiprocessing (core.ml:76; pplacer_run.ml:618-649):
 *                             ONE blocking call drives every GPU named at creation (pp_multi_place)
 *   Gpu_place.export_glvs <-  darr / parr / snodes for --check-like and --map-mrca (pplacer_run.ml:344-377,446-464)
 *
 * No OCaml toolchain exists in the build image: tests/test_ocaml_stub.py compiles this file against the mock
 * caml/ headers of oracle/caml_shim and drives it end to end on the GPU through that mock runtime. */
#include <stdlib.h>
#include <string.h>
#include <caml/mlvalues.h>
#include <caml/memory.h>
#include <caml/alloc.h>
#include <caml/custom.h>
#include <caml/fail.h>
#include <caml/bigarray.h>
#include "pplacer_b200.h"

#define Multi_val(v) (*((pp_multi **) Data_custom_val(v)))
static void ctx_finalize(value v) { pp_destroy_multi(Multi_val(v)); Multi_val(v) = NULL; }
static struct custom_operations ctx_ops = { "org.pplacer.b200.ctx", ctx_finalize,
  custom_compare_default, custom_hash_default, custom_serialize_default, custom_deserialize_default,
  custom_compare_ext_default, custom_fixed_length_default };
/* 'a option of a Bigarray: None -> NULL */
static void *opt_data(value o) { return Is_block(o) ? Caml_ba_data_val(Field(o, 0)) : NULL; }

CAMLprim value ppb_create(value n_states, value model_class, value exch, value statd, value statd_chars,
                          value rates, value alpha, value n_rates, value site_cat, value root,
                          value child_ptr, value child_idx, value bl, value leaf_row, value ref_chars, value devices)
{
  CAMLparam5(exch, statd, statd_chars, rates, site_cat);
  CAMLxparam5(child_ptr, child_idx, bl, leaf_row, ref_chars);
  CAMLxparam1(devices);
  CAMLlocal1(res);
  pp_model m;
  pp_tree t;
  pp_multi *h = NULL;
  memset(&m, 0, sizeof m);
  m.n_states = Int_val(n_states); m.model_class = Int_val(model_class); m.n_rates = Int_val(n_rates);
  m.exch = Caml_ba_data_val(exch); m.statd = opt_data(statd); m.statd_chars = opt_data(statd_chars);
  m.n_statd_chars = Is_block(statd_chars) ? (int64_t) caml_ba_num_elts(Caml_ba_array_val(Field(statd_chars, 0))) : 0;
  m.rates = opt_data(rates); m.gamma_alpha = Double_val(alpha); m.site_cat = opt_data(site_cat);
  t.n_nodes = (int32_t) Caml_ba_array_val(bl)->dim[0]; t.root_id = Int_val(root);
  t.child_ptr = Caml_ba_data_val(child_ptr); t.child_idx = Caml_ba_data_val(child_idx);
  t.bl = Caml_ba_data_val(bl); t.leaf_row = Caml_ba_data_val(leaf_row);
  if (pp_create_multi(&m, &t, Caml_ba_data_val(ref_chars), (int32_t) Caml_ba_array_val(ref_chars)->dim[0],
                      (int32_t) Caml_ba_array_val(ref_chars)->dim[1], (int32_t) Caml_ba_array_val(devices)->dim[0],
                      Caml_ba_data_val(devices), &h) != PP_OK)
    caml_failwith(pp_last_error());
  res = caml_alloc_custom(&ctx_ops, sizeof(pp_multi *), 0, 1);
  Multi_val(res) = h;
  CAMLreturn(res);
}
CAMLprim value ppb_create_bytecode(value *a, int n) {
  (void) n;
  return ppb_create(a[0],a[1],a[2],a[3],a[4],a[5],a[6],a[7],a[8],a[9],a[10],a[11],a[12],a[13],a[14],a[15]); }

CAMLprim value ppb_precompute(value ctx)
{
  CAMLparam1(ctx);
  if (pp_multi_precompute(Multi_val(ctx)) != PP_OK) caml_failwith(pp_last_error());
  CAMLreturn(Val_unit);
}

/* prefs: float64 vector in pp_prefs field order (filled from Prefs.* by gpu_place.ml); prior_means: float64 vector option.
 * edges: int32 [Q x keep_at_most]; out: float64 [Q x keep_at_most x 6] = log_like, pendant, distal, ml_ratio,
 * marginal_like, post_prob (placement.ml:50-65); n_out, status: int32 [Q]. */
CAMLprim value ppb_place(value ctx, value prefs, value prior_means, value queries, value edges, value out, value n_out,
                         value status)
{
  CAMLparam5(ctx, prefs, prior_means, queries, edges);
  CAMLxparam3(out, n_out, status);
  const double *p = Caml_ba_data_val(prefs);
  pp_prefs pf;
  int64_t Q = Caml_ba_array_val(queries)->dim[0], i, n;
  int32_t *ed = Caml_ba_data_val(edges);
  double *o = Caml_ba_data_val(out);
  pp_result *tmp;
  int rc;
  pp_default_prefs(&pf);
  pf.start_pend = p[0]; pf.max_pend = p[1]; pf.initial_tolerance = p[2]; pf.final_tolerance = p[3];
  pf.max_iter = (int32_t) p[4]; pf.max_strikes = (int32_t) p[5]; pf.strike_box = p[6]; pf.max_pitches = (int32_t) p[7];
  pf.keep_at_most = (int32_t) p[8]; pf.keep_factor = p[9]; pf.calc_pp = (int32_t) p[10]; pf.prior_kind = (int32_t) p[11];
  pf.pp_rel_err = p[12]; pf.prior_mean = p[13]; pf.prior_means = opt_data(prior_means);
  n = Q * pf.keep_at_most;
  tmp = (pp_result *) malloc((size_t) (n > 0 ? n : 1) * sizeof(pp_result));
  if (!tmp) caml_failwith("ppb_place: out of memory");
  rc = pp_multi_place(Multi_val(ctx), &pf, Q, Caml_ba_data_val(queries), tmp, Caml_ba_data_val(n_out), Caml_ba_data_val(status));
  if (rc == PP_OK)
    for (i = 0; i < n; ++i) {
      ed[i] = tmp[i].edge;
      o[6 * i] = tmp[i].log_like; o[6 * i + 1] = tmp[i].pendant; o[6 * i + 2] = tmp[i].distal;
      o[6 * i + 3] = tmp[i].ml_ratio; o[6 * i + 4] = tmp[i].marginal_like; o[6 * i + 5] = tmp[i].post_prob;
    }
  free(tmp);
  if (rc != PP_OK) caml_failwith(pp_last_error());
  CAMLreturn(Val_unit);
}
CAMLprim value ppb_place_bytecode(value *a, int n) { (void) n; return ppb_place(a[0],a[1],a[2],a[3],a[4],a[5],a[6],a[7]); }

/* GLV of one node in the reference layout (Gmix a[R][L][S], Gcat a[L][S]; exponents as int64): glv.ml's {e; a} */
CAMLprim value ppb_export_glv(value ctx, value which, value node, value a, value e)
{
  CAMLparam5(ctx, which, node, a, e);
  if (pp_export_glvs(pp_multi_ctx(Multi_val(ctx), 0), Int_val(which), Int_val(node), Int_val(node) + 1,
                     Caml_ba_data_val(a), Caml_ba_data_val(e)) != PP_OK)
    caml_failwith(pp_last_error());
  CAMLreturn(Val_unit);
}
