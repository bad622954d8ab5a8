(* OCaml side of the binding (the file a maintainer adds as pplacer_src/gpu_place.ml).  No OCaml toolchain exists in
   the build image, so this file is documentation that is kept in step with pplacer_b200_stubs.c by hand; the C side
   is compiled and driven end to end by tests/test_ocaml_stub.py through a mock runtime. *)
type ctx  (* custom block holding pp_multi*, finaliser = pp_destroy_multi *)
type u8_mat  = (int, Bigarray.int8_unsigned_elt, Bigarray.c_layout) Bigarray.Array2.t
type f64_vec = (float, Bigarray.float64_elt, Bigarray.c_layout) Bigarray.Array1.t
type i32_vec = (int32, Bigarray.int32_elt, Bigarray.c_layout) Bigarray.Array1.t
type i64_vec = (int64, Bigarray.int64_elt, Bigarray.c_layout) Bigarray.Array1.t

external create :
  n_states:int -> model_class:int -> exch:f64_vec -> statd:f64_vec option -> statd_chars:u8_mat option ->
  rates:f64_vec option -> gamma_alpha:float -> n_rates:int -> site_cat:i32_vec option ->
  root:int -> child_ptr:i32_vec -> child_idx:i32_vec -> bl:f64_vec -> leaf_row:i32_vec ->
  ref_chars:u8_mat -> devices:i32_vec -> ctx = "ppb_create_bytecode" "ppb_create"
external precompute : ctx -> unit = "ppb_precompute"
(* edges: Q x keep_at_most int32; out: Q x keep_at_most x 6 floats (log_like, pendant, distal, ml_ratio, marginal, post_prob) *)
external place :
  ctx -> prefs:f64_vec -> prior_means:f64_vec option -> queries:u8_mat -> edges:i32_vec -> out:f64_vec ->
  n_out:i32_vec -> status:i32_vec -> unit = "ppb_place_bytecode" "ppb_place"
external export_glv : ctx -> which:int -> node:int -> f64_vec -> i64_vec -> unit = "ppb_export_glv"

(* Pplacer_run.run_placements, in place of Multiprocessing.try_fork (pplacer_run.ml:618-649):
     let gpu = Gpu_place.create ... ~devices:(all_gpus ()) in
     Gpu_place.precompute gpu;
     Gpu_place.place gpu ~prefs:(prefs_vector prefs) ~prior_means ~queries ~edges ~out ~n_out ~status;
     results_of_arrays edges out n_out status |> List.iter donefunc                      (* core.ml:58-62 *) *)
