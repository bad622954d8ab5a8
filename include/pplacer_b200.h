/* pplacer_b200 — C ABI of libpplacer_b200.so: the B200-native (sm_100a) implementation of
 * pplacer's per-query placement hot path.
 *
 * This header is the drop-in boundary.  Plain C, plain pointers and sizes, caller-owned HOST
 * buffers unless a function says "device".  It replaces, for this path, what the reference binds
 * through OCaml `external`s to pplacer_src/linear_c.c (declared in pplacer_src/linear.ml:13-89)
 * plus the OCaml loops that call them; INTEGRATION.md shows the OCaml stub a maintainer adds.
 *
 *   reference interface replaced                          entry point here
 *   ---------------------------------------------------   ---------------------------------
 *   Model.build / Diagd / Gamma          (gmix_model.ml:25-39, gcat_model.ml:52-69,
 *                                         diagd.ml:169, gamma.ml:52)        pp_create
 *   Like_stree.calc_distal_and_proximal + pull(-10) + Glv_arr.prep_supernodes
 *                                        (pplacer_run.ml:322-340)           pp_precompute
 *   darr / parr / snodes Bigarrays       (pplacer_run.ml:289-309)           pp_export_glvs
 *   Core.pplacer_core (name, seq) -> result list   (core.ml:76,113-409)     pp_place
 *   Glv.masked_logdot over all edges     (core.ml:187-194; linear_c.c:251,465)   pp_prefilter_scores
 *   Three_tax.log_like                   (three_tax.ml:30-36; linear_c.c:170,367) pp_three_tax_log_like
 *   Model.evolve_into                    (gmix_model.ml:308, gcat_model.ml:327;
 *                                         linear_c.c:29,92)                 pp_evolve
 *   Gamma.discrete_gamma                 (gamma.ml:52)                      pp_discrete_gamma
 *   Diagd.normed_of_exchangeable_pair    (diagd.ml:169)                     pp_diagd
 *
 * Every function returns 0 on success or a PP_E* code; pp_last_error() gives the message for the
 * calling thread.  There is NO CPU fallback: without a CUDA device every compute entry point
 * fails with PP_ENODEVICE.
 */
#ifndef PPLACER_B200_H
#define PPLACER_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PP_ABI_VERSION 1

/* error codes */
enum {
  PP_OK = 0,
  PP_EINVAL = 1,     /* bad argument */
  PP_ENODEVICE = 2,  /* no CUDA device / driver */
  PP_ECUDA = 3,      /* CUDA runtime error (message in pp_last_error) */
  PP_ENOMEM = 4,
  PP_ESTATE = 5      /* call order (e.g. pp_place before pp_precompute) */
};

/* per-query status written by pp_place (0 = placed).  The GSL-like codes are handled inside the
 * path exactly as core.ml:272-297 does (retry from defaults, then skip the location); a query
 * status is non-zero only where the reference would abort the run or produce no placement. */
enum {
  PP_Q_OK = 0,
  PP_Q_NO_INFORMATIVE = 103,   /* core.ml:157: "has no informative sites" */
  PP_Q_EMPTY_RESULTS = 104,    /* core.ml:323 */
  PP_Q_BAD_CHAR = 105,         /* character outside nuc_models.ml:28-45 / prot_models.ml:5,40-44 */
  PP_Q_FIND_START_FAILURE = 101, /* Minimization.FindStartFailure  (uncaught in the reference) */
  PP_Q_INVALID_START = 102,      /* Minimization.InvalidStartValues (uncaught in the reference) */
  PP_Q_PP_FAILED = 107           /* find_upper_limit raised ExceededMaxIter (three_tax.ml:154) */
};

/* Substitution model inputs (what Model.build consumes, glvm.ml:5-9). */
typedef struct pp_model {
  int32_t n_states;     /* 4 (nucleotide) or 20 (amino acid) */
  int32_t model_class;  /* 0 = Gmix_model (Gamma mixture), 1 = Gcat_model (per-site CAT) */
  int32_t n_rates;      /* Gamma categories, or number of CAT rate categories */
  int32_t _pad;
  const double *exch;   /* [S*S] symmetric exchangeabilities, diagonal ignored (GTR: nuc_models.ml:78; LG/WAG/JTT) */
  const double *statd;  /* [S] stationary distribution, or NULL = empirical from statd_chars (alignment.ml:186) */
  const uint8_t *statd_chars; /* UN-premasked reference alignment characters, used when statd == NULL */
  int64_t n_statd_chars;
  const double *rates;  /* [n_rates] explicit rates (CAT), or NULL = discrete_gamma(n_rates, gamma_alpha) */
  double gamma_alpha;
  const int32_t *site_cat; /* Gcat: [L] zero-based category per (pre-masked) site; NULL for Gmix */
} pp_model;

/* Reference tree: node ids exactly as numbered by the reference's Newick parser (post-order,
 * newick_parser.mly:49-62) = jplace edge numbers.  Children in CSR form, in Newick order. */
typedef struct pp_tree {
  int32_t n_nodes;          /* E + 1 */
  int32_t root_id;
  const int32_t *child_ptr; /* [n_nodes + 1] */
  const int32_t *child_idx; /* [child_ptr[n_nodes]] */
  const double *bl;         /* [n_nodes]; root 0 (newick_gtree.ml:145-157) */
  const int32_t *leaf_row;  /* [n_nodes] row of ref_chars for leaves, -1 otherwise */
} pp_tree;

/* Prefs fields on this path (prefs.ml:65-120; core.ml:10,14). */
typedef struct pp_prefs {
  double start_pend;        /* 0.1  */
  double max_pend;          /* 2.0  */
  double initial_tolerance; /* 0.01 */
  double final_tolerance;   /* 1e-5 */
  int32_t max_iter;         /* 200  */
  int32_t max_strikes;      /* 6  (0 = evaluate every edge) */
  double strike_box;        /* 3.0  */
  int32_t max_pitches;      /* 40   */
  int32_t keep_at_most;     /* 7    */
  double keep_factor;       /* 0.01 */
  int32_t calc_pp;          /* 0    (-p sets 1 and keep_at_most 20, keep_factor 0.001, max_strikes 20) */
  int32_t prior_kind;       /* 0 uniform, 1 flat exponential, 2 informative exponential */
  double pp_rel_err;        /* 0.01 */
  double prior_mean;        /* flat prior mean = tree length / n_edges (pplacer_run.ml:387-390) */
  const double *prior_means;/* [n_nodes] informative prior means (core.ml:34-52) or NULL */
} pp_prefs;

/* One placement (core.ml:370-404, placement.ml:50-65). */
typedef struct pp_result {
  int32_t edge;       /* location = jplace edge_num */
  int32_t _pad;
  double log_like;    /* "likelihood" */
  double pendant;     /* pendant_length */
  double distal;      /* distal_length */
  double ml_ratio;    /* like_weight_ratio */
  double marginal_like; /* NaN unless calc_pp */
  double post_prob;     /* NaN unless calc_pp */
} pp_result;

/* Optional stage outputs of pp_place for parity tests / --timing (any pointer may be NULL). */
typedef struct pp_debug {
  double *scores;        /* [Q * n_nodes] prefilter score per edge, NaN at the root */
  int32_t *n_pitches;    /* [Q] number of baseball results */
  int32_t *pitch_edges;  /* [Q * (max_pitches+1)] in play order */
  double *pitch_ll, *pitch_pend, *pitch_dist; /* [Q * (max_pitches+1)] */
  int64_t *n_like_calls; /* [Q] objective evaluations (three_tax.ml:8) */
  double *timing;        /* [4] seconds: ranking, ML calculation, PP calculation, query prep (core.ml timers) */
} pp_debug;

typedef struct pp_ctx pp_ctx;

const char *pp_last_error(void);
int pp_abi_version(void);
void pp_default_prefs(pp_prefs *);

/* Builds the model (a2, a3), uploads tree / alignment, allocates darr, parr, snodes on `device`. */
int pp_create(const pp_model *, const pp_tree *, const uint8_t *ref_chars /* n_leaves x L, uppercase, pre-masked */,
              int32_t n_leaves, int32_t L, int32_t device, pp_ctx **out);
void pp_destroy(pp_ctx *);
/* a4-a9 on device (level-order pruning, exponent pulls at -50 then -10, supernodes). */
int pp_precompute(pp_ctx *);
/* Gcat_model.refine (gcat_model.ml:387-419; pplacer_run.ml:316-317, --always-refine): re-derive the Price-CAT site
 * categories on the device - one tree-likelihood attempt per rate category with every site in that category
 * (Like_stree.site_log_like_arr, like_stree.ml:113-131), best category per site under FastTree's Gamma(3, 1/3) prior.
 * Replaces the context's site categories; call before pp_precompute.  site_cat_out [L] and site_ll_out [R x L]
 * (log-likelihood of every site under every category) may be NULL. */
int pp_refine_site_categories(pp_ctx *, int32_t *site_cat_out, double *site_ll_out);
/* --check-like (pplacer_run.ml:344-377): out [n_nodes x 2] = (tree_likelihood, supernode_like) seen from every edge. */
int pp_check_like(pp_ctx *, double *out);
/* Maximum a posteriori states on the proximal side of internal nodes, what --map-mrca / --map-identity read
 * (Map_seq.of_map, map_seq.ml:17-37 -> Seq_post.get_summary Proximal, seq_post.ml:14-31): out [n_ids x L] receives
 * per site the index of the most likely state of statd .* evolve(darr[id], bl[id]) .* parr[id] summed over the rates,
 * or -1 when its posterior share is below `cutoff` (Prefs.map_cutoff).  The MRCA bookkeeping around it
 * (Refpkg.get_mrcam, Map_seq.mrca_map_seq_map) needs the refpkg taxonomy and stays with the host. */
int pp_map_states(pp_ctx *, int32_t n_ids, const int32_t *ids, double cutoff, int8_t *out);
/* Copy GLVs [node_begin, node_end) of darr (which=0) / parr (1) / snodes (2) back in the REFERENCE
 * layout: Gmix a[R][L][S], Gcat a[L][S] per node (linear_c.c:5-12); e[L] as int64 per node. */
int pp_export_glvs(pp_ctx *, int32_t which, int32_t node_begin, int32_t node_end, double *a, int64_t *e);
/* Model numerics as built (statd[S], x[S*S], lambda[S], xit[S*S], rates[R]); any may be NULL. */
int pp_get_model(pp_ctx *, double *statd, double *x, double *lambda, double *xit, double *rates);

/* The path: place Q pre-masked, upper-cased queries (Q x L chars).  out is Q x keep_at_most. */
int pp_place(pp_ctx *, const pp_prefs *, int64_t Q, const uint8_t *query_chars, pp_result *out,
             int32_t *n_out, int32_t *status, pp_debug *dbg_or_null);

/* Figs (--fig-cutoff; fig.ml): with n_figs > 0 the ranking of pp_place follows Fig._enum_by_score (fig.ml:120-138) -
 * one representative edge per fig is scored for every query, the figs are taken in the order of those scores and each
 * GROUP (a fig and the following ones within strike_box of its score) is expanded, its edges scored and sorted, only
 * when baseball needs more locations.  The fig list is built by the host exactly as Fig._figs_of_gtree does
 * (fig.ml:70-99; pplacer_b200/fig.py): rep[n_figs], CSR edge lists ptr[n_figs+1] / idx.  n_figs == 0 restores the plain
 * ranking of every non-root edge (Fig.Dummy).  max_strikes == 0 (every edge is optimised) ignores figs. */
int pp_set_figs(pp_ctx *, int32_t n_figs, const int32_t *rep, const int32_t *ptr, const int32_t *idx);
int pp_n_figs(pp_ctx *);

/* ---- all the GPUs of one box behind one handle ------------------------------------------------------------
 * Replaces Multiprocessing's fork-per-job (pplacer_run.ml:618-649, multiprocessing.ml:279): the reference GLVs are
 * REPLICATED on n_devices GPUs (each runs the device precompute itself) and pp_multi_place shards the queries over
 * them - contiguous chunks handed out from a shared counter, one host thread per device, no inter-GPU traffic -
 * and returns when every record is in the caller's host buffers, in query order.  A placement does not depend on
 * the batch it is computed in, so the bytes equal pp_place on one device (tested).  device_ids == NULL means
 * devices 0 .. n_devices-1; an id may repeat (several contexts on one GPU: for tests on a single-GPU box).
 * A single-threaded host (the OCaml runtime) drives all GPUs with this one blocking call. */
typedef struct pp_multi pp_multi;
int pp_create_multi(const pp_model *, const pp_tree *, const uint8_t *ref_chars, int32_t n_leaves, int32_t L,
                    int32_t n_devices, const int32_t *device_ids, pp_multi **out);
void pp_destroy_multi(pp_multi *);
int pp_multi_precompute(pp_multi *);
int pp_multi_place(pp_multi *, const pp_prefs *, int64_t Q, const uint8_t *query_chars, pp_result *out,
                   int32_t *n_out, int32_t *status);
int pp_multi_n_devices(pp_multi *);
pp_ctx *pp_multi_ctx(pp_multi *, int32_t i); /* the i-th per-device context (stage-level entry points, tuning) */
/* queries placed by each device in the last pp_multi_place (n_devices entries) */
int pp_multi_last_counts(pp_multi *, int64_t *per_device);

/* Same work with inputs/outputs already resident in HBM (device pointers): used to time the
 * kernels without host<->device copies.  d_out: Q x keep_at_most pp_result, d_n_out/d_status: Q. */
int pp_place_device(pp_ctx *, const pp_prefs *, int64_t Q, const uint8_t *d_query_chars, pp_result *d_out,
                    int32_t *d_n_out, int32_t *d_status);

/* Stage-level entry points (host buffers) for parity tests against the oracle / linear_c.c. */
int pp_prefilter_scores(pp_ctx *, double start_pend, int64_t Q, const uint8_t *query_chars,
                        double *scores /* Q x n_nodes, NaN at root */, int32_t kernel /* 0 default, 1 direct-FMA, 2 table */);
/* The same prefilter pass timed on the device: after one untimed launch, `reps` launches of the chosen kernel
 * form over the same Q queries bracketed by CUDA events on the context's stream; *ms_per_launch is their mean.
 * (bench.py reports the direct-FMA form against the FP64 roofline next to the default table form.) */
int pp_prefilter_time(pp_ctx *, double start_pend, int64_t Q, const uint8_t *query_chars, int32_t kernel, int32_t reps,
                      double *ms_per_launch);
int pp_three_tax_log_like(pp_ctx *, int64_t n, const uint8_t *query_chars /* n x L */, const int32_t *edge,
                          const double *pendant, const double *distal, double *ll_out);
/* dst = evolve(src, bl) on GLVs given in the reference layout (n_sites sites; Gcat uses site_cat of
 * the first n_sites sites). */
int pp_evolve(pp_ctx *, const double *src_a, const int64_t *src_e, int32_t n_sites, double bl, double *dst_a,
              int64_t *dst_e);

/* Host numerics exposed for tests. */
int pp_discrete_gamma(int32_t n_cat, double alpha, double *rates_out);
int pp_diagd(int32_t S, const double *exch, const double *pi, double *x, double *lambda, double *xit);
int pp_emper_freq(int32_t S, const uint8_t *chars, int64_t n, double *out);

/* Run this context's work on a caller-owned CUDA stream (cudaStream_t passed as void*; NULL restores
 * the context's own stream).  Lets the host time the path with CUDA events on the launching stream. */
int pp_set_stream(pp_ctx *, void *cuda_stream);

/* Which prefilter kernel pp_place uses (all forms give ten_masked_logdot_c / mat_masked_logdot_c, linear_c.c:465,251):
 * 0 = default: DNA Gamma-4 model - the pair-of-sites query-lane log table (8) for batches of >= 256 queries, the
 *     four-base product table (2) below that; 20-state models and the other DNA models - the generic query-lane table (7)
 *     for large batches, the direct FMA form (1) otherwise;
 * 1 = direct FMA form (K = rates*states DFMAs per query x edge x site, as the reference writes it; the form the FP64
 *     roofline is defined on), 3 = the same with the evolved-query table in shared memory;
 * 2 = four-base product table; 6 = log table, lanes walk edges; 7 = log table, lanes are queries (TMA-streamed tiles);
 * 8 = the same with one table word per PAIR of adjacent sites (36 columns);
 * 4 = direct form staged through shared memory by TMA bulk copies + an mbarrier ring (DFMA), 5 = the same with the
 *     K contraction on the FP64 tensor cores (DMMA m8n8k4).  2, 4-6 and 8 exist for the DNA Gamma-4 model only.
 * pp_prefilter_kernel_used reports what the last prefilter launch ran (1, 2, 4, 5, 6, 7 or 8). */
int pp_set_prefilter_kernel(pp_ctx *, int32_t kernel);
int pp_prefilter_kernel_used(pp_ctx *);

/* Ranking (Fig.enum_by_score with fig_cutoff 0, fig.ml:143-148): 0 = on trees with more than 1024 nodes keep only
 * the first 256 entries of the stable descending order (all baseball reads unless a query runs through more
 * than 256 locations - then the batch is redone on the full order, counted by pp_rank_fallbacks), 1 = always
 * the full segmented sort.  Results are identical either way. */
int pp_set_rank_mode(pp_ctx *, int32_t mode);
int pp_rank_fallbacks(pp_ctx *);

/* Which engine runs the second stage (Three_tax.optimize / calc_marg_prob, three_tax.ml:70-83,169-215):
 * 0 = default (engine 4 for the DNA Gamma-4 model; engine 1 otherwise),
 * 1 = block-per-candidate (every model), 2 = warp-per-candidate, 3 = block-per-candidate with the cached
 * (statd*prox*dist) product in shared memory, 4 = block-per-candidate with the pendant objective in the
 * eigenbasis of the rate matrix (3R multiply-adds per site and call, everything in shared memory).
 * blocks_per_sm > 0 caps the resident thread blocks of engine 2 (8 candidates each); 0 = default.
 * pp_three_tax_engine_used: 1, 2, 3 or 4. */
int pp_set_three_tax_engine(pp_ctx *, int32_t engine, int32_t blocks_per_sm);
int pp_three_tax_engine_used(pp_ctx *);

/* Counters for bench.py: kernels launched by this library on this context since the last reset,
 * and accumulated device time (ms, CUDA events) of the named phase kernels. */
int64_t pp_kernel_launches(pp_ctx *, int32_t reset);
int pp_phase_times_ms(pp_ctx *, double *out4 /* prefilter, rank, optimise, pp */, int32_t reset);
/* Second-stage work done by engines 2 and 3 since the last reset (engine 1 does not count): objective calls (three_tax.ml:51-54)
 * with the pendant length moving, with the distal length moving, and each weighted by the query's number
 * of informative sites (the unit of SURVEY.md 8d's second-stage roofline).  --pp integrands not counted. */
int pp_three_tax_work(pp_ctx *, double *out4 /* pendant calls, distal calls, pendant call-sites, distal call-sites */, int32_t reset);
/* Per-phase cycle counters of the block engines (only filled when the library is built with -DPPB_TT_PROFILE, see
 * tools/tt_profile.py): out32[4] scalar cycles, [5] matrix-build cycles, [6] passes with a build, [7] build jobs,
 * [8+cmd] pass cycles by command, [16+cmd] passes by command, [24] candidates, [25] cycles per candidate summed. */
int pp_three_tax_profile(pp_ctx *, double *out32, int32_t reset);
/* FP64 FMA micro-benchmark on the context's device: returns measured DFMA throughput in TFLOP/s. */
int pp_measure_fp64_peak(pp_ctx *, double *tflops_out);

#ifdef __cplusplus
}
#endif
#endif
