/* TEST INFRASTRUCTURE ONLY.  CPU (FP64, plain C) restatement of pplacer's per-query placement hot
 * path, used exclusively as the parity checker by tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs.  The product (pplacer_b200/csrc) never links, imports or
 * calls anything in oracle/.
 *
 * Pinning status (see DESIGN.md "Oracle"):
 *   pinned   - model build, discrete Gamma, diagonalisation, evolve, pruning, exponent pulls,
 *              supernodes, logdot/log_like3 : reference tests/pplacer/test_like.ml (PhyML LL at
 *              every node, rel 1e-7), test_linear.ml toys, gamma.ml:47-50 vector, and bit-for-bit
 *              against the reference's own linear_c.c compiled into oracle/_ref.
 *   pinned   - Gcat model + Brent + baseball + refinement: the three ar53 GOLDEN placements
 *              (donovan-test/ar53_pplacer_testset/ar53_test_genomes.GOLDEN.json).
 *   UNPINNED - --pp (calc_marg_prob / QNG) and multi-placement keep sets: no reference test or
 *              golden touches them ("parity unpinned"); GSL's Brent/QNG are restated from the
 *              published algorithm (GSL 2.7 min/brent.c, integration/qng.c).
 */
#ifndef PPLACER_ORACLE_H
#define PPLACER_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_MAX_STATES 20

/* status codes: 0 ok; GSL-like errors are recoverable per location (core.ml:272-297), the
 * "fatal" ones abort the run in the reference (uncaught OCaml exceptions). */
enum {
  ORC_OK = 0,
  ORC_GSL_EINVAL = 4,     /* gsl_errno.h GSL_EINVAL: endpoints do not enclose a minimum */
  ORC_GSL_EBADFUNC = 9,   /* GSL_EBADFUNC: function value is infinite or NaN */
  ORC_GSL_ETOL = 14,      /* GSL_ETOL: failed to reach tolerance */
  ORC_GSL_EBADTOL = 13,
  ORC_EXCEEDED_MAX_ITER = 100, /* Minimization.ExceededMaxIter */
  ORC_FIND_START_FAILURE = 101,/* Minimization.FindStartFailure  (fatal) */
  ORC_INVALID_START = 102,     /* Minimization.InvalidStartValues (fatal) */
  ORC_NO_INFORMATIVE = 103,    /* core.ml:157 */
  ORC_EMPTY_RESULTS = 104,     /* core.ml:323 */
  ORC_BAD_CHAR = 105,
  ORC_ASSERT = 106
};

typedef struct orc_prefs {
  double start_pend;        /* 0.1    prefs.ml:76 */
  double max_pend;          /* 2.0 */
  double initial_tolerance; /* 0.01 */
  double final_tolerance;   /* 1e-5   core.ml:14 */
  int max_iter;             /* 200    core.ml:10 */
  double strike_box;        /* 3.0 */
  int max_strikes;          /* 6 */
  int max_pitches;          /* 40 */
  int keep_at_most;         /* 7 */
  double keep_factor;       /* 0.01 */
  int calc_pp;              /* 0 */
  double pp_rel_err;        /* 0.01 */
  int prior_kind;           /* 0 uniform, 1 flat exponential (mean = prior_mean), 2 informative */
  double prior_mean;
  const double *prior_means;/* [n_nodes] when prior_kind == 2 */
} orc_prefs;

typedef struct orc_result {
  int32_t edge;
  int32_t _pad;
  double log_like, pendant, distal, ml_ratio, marginal_like, post_prob;
} orc_result;

typedef struct orc_ctx orc_ctx;

/* ---- numerical building blocks (each individually testable) ---- */
void orc_default_prefs(orc_prefs *p);
/* gamma.ml:52-70 */
void orc_discrete_gamma(int n_cat, double alpha, double *rates_out);
double orc_gamma_inc_P(double a, double x);            /* gsl_sf_gamma_inc_P */
double orc_gamma_Pinv(double p, double a);             /* gsl_cdf_gamma_Pinv (scale 1) */
/* diagd.ml:123-172: exch is S x S symmetric (diagonal ignored), pi is S.  Outputs x, lambda, xit. */
void orc_diagd_normed_of_exchangeable_pair(int S, const double *exch, const double *pi, double *x,
                                           double *lambda, double *xit);
/* symmetric eigendecomposition (cyclic Jacobi); columns of evec are eigenvectors */
void orc_symm_eigs(int n, const double *m, double *eval, double *evec);
/* diagd.ml:79-98 */
void orc_multi_exp(int S, int R, const double *x, const double *lambda, const double *xit,
                   const double *rates, const uint8_t *mask_or_null, double bl, double *tensor);
/* alignment.ml:186-217 */
int orc_emper_freq(int S, const uint8_t *chars, int64_t n_chars, double *out);
/* nuc_models.ml:24-54 / prot_models.ml:5-51 : likelihood vector for a character; returns 0 if unknown */
int orc_lv_of_char(int S, uint8_t c, double *lv);

/* GSL restatements */
typedef double (*orc_fn)(double x, void *ud);
int orc_brent(orc_fn f, void *ud, double raw_start, double left, double right, double tol,
              int max_iters, double *xmin_out, int *n_calls);
int orc_qng(orc_fn f, void *ud, double a, double b, double epsabs, double epsrel, double *result,
            double *abserr, int *neval);

/* ---- the path ---- */
/* Tree: node ids as numbered by the reference parser (post-order); child lists in CSR form, in
 * the order of the Newick string.  leaf_row[id] = row of ref_chars for leaves, -1 otherwise.
 * Model: gcat == 0 -> Gmix (R Gamma categories, rates[R]); gcat == 1 -> Gcat (R CAT categories,
 * site_cat[L] zero-based).  ref_chars is n_leaves x L, uppercase, already pre-masked. */
orc_ctx *orc_create(int S, int R, int gcat, const double *statd, const double *x,
                    const double *lambda, const double *xit, const double *rates,
                    const int32_t *site_cat, int n_nodes, int root_id, const int32_t *child_ptr,
                    const int32_t *child_idx, const double *bl, const int32_t *leaf_row,
                    const uint8_t *ref_chars, int n_leaves, int L, const char *mmap_dir_or_null);
void orc_destroy(orc_ctx *);
/* pplacer_run.ml:322-340 (calc_distal_and_proximal, pull at -10, prep_supernodes) */
int orc_precompute(orc_ctx *);
/* pointers into the reference-layout arrays: Gmix a[R][L][S], Gcat a[L][S]; e[L] int64 */
const double *orc_glv_a(orc_ctx *, int which /*0 darr,1 parr,2 snodes*/, int node);
const int64_t *orc_glv_e(orc_ctx *, int which, int node);
int64_t orc_glv_a_len(orc_ctx *);
/* test_like.ml:57-66 quantities at a node: slow_log_like3(evolve(d,bl/2),evolve(p,bl/2),1) and logdot(snode,1) */
void orc_node_tree_likes(orc_ctx *, int node, double *slow_ll3, double *sn_logdot);

/* Place Q queries (each L uppercase chars).  out is Q x keep_at_most, n_out[Q], status[Q].
 * Optional debug outputs may be NULL: scores (Q x n_nodes, NaN at the root), n_pitches[Q],
 * pitch_edges / pitch_ll / pitch_pend / pitch_dist (Q x (max_pitches+1)) = the baseball results
 * before sorting, n_like_calls[Q].  n_threads > 1 mirrors the reference's -j (fork per job ->
 * one worker thread per core over a shared query counter). */
typedef struct orc_debug {
  double *scores;
  int32_t *n_pitches;
  int32_t *pitch_edges;
  double *pitch_ll, *pitch_pend, *pitch_dist;
  int64_t *n_like_calls;
  double *timing; /* [4] summed seconds: ranking, ML, PP, query prep */
} orc_debug;
int orc_place(orc_ctx *, const orc_prefs *, int64_t Q, const uint8_t *query_chars, orc_result *out,
              int32_t *n_out, int32_t *status, orc_debug *dbg_or_null, int n_threads);

/* single three-taxon objective for kernel tests: LL of query at (edge, pendant, distal) */
int orc_three_tax_ll(orc_ctx *, const uint8_t *query_chars, int edge, double pendant, double distal,
                     double *ll_out);

/* Gcat_model.refine (gcat_model.ml:387-419): replaces the context's site categories; call before orc_precompute */
int orc_refine_site_categories(orc_ctx *, int32_t *site_cat_out, double *site_ll_out);

/* figs for the ranking (fig.ml:120-148); m == 0: plain ranking */
int orc_set_figs(orc_ctx *, int m, const int32_t *rep, const int32_t *ptr, const int32_t *idx);

/* kernel-level hooks (inner loops on caller arrays, zero exponents) for the bit-for-bit comparison
 * with the reference's own linear_c.c in tests/test_oracle_vs_ref.py */
void orc_k_gemmish(double *dst, const double *a, const double *b, int S, int n_sites);
void orc_k_dediagonalize(double *dst, const double *u, const double *lambda, const double *uit, int n);
void orc_k_pairwise_prod(double *dst, const double *x, const double *y, int R, int gcat, int n, int S);
void orc_k_statd_pairwise_prod(const double *statd, double *dst, const double *a, const double *b, int R, int gcat,
                               int n, int S);
double orc_k_log_like3(const double *statd, const double *x, const double *y, const double *z, int R, int gcat, int n,
                       int S);
double orc_k_masked_logdot(const double *x, const double *y, const uint16_t *mask, int R, int gcat, int n, int S);
double orc_k_logdot(const double *x, const double *y, int R, int gcat, int n, int S);

#ifdef __cplusplus
}
#endif
#endif
