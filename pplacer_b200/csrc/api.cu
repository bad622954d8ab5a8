// libpplacer_b200.so — context, host orchestration and the C ABI (include/pplacer_b200.h).
// Host-side control flow mirrors Pplacer_run.run_placements (pplacer_src/pplacer_run.ml:289-340)
// for the precompute and Core.pplacer_core (pplacer_src/core.ml:113-409) for the per-query path,
// batched over queries.  All arithmetic runs in the CUDA kernels of this directory; there is no
// CPU fallback.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <atomic>
#include <string>
#include <thread>
#include <vector>

#include <cstdlib>
#include <cub/device/device_segmented_sort.cuh>
#include <cub/device/device_radix_sort.cuh>

#include "../../include/pplacer_b200.h"
#include "baseball.cuh"
#include "fig.cuh"
#include "device_common.cuh"
#include "model_host.hpp"
#include "precompute.cuh"
#include "prefilter.cuh"
#include "prefilter_staged.cuh"
#include "prefilter_qlane.cuh"
#include "query.cuh"
#include "threetax.cuh"
#include "threetax_warp.cuh"
#include "threetax_group.cuh"
#include "threetax_eigen.cuh"

using namespace ppb;

static thread_local std::string g_err;
static int fail(int code, const std::string &msg) { g_err = msg; return code; }

#define CK(call)                                                                                   \
  do {                                                                                             \
    cudaError_t e_ = (call);                                                                       \
    if (e_ != cudaSuccess) {                                                                       \
      return fail(e_ == cudaErrorMemoryAllocation ? PP_ENOMEM : PP_ECUDA,                          \
                  std::string(#call) + ": " + cudaGetErrorString(e_) + " (" __FILE__ ":" +         \
                      std::to_string(__LINE__) + ")");                                             \
    }                                                                                              \
  } while (0)

// (S, GCAT) dispatch for the templated kernels
#define DISPATCH_MODEL(ctx, ...)                                          \
  do {                                                                    \
    if ((ctx)->S == 4 && !(ctx)->gcat) { constexpr int S_ = 4; constexpr bool G_ = false; __VA_ARGS__; }       \
    else if ((ctx)->S == 4 && (ctx)->gcat) { constexpr int S_ = 4; constexpr bool G_ = true; __VA_ARGS__; }    \
    else if ((ctx)->S == 20 && !(ctx)->gcat) { constexpr int S_ = 20; constexpr bool G_ = false; __VA_ARGS__; }\
    else { constexpr int S_ = 20; constexpr bool G_ = true; __VA_ARGS__; }                                     \
  } while (0)

namespace {

template <class T>
struct DevBuf {
  T *p = nullptr;
  size_t n = 0;
  int ensure(size_t want) {
    if (want <= n) return PP_OK;
    if (p) cudaFree(p);
    p = nullptr; n = 0;
    size_t cap = want + want / 8 + 16;
    cudaError_t e = cudaMalloc(&p, cap * sizeof(T));
    if (e != cudaSuccess) { cudaGetLastError(); return fail(PP_ENOMEM, std::string("cudaMalloc of ") + std::to_string(cap * sizeof(T)) + " bytes: " + cudaGetErrorString(e)); }
    n = cap;
    return PP_OK;
  }
  void release() { if (p) cudaFree(p); p = nullptr; n = 0; }
};

// scope guards: a failing CK(...) returns from the middle of a function; what was allocated before it is released here
struct DevFree {
  void **pp;
  explicit DevFree(void **p) : pp(p) {}
  ~DevFree() { if (*pp) { cudaFree(*pp); *pp = nullptr; } }
  DevFree(const DevFree &) = delete;
};
struct CtxGuard {  // pp_create: destroys the half-built context unless release() was reached
  pp_ctx *c;
  explicit CtxGuard(pp_ctx *c_) : c(c_) {}
  ~CtxGuard();
  void release() { c = nullptr; }
  CtxGuard(const CtxGuard &) = delete;
};

constexpr int kTTThreadsDNA = 128;
constexpr int kTTThreadsAA = 256;

}  // namespace

struct pp_ctx {
  int device = 0;
  cudaStream_t stream = nullptr, own_stream = nullptr;
  int sm_count = 148;
  Alphabet al;
  DevModel hm;
  DevModel *dm = nullptr;
  int S = 0, R = 0, Rg = 0, K = 0, gcat = 0, L = 0, Lp = 0, n_nodes = 0, root_id = 0, n_leaves = 0;
  std::vector<int32_t> child_ptr, child_idx, parent, leaf_row, depth;
  std::vector<double> bl;
  GlvArrays g{};
  TreeDev t{};
  double *d_bl = nullptr;
  uint8_t *d_code_of_char = nullptr, *d_ref_codes = nullptr;
  int32_t *d_site_cat = nullptr, *d_child_ptr = nullptr, *d_child_idx = nullptr, *d_parent = nullptr, *d_leaf_row = nullptr;
  bool precomputed = false;
  // supernodes in prefilter layout (edge fastest), see prefilter.cuh
  double *snT_a = nullptr;
  int32_t *snT_e = nullptr;
  int64_t Np = 0;
  double tq_pend = -1.0;   // pendant length the evolved-query table Tq was built for
  double mt_pend = -1.0;   // ... and the four-base table MT
  TqPrimary tqp;
  DevBuf<uint8_t> qc;
  DevBuf<double> LTq;               // query-lane log table [Np/32][L][32][6]
  DevBuf<double> LTq2;              // pair-of-sites table [Np/32][ceil(L/2)][32][36]
  DevBuf<int32_t> tile_lo2, tile_hi2;
  double ltq2_pend = -1.0;
  DevBuf<uint32_t> qcol;            // column bytes of the sorted queries, four sites per word
  DevBuf<int32_t> slow_flag;
  double ltq_pend = -1.0;
  DevBuf<double> partial, MT, LT;  // LT: log-table form [L][5][Np]
  TqFive tq5;
  double lt_pend = -1.0;
  DevBuf<int32_t> MTe;  // per (site, edge): (supernode exponent + table shift) << 1 | out-of-range flag
  int prefilter_choice = 0;       // pp_set_prefilter_kernel: 0 default, 1 direct, 2 table
  int prefilter_kernel_used = 0;  // 1 direct, 2 table (last run_prefilter)
  // workspace
  DevBuf<uint8_t> chars, codes, sort_tmp;
  DevBuf<int32_t> qsite_cat, nqc, cat_perm, cat_start;  // Gcat: site lists in (category, site) order for three_tax_kernel (query.cuh)
  bool cat_perm_ready = false;
  int64_t LpC = 0;
  DevBuf<int32_t> qsite, nq, prep_status, rank_in, rank_loc, seg_off, i32pool, n_out;
  DevBuf<double> Tq, scores, keys_in, keys_out, wbuf, f64pool;
  DevBuf<CandIn> cand;
  DevBuf<CandOut> cand_out, pp_out;
  DevBuf<PlaceOut> place_out;
  DevBuf<int64_t> n_calls;
  DevBuf<int> counters;  // [0] queue head, [1] wave count, [2] active queries, [3] keep count
  DevBuf<unsigned long long> tt_work;  // second-stage work done (warp engine): pendant calls, distal calls, each x sites
  int *h_counters = nullptr;  // pinned
  int64_t launches = 0;
  double phase_ms[4] = {0, 0, 0, 0};
  cudaEvent_t ev[2] = {nullptr, nullptr};
  int tt_blocks = 0;
  size_t tt_smem = 0;
  // warp-per-candidate engine (threetax_warp.cuh): DNA Gamma-4 only
  int tt_engine = 0;          // pp_set_three_tax_engine: 0 default, 1 block-per-candidate, 2 warp-per-candidate, 3 block + shared-memory w
  int ttg_blocks = 0, ttg_variant = 0;
  int rank_mode = -1;  // pp_set_rank_mode: -1 = default (PPB_RANK_FULL), 0 truncated ranking when possible, 1 always the full sort
  int rank_fallbacks = 0;  // batches that had to be redone on the full ranking
  double batch_budget = 0.0;  // bytes for the per-batch Q x n_nodes buffers (batch_size)
  bool ttg_pp = true;  // --pp integrals on the block engine (PPB_TTG_PP=0: on the warp engine)
  // engine 4 (threetax_eigen.cuh): pendant objective in the eigenbasis; needs the zero eigenvalue of the rate matrix
  EigModel<4> em4{};
  bool eig_ok = false;
  int tte_variant = 0;
  int ttw_blocks_per_sm = 0;  // 0 = occupancy maximum
  int ttw_max_blocks = 0;
  DevBuf<int32_t> ord_keys, ord_vals, ord_keys2, ord_vals2;
  DevBuf<uint8_t> ord_tmp;
  // figs (fig.cuh): host CSR, device copies, the representatives' supernode columns in prefilter layout
  std::vector<int32_t> fig_rep, fig_ptr, fig_idx;
  DevBuf<int32_t> d_fig_rep, d_fig_ptr, d_fig_idx, rep_e, fig_order, fig_i32, fig_pair_q, fig_pair_e, fig_pair_e2, fig_seg;
  DevBuf<double> rep_a, rep_scores, rep_MT, rep_LT, rep_LTq, rep_LTq2;
  DevBuf<int32_t> rep_MTe;
  DevBuf<uint32_t> fig_bits;
  DevBuf<unsigned long long> fig_keys, fig_keys2;
  DevBuf<uint8_t> fig_tmp;
  int64_t rep_Np = 0;
  bool rep_ready = false;
  double rep_mt_pend = -1.0, rep_lt_pend = -1.0, rep_ltq_pend = -1.0, rep_ltq2_pend = -1.0;
  // prefilter tiling: queries sorted by the window of sites they cover (prefilter.cuh window_keys_kernel)
  DevBuf<unsigned long long> win_keys, win_keys2;
  DevBuf<int32_t> win_vals, qorder, tile_lo, tile_hi;
  DevBuf<uint8_t> win_tmp;
};

constexpr int kTTWarpsPerBlock = 8;
constexpr int kTTGroupDefaultVariant = 3;  // 0: 8 warps x 2 blocks/SM, 768 sites of w in shared memory; 1: 4 warps x 4 blocks/SM, 256 sites, one site per thread per iteration; 2: 4 warps x 3 blocks/SM, 512 sites; 3: as 1, two sites per thread per iteration (default); 4: as 1, rate loop unrolled
constexpr int kTTWarpDefaultBlocks = 2;
constexpr int kTTWarpDefaultFlags = 2;
static bool ttw_supported(const pp_ctx *c) { return c->S == 4 && !c->gcat && c->R == 4; }

template <int S, bool G>
static size_t tt_smem_bytes(const pp_ctx *c) {
  if (S == 4) return ThreeTax<S, G, kTTThreadsDNA>::smem_doubles(c->R, c->al.n_codes) * sizeof(double);
  return ThreeTax<S, G, kTTThreadsAA>::smem_doubles(c->R, c->al.n_codes) * sizeof(double);
}

extern "C" const char *pp_last_error(void) { return g_err.c_str(); }
extern "C" int pp_abi_version(void) { return PP_ABI_VERSION; }

extern "C" void pp_default_prefs(pp_prefs *p) {
  std::memset(p, 0, sizeof *p);
  p->start_pend = 0.1; p->max_pend = 2.0; p->initial_tolerance = 0.01; p->final_tolerance = 1e-5;
  p->max_iter = 200; p->max_strikes = 6; p->strike_box = 3.0; p->max_pitches = 40; p->keep_at_most = 7;
  p->keep_factor = 0.01; p->calc_pp = 0; p->prior_kind = 1; p->pp_rel_err = 0.01; p->prior_mean = 1.0;
  p->prior_means = nullptr;
}

extern "C" int pp_discrete_gamma(int32_t n_cat, double alpha, double *out) {
  if (n_cat < 1 || !(alpha > 0) || !out) return fail(PP_EINVAL, "pp_discrete_gamma: bad arguments");
  discrete_gamma(n_cat, alpha, out);
  return PP_OK;
}
extern "C" int pp_diagd(int32_t S, const double *exch, const double *pi, double *x, double *lambda, double *xit) {
  if (S < 2 || S > kMaxStates) return fail(PP_EINVAL, "pp_diagd: bad S");
  diagd_normed(S, exch, pi, x, lambda, xit);
  return PP_OK;
}
extern "C" int pp_emper_freq(int32_t S, const uint8_t *chars, int64_t n, double *out) {
  if (S != 4 && S != 20) return fail(PP_EINVAL, "pp_emper_freq: S must be 4 or 20");
  if (!emper_freq(S, chars, n, out)) return fail(PP_EINVAL, "pp_emper_freq: unknown base in alignment");
  return PP_OK;
}

template <class T>
static int upload(T **dst, const T *src, size_t n) {
  CK(cudaMalloc(dst, std::max<size_t>(n, 1) * sizeof(T)));
  if (n) CK(cudaMemcpy(*dst, src, n * sizeof(T), cudaMemcpyHostToDevice));
  return PP_OK;
}

extern "C" void pp_destroy(pp_ctx *c) {
  if (!c) return;
  cudaSetDevice(c->device);
  if (c->stream) cudaStreamSynchronize(c->stream);
  for (int w = 0; w < 3; ++w) { if (c->g.a[w]) cudaFree(c->g.a[w]); if (c->g.e[w]) cudaFree(c->g.e[w]); }
  for (int w = 0; w < 2; ++w) if (c->g.ea[w]) cudaFree(c->g.ea[w]);
  c->qsite_cat.release(); c->nqc.release(); c->cat_perm.release(); c->cat_start.release(); c->qc.release(); c->partial.release(); c->MT.release(); c->LT.release(); c->LTq.release(); c->LTq2.release(); c->tile_lo2.release(); c->tile_hi2.release(); c->qcol.release(); c->slow_flag.release(); c->MTe.release();
  void *ptrs[] = {c->snT_a, c->snT_e, c->dm, c->d_bl, c->d_code_of_char, c->d_ref_codes, c->d_site_cat, c->d_child_ptr, c->d_child_idx, c->d_parent, c->d_leaf_row};
  for (void *p : ptrs) if (p) cudaFree(p);
  c->chars.release(); c->codes.release(); c->sort_tmp.release(); c->qsite.release(); c->nq.release();
  c->prep_status.release(); c->rank_in.release(); c->rank_loc.release(); c->seg_off.release(); c->i32pool.release();
  c->n_out.release(); c->Tq.release(); c->scores.release(); c->keys_in.release(); c->keys_out.release();
  c->ord_keys.release(); c->ord_vals.release(); c->ord_keys2.release(); c->ord_vals2.release(); c->ord_tmp.release();
  c->d_fig_rep.release(); c->d_fig_ptr.release(); c->d_fig_idx.release(); c->rep_e.release(); c->fig_order.release(); c->fig_i32.release();
  c->fig_pair_q.release(); c->fig_pair_e.release(); c->fig_pair_e2.release(); c->fig_seg.release(); c->rep_a.release(); c->rep_scores.release();
  c->rep_MT.release(); c->rep_LT.release(); c->rep_LTq.release(); c->rep_LTq2.release(); c->rep_MTe.release(); c->fig_bits.release(); c->fig_keys.release(); c->fig_keys2.release(); c->fig_tmp.release();
  c->win_keys.release(); c->win_keys2.release(); c->win_vals.release(); c->qorder.release(); c->tile_lo.release(); c->tile_hi.release(); c->win_tmp.release();
  c->wbuf.release(); c->f64pool.release(); c->cand.release(); c->cand_out.release(); c->pp_out.release();
  c->place_out.release(); c->n_calls.release(); c->counters.release(); c->tt_work.release();
  if (c->h_counters) cudaFreeHost(c->h_counters);
  if (c->ev[0]) cudaEventDestroy(c->ev[0]);
  if (c->ev[1]) cudaEventDestroy(c->ev[1]);
  if (c->own_stream) cudaStreamDestroy(c->own_stream);
  delete c;
}

CtxGuard::~CtxGuard() { if (c) pp_destroy(c); }

extern "C" int pp_create(const pp_model *m, const pp_tree *t, const uint8_t *ref_chars, int32_t n_leaves, int32_t L,
                         int32_t device, pp_ctx **out) {
  if (!m || !t || !ref_chars || !out) return fail(PP_EINVAL, "pp_create: null argument");
  *out = nullptr;
  if (m->n_states != 4 && m->n_states != 20) return fail(PP_EINVAL, "pp_create: n_states must be 4 or 20");
  if (m->n_rates < 1 || m->n_rates > kMaxRates) return fail(PP_EINVAL, "pp_create: n_rates out of range");
  if (m->model_class == 1 && (!m->site_cat || !m->rates)) return fail(PP_EINVAL, "pp_create: Gcat needs rates and site_cat");
  if (L < 1 || t->n_nodes < 2) return fail(PP_EINVAL, "pp_create: empty problem");
  if (L > kSiteMask) return fail(PP_EINVAL, "pp_create: alignment longer than 2^24 - 1 columns");
  if (t->n_nodes > 2000000) return fail(PP_EINVAL, "pp_create: more than 2,000,000 tree nodes (the edge-major transpose indexes edge tiles by grid.y)");
  if (t->root_id < 0 || t->root_id >= t->n_nodes) return fail(PP_EINVAL, "pp_create: root_id out of range");
  if (!t->child_ptr || !t->child_idx || !t->bl || !t->leaf_row || t->child_ptr[0] != 0) return fail(PP_EINVAL, "pp_create: malformed tree arrays");
  for (int i = 0; i < t->n_nodes; ++i)
    if (t->child_ptr[i + 1] < t->child_ptr[i] || t->child_ptr[i + 1] > t->n_nodes) return fail(PP_EINVAL, "pp_create: child_ptr must be non-decreasing and bounded by n_nodes");
  for (int i = 0; i < t->n_nodes; ++i)
    if (i != t->root_id && !(t->bl[i] >= 0.0)) return fail(PP_EINVAL, "pp_create: missing or negative branch length on a non-root edge (Gtree.get_bl raises in the reference)");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); return fail(PP_ENODEVICE, "pp_create: no CUDA device (libpplacer_b200 has no CPU fallback)"); }
  if (device < 0 || device >= ndev) return fail(PP_EINVAL, "pp_create: bad device index");
  CK(cudaSetDevice(device));
  pp_ctx *c = new pp_ctx();
  CtxGuard guard(c);  // every early return below destroys the half-built context
  c->device = device;
  cudaDeviceGetAttribute(&c->sm_count, cudaDevAttrMultiProcessorCount, device);
  c->S = m->n_states; c->gcat = m->model_class ? 1 : 0; c->R = m->n_rates; c->Rg = c->gcat ? 1 : c->R; c->K = c->Rg * c->S;
  c->L = L; c->Lp = (L + 31) / 32 * 32; c->n_nodes = t->n_nodes; c->root_id = t->root_id; c->n_leaves = n_leaves;
  c->al = make_alphabet(c->S);
  // ---- model numerics (a2, a3) ----
  DevModel &hm = c->hm;
  std::memset(&hm, 0, sizeof hm);
  hm.S = c->S; hm.R = c->R; hm.Rg = c->Rg; hm.K = c->K; hm.gcat = c->gcat; hm.n_codes = c->al.n_codes; hm.any_code = c->al.any_code;
  if (m->statd) std::memcpy(hm.statd, m->statd, sizeof(double) * c->S);
  else {
    if (!m->statd_chars) return fail(PP_EINVAL, "pp_create: neither statd nor statd_chars given");
    if (!emper_freq(c->S, m->statd_chars, m->n_statd_chars, hm.statd)) return fail(PP_EINVAL, "pp_create: unknown base in reference alignment");
  }
  diagd_normed(c->S, m->exch, hm.statd, hm.x, hm.lam, hm.xit);
  if (m->rates) std::memcpy(hm.rates, m->rates, sizeof(double) * c->R);
  else discrete_gamma(c->R, m->gamma_alpha, hm.rates);
  std::memcpy(hm.mask_of_code, c->al.mask_of_code, sizeof(uint32_t) * kMaxCodesDev);
  // ---- tree ----
  c->child_ptr.assign(t->child_ptr, t->child_ptr + t->n_nodes + 1);
  c->child_idx.assign(t->child_idx, t->child_idx + t->child_ptr[t->n_nodes]);
  c->leaf_row.assign(t->leaf_row, t->leaf_row + t->n_nodes);
  c->bl.assign(t->bl, t->bl + t->n_nodes);
  c->parent.assign(t->n_nodes, -1);
  for (int i = 0; i < t->n_nodes; ++i)
    for (int k = c->child_ptr[i]; k < c->child_ptr[i + 1]; ++k) {
      const int ch = c->child_idx[k];
      if (ch < 0 || ch >= t->n_nodes || c->parent[ch] != -1 || ch == t->root_id) return fail(PP_EINVAL, "pp_create: malformed tree");
      c->parent[ch] = i;
    }
  c->depth.assign(t->n_nodes, -1);
  {
    std::vector<int> stack{t->root_id};
    c->depth[t->root_id] = 0;
    size_t seen = 0;
    while (!stack.empty()) {
      const int v = stack.back(); stack.pop_back(); ++seen;
      for (int k = c->child_ptr[v]; k < c->child_ptr[v + 1]; ++k) { c->depth[c->child_idx[k]] = c->depth[v] + 1; stack.push_back(c->child_idx[k]); }
    }
    if ((int)seen != t->n_nodes) return fail(PP_EINVAL, "pp_create: tree is not connected");
  }
  for (int i = 0; i < t->n_nodes; ++i) {
    const bool leaf = c->child_ptr[i] == c->child_ptr[i + 1];
    if (leaf && (c->leaf_row[i] < 0 || c->leaf_row[i] >= n_leaves)) return fail(PP_EINVAL, "pp_create: leaf without alignment row");
  }
  // ---- reference alignment -> codes ----
  std::vector<uint8_t> codes((size_t)n_leaves * c->Lp, (uint8_t)(c->al.any_code | kGapBit));
  for (int r = 0; r < n_leaves; ++r)
    for (int s = 0; s < L; ++s) {
      const uint8_t code = c->al.code_of_char[ref_chars[(size_t)r * L + s]];
      if (code == kBadCode) return fail(PP_EINVAL, std::string("pp_create: '") + (char)ref_chars[(size_t)r * L + s] + "' is not a known base in the reference alignment");
      codes[(size_t)r * c->Lp + s] = code;
    }
  std::vector<int32_t> cat(c->Lp, 0);
  if (c->gcat)
    for (int s = 0; s < L; ++s) {
      if (m->site_cat[s] < 0 || m->site_cat[s] >= c->R) return fail(PP_EINVAL, "pp_create: site category out of range");
      cat[s] = m->site_cat[s];
    }
  int rc;
#define UP(dst, src, n) if ((rc = upload(&(dst), (src), (n))) != PP_OK) return rc;
  CK(cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking));
  c->stream = c->own_stream;
  CK(cudaEventCreate(&c->ev[0]));
  CK(cudaEventCreate(&c->ev[1]));
  UP(c->dm, &hm, 1);
  UP(c->d_bl, c->bl.data(), c->bl.size());
  UP(c->d_code_of_char, c->al.code_of_char, 256);
  UP(c->d_ref_codes, codes.data(), codes.size());
  UP(c->d_site_cat, cat.data(), cat.size());
  UP(c->d_child_ptr, c->child_ptr.data(), c->child_ptr.size());
  UP(c->d_child_idx, c->child_idx.data(), c->child_idx.size());
  UP(c->d_parent, c->parent.data(), c->parent.size());
  UP(c->d_leaf_row, c->leaf_row.data(), c->leaf_row.size());
#undef UP
  c->g.L = c->L; c->g.Lp = c->Lp; c->g.K = c->K; c->g.n_nodes = c->n_nodes; c->g.a_stride = (int64_t)c->K * c->Lp;
  for (int w = 0; w < 3; ++w) {
    cudaError_t e1 = cudaMalloc(&c->g.a[w], sizeof(double) * (size_t)c->g.a_stride * c->n_nodes);
    cudaError_t e2 = e1 == cudaSuccess ? cudaMalloc(&c->g.e[w], sizeof(int32_t) * (size_t)c->Lp * c->n_nodes) : e1;
    if (e1 != cudaSuccess || e2 != cudaSuccess) { cudaGetLastError(); return fail(PP_ENOMEM, "pp_create: not enough device memory for darr/parr/snodes"); }
    cudaMemsetAsync(c->g.a[w], 0, sizeof(double) * (size_t)c->g.a_stride * c->n_nodes, c->stream);
    cudaMemsetAsync(c->g.e[w], 0, sizeof(int32_t) * (size_t)c->Lp * c->n_nodes, c->stream);
  }
  c->Np = (c->n_nodes + 63) / 64 * 64;
  {
    cudaError_t e1 = cudaMalloc(&c->snT_a, sizeof(double) * (size_t)c->L * c->K * c->Np);
    cudaError_t e2 = e1 == cudaSuccess ? cudaMalloc(&c->snT_e, sizeof(int32_t) * (size_t)c->L * c->Np) : e1;
    if (e1 != cudaSuccess || e2 != cudaSuccess) { cudaGetLastError(); return fail(PP_ENOMEM, "pp_create: not enough device memory for the prefilter copy of the supernodes"); }
  }
  c->t.child_ptr = c->d_child_ptr; c->t.child_idx = c->d_child_idx; c->t.parent = c->d_parent; c->t.leaf_row = c->d_leaf_row;
  c->t.ref_codes = c->d_ref_codes; c->t.site_cat = c->d_site_cat;
  CK(cudaMallocHost(&c->h_counters, 8 * sizeof(int)));
  if ((rc = c->counters.ensure(8)) != PP_OK || (rc = c->tt_work.ensure(32)) != PP_OK) return rc;
  CK(cudaMemsetAsync(c->tt_work.p, 0, 32 * sizeof(unsigned long long), c->stream));
  // QNG tables -> constant memory
  cudaMemcpyToSymbol(c_qng_x1, ppb_qng_x1, sizeof ppb_qng_x1); cudaMemcpyToSymbol(c_qng_w10, ppb_qng_w10, sizeof ppb_qng_w10);
  cudaMemcpyToSymbol(c_qng_x2, ppb_qng_x2, sizeof ppb_qng_x2); cudaMemcpyToSymbol(c_qng_w21a, ppb_qng_w21a, sizeof ppb_qng_w21a);
  cudaMemcpyToSymbol(c_qng_w21b, ppb_qng_w21b, sizeof ppb_qng_w21b); cudaMemcpyToSymbol(c_qng_x3, ppb_qng_x3, sizeof ppb_qng_x3);
  cudaMemcpyToSymbol(c_qng_w43a, ppb_qng_w43a, sizeof ppb_qng_w43a); cudaMemcpyToSymbol(c_qng_w43b, ppb_qng_w43b, sizeof ppb_qng_w43b);
  cudaMemcpyToSymbol(c_qng_x4, ppb_qng_x4, sizeof ppb_qng_x4); cudaMemcpyToSymbol(c_qng_w87a, ppb_qng_w87a, sizeof ppb_qng_w87a);
  cudaMemcpyToSymbol(c_qng_w87b, ppb_qng_w87b, sizeof ppb_qng_w87b);
  // three-taxon kernel configuration
  DISPATCH_MODEL(c, {
    c->tt_smem = tt_smem_bytes<S_, G_>(c);
    constexpr int TH = (S_ == 4) ? kTTThreadsDNA : kTTThreadsAA;
    auto kern = three_tax_kernel<S_, G_, TH>;
    if (c->tt_smem > 227 * 1024) return fail(PP_EINVAL, "pp_create: model too large for shared memory (rates x states^2)");
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->tt_smem));
    int per_sm = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, TH, c->tt_smem));
    if (per_sm < 1) per_sm = 1;
    c->tt_blocks = per_sm * c->sm_count;
  });
  size_t w_slots = (size_t)c->tt_blocks;
  if (ttw_supported(c)) {
    const size_t smw = TTWarp<4, 1>::smem_bytes(kTTWarpsPerBlock);
    CK(cudaFuncSetAttribute(tt_warp_kernel<4, kTTWarpsPerBlock, 2, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smw));
    CK(cudaFuncSetAttribute(tt_warp_kernel<4, kTTWarpsPerBlock, 3, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smw));
    int per_sm = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, tt_warp_kernel<4, kTTWarpsPerBlock, 3, 1>, kTTWarpsPerBlock * 32, smw));
    c->ttw_max_blocks = std::max(per_sm, 2) * c->sm_count;
    w_slots = std::max(w_slots, (size_t)c->ttw_max_blocks * kTTWarpsPerBlock);
    static const int variant = [] { const char *e = std::getenv("PPB_TTG_VARIANT"); return e ? std::atoi(e) : kTTGroupDefaultVariant; }();
    c->ttg_variant = variant;
    static const bool pp_on_group = [] { const char *e = std::getenv("PPB_TTG_PP"); return e ? std::atoi(e) != 0 : true; }();
    c->ttg_pp = pp_on_group;
    w_slots = std::max(w_slots, (size_t)6 * c->sm_count);
    c->ttg_blocks = 4 * c->sm_count;
    w_slots = std::max(w_slots, (size_t)c->ttg_blocks);
  }
  if (ttw_supported(c)) {
    // engine 4: split off the zero eigenvalue (the rows of a rate matrix sum to zero, so one eigenvalue is 0 up to rounding)
    int k0 = 0;
    double lmax = 0.0;
    for (int k = 0; k < 4; ++k) { if (std::fabs(hm.lam[k]) < std::fabs(hm.lam[k0])) k0 = k; lmax = std::max(lmax, std::fabs(hm.lam[k])); }
    c->eig_ok = std::fabs(hm.lam[k0]) <= 1e-9 * lmax;
    EigModel<4> &em = c->em4;
    std::memcpy(em.statd, hm.statd, sizeof em.statd); std::memcpy(em.lam, hm.lam, sizeof em.lam);
    std::memcpy(em.x, hm.x, sizeof em.x); std::memcpy(em.xit, hm.xit, sizeof em.xit); std::memcpy(em.rates, hm.rates, sizeof em.rates);
    int kk[3], nk = 0;
    for (int k = 0; k < 4; ++k) if (k != k0) kk[nk++] = k;
    for (int j = 0; j < 3; ++j) em.lamn[j] = hm.lam[kk[j]];
    for (int s = 0; s < 4; ++s) { em.xz[s] = hm.x[s * 4 + k0]; for (int j = 0; j < 3; ++j) em.xn[s * 3 + j] = hm.x[s * 4 + kk[j]]; }
    for (int mask = 0; mask < 16; ++mask)
      for (int j = 0; j < 4; ++j) {
        const int k = j == 0 ? k0 : kk[j - 1];
        double acc = 0.0;
        for (int cc = 0; cc < 4; ++cc) if (mask >> cc & 1) acc += hm.xit[cc * 4 + k];
        em.xq[mask * 4 + j] = acc;
      }
    static const int tte = [] { const char *e = std::getenv("PPB_TTE_VARIANT"); return e ? std::atoi(e) : 0; }();
    c->tte_variant = tte;
  }
  // Gcat: per block w + the compact distal / proximal copies, rows of LpC = the padded (category, site) list capacity
  c->LpC = c->gcat ? ((c->Lp + 32 * (int64_t)c->R + 31) / 32) * 32 : c->Lp;
  if ((rc = c->wbuf.ensure(c->gcat ? w_slots * 3 * c->K * c->LpC : w_slots * c->K * c->Lp)) != PP_OK) return rc;
  CK(cudaStreamSynchronize(c->stream));
  CK(cudaGetLastError());
  guard.release();
  *out = c;
  return PP_OK;
}

extern "C" int pp_get_model(pp_ctx *c, double *statd, double *x, double *lambda, double *xit, double *rates) {
  if (!c) return fail(PP_EINVAL, "null ctx");
  if (statd) std::memcpy(statd, c->hm.statd, sizeof(double) * c->S);
  if (x) std::memcpy(x, c->hm.x, sizeof(double) * c->S * c->S);
  if (lambda) std::memcpy(lambda, c->hm.lam, sizeof(double) * c->S);
  if (xit) std::memcpy(xit, c->hm.xit, sizeof(double) * c->S * c->S);
  if (rates) std::memcpy(rates, c->hm.rates, sizeof(double) * c->R);
  return PP_OK;
}

#define LAUNCH_CHECK(c) do { (c)->launches++; CK(cudaGetLastError()); } while (0)

// ---------------------------------------------------------------------------------------------
// precompute (pplacer_run.ml:322-340)
// ---------------------------------------------------------------------------------------------
struct LevelLists {
  std::vector<int32_t> flat;
  std::vector<int> off;
  int max_depth = 0;
};
static LevelLists level_lists(const pp_ctx *c) {
  LevelLists ll;
  for (int d : c->depth) ll.max_depth = std::max(ll.max_depth, d);
  std::vector<std::vector<int32_t>> levels(ll.max_depth + 1);
  for (int i = 0; i < c->n_nodes; ++i) levels[c->depth[i]].push_back(i);
  ll.off.assign(ll.max_depth + 2, 0);
  for (int d = 0; d <= ll.max_depth; ++d) { ll.off[d] = (int)ll.flat.size(); ll.flat.insert(ll.flat.end(), levels[d].begin(), levels[d].end()); }
  ll.off[ll.max_depth + 1] = (int)ll.flat.size();
  return ll;
}

// Gcat_model.refine (gcat_model.ml:387-419; called from pplacer_run.ml:316-317 when the model does not come straight from
// the reference package or with --always-refine): R attempts of the per-site tree likelihood with every site in one
// category, best category per site under the FastTree prior.  The context's site categories are REPLACED (call before
// pp_precompute, as the reference does); site_cat_out [L] and site_ll_out [R x L] (optional) receive the result.
extern "C" int pp_refine_site_categories(pp_ctx *c, int32_t *site_cat_out, double *site_ll_out) {
  if (!c) return fail(PP_EINVAL, "null ctx");
  if (!c->gcat) return fail(PP_EINVAL, "pp_refine_site_categories: only the Price-CAT model has site categories (Gmix_model.refine is a no-op)");
  CK(cudaSetDevice(c->device));
  const int Rt = c->R;
  const LevelLists ll = level_lists(c);
  double *Pfull = nullptr, *best = nullptr, *site_ll = nullptr;
  int32_t *d_levels = nullptr, *best_cat = nullptr;
  int *n_bad = nullptr;
  auto cleanup = [&] { cudaFree(Pfull); cudaFree(best); cudaFree(site_ll); cudaFree(d_levels); cudaFree(best_cat); cudaFree(n_bad); };
  int rc = PP_OK;
  auto body = [&]() -> int {
    CK(cudaMalloc(&Pfull, (size_t)c->n_nodes * Rt * c->S * c->S * sizeof(double)));
    CK(cudaMalloc(&best, (size_t)c->L * sizeof(double)));
    CK(cudaMalloc(&best_cat, (size_t)c->L * sizeof(int32_t)));
    CK(cudaMalloc(&n_bad, sizeof(int)));
    if (site_ll_out) CK(cudaMalloc(&site_ll, (size_t)Rt * c->L * sizeof(double)));
    CK(cudaMalloc(&d_levels, ll.flat.size() * sizeof(int32_t)));
    CK(cudaMemcpyAsync(d_levels, ll.flat.data(), ll.flat.size() * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemsetAsync(n_bad, 0, sizeof(int), c->stream));
    fill_f64_kernel<<<64, 256, 0, c->stream>>>(best, -INFINITY, c->L); LAUNCH_CHECK(c);
    fill_i32_kernel<<<64, 256, 0, c->stream>>>(best_cat, -1, c->L); LAUNCH_CHECK(c);
    const int tiles = (c->L + kPreThreads - 1) / kPreThreads;
    if (c->child_ptr[c->root_id + 1] - c->child_ptr[c->root_id] < 2) return fail(PP_EINVAL, "pp_refine_site_categories: degree two node at the root");
    DISPATCH_MODEL(c, {
      if (G_) {
        const size_t sm_p = (size_t)(Rt * S_ + Rt * S_ * S_) * sizeof(double);
        CK(cudaFuncSetAttribute(pmat_kernel<S_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_p));
        pmat_kernel<S_><<<c->n_nodes, 128, sm_p, c->stream>>>(c->dm, c->d_bl, c->n_nodes, 1.0, Rt, Pfull); LAUNCH_CHECK(c);
        const size_t sm_l = (size_t)Rt * p_rstride<S_, true>() * sizeof(double);
        CK(cudaFuncSetAttribute(distal_level_kernel<S_, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_l));
        for (int cat = 0; cat < Rt; ++cat) {
          fill_i32_kernel<<<64, 256, 0, c->stream>>>(c->d_site_cat, cat, c->Lp); LAUNCH_CHECK(c);
          for (int d = ll.max_depth; d >= 1; --d) {  // the root's own distal vector is not needed
            const int n = ll.off[d + 1] - ll.off[d];
            for (int s0 = 0; s0 < n; s0 += 65535) {
              dim3 grid(tiles, std::min(65535, n - s0));
              distal_level_kernel<S_, true><<<grid, kPreThreads, sm_l, c->stream>>>(c->dm, c->g, c->t, d_levels + ll.off[d] + s0, Pfull, Rt); LAUNCH_CHECK(c);
            }
          }
          const double prior = 2.0 * std::log(c->hm.rates[cat]) - 3.0 * c->hm.rates[cat];
          refine_attempt_kernel<S_><<<tiles, kPreThreads, 0, c->stream>>>(c->dm, c->g, c->t, c->root_id, cat, prior, best, best_cat, site_ll, n_bad); LAUNCH_CHECK(c);
        }
      }
    });
    // pad sites keep category 0
    fill_i32_kernel<<<64, 256, 0, c->stream>>>(c->d_site_cat, 0, c->Lp); LAUNCH_CHECK(c);
    CK(cudaMemcpyAsync(c->d_site_cat, best_cat, (size_t)c->L * sizeof(int32_t), cudaMemcpyDeviceToDevice, c->stream));
    int h_bad = 0;
    CK(cudaMemcpyAsync(&h_bad, n_bad, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    if (site_cat_out) CK(cudaMemcpyAsync(site_cat_out, best_cat, (size_t)c->L * sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
    if (site_ll_out) CK(cudaMemcpyAsync(site_ll_out, site_ll, (size_t)Rt * c->L * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    if (h_bad > 0) return fail(PP_EINVAL, "pp_refine_site_categories: a site has zero likelihood (gcat_model.ml:368-369)");
    c->precomputed = false;  // the GLVs on the device are those of the last attempt
    c->cat_perm_ready = false;
    return PP_OK;
  };
  rc = body();
  cleanup();
  return rc;
}

extern "C" int pp_precompute(pp_ctx *c) {
  if (!c) return fail(PP_EINVAL, "null ctx");
  CK(cudaSetDevice(c->device));
  const int Rt = c->R;
  const size_t pm = (size_t)c->n_nodes * Rt * c->S * c->S;
  double *Pfull = nullptr, *Phalf = nullptr;
  int32_t *d_levels = nullptr;
  DevFree g_pfull(reinterpret_cast<void **>(&Pfull)), g_phalf(reinterpret_cast<void **>(&Phalf)), g_levels(reinterpret_cast<void **>(&d_levels));
  CK(cudaMalloc(&Pfull, pm * sizeof(double)));
  CK(cudaMalloc(&Phalf, pm * sizeof(double)));
  // level lists
  int max_depth = 0;
  for (int d : c->depth) max_depth = std::max(max_depth, d);
  std::vector<std::vector<int32_t>> levels(max_depth + 1);
  for (int i = 0; i < c->n_nodes; ++i) levels[c->depth[i]].push_back(i);
  std::vector<int32_t> flat; std::vector<int> off(max_depth + 2, 0);
  for (int d = 0; d <= max_depth; ++d) { off[d] = (int)flat.size(); flat.insert(flat.end(), levels[d].begin(), levels[d].end()); }
  off[max_depth + 1] = (int)flat.size();
  CK(cudaMalloc(&d_levels, flat.size() * sizeof(int32_t)));
  CK(cudaMemcpyAsync(d_levels, flat.data(), flat.size() * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
  const int tiles = (c->L + kPreThreads - 1) / kPreThreads;
  int rc = PP_OK;
  DISPATCH_MODEL(c, {
    const size_t sm_p = (size_t)(Rt * S_ + Rt * S_ * S_) * sizeof(double);
    CK(cudaFuncSetAttribute(pmat_kernel<S_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_p));
    pmat_kernel<S_><<<c->n_nodes, 128, sm_p, c->stream>>>(c->dm, c->d_bl, c->n_nodes, 1.0, Rt, Pfull); LAUNCH_CHECK(c);
    pmat_kernel<S_><<<c->n_nodes, 128, sm_p, c->stream>>>(c->dm, c->d_bl, c->n_nodes, 0.5, Rt, Phalf); LAUNCH_CHECK(c);
    const size_t sm_l = (size_t)Rt * p_rstride<S_, G_>() * sizeof(double);
    CK(cudaFuncSetAttribute(distal_level_kernel<S_, G_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_l));
    CK(cudaFuncSetAttribute(proximal_level_kernel<S_, G_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_l));
    CK(cudaFuncSetAttribute(supernode_kernel<S_, G_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_l));
    for (int d = max_depth; d >= 0; --d) {  // post-order by level: deepest first
      const int n = off[d + 1] - off[d];
      for (int s0 = 0; s0 < n; s0 += 65535) {
        dim3 grid(tiles, std::min(65535, n - s0));
        distal_level_kernel<S_, G_><<<grid, kPreThreads, sm_l, c->stream>>>(c->dm, c->g, c->t, d_levels + off[d] + s0, Pfull, Rt); LAUNCH_CHECK(c);
      }
    }
    set_unit_kernel<<<64, 256, 0, c->stream>>>(c->g, 1, c->root_id); LAUNCH_CHECK(c);
    for (int d = 1; d <= max_depth; ++d) {  // pre-order by level: children of level d-1
      const int n = off[d + 1] - off[d];
      for (int s0 = 0; s0 < n; s0 += 65535) {
        dim3 grid(tiles, std::min(65535, n - s0));
        proximal_level_kernel<S_, G_><<<grid, kPreThreads, sm_l, c->stream>>>(c->dm, c->g, c->t, d_levels + off[d] + s0, Pfull, Rt); LAUNCH_CHECK(c);
      }
    }
    for (int s0 = 0; s0 < c->n_nodes; s0 += 65535) {
      dim3 grid(tiles, std::min(65535, c->n_nodes - s0));
      supernode_kernel<S_, G_><<<grid, kPreThreads, sm_l, c->stream>>>(c->dm, c->g, c->t, Phalf, Rt, s0); LAUNCH_CHECK(c);
    }
  });
  if (rc == PP_OK) {
    // supernodes -> prefilter layout snT[site][k][edge] (prefilter.cuh)
    dim3 grid((c->L + 31) / 32, (unsigned)(c->Np / 32), c->K);
    transpose_to_edge_major_kernel<double><<<grid, 256, 0, c->stream>>>(c->g.a[2], c->snT_a, c->n_nodes, c->K, c->L, c->Lp, c->Np, 1.0); LAUNCH_CHECK(c);
    dim3 grid_e((c->L + 31) / 32, (unsigned)(c->Np / 32), 1);
    transpose_to_edge_major_kernel<int32_t><<<grid_e, 256, 0, c->stream>>>(c->g.e[2], c->snT_e, c->n_nodes, 1, c->L, c->Lp, c->Np, 0); LAUNCH_CHECK(c);
    c->tq_pend = c->mt_pend = c->lt_pend = c->ltq_pend = c->ltq2_pend = -1.0;
    c->rep_ready = false; c->rep_mt_pend = c->rep_lt_pend = c->rep_ltq_pend = c->rep_ltq2_pend = -1.0;
  }
  if (rc == PP_OK && ttw_supported(c) && c->eig_ok) {
    // engine 4 reads darr / parr in the eigenbasis of the rate matrix (threetax_eigen.cuh); without room for the two
    // copies the second stage falls back to engine 3
    const size_t bytes = sizeof(double) * (size_t)c->g.a_stride * c->n_nodes;
    for (int w = 0; w < 2 && c->eig_ok; ++w)
      if (!c->g.ea[w] && cudaMalloc(&c->g.ea[w], bytes) != cudaSuccess) { cudaGetLastError(); c->g.ea[w] = nullptr; c->eig_ok = false; }
    if (c->eig_ok) {
      XitMat xm;
      std::memcpy(xm.v, c->em4.xit, sizeof xm.v);
      const dim3 grid((unsigned)((c->Lp + 255) / 256), (unsigned)std::min<int64_t>((int64_t)c->n_nodes * c->R, 32768));
      for (int w = 0; w < 2; ++w) { eigen_glv_kernel<<<grid, 256, 0, c->stream>>>(c->g.a[w], c->g.ea[w], (int64_t)c->n_nodes * c->R, c->Lp, xm); LAUNCH_CHECK(c); }
    }
  }
  cudaError_t e = cudaStreamSynchronize(c->stream);
  // Pfull, Phalf, d_levels: released by their guards
  if (e != cudaSuccess) return fail(PP_ECUDA, std::string("pp_precompute: ") + cudaGetErrorString(e));
  if (rc != PP_OK) return rc;
  c->precomputed = true;
  return PP_OK;
}

extern "C" int pp_export_glvs(pp_ctx *c, int32_t which, int32_t nb, int32_t ne, double *a, int64_t *e) {
  if (!c || which < 0 || which > 2 || nb < 0 || ne > c->n_nodes || nb > ne) return fail(PP_EINVAL, "pp_export_glvs: bad arguments");
  CK(cudaSetDevice(c->device));
  const size_t per = (size_t)c->g.a_stride;
  std::vector<double> tmp(per);
  std::vector<int32_t> te(c->Lp);
  for (int node = nb; node < ne; ++node) {
    if (a) {
      CK(cudaMemcpy(tmp.data(), c->g.a[which] + (size_t)node * per, per * sizeof(double), cudaMemcpyDeviceToHost));
      double *dst = a + (size_t)(node - nb) * c->K * c->L;
      // device [r*S+s][Lp] -> reference [r][L][S]
      for (int r = 0; r < c->Rg; ++r)
        for (int s = 0; s < c->S; ++s) {
          const double *pl = tmp.data() + (size_t)(r * c->S + s) * c->Lp;
          for (int site = 0; site < c->L; ++site) dst[((size_t)r * c->L + site) * c->S + s] = pl[site];
        }
    }
    if (e) {
      CK(cudaMemcpy(te.data(), c->g.e[which] + (size_t)node * c->Lp, c->Lp * sizeof(int32_t), cudaMemcpyDeviceToHost));
      for (int site = 0; site < c->L; ++site) e[(size_t)(node - nb) * c->L + site] = te[site];
    }
  }
  return PP_OK;
}

// ---------------------------------------------------------------------------------------------
// query preparation and prefilter
// ---------------------------------------------------------------------------------------------
static int prep_queries(pp_ctx *c, int64_t Q, const uint8_t *d_chars) {
  int rc;
  if ((rc = c->codes.ensure((size_t)Q * c->Lp)) || (rc = c->qsite.ensure((size_t)Q * c->Lp)) || (rc = c->nq.ensure(Q)) ||
      (rc = c->prep_status.ensure(Q)))
    return rc;
  query_prep_kernel<<<(unsigned)Q, kPrepThreads, 0, c->stream>>>(d_chars, c->d_code_of_char, c->L, c->Lp, c->codes.p, c->qsite.p,
                                                                  c->nq.p, c->prep_status.p, PP_Q_BAD_CHAR, PP_Q_NO_INFORMATIVE);
  LAUNCH_CHECK(c);
  if (c->gcat) {  // the second stage walks the sites category by category
    if (!c->cat_perm_ready) {
      std::vector<int32_t> cat(c->L), perm(c->L), start(c->R + 1, 0);
      CK(cudaMemcpyAsync(cat.data(), c->d_site_cat, sizeof(int32_t) * c->L, cudaMemcpyDeviceToHost, c->stream));
      CK(cudaStreamSynchronize(c->stream));
      for (int i = 0; i < c->L; ++i) { perm[i] = i; start[cat[i] + 1]++; }
      for (int r = 0; r < c->R; ++r) start[r + 1] += start[r];
      std::stable_sort(perm.begin(), perm.end(), [&](int32_t a, int32_t b) { return cat[a] < cat[b]; });
      if ((rc = c->cat_perm.ensure(c->L)) || (rc = c->cat_start.ensure(c->R + 1))) return rc;
      CK(cudaMemcpyAsync(c->cat_perm.p, perm.data(), sizeof(int32_t) * c->L, cudaMemcpyHostToDevice, c->stream));
      CK(cudaMemcpyAsync(c->cat_start.p, start.data(), sizeof(int32_t) * (c->R + 1), cudaMemcpyHostToDevice, c->stream));
      CK(cudaStreamSynchronize(c->stream));
      c->cat_perm_ready = true;
    }
    if ((rc = c->qsite_cat.ensure((size_t)Q * c->LpC)) || (rc = c->nqc.ensure(Q))) return rc;
    qsite_by_category_kernel<<<(unsigned)Q, kPrepThreads, 0, c->stream>>>(c->codes.p, c->cat_perm.p, c->cat_start.p, c->R, c->Lp, c->LpC,
                                                                         c->qsite_cat.p, c->nqc.p); LAUNCH_CHECK(c);
  }
  // tile order of the prefilter: by the window of sites each query covers (library sort of Q keys: plumbing)
  if ((rc = c->win_keys.ensure(Q)) || (rc = c->win_keys2.ensure(Q)) || (rc = c->win_vals.ensure(Q)) || (rc = c->qorder.ensure(Q))) return rc;
  window_keys_kernel<<<(unsigned)((Q + 255) / 256), 256, 0, c->stream>>>(c->qsite.p, c->nq.p, (int)Q, c->Lp, c->win_keys.p, c->win_vals.p); LAUNCH_CHECK(c);
  size_t tmp = 0;
  CK(cub::DeviceRadixSort::SortPairs(nullptr, tmp, c->win_keys.p, c->win_keys2.p, c->win_vals.p, c->qorder.p, (int)Q, 0, 64, c->stream));
  if ((rc = c->win_tmp.ensure(tmp))) return rc;
  CK(cub::DeviceRadixSort::SortPairs(c->win_tmp.p, tmp, c->win_keys.p, c->win_keys2.p, c->win_vals.p, c->qorder.p, (int)Q, 0, 64, c->stream));
  c->launches += 1;
  return PP_OK;
}

static int build_query_table(pp_ctx *c, double t) {
  int rc;
  if (c->tq_pend == t) return PP_OK;  // the table only depends on the model and the pendant length
  const size_t n = (size_t)c->R * c->al.n_codes * c->S;
  if ((rc = c->Tq.ensure(n))) return rc;
  DISPATCH_MODEL(c, {
    const size_t sm = (size_t)(c->R * S_ + c->R * S_ * S_) * sizeof(double);
    CK(cudaFuncSetAttribute(query_table_kernel<S_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    query_table_kernel<S_><<<1, 256, sm, c->stream>>>(c->dm, t, c->Tq.p);
    (void)G_;
    LAUNCH_CHECK(c);
  });
  std::memset(&c->tqp, 0, sizeof c->tqp);
  if (c->S == 4 && !c->gcat && c->K <= 16) {
    // the four base columns go to the kernels by value (parameter constant bank)
    std::vector<double> h(n);
    CK(cudaMemcpyAsync(h.data(), c->Tq.p, n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    for (int j = 0; j < 4; ++j)
      for (int k = 0; k < c->K; ++k) c->tqp.v[j][k] = h[(size_t)(1 << j) * c->K + k];
    std::memset(&c->tq5, 0, sizeof c->tq5);
    for (int j = 0; j < 5; ++j)
      for (int k = 0; k < c->K; ++k) c->tq5.v[j][k] = h[(size_t)(j < 4 ? (1 << j) : 15) * c->K + k];
  }
  c->tq_pend = t;
  return PP_OK;
}

// ---- prefilter launch plumbing (prefilter.cuh) ----
struct PreGeom { int nqt, gx, gy, zs, sites_per_z; };

static PreGeom prefilter_geometry(const pp_ctx *c, int64_t Q, int QT, int TE, int W, int blocks_per_sm) {
  PreGeom g;
  g.nqt = (int)((Q + QT - 1) / QT);
  g.gx = (g.nqt + W - 1) / W;
  g.gy = (int)(c->Np / (32 * TE));
  // few queries on a long alignment: split the site range so every SM holds its share of blocks
  const int64_t want = (int64_t)c->sm_count * blocks_per_sm * 2;
  int zs = (int)std::min<int64_t>((want + (int64_t)g.gx * g.gy - 1) / ((int64_t)g.gx * g.gy), std::max(1, c->L / 256));
  zs = std::max(1, std::min(zs, 64));
  g.sites_per_z = (c->L + zs - 1) / zs;
  g.zs = (c->L + g.sites_per_z - 1) / g.sites_per_z;
  return g;
}

template <int QT>
static int pack_codes(pp_ctx *c, int64_t Q, int nqt) {
  int rc;
  if ((rc = c->qc.ensure((size_t)nqt * c->L * QT + 16))) return rc;
  if ((rc = c->tile_lo.ensure(nqt)) || (rc = c->tile_hi.ensure(nqt))) return rc;
  dim3 grid((c->L + 255) / 256, nqt);
  pack_codes_kernel<QT><<<grid, 256, 0, c->stream>>>(c->codes.p, c->qorder.p, c->qc.p, (int)Q, c->L, c->Lp); LAUNCH_CHECK(c);
  tile_range_kernel<QT><<<(nqt + 127) / 128, 128, 0, c->stream>>>(c->qsite.p, c->nq.p, c->qorder.p, (int)Q, nqt, c->Lp, c->tile_lo.p, c->tile_hi.p); LAUNCH_CHECK(c);
  return PP_OK;
}

static int prefilter_args(pp_ctx *c, int64_t Q, const PreGeom &g, PreArgs *A) {
  int rc;
  A->sn.a = c->snT_a; A->sn.e = c->snT_e; A->sn.Np = c->Np; A->sn.L = c->L; A->sn.K = c->K; A->sn.n_nodes = c->n_nodes;
  A->qc = c->qc.p; A->Tq = c->Tq.p; A->site_cat = c->d_site_cat; A->nq = c->nq.p; A->bad_edge = nullptr;
  A->order = c->qorder.p; A->tile_lo = c->tile_lo.p; A->tile_hi = c->tile_hi.p;
  A->Q = (int)Q; A->nqt = g.nqt; A->root_id = c->root_id; A->R = c->R; A->nc = c->al.n_codes; A->zs = g.zs; A->sites_per_z = g.sites_per_z;
  if (g.zs > 1) {
    if ((rc = c->partial.ensure((size_t)g.zs * Q * c->n_nodes))) return rc;
    A->out = c->partial.p;
  } else
    A->out = c->scores.p;
  return PP_OK;
}

static int prefilter_finish(pp_ctx *c, int64_t Q, const PreGeom &g) {
  if (g.zs > 1) {
    const int64_t total = Q * (int64_t)c->n_nodes;
    combine_partials_kernel<<<c->sm_count * 4, 256, 0, c->stream>>>(c->partial.p, c->scores.p, g.zs, c->n_nodes, c->root_id, total); LAUNCH_CHECK(c);
  }
  return PP_OK;
}

template <int S, bool G, int NRC, int TE, int QT, int W, int TQSRC>
static int launch_prefilter_direct(pp_ctx *c, int64_t Q, int blocks_per_sm) {
  int rc;
  const PreGeom g = prefilter_geometry(c, Q, QT, TE, W, blocks_per_sm);
  PreArgs A;
  if ((rc = pack_codes<QT>(c, Q, g.nqt)) || (rc = prefilter_args(c, Q, g, &A))) return rc;
  const size_t tq_bytes = (size_t)c->al.n_codes * c->K * sizeof(double);
  const size_t sm = TQSRC == TQ_SMEM ? tq_bytes : 0;
  auto kern = prefilter_direct_kernel<S, G, NRC, TE, QT, W, TQSRC>;
  if (TQSRC == TQ_CONST)  // stream-ordered refresh of the constant-bank copy of the evolved-query table
    CK(cudaMemcpyToSymbolAsync(c_tq, c->Tq.p, tq_bytes, 0, cudaMemcpyDeviceToDevice, c->stream));
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
  kern<<<dim3(g.gx, g.gy, g.zs), W * 32, sm, c->stream>>>(A); LAUNCH_CHECK(c);
  c->prefilter_kernel_used = 1;
  return prefilter_finish(c, Q, g);
}

template <int TE, int QT, int W, int MINB = 1>
static int launch_prefilter_table(pp_ctx *c, int64_t Q, double pend, int blocks_per_sm) {
  int rc;
  if (c->mt_pend != pend) {  // four-base table: once per (reference, pendant length)
    if ((rc = c->MT.ensure((size_t)c->L * 4 * c->Np)) || (rc = c->MTe.ensure((size_t)c->L * c->Np))) return rc;
    SnT sn; sn.a = c->snT_a; sn.e = c->snT_e; sn.Np = c->Np; sn.L = c->L; sn.K = c->K; sn.n_nodes = c->n_nodes;
    for (int s0 = 0; s0 < c->L; s0 += 65535) {  // grid.y is capped at 65535 sites per launch
      mtable_kernel<16><<<dim3((unsigned)((c->Np + 127) / 128), std::min(65535, c->L - s0)), 128, 0, c->stream>>>(sn, c->tqp, c->MT.p, c->MTe.p, s0); LAUNCH_CHECK(c);
    }
    c->mt_pend = pend;
  }
  const PreGeom g = prefilter_geometry(c, Q, QT, TE, W, blocks_per_sm);
  PreArgs A;
  if ((rc = pack_codes<QT>(c, Q, g.nqt)) || (rc = prefilter_args(c, Q, g, &A))) return rc;
  prefilter_table_kernel<TE, QT, W, MINB><<<dim3(g.gx, g.gy, g.zs), W * 32, 0, c->stream>>>(A, c->MT.p, c->MTe.p); LAUNCH_CHECK(c);
  c->prefilter_kernel_used = 2;
  return prefilter_finish(c, Q, g);
}

template <int QT, int W, int MINB>
static int launch_prefilter_logtable(pp_ctx *c, int64_t Q, double pend, int blocks_per_sm) {
  int rc;
  if (c->lt_pend != pend) {  // log table: once per (reference, pendant length)
    if ((rc = c->LT.ensure((size_t)c->L * 5 * c->Np))) return rc;
    SnT sn; sn.a = c->snT_a; sn.e = c->snT_e; sn.Np = c->Np; sn.L = c->L; sn.K = c->K; sn.n_nodes = c->n_nodes;
    for (int s0 = 0; s0 < c->L; s0 += 65535) {
      logtable_kernel<16><<<dim3((unsigned)((c->Np + 127) / 128), std::min(65535, c->L - s0)), 128, 0, c->stream>>>(sn, c->tq5, c->LT.p, s0); LAUNCH_CHECK(c);
    }
    c->lt_pend = pend;
  }
  const PreGeom g = prefilter_geometry(c, Q, QT, 2, W, blocks_per_sm);
  PreArgs A;
  if ((rc = pack_codes<QT>(c, Q, g.nqt)) || (rc = prefilter_args(c, Q, g, &A))) return rc;
  prefilter_logtable_kernel<QT, W, MINB><<<dim3(g.gx, g.gy, g.zs), W * 32, 0, c->stream>>>(A, c->LT.p); LAUNCH_CHECK(c);
  c->prefilter_kernel_used = 6;
  return prefilter_finish(c, Q, g);
}

// r5a: C2 10.0 -> 3.8 ms, C4 122 -> 44.6 ms per 10k queries against the product-table form: the default for >= 256 queries (PPB_PF_QLANE=0 restores the table form)
static const bool kQlaneDefault = [] { const char *e = std::getenv("PPB_PF_QLANE"); return !e || std::atoi(e) != 0; }();
// query-lane log-table form (prefilter_qlane.cuh): lane = query, warp = 32 queries x 32 edges, TMA ring over the table tile
template <int W, int SPS, int NS, int MINB>
static int launch_prefilter_qlane(pp_ctx *c, int64_t Q, double pend, int blocks_per_sm) {
  int rc;
  using G = QlaneGeom<W, SPS, NS>;
  SnT sn; sn.a = c->snT_a; sn.e = c->snT_e; sn.Np = c->Np; sn.L = c->L; sn.K = c->K; sn.n_nodes = c->n_nodes;
  if (c->ltq_pend != pend) {  // once per (reference, pendant length)
    if ((rc = c->LTq.ensure((size_t)c->L * c->Np * kQlCols))) return rc;
    for (int s0 = 0; s0 < c->L; s0 += 65535) {
      qlane_table_kernel<16><<<dim3((unsigned)((c->Np + 127) / 128), std::min(65535, c->L - s0)), 128, 0, c->stream>>>(sn, c->tq5, c->LTq.p, s0); LAUNCH_CHECK(c);
    }
    c->ltq_pend = pend;
  }
  const PreGeom g = prefilter_geometry(c, Q, 32, 1, W, blocks_per_sm);
  const int n4 = (c->L + 3) / 4;
  if ((rc = c->tile_lo.ensure(g.nqt)) || (rc = c->tile_hi.ensure(g.nqt)) || (rc = c->qcol.ensure((size_t)g.nqt * n4 * 32)) || (rc = c->slow_flag.ensure(Q))) return rc;
  CK(cudaMemsetAsync(c->slow_flag.p, 0, sizeof(int32_t) * Q, c->stream));
  tile_range_kernel<32><<<(g.nqt + 127) / 128, 128, 0, c->stream>>>(c->qsite.p, c->nq.p, c->qorder.p, (int)Q, g.nqt, c->Lp, c->tile_lo.p, c->tile_hi.p); LAUNCH_CHECK(c);
  qlane_pack_kernel<<<dim3((n4 + 7) / 8, g.nqt), 256, 0, c->stream>>>(c->codes.p, c->qorder.p, c->qcol.p, (int)Q, c->L, c->Lp, g.nqt, c->slow_flag.p); LAUNCH_CHECK(c);
  PreArgs A;
  if ((rc = prefilter_args(c, Q, g, &A))) return rc;
  const size_t sm = G::smem_bytes();
  auto kern = prefilter_qlane_kernel<W, SPS, NS, MINB>;
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
  kern<<<dim3(g.gx, g.gy, g.zs), (W + 1) * 32, sm, c->stream>>>(A, c->LTq.p, c->qcol.p); LAUNCH_CHECK(c);
  c->prefilter_kernel_used = 7;
  if ((rc = prefilter_finish(c, Q, g))) return rc;
  A.out = c->scores.p;
  prefilter_qlane_fix_kernel<<<dim3((unsigned)Q, (c->n_nodes + 127) / 128), 128, 0, c->stream>>>(A, c->qsite.p, c->Lp, c->slow_flag.p, c->scores.p); LAUNCH_CHECK(c);
  return PP_OK;
}

// pair-of-sites table (prefilter_qlane.cuh): the query-lane kernel with NC = 36 walking pairs of sites
template <int W, int SPS, int NS, int MINB>
static int launch_prefilter_qlane_pairs(pp_ctx *c, int64_t Q, double pend, int blocks_per_sm) {
  int rc;
  constexpr int NC = kQlPairCols;
  using G = QlaneGeom<W, SPS, NS, NC>;
  const int L2 = (c->L + 1) / 2;
  SnT sn; sn.a = c->snT_a; sn.e = c->snT_e; sn.Np = c->Np; sn.L = c->L; sn.K = c->K; sn.n_nodes = c->n_nodes;
  if (c->ltq2_pend != pend) {  // once per (reference, pendant length)
    if ((rc = c->LTq2.ensure((size_t)L2 * c->Np * NC))) return rc;
    for (int p0 = 0; p0 < L2; p0 += 65535) {
      qlane_pair_table_kernel<16><<<dim3((unsigned)((c->Np + 127) / 128), std::min(65535, L2 - p0)), 128, 0, c->stream>>>(sn, c->tq5, c->LTq2.p, p0); LAUNCH_CHECK(c);
    }
    c->ltq2_pend = pend;
  }
  PreGeom g = prefilter_geometry(c, Q, 32, 1, W, blocks_per_sm);
  g.sites_per_z = (L2 + g.zs - 1) / g.zs;             // the kernel's "sites" are pairs
  g.zs = (L2 + g.sites_per_z - 1) / g.sites_per_z;
  const int n4 = (L2 + 3) / 4;
  if ((rc = c->tile_lo.ensure(g.nqt)) || (rc = c->tile_hi.ensure(g.nqt)) || (rc = c->tile_lo2.ensure(g.nqt)) || (rc = c->tile_hi2.ensure(g.nqt)) ||
      (rc = c->qcol.ensure((size_t)g.nqt * n4 * 32)) || (rc = c->slow_flag.ensure(Q)))
    return rc;
  CK(cudaMemsetAsync(c->slow_flag.p, 0, sizeof(int32_t) * Q, c->stream));
  tile_range_kernel<32><<<(g.nqt + 127) / 128, 128, 0, c->stream>>>(c->qsite.p, c->nq.p, c->qorder.p, (int)Q, g.nqt, c->Lp, c->tile_lo.p, c->tile_hi.p); LAUNCH_CHECK(c);
  qlane_pair_ranges_kernel<<<(g.nqt + 127) / 128, 128, 0, c->stream>>>(c->tile_lo.p, c->tile_hi.p, g.nqt, c->tile_lo2.p, c->tile_hi2.p); LAUNCH_CHECK(c);
  qlane_pair_pack_kernel<<<dim3((n4 + 7) / 8, g.nqt), 256, 0, c->stream>>>(c->codes.p, c->qorder.p, c->qcol.p, (int)Q, c->L, c->Lp, g.nqt, c->slow_flag.p); LAUNCH_CHECK(c);
  PreArgs A;
  if ((rc = prefilter_args(c, Q, g, &A))) return rc;
  PreArgs A2 = A;                                       // the pair view the main kernel walks
  A2.sn.L = L2; A2.tile_lo = c->tile_lo2.p; A2.tile_hi = c->tile_hi2.p;
  const size_t sm = G::smem_bytes();
  auto kern = prefilter_qlane_kernel<W, SPS, NS, MINB, NC, false>;
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
  kern<<<dim3(g.gx, g.gy, g.zs), (W + 1) * 32, sm, c->stream>>>(A2, c->LTq2.p, c->qcol.p); LAUNCH_CHECK(c);
  c->prefilter_kernel_used = 8;
  if ((rc = prefilter_finish(c, Q, g))) return rc;
  A.out = c->scores.p;
  prefilter_qlane_fix_kernel<<<dim3((unsigned)Q, (c->n_nodes + 127) / 128), 128, 0, c->stream>>>(A, c->qsite.p, c->Lp, c->slow_flag.p, c->scores.p); LAUNCH_CHECK(c);
  return PP_OK;
}

// the same form for the 20-state models (Gamma mixtures and Price-CAT): 22 columns (prefilter_qlane.cuh, second half)
template <int S, bool G, int W, int SPS, int NS, int MINB>
static int launch_prefilter_qlane_generic(pp_ctx *c, int64_t Q, double pend, int blocks_per_sm) {
  int rc;
  constexpr int NC = S + 2;
  using Geo = QlaneGeom<W, SPS, NS, NC>;
  QlColMap map;
  std::memset(&map, 0, sizeof map);
  for (int code = 0; code < kMaxCodesDev; ++code) {
    const uint32_t m = code < c->al.n_codes ? c->al.mask_of_code[code] : 0u;
    map.col[code] = 255;
    if (m && !(m & (m - 1))) { int st = 0; while (!((m >> st) & 1u)) ++st; map.col[code] = (uint8_t)st; map.code_of_state[st] = (uint8_t)code; }
    else if (m == (S == 32 ? 0xffffffffu : (1u << S) - 1u)) { map.col[code] = (uint8_t)S; map.code_of_state[S] = (uint8_t)code; }  // X / N
  }
  SnT sn; sn.a = c->snT_a; sn.e = c->snT_e; sn.Np = c->Np; sn.L = c->L; sn.K = c->K; sn.n_nodes = c->n_nodes;
  if (c->ltq_pend != pend) {  // once per (reference, pendant length): 19.7 GB for ar53
    if ((rc = c->LTq.ensure((size_t)c->L * c->Np * NC))) return rc;
    for (int s0 = 0; s0 < c->L; s0 += 65535) {
      qlane_table_generic_kernel<S, G, NC><<<dim3((unsigned)((c->Np + 127) / 128), std::min(65535, c->L - s0)), 128, 0, c->stream>>>(
          sn, c->Tq.p, c->d_site_cat, c->al.n_codes, c->Rg, map, c->LTq.p, s0); LAUNCH_CHECK(c);
    }
    c->ltq_pend = pend;
  }
  const PreGeom g = prefilter_geometry(c, Q, 32, 1, W, blocks_per_sm);
  const int n4 = (c->L + 3) / 4;
  if ((rc = c->tile_lo.ensure(g.nqt)) || (rc = c->tile_hi.ensure(g.nqt)) || (rc = c->qcol.ensure((size_t)g.nqt * n4 * 32)) || (rc = c->slow_flag.ensure(Q))) return rc;
  CK(cudaMemsetAsync(c->slow_flag.p, 0, sizeof(int32_t) * Q, c->stream));
  tile_range_kernel<32><<<(g.nqt + 127) / 128, 128, 0, c->stream>>>(c->qsite.p, c->nq.p, c->qorder.p, (int)Q, g.nqt, c->Lp, c->tile_lo.p, c->tile_hi.p); LAUNCH_CHECK(c);
  qlane_pack_generic_kernel<<<dim3((n4 + 7) / 8, g.nqt), 256, 0, c->stream>>>(c->codes.p, c->qorder.p, c->qcol.p, (int)Q, c->L, c->Lp, g.nqt, map, (unsigned)(S + 1), c->slow_flag.p); LAUNCH_CHECK(c);
  PreArgs A;
  if ((rc = prefilter_args(c, Q, g, &A))) return rc;
  const size_t sm = Geo::smem_bytes();
  auto kern = prefilter_qlane_kernel<W, SPS, NS, MINB, NC, G>;
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
  kern<<<dim3(g.gx, g.gy, g.zs), (W + 1) * 32, sm, c->stream>>>(A, c->LTq.p, c->qcol.p); LAUNCH_CHECK(c);
  c->prefilter_kernel_used = 7;
  if ((rc = prefilter_finish(c, Q, g))) return rc;
  A.out = c->scores.p;
  prefilter_qlane_fix_generic_kernel<S, G><<<dim3((unsigned)Q, (c->n_nodes + 127) / 128), 128, 0, c->stream>>>(A, c->qsite.p, c->Lp, map, c->Rg, c->slow_flag.p, c->scores.p); LAUNCH_CHECK(c);
  return PP_OK;
}

// staged form (prefilter_staged.cuh): TMA bulk copies + mbarrier ring; MMA = the K-contraction on the FP64 tensor cores
template <int W, int SPS, int NS, bool MMA, int MINB>
static int launch_prefilter_staged(pp_ctx *c, int64_t Q, int blocks_per_sm) {
  int rc;
  using G = StagedGeom<W, SPS, NS>;
  const PreGeom g = prefilter_geometry(c, Q, G::QT, 2, W, blocks_per_sm);
  PreArgs A;
  if ((rc = pack_codes<G::QT>(c, Q, g.nqt)) || (rc = prefilter_args(c, Q, g, &A))) return rc;
  const size_t sm = G::smem_bytes(c->al.n_codes);
  auto kern = prefilter_staged_kernel<W, SPS, NS, MMA, MINB>;
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
  kern<<<dim3(g.gx, g.gy, g.zs), (W + 1) * 32, sm, c->stream>>>(A); LAUNCH_CHECK(c);
  c->prefilter_kernel_used = MMA ? 5 : 4;
  return prefilter_finish(c, Q, g);
}

// kernel_choice: 0 default (table form for DNA Gamma-4, direct otherwise), 1 direct FMA form, 2 table form,
// 4 staged FMA form, 5 staged DMMA form (DNA Gamma-4)
static int run_prefilter(pp_ctx *c, int64_t Q, int kernel_choice, double pend) {
  int rc;
  if ((rc = c->scores.ensure((size_t)Q * c->n_nodes))) return rc;
  const bool few = Q <= 64;  // one warp per block when there are too few query tiles to fill a block
  const bool const_ok = (size_t)c->al.n_codes * c->K <= (size_t)kConstTqDoubles;
  if (c->S == 4 && !c->gcat && c->K == 16) {
    // direct form: 8 queries per warp x 4 warps (157 registers, 3 blocks/SM) measured 35.9 ms against 42.9 ms for 16 x 8 (r2q)
    if (kernel_choice == 1) return few ? launch_prefilter_direct<4, false, 4, 2, 8, 1, TQ_CONST>(c, Q, 8) : launch_prefilter_direct<4, false, 4, 2, 8, 4, TQ_CONST>(c, Q, 3);
    if (kernel_choice == 3) return launch_prefilter_direct<4, false, 4, 2, 16, 8, TQ_SMEM>(c, Q, 1);
    if (kernel_choice == 4 || kernel_choice == 5) {
      static const int sv = [] { const char *e = std::getenv("PPB_PFS_VARIANT"); return e ? std::atoi(e) : 0; }();
      if (kernel_choice == 4) {
        if (sv == 1) return launch_prefilter_staged<8, 4, 4, false, 1>(c, Q, 1);
        if (sv == 2) return launch_prefilter_staged<4, 2, 4, false, 2>(c, Q, 2);
        return launch_prefilter_staged<4, 4, 3, false, 2>(c, Q, 2);
      }
      if (sv == 1) return launch_prefilter_staged<8, 4, 3, true, 2>(c, Q, 2);
      if (sv == 2) return launch_prefilter_staged<4, 4, 3, true, 2>(c, Q, 2);
      if (sv == 3) return launch_prefilter_staged<8, 4, 6, true, 1>(c, Q, 1);
      return launch_prefilter_staged<4, 2, 4, true, 3>(c, Q, 3);
    }
    // Geometry of the table form (measured, profiles/README.md r2k): the kernel is latency-bound, so FEW
    // queries per warp and many resident warps win over register-level reuse of the table entries
    // (16 queries per warp, 3 blocks/SM: 18.1 ms; 4 queries per warp, 8 blocks/SM: 13.3 ms for C2).
    static const int geom = [] { const char *e = std::getenv("PPB_PF_GEOM"); return e ? std::atoi(e) : 0; }();
    // log-table form: one load + one add per (query, edge, site); measured (r4f) 11.8 ms against 9.9 ms for the product
    // table on C2 - it executes half the instructions but one L1 request per (query, site) instead of 5 per site
    // (ncu r4g: l1tex 77 %, long_scoreboard 6.4 warps per issue) - so it is selectable (6), not the default
    // r5y: the pair-of-sites table is the default for large batches (C2 3.84 -> 2.90 ms, C4 42.7 -> 30.0 ms per 10k queries); PPB_PF_PAIRS=0
    // goes back to one word per site (form 7), which is also what an out-of-memory table allocation falls back to
    static const bool pairs_default = [] { const char *e = std::getenv("PPB_PF_PAIRS"); return !e || std::atoi(e) != 0; }();
    if (kernel_choice == 8 || (kernel_choice == 0 && pairs_default && Q >= 256)) {
      rc = launch_prefilter_qlane_pairs<8, 4, 4, 1>(c, Q, pend, 1);
      if (!(rc == PP_ENOMEM && kernel_choice == 0)) return rc;
    }
    if (kernel_choice == 7 || (kernel_choice == 0 && kQlaneDefault && Q >= 256)) {
      static const int qv = [] { const char *e = std::getenv("PPB_PFQ_VARIANT"); return e ? std::atoi(e) : 0; }();
      if (qv == 1) return launch_prefilter_qlane<4, 16, 4, 2>(c, Q, pend, 2);
      if (qv == 2) return launch_prefilter_qlane<8, 8, 6, 1>(c, Q, pend, 1);
      if (qv == 3) return launch_prefilter_qlane<2, 8, 6, 3>(c, Q, pend, 3);
      return launch_prefilter_qlane<4, 8, 6, 2>(c, Q, pend, 2);
    }
    if (kernel_choice == 6) {
      if (few) return launch_prefilter_logtable<8, 1, 1>(c, Q, pend, 12);
      if (geom == 1) return launch_prefilter_logtable<4, 4, 8>(c, Q, pend, 8);
      if (geom == 2) return launch_prefilter_logtable<16, 4, 4>(c, Q, pend, 4);
      if (geom == 3) return launch_prefilter_logtable<8, 8, 2>(c, Q, pend, 4);
      return launch_prefilter_logtable<8, 4, 6>(c, Q, pend, 8);
    }
    if (!few && geom == 1) return launch_prefilter_table<2, 16, 4, 3>(c, Q, pend, 3);
    if (!few && geom == 2) return launch_prefilter_table<2, 8, 4, 6>(c, Q, pend, 6);
    if (!few && geom == 3) return launch_prefilter_table<2, 4, 8, 4>(c, Q, pend, 4);
    return few ? launch_prefilter_table<2, 8, 1, 1>(c, Q, pend, 12) : launch_prefilter_table<2, 4, 4, 8>(c, Q, pend, 8);
  }
  if (kernel_choice == 2) return fail(PP_EINVAL, "prefilter: the table form exists for the DNA Gamma-4 model only");
  if (c->S == 4 && (kernel_choice == 7 || (kernel_choice == 0 && kQlaneDefault && Q >= 256))) {
    // DNA with Price-CAT (FastTree reference trees) or a Gamma mixture other than four rates: six-column generic table
    rc = c->gcat ? launch_prefilter_qlane_generic<4, true, 4, 8, 6, 2>(c, Q, pend, 2) : launch_prefilter_qlane_generic<4, false, 4, 8, 6, 2>(c, Q, pend, 2);
    if (!(rc == PP_ENOMEM && kernel_choice == 0)) return rc;
  }
  if (c->S == 20 && (kernel_choice == 7 || (kernel_choice == 0 && kQlaneDefault && Q >= 64))) {
    // proteins: 22-column query-lane log table; from one full warp of queries on it beats the direct form, which has to
    // run S (Price-CAT) or R*S DFMAs per unit (C3, 256 genomes: 101 ms direct)
    rc = c->gcat ? launch_prefilter_qlane_generic<20, true, 8, 4, 4, 1>(c, Q, pend, 1) : launch_prefilter_qlane_generic<20, false, 8, 4, 4, 1>(c, Q, pend, 1);
    if (!(rc == PP_ENOMEM && kernel_choice == 0)) return rc;  // no room for the table: the direct form below needs none
  }
  if (c->gcat) {
    if (c->S == 4) return few ? launch_prefilter_direct<4, true, 1, 2, 8, 1, TQ_GLOBAL>(c, Q, 12) : launch_prefilter_direct<4, true, 1, 2, 16, 8, TQ_GLOBAL>(c, Q, 2);
    return few ? launch_prefilter_direct<20, true, 1, 2, 4, 1, TQ_GLOBAL>(c, Q, 12) : launch_prefilter_direct<20, true, 1, 2, 8, 4, TQ_GLOBAL>(c, Q, 3);
  }
  if (c->S == 4) {
    if (!const_ok) return launch_prefilter_direct<4, false, 0, 2, 8, 1, TQ_SMEM>(c, Q, 8);
    return few ? launch_prefilter_direct<4, false, 0, 2, 8, 1, TQ_CONST>(c, Q, 8) : launch_prefilter_direct<4, false, 0, 2, 8, 8, TQ_CONST>(c, Q, 1);
  }
  if (!const_ok) return launch_prefilter_direct<20, false, 0, 2, 4, 1, TQ_SMEM>(c, Q, 8);
  return few ? launch_prefilter_direct<20, false, 0, 2, 4, 1, TQ_CONST>(c, Q, 8) : launch_prefilter_direct<20, false, 0, 2, 8, 4, TQ_CONST>(c, Q, 2);
}

extern "C" int pp_prefilter_scores(pp_ctx *c, double start_pend, int64_t Q, const uint8_t *query_chars, double *scores,
                                   int32_t kernel) {
  if (!c || !query_chars || !scores || Q < 1) return fail(PP_EINVAL, "pp_prefilter_scores: bad arguments");
  if (!c->precomputed) return fail(PP_ESTATE, "pp_prefilter_scores: call pp_precompute first");
  CK(cudaSetDevice(c->device));
  int rc;
  if ((rc = c->chars.ensure((size_t)Q * c->L))) return rc;
  CK(cudaMemcpyAsync(c->chars.p, query_chars, (size_t)Q * c->L, cudaMemcpyHostToDevice, c->stream));
  if ((rc = prep_queries(c, Q, c->chars.p)) || (rc = build_query_table(c, start_pend)) || (rc = run_prefilter(c, Q, kernel, start_pend))) return rc;
  CK(cudaMemcpyAsync(scores, c->scores.p, sizeof(double) * (size_t)Q * c->n_nodes, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  return PP_OK;
}

extern "C" int pp_prefilter_time(pp_ctx *c, double start_pend, int64_t Q, const uint8_t *query_chars, int32_t kernel, int32_t reps,
                                 double *ms_per_launch) {
  if (!c || !query_chars || !ms_per_launch || Q < 1 || reps < 1) return fail(PP_EINVAL, "pp_prefilter_time: bad arguments");
  if (!c->precomputed) return fail(PP_ESTATE, "pp_prefilter_time: call pp_precompute first");
  CK(cudaSetDevice(c->device));
  int rc;
  if ((rc = c->chars.ensure((size_t)Q * c->L))) return rc;
  CK(cudaMemcpyAsync(c->chars.p, query_chars, (size_t)Q * c->L, cudaMemcpyHostToDevice, c->stream));
  if ((rc = prep_queries(c, Q, c->chars.p)) || (rc = build_query_table(c, start_pend)) || (rc = run_prefilter(c, Q, kernel, start_pend))) return rc;
  CK(cudaEventRecord(c->ev[0], c->stream));
  for (int i = 0; i < reps; ++i)
    if ((rc = run_prefilter(c, Q, kernel, start_pend))) return rc;
  CK(cudaEventRecord(c->ev[1], c->stream));
  CK(cudaEventSynchronize(c->ev[1]));
  float ms = 0;
  CK(cudaEventElapsedTime(&ms, c->ev[0], c->ev[1]));
  *ms_per_launch = (double)ms / reps;
  return PP_OK;
}

// ---------------------------------------------------------------------------------------------
// three-taxon launches
// ---------------------------------------------------------------------------------------------
static int launch_three_tax_warp(pp_ctx *c, const CandIn *in, CandOut *out, int n, const TTParams &prm) {
  int rc;
  // queue order: candidates sorted by edge so that concurrently running warps share darr/parr planes in L2
  const int32_t *order = nullptr;
  if (n > 1) {
    if ((rc = c->ord_keys.ensure(n)) || (rc = c->ord_vals.ensure(n)) || (rc = c->ord_keys2.ensure(n)) || (rc = c->ord_vals2.ensure(n))) return rc;
    int loc_bits = 1;
    while ((1 << loc_bits) < c->n_nodes && loc_bits < 28) ++loc_bits;
    const int end_bit = loc_bits + 2;
    cand_keys_kernel<<<(n + 255) / 256, 256, 0, c->stream>>>(in, c->nq.p, n, c->L, loc_bits, c->ord_keys.p, c->ord_vals.p); LAUNCH_CHECK(c);
    size_t tmp = 0;
    CK(cub::DeviceRadixSort::SortPairs(nullptr, tmp, c->ord_keys.p, c->ord_keys2.p, c->ord_vals.p, c->ord_vals2.p, n, 0, end_bit, c->stream));
    if ((rc = c->ord_tmp.ensure(tmp))) return rc;
    CK(cub::DeviceRadixSort::SortPairs(c->ord_tmp.p, tmp, c->ord_keys.p, c->ord_keys2.p, c->ord_vals.p, c->ord_vals2.p, n, 0, end_bit, c->stream));
    c->launches += 1;
    order = c->ord_vals2.p;
  }
  // two builds of the kernel: <=128 registers (two blocks = 16 candidates per SM) and <=80 registers
  // (three blocks = 24 candidates per SM); ttw_blocks_per_sm picks (default kTTWarpDefaultBlocks)
  const int per_sm = c->ttw_blocks_per_sm > 0 ? std::min(c->ttw_blocks_per_sm, 3) : kTTWarpDefaultBlocks;
  if (c->tt_engine != 2 && prm.mode == TT_MODE_MARGINAL && c->ttg_pp) {  // --pp on the block engine
    const int gblocks = std::min(4 * c->sm_count, n);
    const size_t smg = TTGroup<4, 4, 1, 256>::smem_bytes();
    CK(cudaFuncSetAttribute(tt_group_pp_kernel<4, 4, 4, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smg));
    tt_group_pp_kernel<4, 4, 4, 256><<<gblocks, 128, smg, c->stream>>>(c->dm, c->g, c->d_bl, c->qsite.p, c->nq.p, in, order, out, n, prm,
                                                                       c->wbuf.p, c->counters.p);
    LAUNCH_CHECK(c);
    return PP_OK;
  }
  if ((c->tt_engine == 0 || c->tt_engine == 4) && c->eig_ok && prm.mode != TT_MODE_MARGINAL) {  // engine 4: pendant objective in the eigenbasis
#define PPB_LAUNCH_EIGEN(NW_, MINB_, WS_)                                                                                         \
  do {                                                                                                                            \
    const size_t sme = TTEig<4, NW_, WS_>::smem_bytes();                                                                          \
    static const int cap = [] { const char *e = std::getenv("PPB_TTE_BLOCKS"); return e ? std::atoi(e) : 0; }();                     \
    const int eblocks = std::min((cap > 0 ? std::min(cap, MINB_) : MINB_) * c->sm_count, n);                                         \
    CK(cudaFuncSetAttribute(tt_eigen_kernel<4, NW_, MINB_, WS_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sme));           \
    tt_eigen_kernel<4, NW_, MINB_, WS_><<<eblocks, NW_ * 32, sme, c->stream>>>(c->em4, c->g, c->d_bl, c->qsite.p, c->nq.p, in, order, out, n, \
                                                                           prm, c->wbuf.p, c->counters.p, c->tt_work.p);       \
  } while (0)
    // r4j: 2 warps x 8 blocks/SM (160 sites of y in shared memory) 85.3 ms, 4 x 5 (320 sites) 73.1, 2 x 6 (256) 102.1 - the default stays 4 x 4 (448): 67.2
    if (c->tte_variant == 1) PPB_LAUNCH_EIGEN(4, 3, 608);
    else if (c->tte_variant == 2) PPB_LAUNCH_EIGEN(8, 2, 960);
    else if (c->tte_variant == 3) PPB_LAUNCH_EIGEN(4, 2, 960);
    else PPB_LAUNCH_EIGEN(4, 4, 448);
#undef PPB_LAUNCH_EIGEN
    LAUNCH_CHECK(c);
    return PP_OK;
  }
  if (c->tt_engine != 2 && prm.mode != TT_MODE_MARGINAL) {  // block per candidate, w in shared memory
    const int gblocks = std::min(c->ttg_blocks, n);
    // the shared-memory attribute is per function, not per context: set it at every launch
#define PPB_LAUNCH_GROUP(G_, MINB_, TS_, WS_, RU_)                                                                              \
  do {                                                                                                                          \
    const size_t smg = TTGroup<4, G_, TS_, WS_, RU_>::smem_bytes();                                                             \
    CK(cudaFuncSetAttribute(tt_group_kernel<4, G_, MINB_, TS_, WS_, RU_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smg)); \
    tt_group_kernel<4, G_, MINB_, TS_, WS_, RU_><<<gblocks, G_ * 32, smg, c->stream>>>(c->dm, c->g, c->d_bl, c->qsite.p, c->nq.p, in, order, out, \
                                                                                      n, prm, c->wbuf.p, c->counters.p, c->tt_work.p); \
  } while (0)
    PPB_LAUNCH_GROUP(4, 4, 2, 256, false);  // the r2 geometry sweep (profiles/README.md) settled on 4 warps x 4 blocks/SM, two sites per thread
#undef PPB_LAUNCH_GROUP
    LAUNCH_CHECK(c);
    return PP_OK;
  }
  int blocks = std::min(c->ttw_max_blocks, per_sm * c->sm_count);
  blocks = std::min(blocks, (n + kTTWarpsPerBlock - 1) / kTTWarpsPerBlock);
  const size_t smw = TTWarp<4, 1>::smem_bytes(kTTWarpsPerBlock);
  if (per_sm >= 3)
    tt_warp_kernel<4, kTTWarpsPerBlock, 3, 1><<<blocks, kTTWarpsPerBlock * 32, smw, c->stream>>>(
        c->dm, c->g, c->d_bl, c->codes.p, c->qsite.p, c->nq.p, in, order, out, n, prm, c->wbuf.p, c->counters.p, c->tt_work.p);
  else
    tt_warp_kernel<4, kTTWarpsPerBlock, 2, 2><<<blocks, kTTWarpsPerBlock * 32, smw, c->stream>>>(
        c->dm, c->g, c->d_bl, c->codes.p, c->qsite.p, c->nq.p, in, order, out, n, prm, c->wbuf.p, c->counters.p, c->tt_work.p);
  LAUNCH_CHECK(c);
  return PP_OK;
}

static int launch_three_tax(pp_ctx *c, const CandIn *in, CandOut *out, int n, const TTParams &prm) {
  if (n <= 0) return PP_OK;
  CK(cudaMemsetAsync(c->counters.p, 0, sizeof(int), c->stream));
  if (ttw_supported(c) && c->tt_engine != 1) return launch_three_tax_warp(c, in, out, n, prm);
  if (c->tt_engine == 2) return fail(PP_EINVAL, "three-taxon warp engine supports the four-state Gamma-4 mixture only");
  const int blocks = std::min(c->tt_blocks, n);
  DISPATCH_MODEL(c, {
    constexpr int TH = (S_ == 4) ? kTTThreadsDNA : kTTThreadsAA;
    three_tax_kernel<S_, G_, TH><<<blocks, TH, c->tt_smem, c->stream>>>(c->dm, c->g, c->d_bl, c->d_site_cat, c->codes.p, G_ ? c->qsite_cat.p : c->qsite.p,
                                                                        c->nq.p, in, out, n, prm, c->wbuf.p, c->counters.p, c->nqc.p, c->LpC);
    LAUNCH_CHECK(c);
  });
  return PP_OK;
}

static TTParams make_params(const pp_prefs *pf, int mode, double tol, int use_defaults) {
  TTParams p;
  p.start_pend = pf->start_pend; p.max_pend = pf->max_pend; p.tol = tol; p.pp_rel_err = pf->pp_rel_err;
  p.prior_mean = pf->prior_mean; p.prior_means = nullptr; p.max_iter = pf->max_iter; p.mode = mode;
  p.prior_kind = pf->prior_kind; p.use_defaults = use_defaults;
  static const int env_flags = [] { const char *e = std::getenv("PPB_TT_FLAGS"); return e ? std::atoi(e) : kTTWarpDefaultFlags; }();
  p.flags = env_flags;
  return p;
}

extern "C" int pp_three_tax_log_like(pp_ctx *c, int64_t n, const uint8_t *query_chars, const int32_t *edge,
                                     const double *pendant, const double *distal, double *ll_out) {
  if (!c || n < 1 || !query_chars || !edge || !pendant || !distal || !ll_out) return fail(PP_EINVAL, "pp_three_tax_log_like: bad arguments");
  if (!c->precomputed) return fail(PP_ESTATE, "pp_three_tax_log_like: call pp_precompute first");
  CK(cudaSetDevice(c->device));
  int rc;
  if ((rc = c->chars.ensure((size_t)n * c->L)) || (rc = c->cand.ensure(n)) || (rc = c->cand_out.ensure(n))) return rc;
  CK(cudaMemcpyAsync(c->chars.p, query_chars, (size_t)n * c->L, cudaMemcpyHostToDevice, c->stream));
  if ((rc = prep_queries(c, n, c->chars.p))) return rc;
  std::vector<CandIn> in(n);
  for (int64_t i = 0; i < n; ++i) {
    if (edge[i] < 0 || edge[i] >= c->n_nodes) return fail(PP_EINVAL, "pp_three_tax_log_like: edge out of range");
    in[i].q = (int32_t)i; in[i].loc = edge[i]; in[i].pend0 = pendant[i]; in[i].dist0 = distal[i];
  }
  CK(cudaMemcpyAsync(c->cand.p, in.data(), n * sizeof(CandIn), cudaMemcpyHostToDevice, c->stream));
  pp_prefs pf; pp_default_prefs(&pf);
  if ((rc = launch_three_tax(c, c->cand.p, c->cand_out.p, (int)n, make_params(&pf, TT_MODE_LOGLIKE, 0.0, 0)))) return rc;
  std::vector<CandOut> o(n);
  CK(cudaMemcpyAsync(o.data(), c->cand_out.p, n * sizeof(CandOut), cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  std::vector<int32_t> st(n);
  CK(cudaMemcpy(st.data(), c->prep_status.p, n * sizeof(int32_t), cudaMemcpyDeviceToHost));
  for (int64_t i = 0; i < n; ++i) ll_out[i] = st[i] ? NAN : o[i].ll;
  return PP_OK;
}

// ---------------------------------------------------------------------------------------------
// figs (fig.ml): ranking by fig representatives, groups expanded lazily per baseball wave (fig.cuh)
// ---------------------------------------------------------------------------------------------
extern "C" int pp_set_figs(pp_ctx *c, int32_t n_figs, const int32_t *rep, const int32_t *ptr, const int32_t *idx) {
  if (!c || n_figs < 0) return fail(PP_EINVAL, "pp_set_figs: bad arguments");
  CK(cudaSetDevice(c->device));
  c->fig_rep.clear(); c->fig_ptr.clear(); c->fig_idx.clear();
  c->rep_ready = false; c->rep_mt_pend = c->rep_lt_pend = c->rep_ltq_pend = c->rep_ltq2_pend = -1.0;
  if (n_figs == 0) return PP_OK;  // Fig.Dummy: the plain ranking of every non-root edge
  if (!rep || !ptr || !idx || ptr[0] != 0) return fail(PP_EINVAL, "pp_set_figs: bad arguments");
  for (int f = 0; f < n_figs; ++f) {
    if (rep[f] < 0 || rep[f] >= c->n_nodes || rep[f] == c->root_id || ptr[f + 1] < ptr[f]) return fail(PP_EINVAL, "pp_set_figs: bad representative or edge list");
  }
  for (int k = 0; k < ptr[n_figs]; ++k)
    if (idx[k] < 0 || idx[k] >= c->n_nodes || idx[k] == c->root_id) return fail(PP_EINVAL, "pp_set_figs: edge out of range");
  c->fig_rep.assign(rep, rep + n_figs); c->fig_ptr.assign(ptr, ptr + n_figs + 1); c->fig_idx.assign(idx, idx + ptr[n_figs]);
  int rc;
  if ((rc = c->d_fig_rep.ensure(n_figs)) || (rc = c->d_fig_ptr.ensure(n_figs + 1)) || (rc = c->d_fig_idx.ensure(std::max(ptr[n_figs], 1)))) return rc;
  CK(cudaMemcpyAsync(c->d_fig_rep.p, rep, sizeof(int32_t) * n_figs, cudaMemcpyHostToDevice, c->stream));
  CK(cudaMemcpyAsync(c->d_fig_ptr.p, ptr, sizeof(int32_t) * (n_figs + 1), cudaMemcpyHostToDevice, c->stream));
  CK(cudaMemcpyAsync(c->d_fig_idx.p, idx, sizeof(int32_t) * ptr[n_figs], cudaMemcpyHostToDevice, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  return PP_OK;
}
extern "C" int pp_n_figs(pp_ctx *c) { return c ? (int)c->fig_rep.size() : -1; }

// the prefilter over the fig representatives only: the same kernels on a compact copy of the representatives' columns
static int run_prefilter_reps(pp_ctx *c, int64_t Q, int kernel_choice, double pend) {
  int rc;
  const int m = (int)c->fig_rep.size();
  if (!c->rep_ready) {
    c->rep_Np = ((int64_t)m + 63) / 64 * 64;
    if ((rc = c->rep_a.ensure((size_t)c->L * c->K * c->rep_Np)) || (rc = c->rep_e.ensure((size_t)c->L * c->rep_Np))) return rc;
    gather_rep_columns_kernel<double><<<c->sm_count * 8, 256, 0, c->stream>>>(c->snT_a, c->rep_a.p, c->d_fig_rep.p, m, (int64_t)c->L * c->K, c->Np, c->rep_Np, 1.0); LAUNCH_CHECK(c);
    gather_rep_columns_kernel<int32_t><<<c->sm_count * 8, 256, 0, c->stream>>>(c->snT_e, c->rep_e.p, c->d_fig_rep.p, m, (int64_t)c->L, c->Np, c->rep_Np, 0); LAUNCH_CHECK(c);
    c->rep_ready = true;
  }
  // view swap: the prefilter launchers read the supernode arrays, the table caches and the score buffer from the context
  std::swap(c->snT_a, c->rep_a.p); std::swap(c->snT_e, c->rep_e.p); std::swap(c->Np, c->rep_Np);
  std::swap(c->scores, c->rep_scores); std::swap(c->MT, c->rep_MT); std::swap(c->MTe, c->rep_MTe); std::swap(c->LT, c->rep_LT); std::swap(c->LTq, c->rep_LTq); std::swap(c->LTq2, c->rep_LTq2);
  std::swap(c->mt_pend, c->rep_mt_pend); std::swap(c->lt_pend, c->rep_lt_pend); std::swap(c->ltq_pend, c->rep_ltq_pend); std::swap(c->ltq2_pend, c->rep_ltq2_pend);
  const int nn = c->n_nodes, root = c->root_id;
  c->n_nodes = m; c->root_id = -1;
  rc = run_prefilter(c, Q, kernel_choice, pend);
  c->n_nodes = nn; c->root_id = root;
  std::swap(c->snT_a, c->rep_a.p); std::swap(c->snT_e, c->rep_e.p); std::swap(c->Np, c->rep_Np);
  std::swap(c->scores, c->rep_scores); std::swap(c->MT, c->rep_MT); std::swap(c->MTe, c->rep_MTe); std::swap(c->LT, c->rep_LT); std::swap(c->LTq, c->rep_LTq); std::swap(c->LTq2, c->rep_LTq2);
  std::swap(c->mt_pend, c->rep_mt_pend); std::swap(c->lt_pend, c->rep_lt_pend); std::swap(c->ltq_pend, c->rep_ltq_pend); std::swap(c->ltq2_pend, c->rep_ltq2_pend);
  return rc;
}

// ---------------------------------------------------------------------------------------------
// the path: pp_place
// ---------------------------------------------------------------------------------------------
struct PhaseTimer {
  pp_ctx *c; int phase;
  PhaseTimer(pp_ctx *c_, int ph) : c(c_), phase(ph) { cudaEventRecord(c->ev[0], c->stream); }
  void stop() { cudaEventRecord(c->ev[1], c->stream); cudaEventSynchronize(c->ev[1]); float ms = 0; cudaEventElapsedTime(&ms, c->ev[0], c->ev[1]); c->phase_ms[phase] += ms; }
};

static int place_batch(pp_ctx *c, const pp_prefs *pf, int64_t Q, const uint8_t *d_chars, PlaceOut *d_out, int32_t *d_n_out,
                       int32_t *d_status, pp_debug *dbg, int64_t q_off) {
  int rc;
  const int n_nodes = c->n_nodes;
  const bool all_edges = pf->max_strikes == 0;
  const int P = all_edges ? n_nodes - 1 : pf->max_pitches + 1;
  BallPrefs bp;
  bp.t_max_pitches = all_edges ? 0x7fffffff : pf->max_pitches;
  bp.t_max_strikes = all_edges ? 0x7fffffff : pf->max_strikes;
  bp.strike_box = pf->strike_box; bp.keep_at_most = pf->keep_at_most; bp.log_keep_factor = std::log(pf->keep_factor);
  // ---- stage 0: query prep, evolved-query table (core.ml:121-158) ----
  if ((rc = prep_queries(c, Q, d_chars)) || (rc = build_query_table(c, pf->start_pend))) return rc;
  // ---- stage 1: prefilter + ranking (core.ml:187-197) ----
  const int n_figs = (int)c->fig_rep.size();
  const bool figs = n_figs > 0 && !all_edges;  // max_strikes 0 optimises every edge: the order is irrelevant
  {
    PhaseTimer t(c, 0);
    if (figs) {
      // Fig._enum_by_score (fig.ml:120-138): only the representatives are scored up front
      if ((rc = run_prefilter_reps(c, Q, c->prefilter_choice, pf->start_pend)) || (rc = c->scores.ensure((size_t)Q * n_nodes))) return rc;
      fill_f64_kernel<<<c->sm_count * 8, 256, 0, c->stream>>>(c->scores.p, std::nan(""), Q * (int64_t)n_nodes); LAUNCH_CHECK(c);
    } else if ((rc = run_prefilter(c, Q, c->prefilter_choice, pf->start_pend))) return rc;
    t.stop();
  }
  const int64_t total = Q * (int64_t)n_nodes;
  // Ranking: baseball reads only the head of Fig.enum_by_score's order, so large trees get a truncated
  // ranking (first kTopK entries, topk_rank_kernel); a query that runs through all of them sends the
  // batch back through the full segmented sort (second attempt below).
  static const bool rank_full_only = [] { const char *e = std::getenv("PPB_RANK_FULL"); return e && std::atoi(e) != 0; }();
  const bool full_only = c->rank_mode < 0 ? rank_full_only : c->rank_mode == 1;
  const bool can_truncate = !all_edges && !full_only && !figs && n_nodes > 4 * kTopK;
  // stable descending order of Q segments of n scores each (nan and `root` last): out[q][n] = indices
  auto rank_sorted = [&](const double *sc, int n, int root, int32_t *out) -> int {
    int rc2;
    const int64_t tot = Q * (int64_t)n;
    if ((rc2 = c->keys_in.ensure(tot)) || (rc2 = c->keys_out.ensure(tot)) || (rc2 = c->rank_in.ensure(tot)) || (rc2 = c->seg_off.ensure(Q + 1)))
      return rc2;
    rank_keys_kernel<<<c->sm_count * 8, 256, 0, c->stream>>>(sc, c->keys_in.p, n, root, tot); LAUNCH_CHECK(c);
    iota_segments_kernel<<<c->sm_count * 8, 256, 0, c->stream>>>(c->rank_in.p, n, tot); LAUNCH_CHECK(c);
    std::vector<int32_t> off(Q + 1);
    if (tot > 0x7fffffffLL) return fail(PP_EINVAL, "place_batch: batch too large for 32-bit segment offsets");
    for (int64_t q = 0; q <= Q; ++q) off[q] = (int32_t)(q * n);
    CK(cudaMemcpyAsync(c->seg_off.p, off.data(), (Q + 1) * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
    size_t tmp_bytes = 0;
    CK(cub::DeviceSegmentedSort::StableSortPairsDescending(nullptr, tmp_bytes, c->keys_in.p, c->keys_out.p, c->rank_in.p, out,
                                                           tot, (int)Q, c->seg_off.p, c->seg_off.p + 1, c->stream));
    if ((rc2 = c->sort_tmp.ensure(tmp_bytes))) return rc2;
    CK(cub::DeviceSegmentedSort::StableSortPairsDescending(c->sort_tmp.p, tmp_bytes, c->keys_in.p, c->keys_out.p, c->rank_in.p, out,
                                                           tot, (int)Q, c->seg_off.p, c->seg_off.p + 1, c->stream));
    c->launches += 3;
    CK(cudaStreamSynchronize(c->stream));  // `off` must outlive the async copy
    return PP_OK;
  };
  auto rank_full = [&]() -> int {
    int rc2;
    if ((rc2 = c->rank_loc.ensure(total))) return rc2;
    return rank_sorted(c->scores.p, n_nodes, c->root_id, c->rank_loc.p);
  };
  // ---- fig ranking state (fig.cuh) ----
  FigState fs{};
  FigDev fd{};
  int32_t *fig_seg_begin = nullptr, *fig_seg_end = nullptr;
  auto fig_setup = [&]() -> int {
    int rc2;
    const int words = (n_nodes + 31) / 32;
    if ((rc2 = c->rank_loc.ensure(total)) || (rc2 = c->fig_order.ensure((size_t)Q * n_figs)) || (rc2 = c->fig_i32.ensure((size_t)6 * Q)) ||
        (rc2 = c->fig_bits.ensure((size_t)Q * words)))
      return rc2;
    if ((rc2 = rank_sorted(c->rep_scores.p, n_figs, -1, c->fig_order.p))) return rc2;
    fs.order = c->fig_order.p; fs.rep_sc = c->rep_scores.p;
    fs.fig_next = c->fig_i32.p; fs.n_ranked = c->fig_i32.p + Q; fs.n_new = c->fig_i32.p + 2 * Q; fs.need = c->fig_i32.p + 3 * Q;
    fig_seg_begin = c->fig_i32.p + 4 * Q; fig_seg_end = c->fig_i32.p + 5 * Q;
    fs.yielded = c->fig_bits.p; fs.words = words;
    fd.ptr = c->d_fig_ptr.p; fd.idx = c->d_fig_idx.p; fd.m = n_figs;
    fig_init_kernel<<<c->sm_count * 4, 256, 0, c->stream>>>(fs, (int)Q); LAUNCH_CHECK(c);
    return PP_OK;
  };
  // one expansion round for the flagged queries: next group -> score the new members -> sort -> append
  auto fig_expand_round = [&]() -> int {
    int rc2;
    CK(cudaMemsetAsync(c->counters.p + 6, 0, 2 * sizeof(int), c->stream));
    fig_expand_kernel<<<(unsigned)Q, 128, 0, c->stream>>>(fs, fd, pf->strike_box, c->rank_loc.p, n_nodes, (int)Q, c->counters.p + 6); LAUNCH_CHECK(c);
    CK(cudaMemcpyAsync(c->h_counters, c->counters.p, 8 * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    const int n_pairs = c->h_counters[6];
    if (n_pairs > 0) {
      if ((rc2 = c->fig_pair_q.ensure(n_pairs)) || (rc2 = c->fig_pair_e.ensure(n_pairs)) || (rc2 = c->fig_pair_e2.ensure(n_pairs)) ||
          (rc2 = c->fig_keys.ensure(n_pairs)) || (rc2 = c->fig_keys2.ensure(n_pairs)))
        return rc2;
      fig_pairs_kernel<<<(unsigned)Q, 128, 0, c->stream>>>(fs, c->rank_loc.p, n_nodes, (int)Q, c->fig_pair_q.p, c->fig_pair_e.p, fig_seg_begin, fig_seg_end,
                                                             c->counters.p + 7); LAUNCH_CHECK(c);
      DISPATCH_MODEL(c, {
        fig_score_kernel<S_, G_><<<(unsigned)(((int64_t)n_pairs * 32 + 127) / 128), 128, 0, c->stream>>>(c->dm, c->g, c->Tq.p, c->d_site_cat, c->qsite.p, c->nq.p,
                                                                                                       c->fig_pair_q.p, c->fig_pair_e.p, n_pairs, c->scores.p);
        LAUNCH_CHECK(c);
      });
      // the reference's order inside a group: IntSet.elements (edge ascending), then a stable sort by score descending
      size_t t1 = 0, t2 = 0;
      CK(cub::DeviceSegmentedSort::SortKeys(nullptr, t1, c->fig_pair_e.p, c->fig_pair_e2.p, n_pairs, (int)Q, fig_seg_begin, fig_seg_end, c->stream));
      CK(cub::DeviceSegmentedSort::StableSortPairsDescending(nullptr, t2, c->fig_keys.p, c->fig_keys2.p, c->fig_pair_e2.p, c->fig_pair_e.p, n_pairs, (int)Q,
                                                             fig_seg_begin, fig_seg_end, c->stream));
      if ((rc2 = c->fig_tmp.ensure(std::max(t1, t2)))) return rc2;
      CK(cub::DeviceSegmentedSort::SortKeys(c->fig_tmp.p, t1, c->fig_pair_e.p, c->fig_pair_e2.p, n_pairs, (int)Q, fig_seg_begin, fig_seg_end, c->stream));
      fig_keys_kernel<<<(n_pairs + 255) / 256, 256, 0, c->stream>>>(c->fig_pair_q.p, c->fig_pair_e2.p, n_pairs, c->scores.p, n_nodes, c->fig_keys.p); LAUNCH_CHECK(c);
      CK(cub::DeviceSegmentedSort::StableSortPairsDescending(c->fig_tmp.p, t2, c->fig_keys.p, c->fig_keys2.p, c->fig_pair_e2.p, c->fig_pair_e.p, n_pairs, (int)Q,
                                                             fig_seg_begin, fig_seg_end, c->stream));
      c->launches += 2;
    } else {
      CK(cudaMemsetAsync(fig_seg_begin, 0, 2 * Q * sizeof(int32_t), c->stream));
    }
    fig_append_kernel<<<(unsigned)Q, 128, 0, c->stream>>>(fs, fig_seg_begin, c->fig_pair_e.p, c->rank_loc.p, n_nodes, (int)Q); LAUNCH_CHECK(c);
    return PP_OK;
  };
  auto rank_truncated = [&]() -> int {
    int rc2;
    if ((rc2 = c->rank_loc.ensure(total))) return rc2;
    topk_rank_kernel<<<(unsigned)Q, 256, 0, c->stream>>>(c->scores.p, c->rank_loc.p, n_nodes, c->root_id); LAUNCH_CHECK(c);
    return PP_OK;
  };
  // ---- stage 2: baseball in waves (core.ml:283-318) ----
  const size_t per_q_i32 = 7, per_q_f64 = 1;
  if ((rc = c->i32pool.ensure((size_t)Q * (per_q_i32 + 2) + (size_t)Q * P)) || (rc = c->f64pool.ensure((size_t)Q * per_q_f64 + 3 * (size_t)Q * P)) ||
      (rc = c->n_calls.ensure(Q)))
    return rc;
  BallState b;
  {
    int32_t *ip = c->i32pool.p;
    b.next_rank = ip; ip += Q; b.n_results = ip; ip += Q; b.n_strikes = ip; ip += Q; b.done = ip; ip += Q;
    b.wave_begin = ip; ip += Q; b.wave_n = ip; ip += Q; b.qstatus = ip; ip += Q;
    KeepState dummy; (void)dummy;
    b.res_loc = ip + 2 * Q;  // keep_begin / n_keep live in the two Q-sized slots before
    double *fp = c->f64pool.p;
    b.like_record = fp; fp += Q; b.res_ll = fp; fp += (size_t)Q * P; b.res_pend = fp; fp += (size_t)Q * P; b.res_dist = fp;
    b.n_calls = c->n_calls.p; b.P = P;
  }
  KeepState ks; ks.keep_begin = c->i32pool.p + 7 * Q; ks.n_keep = c->i32pool.p + 8 * Q;
  const int qb = (int)((Q + 255) / 256);
  const size_t max_wave = all_edges ? (size_t)Q * (n_nodes - 1) : (size_t)Q * std::max(std::min(pf->max_strikes, pf->max_pitches + 1), 1);
  if ((rc = c->cand.ensure(std::max(max_wave, (size_t)Q * pf->keep_at_most))) || (rc = c->cand_out.ensure(std::max(max_wave, (size_t)Q * pf->keep_at_most))))
    return rc;
  {
    const TTParams prm = make_params(pf, TT_MODE_OPTIMIZE, pf->initial_tolerance, 1);
    for (int attempt = 0;; ++attempt) {
      const bool truncated = can_truncate && attempt == 0;
      const int n_rank = truncated ? kTopK : n_nodes;
      {
        PhaseTimer tr(c, 1);
        if ((rc = figs ? fig_setup() : (truncated ? rank_truncated() : rank_full()))) return rc;
        tr.stop();
      }
      PhaseTimer tw(c, 2);
      ball_init_kernel<<<qb, 256, 0, c->stream>>>(b, c->prep_status.p, (int)Q); LAUNCH_CHECK(c);
      CK(cudaMemsetAsync(c->counters.p + 5, 0, sizeof(int), c->stream));
      for (int wave = 0;; ++wave) {
        if (figs) {  // expand groups until every unfinished query has the locations this wave needs (or no fig left)
          for (int round = 0;; ++round) {
            CK(cudaMemsetAsync(c->counters.p + 4, 0, sizeof(int), c->stream));
            fig_need_kernel<<<qb, 256, 0, c->stream>>>(b, bp, fs, n_figs, (int)Q, c->counters.p + 4); LAUNCH_CHECK(c);
            CK(cudaMemcpyAsync(c->h_counters, c->counters.p, 8 * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
            CK(cudaStreamSynchronize(c->stream));
            if (c->h_counters[4] == 0) break;
            if ((rc = fig_expand_round())) return rc;
            if (round > n_figs + 8) return fail(PP_ECUDA, "place_batch: fig expansion did not terminate");
          }
        }
        CK(cudaMemsetAsync(c->counters.p + 1, 0, 2 * sizeof(int), c->stream));
        ball_plan_kernel<<<qb, 256, 0, c->stream>>>(b, bp, c->rank_loc.p, n_nodes, n_rank, c->root_id, (int)Q, c->cand.p, c->counters.p + 1,
                                                    c->counters.p + 5, figs ? fs.n_ranked : nullptr); LAUNCH_CHECK(c);
        CK(cudaMemcpyAsync(c->h_counters, c->counters.p, 8 * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
        CK(cudaStreamSynchronize(c->stream));
        const int n_cand = c->h_counters[1];
        if (n_cand == 0) break;
        if ((rc = launch_three_tax(c, c->cand.p, c->cand_out.p, n_cand, prm))) return rc;
        ball_update_kernel<<<qb, 256, 0, c->stream>>>(b, bp, c->cand.p, c->cand_out.p, n_rank, n_nodes, (int)Q, c->counters.p + 2,
                                                      c->counters.p + 5, figs ? fs.n_ranked : nullptr, figs ? fs.fig_next : nullptr, n_figs); LAUNCH_CHECK(c);
        if (wave > n_nodes + 64) return fail(PP_ECUDA, "place_batch: baseball did not terminate");
      }
      CK(cudaMemcpyAsync(c->h_counters, c->counters.p, 8 * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
      CK(cudaStreamSynchronize(c->stream));
      tw.stop();
      if (!(truncated && c->h_counters[5] > 0)) break;  // else: some query used up the truncated ranking - redo on the full one
      c->rank_fallbacks++;
    }
    PhaseTimer t(c, 2);
    // ---- keep + refinement (core.ml:336-369) ----
    CK(cudaMemsetAsync(c->counters.p + 3, 0, sizeof(int), c->stream));
    keep_select_kernel<<<qb, 256, 0, c->stream>>>(b, bp, ks, (int)Q, c->cand.p, c->counters.p + 3, PP_Q_EMPTY_RESULTS); LAUNCH_CHECK(c);
    CK(cudaMemcpyAsync(c->h_counters, c->counters.p, 4 * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    const int n_keep = c->h_counters[3];
    if ((rc = launch_three_tax(c, c->cand.p, c->cand_out.p, n_keep, make_params(pf, TT_MODE_OPTIMIZE, pf->final_tolerance, 0)))) return rc;
    refine_finish_kernel<<<qb, 256, 0, c->stream>>>(b, ks, (int)Q, c->cand.p, c->cand_out.p, d_out, pf->keep_at_most, d_n_out); LAUNCH_CHECK(c);
    t.stop();
    // ---- --pp (core.ml:379-404) ----
    PhaseTimer tp(c, 3);
    if (pf->calc_pp) {
      if ((rc = c->pp_out.ensure(std::max(n_keep, 1)))) return rc;
      TTParams pp = make_params(pf, TT_MODE_MARGINAL, 0.0, 0);
      DevBuf<double> prior_means;
      if (pf->prior_kind == 2) {
        if (!pf->prior_means) return fail(PP_EINVAL, "pp_place: informative prior needs prior_means");
        if ((rc = prior_means.ensure(n_nodes))) return rc;
        CK(cudaMemcpyAsync(prior_means.p, pf->prior_means, sizeof(double) * n_nodes, cudaMemcpyHostToDevice, c->stream));
        pp.prior_means = prior_means.p;
      }
      rc = launch_three_tax(c, c->cand.p, c->pp_out.p, n_keep, pp);
      if (rc == PP_OK) { final_sort_kernel<<<qb, 256, 0, c->stream>>>(b, ks, (int)Q, c->pp_out.p, 1, d_out, pf->keep_at_most, d_n_out, PP_Q_PP_FAILED); c->launches++; }
      cudaStreamSynchronize(c->stream);
      prior_means.release();
      if (rc) return rc;
      CK(cudaGetLastError());
    } else {
      final_sort_kernel<<<qb, 256, 0, c->stream>>>(b, ks, (int)Q, nullptr, 0, d_out, pf->keep_at_most, d_n_out, PP_Q_PP_FAILED); LAUNCH_CHECK(c);
    }
    tp.stop();
  }
  CK(cudaMemcpyAsync(d_status, b.qstatus, Q * sizeof(int32_t), cudaMemcpyDeviceToDevice, c->stream));
  // ---- optional stage outputs ----
  if (dbg) {
    if (dbg->scores) CK(cudaMemcpyAsync(dbg->scores + q_off * n_nodes, c->scores.p, sizeof(double) * total, cudaMemcpyDeviceToHost, c->stream));
    if (dbg->n_pitches) CK(cudaMemcpyAsync(dbg->n_pitches + q_off, b.n_results, sizeof(int32_t) * Q, cudaMemcpyDeviceToHost, c->stream));
    if (dbg->n_like_calls) CK(cudaMemcpyAsync(dbg->n_like_calls + q_off, b.n_calls, sizeof(int64_t) * Q, cudaMemcpyDeviceToHost, c->stream));
    // NOTE: pitch_* arrays come back sorted by likelihood (keep_select sorts in place), not in play order
    const int PK = pf->max_pitches + 1;
    if (!all_edges && PK == P) {
      if (dbg->pitch_edges) CK(cudaMemcpyAsync(dbg->pitch_edges + q_off * PK, b.res_loc, sizeof(int32_t) * Q * P, cudaMemcpyDeviceToHost, c->stream));
      if (dbg->pitch_ll) CK(cudaMemcpyAsync(dbg->pitch_ll + q_off * PK, b.res_ll, sizeof(double) * Q * P, cudaMemcpyDeviceToHost, c->stream));
      if (dbg->pitch_pend) CK(cudaMemcpyAsync(dbg->pitch_pend + q_off * PK, b.res_pend, sizeof(double) * Q * P, cudaMemcpyDeviceToHost, c->stream));
      if (dbg->pitch_dist) CK(cudaMemcpyAsync(dbg->pitch_dist + q_off * PK, b.res_dist, sizeof(double) * Q * P, cudaMemcpyDeviceToHost, c->stream));
    }
  }
  CK(cudaStreamSynchronize(c->stream));
  return PP_OK;
}

static int64_t batch_size(pp_ctx *c, const pp_prefs *pf) {
  // bound the Q x n_nodes score / ranking buffers (~40 B per entry) to a third of the memory that is free once the
  // reference is resident (6 ... 24 GB) and to 2^31 entries: larger batches mean fuller baseball waves (C4: 7 500 -> 30 000 queries)
  if (c->batch_budget <= 0.0) {
    size_t free_b = 0, total_b = 0;
    if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) { cudaGetLastError(); free_b = 0; }
    c->batch_budget = std::min(24.0e9, std::max(6.0e9, (double)free_b / 3.0));
  }
  const int64_t by_mem = (int64_t)(c->batch_budget / (40.0 * c->n_nodes));
  const int64_t by_idx = 0x7fffffffLL / c->n_nodes;
  int64_t b = std::min<int64_t>(std::min(by_mem, by_idx), 1 << 16);
  if (pf->max_strikes == 0) b = std::min<int64_t>(b, 4096);
  return std::max<int64_t>(b, 1);
}

static int check_prefs(const pp_prefs *pf) {
  if (!pf) return fail(PP_EINVAL, "null prefs");
  if (pf->keep_at_most < 1 || pf->max_pitches < 1 || pf->max_strikes < 0 || !(pf->keep_factor > 0) || pf->max_iter < 1)
    return fail(PP_EINVAL, "pp_place: bad prefs");
  return PP_OK;
}

extern "C" int pp_place_device(pp_ctx *c, const pp_prefs *pf, int64_t Q, const uint8_t *d_query_chars, pp_result *d_out,
                               int32_t *d_n_out, int32_t *d_status) {
  if (!c || !d_query_chars || !d_out || !d_n_out || !d_status || Q < 0) return fail(PP_EINVAL, "pp_place_device: bad arguments");
  if (!c->precomputed) return fail(PP_ESTATE, "pp_place_device: call pp_precompute first");
  int rc;
  if ((rc = check_prefs(pf))) return rc;
  CK(cudaSetDevice(c->device));
  static_assert(sizeof(PlaceOut) == sizeof(pp_result), "PlaceOut must match pp_result");
  const int64_t B = batch_size(c, pf);
  for (int64_t q0 = 0; q0 < Q; q0 += B) {
    const int64_t n = std::min(B, Q - q0);
    if ((rc = place_batch(c, pf, n, d_query_chars + q0 * c->L, reinterpret_cast<PlaceOut *>(d_out) + q0 * pf->keep_at_most,
                          d_n_out + q0, d_status + q0, nullptr, q0)))
      return rc;
  }
  return PP_OK;
}

extern "C" int pp_place(pp_ctx *c, const pp_prefs *pf, int64_t Q, const uint8_t *query_chars, pp_result *out, int32_t *n_out,
                        int32_t *status, pp_debug *dbg) {
  if (!c || !query_chars || !out || !n_out || !status || Q < 0) return fail(PP_EINVAL, "pp_place: bad arguments");
  if (!c->precomputed) return fail(PP_ESTATE, "pp_place: call pp_precompute first");
  int rc;
  if ((rc = check_prefs(pf))) return rc;
  CK(cudaSetDevice(c->device));
  const int64_t B = batch_size(c, pf);
  const int64_t Bm = std::min(B, std::max<int64_t>(Q, 1));
  if ((rc = c->chars.ensure((size_t)Bm * c->L)) || (rc = c->place_out.ensure((size_t)Bm * pf->keep_at_most)) || (rc = c->n_out.ensure(2 * Bm)))
    return rc;
  if (dbg && dbg->timing) for (int i = 0; i < 4; ++i) dbg->timing[i] = 0.0;
  double ms0[4]; std::memcpy(ms0, c->phase_ms, sizeof ms0);
  for (int64_t q0 = 0; q0 < Q; q0 += B) {
    const int64_t n = std::min(B, Q - q0);
    CK(cudaMemcpyAsync(c->chars.p, query_chars + q0 * c->L, (size_t)n * c->L, cudaMemcpyHostToDevice, c->stream));
    if ((rc = place_batch(c, pf, n, c->chars.p, c->place_out.p, c->n_out.p, c->n_out.p + Bm, dbg, q0))) return rc;
    CK(cudaMemcpyAsync(out + q0 * pf->keep_at_most, c->place_out.p, sizeof(pp_result) * (size_t)n * pf->keep_at_most, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaMemcpyAsync(n_out + q0, c->n_out.p, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaMemcpyAsync(status + q0, c->n_out.p + Bm, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
  }
  if (dbg && dbg->timing) {
    dbg->timing[0] = (c->phase_ms[0] - ms0[0] + c->phase_ms[1] - ms0[1]) * 1e-3;  // "ranking"
    dbg->timing[1] = (c->phase_ms[2] - ms0[2]) * 1e-3;                             // "ML calculation"
    dbg->timing[2] = (c->phase_ms[3] - ms0[3]) * 1e-3;                             // "PP calculation"
    dbg->timing[3] = 0.0;
  }
  return PP_OK;
}

// ---------------------------------------------------------------------------------------------
// misc
// ---------------------------------------------------------------------------------------------
template <int S, bool G>
__global__ void evolve_test_kernel(const DevModel *__restrict__ m, const double *__restrict__ src, double *__restrict__ dst, int n,
                                   int np, const int32_t *__restrict__ site_cat, double bl, int Rt) {
  extern __shared__ double sm[];
  constexpr int RS = p_rstride<S, G>();
  double *ex = sm, *P = sm + Rt * S;
  block_build_P<S>(m, Rt, bl, P, RS, ex);
  const int Rg = G ? 1 : Rt;
  for (int site = threadIdx.x + blockIdx.x * blockDim.x; site < n; site += gridDim.x * blockDim.x)
    for (int r = 0; r < Rg; ++r) {
      const double *Pr = P + (G ? site_cat[site] : r) * RS;
      double v[S];
      for (int s = 0; s < S; ++s) v[s] = src[(size_t)(r * S + s) * np + site];
      for (int i = 0; i < S; ++i) {
        double acc = 0.0;
        for (int j = 0; j < S; ++j) acc = fma(Pr[i * S + j], v[j], acc);
        dst[(size_t)(r * S + i) * np + site] = acc;
      }
    }
}

extern "C" int pp_evolve(pp_ctx *c, const double *src_a, const int64_t *src_e, int32_t n, double bl, double *dst_a, int64_t *dst_e) {
  if (!c || !src_a || !dst_a || n < 1 || n > c->L) return fail(PP_EINVAL, "pp_evolve: bad arguments");
  CK(cudaSetDevice(c->device));
  const int np = (n + 31) / 32 * 32;
  std::vector<double> h((size_t)c->K * np, 0.0);
  for (int r = 0; r < c->Rg; ++r)
    for (int site = 0; site < n; ++site)
      for (int s = 0; s < c->S; ++s) h[(size_t)(r * c->S + s) * np + site] = src_a[((size_t)r * n + site) * c->S + s];
  double *d_src = nullptr, *d_dst = nullptr;
  DevFree g_src(reinterpret_cast<void **>(&d_src)), g_dst(reinterpret_cast<void **>(&d_dst));
  CK(cudaMalloc(&d_src, h.size() * sizeof(double)));
  CK(cudaMalloc(&d_dst, h.size() * sizeof(double)));
  CK(cudaMemcpy(d_src, h.data(), h.size() * sizeof(double), cudaMemcpyHostToDevice));
  DISPATCH_MODEL(c, {
    const size_t sm = (size_t)(c->R * S_ + c->R * p_rstride<S_, G_>()) * sizeof(double);
    CK(cudaFuncSetAttribute(evolve_test_kernel<S_, G_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    evolve_test_kernel<S_, G_><<<(n + 127) / 128, 128, sm, c->stream>>>(c->dm, d_src, d_dst, n, np, c->d_site_cat, bl, c->R);
    LAUNCH_CHECK(c);
  });
  CK(cudaStreamSynchronize(c->stream));
  CK(cudaMemcpy(h.data(), d_dst, h.size() * sizeof(double), cudaMemcpyDeviceToHost));
  for (int r = 0; r < c->Rg; ++r)
    for (int site = 0; site < n; ++site)
      for (int s = 0; s < c->S; ++s) dst_a[((size_t)r * n + site) * c->S + s] = h[(size_t)(r * c->S + s) * np + site];
  if (src_e && dst_e) std::memcpy(dst_e, src_e, sizeof(int64_t) * n);  // exponents are copied (gmix_model.ml:311)
  return PP_OK;
}

extern "C" int pp_set_stream(pp_ctx *c, void *cuda_stream) {
  if (!c) return fail(PP_EINVAL, "null ctx");
  CK(cudaSetDevice(c->device));
  CK(cudaStreamSynchronize(c->stream));
  c->stream = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : c->own_stream;
  return PP_OK;
}

extern "C" int pp_set_prefilter_kernel(pp_ctx *c, int32_t kernel) {
  if (!c || kernel < 0 || kernel > 8) return fail(PP_EINVAL, "pp_set_prefilter_kernel: kernel must be 0 ... 8");
  const bool dna4 = c->S == 4 && !c->gcat && c->K == 16;
  if ((kernel == 2 || (kernel >= 4 && kernel <= 6) || kernel == 8) && !dna4) return fail(PP_EINVAL, "pp_set_prefilter_kernel: the table and staged forms exist for the DNA Gamma-4 model only");

  c->prefilter_choice = kernel;
  return PP_OK;
}
extern "C" int pp_prefilter_kernel_used(pp_ctx *c) { return c ? c->prefilter_kernel_used : -1; }

extern "C" int pp_set_rank_mode(pp_ctx *c, int32_t mode) {
  if (!c || mode < 0 || mode > 1) return fail(PP_EINVAL, "pp_set_rank_mode: mode must be 0 (truncated ranking when possible) or 1 (full sort)");
  c->rank_mode = mode;
  return PP_OK;
}
extern "C" int pp_rank_fallbacks(pp_ctx *c) { return c ? c->rank_fallbacks : -1; }

extern "C" int pp_set_three_tax_engine(pp_ctx *c, int32_t engine, int32_t blocks_per_sm) {
  if (!c || engine < 0 || engine > 4 || blocks_per_sm < 0) return fail(PP_EINVAL, "pp_set_three_tax_engine: engine must be 0, 1, 2, 3 or 4");
  if (engine >= 2 && !ttw_supported(c)) return fail(PP_EINVAL, "pp_set_three_tax_engine: engines 2-4 exist for the DNA Gamma-4 model only");
  if (engine == 4 && !c->eig_ok) return fail(PP_EINVAL, "pp_set_three_tax_engine: engine 4 needs a rate matrix with a zero eigenvalue");
  c->tt_engine = engine;
  c->ttw_blocks_per_sm = blocks_per_sm;
  return PP_OK;
}
extern "C" int pp_three_tax_engine_used(pp_ctx *c) {
  if (!c) return -1;
  if (!(ttw_supported(c) && c->tt_engine != 1)) return 1;
  if (c->tt_engine == 2) return 2;
  return ((c->tt_engine == 0 || c->tt_engine == 4) && c->eig_ok) ? 4 : 3;
}

extern "C" int64_t pp_kernel_launches(pp_ctx *c, int32_t reset) {
  if (!c) return -1;
  const int64_t v = c->launches;
  if (reset) c->launches = 0;
  return v;
}
extern "C" int pp_phase_times_ms(pp_ctx *c, double *out4, int32_t reset) {
  if (!c || !out4) return fail(PP_EINVAL, "pp_phase_times_ms: bad arguments");
  std::memcpy(out4, c->phase_ms, sizeof c->phase_ms);
  if (reset) std::memset(c->phase_ms, 0, sizeof c->phase_ms);
  return PP_OK;
}
extern "C" int pp_three_tax_work(pp_ctx *c, double *out4, int32_t reset) {
  if (!c || !out4) return fail(PP_EINVAL, "pp_three_tax_work: bad arguments");
  CK(cudaSetDevice(c->device));
  unsigned long long h[4];
  CK(cudaStreamSynchronize(c->stream));
  CK(cudaMemcpy(h, c->tt_work.p, sizeof h, cudaMemcpyDeviceToHost));
  for (int i = 0; i < 4; ++i) out4[i] = (double)h[i];
  if (reset) CK(cudaMemset(c->tt_work.p, 0, sizeof h));
  return PP_OK;
}
extern "C" int pp_three_tax_profile(pp_ctx *c, double *out32, int32_t reset) {
  if (!c || !out32) return fail(PP_EINVAL, "pp_three_tax_profile: bad arguments");
  CK(cudaSetDevice(c->device));
  unsigned long long h[32];
  CK(cudaStreamSynchronize(c->stream));
  CK(cudaMemcpy(h, c->tt_work.p, sizeof h, cudaMemcpyDeviceToHost));
  for (int i = 0; i < 32; ++i) out32[i] = (double)h[i];
  if (reset) CK(cudaMemset(c->tt_work.p, 0, sizeof h));
  return PP_OK;
}
extern "C" int pp_measure_fp64_peak(pp_ctx *c, double *tflops_out) {
  if (!c || !tflops_out) return fail(PP_EINVAL, "pp_measure_fp64_peak: bad arguments");
  CK(cudaSetDevice(c->device));
  const int blocks = c->sm_count * 8, threads = 256, iters = 1 << 15;
  double *d = nullptr;
  CK(cudaMalloc(&d, sizeof(double) * blocks * threads));
  dfma_probe_kernel<<<blocks, threads, 0, c->stream>>>(d, 1024, 0.999999, 1e-9);  // warm-up
  float best = 1e30f;
  for (int rep = 0; rep < 5; ++rep) {
    cudaEventRecord(c->ev[0], c->stream);
    dfma_probe_kernel<<<blocks, threads, 0, c->stream>>>(d, iters, 0.999999, 1e-9);
    cudaEventRecord(c->ev[1], c->stream);
    CK(cudaEventSynchronize(c->ev[1]));
    float ms = 0; cudaEventElapsedTime(&ms, c->ev[0], c->ev[1]);
    best = std::min(best, ms);
  }
  c->launches += 6;
  cudaFree(d);
  *tflops_out = 2.0 * 8.0 * (double)iters * blocks * threads / (best * 1e-3) / 1e12;
  return PP_OK;
}

// ---------------------------------------------------------------------------------------------
// all the GPUs of one box behind one handle (include/pplacer_b200.h: pp_create_multi ...)
// ---------------------------------------------------------------------------------------------
struct pp_multi {
  std::vector<pp_ctx *> ctx;
  std::vector<int> device;
  std::vector<int64_t> last_counts;
};

// run fn(i) on one host thread per device; the first failing device's code and message win
template <class F>
static int multi_for_each(pp_multi *m, F fn) {
  const int n = (int)m->ctx.size();
  std::vector<int> rc(n, PP_OK);
  std::vector<std::string> msg(n);
  std::vector<std::thread> th;
  for (int i = 0; i < n; ++i)
    th.emplace_back([&, i] {
      rc[i] = fn(i);
      if (rc[i] != PP_OK) msg[i] = g_err;  // the worker's thread-local message
    });
  for (auto &t : th) t.join();
  for (int i = 0; i < n; ++i)
    if (rc[i] != PP_OK) return fail(rc[i], "device " + std::to_string(m->device[i]) + ": " + msg[i]);
  return PP_OK;
}

extern "C" int pp_create_multi(const pp_model *mo, const pp_tree *t, const uint8_t *ref_chars, int32_t n_leaves, int32_t L,
                               int32_t n_devices, const int32_t *device_ids, pp_multi **out) {
  if (!out) return fail(PP_EINVAL, "pp_create_multi: null argument");
  *out = nullptr;
  if (n_devices < 1 || n_devices > 64) return fail(PP_EINVAL, "pp_create_multi: n_devices must be 1 ... 64");
  pp_multi *m = new pp_multi();
  m->ctx.assign(n_devices, nullptr);
  m->last_counts.assign(n_devices, 0);
  for (int i = 0; i < n_devices; ++i) m->device.push_back(device_ids ? device_ids[i] : i);
  const int rc = multi_for_each(m, [&](int i) { return pp_create(mo, t, ref_chars, n_leaves, L, m->device[i], &m->ctx[i]); });
  if (rc != PP_OK) {
    const std::string keep = g_err;
    pp_destroy_multi(m);
    return fail(rc, keep);
  }
  *out = m;
  return PP_OK;
}

extern "C" void pp_destroy_multi(pp_multi *m) {
  if (!m) return;
  for (pp_ctx *c : m->ctx)
    if (c) pp_destroy(c);
  delete m;
}

extern "C" int pp_multi_precompute(pp_multi *m) {
  if (!m) return fail(PP_EINVAL, "pp_multi_precompute: null handle");
  return multi_for_each(m, [&](int i) { return pp_precompute(m->ctx[i]); });
}

extern "C" int pp_multi_place(pp_multi *m, const pp_prefs *pf, int64_t Q, const uint8_t *query_chars, pp_result *out, int32_t *n_out,
                              int32_t *status) {
  if (!m || !query_chars || !out || !n_out || !status || Q < 0) return fail(PP_EINVAL, "pp_multi_place: bad arguments");
  int rc;
  if ((rc = check_prefs(pf))) return rc;
  const int n = (int)m->ctx.size();
  // chunks from a shared counter: one chunk per device when the batch is small, 16k-query chunks otherwise
  const int64_t chunk = std::max<int64_t>(1, std::min<int64_t>((Q + n - 1) / n, 16384));
  std::atomic<int64_t> next{0};
  const int64_t L = m->ctx[0]->L;
  std::fill(m->last_counts.begin(), m->last_counts.end(), 0);
  return multi_for_each(m, [&](int i) {
    for (;;) {
      const int64_t q0 = next.fetch_add(chunk);
      if (q0 >= Q) return (int)PP_OK;
      const int64_t nq = std::min(chunk, Q - q0);
      const int r = pp_place(m->ctx[i], pf, nq, query_chars + q0 * L, out + q0 * pf->keep_at_most, n_out + q0, status + q0, nullptr);
      if (r != PP_OK) return r;
      m->last_counts[i] += nq;
    }
  });
}

extern "C" int pp_multi_n_devices(pp_multi *m) { return m ? (int)m->ctx.size() : -1; }
extern "C" pp_ctx *pp_multi_ctx(pp_multi *m, int32_t i) { return (m && i >= 0 && i < (int)m->ctx.size()) ? m->ctx[i] : nullptr; }
extern "C" int pp_multi_last_counts(pp_multi *m, int64_t *per_device) {
  if (!m || !per_device) return fail(PP_EINVAL, "pp_multi_last_counts: bad arguments");
  for (size_t i = 0; i < m->ctx.size(); ++i) per_device[i] = m->last_counts[i];
  return PP_OK;
}

// --check-like (pplacer_run.ml:344-377): per node i the tree log-likelihood seen from edge i two ways -
// log_like3(evolve(darr[i], bl/2), evolve(parr[i], bl/2), unit) and logdot(snodes[i], unit) - which must agree with each
// other and across nodes.  An all-ones query (N / X) evolves to the unit vector for any pendant length (rows of P sum
// to one), so the two columns are Three_tax.log_like and the prefilter score of that query.  out: n_nodes x 2
// (tree_likelihood, supernode_like); the supernode column is NaN at the root.
// Seq_post.get_summary Proximal (seq_post.ml:14-31) with Map_seq.of_map's bounded_max_index (map_seq.ml:22-25) and
// Glv.summarize_post (gmix_model.ml:241-256 / gcat_model.ml): per site the most likely state of
//   statd .* evolve(darr[id], bl[id]) .* parr[id]   summed over the rates,
// or -1 when its share of the total is below the cutoff.  One block per node.
template <int S, bool GCAT>
__global__ void __launch_bounds__(128)
map_states_kernel(const DevModel *__restrict__ m, GlvArrays g, const double *__restrict__ bl, const int32_t *__restrict__ site_cat,
                  const int32_t *__restrict__ ids, double cutoff, int Rt, int8_t *__restrict__ out) {
  extern __shared__ double sm_map[];
  constexpr int RS = S * S;
  double *ex = sm_map, *P = sm_map + Rt * S;
  const int id = ids[blockIdx.x];
  block_build_P<S>(m, Rt, bl[id], P, RS, ex);
  const double *dbase = g.a[0] + (int64_t)id * g.a_stride, *pbase = g.a[1] + (int64_t)id * g.a_stride;
  for (int site = threadIdx.x; site < g.L; site += blockDim.x) {
    double u[S];
#pragma unroll
    for (int s = 0; s < S; ++s) u[s] = 0.0;
    const int cat = GCAT ? site_cat[site] : 0;
    for (int r = 0; r < (GCAT ? 1 : Rt); ++r) {
      const double *Pr = P + (GCAT ? cat : r) * RS;
      double d[S];
#pragma unroll
      for (int j = 0; j < S; ++j) d[j] = dbase[(int64_t)(r * S + j) * g.Lp + site];
#pragma unroll
      for (int s = 0; s < S; ++s) {
        double ev = 0.0;
#pragma unroll
        for (int j = 0; j < S; ++j) ev = fma(Pr[s * S + j], d[j], ev);
        u[s] += (m->statd[s] * ev) * pbase[(int64_t)(r * S + s) * g.Lp + site];  // statd_pairwise_prod, linear_c.c:148-206
      }
    }
    int best = 0;
    double tot = fabs(u[0]);
#pragma unroll
    for (int s = 1; s < S; ++s) { if (u[s] > u[best]) best = s; tot += fabs(u[s]); }  // Gsl.Vector.max_index: the first maximum
    out[(int64_t)blockIdx.x * g.L + site] = (u[best] / tot < cutoff) ? (int8_t)-1 : (int8_t)best;
  }
}

extern "C" int pp_map_states(pp_ctx *c, int32_t n_ids, const int32_t *ids, double cutoff, int8_t *out) {
  if (!c || !ids || !out || n_ids < 1) return fail(PP_EINVAL, "pp_map_states: bad arguments");
  if (!c->precomputed) return fail(PP_ESTATE, "pp_map_states: call pp_precompute first");
  for (int i = 0; i < n_ids; ++i)
    if (ids[i] < 0 || ids[i] >= c->n_nodes || ids[i] == c->root_id) return fail(PP_EINVAL, "pp_map_states: node id out of range (or the root, which has no proximal side)");
  CK(cudaSetDevice(c->device));
  int rc;
  DevBuf<int32_t> d_ids;
  DevBuf<int8_t> d_out;
  if ((rc = d_ids.ensure(n_ids)) || (rc = d_out.ensure((size_t)n_ids * c->L))) { d_ids.release(); d_out.release(); return rc; }
  cudaMemcpyAsync(d_ids.p, ids, sizeof(int32_t) * n_ids, cudaMemcpyHostToDevice, c->stream);
  const size_t smb = (size_t)(c->R * c->S + c->R * c->S * c->S) * sizeof(double);
  cudaError_t e = cudaSuccess;
  DISPATCH_MODEL(c, {
    e = cudaFuncSetAttribute(map_states_kernel<S_, G_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smb);
    if (e == cudaSuccess) map_states_kernel<S_, G_><<<n_ids, 128, smb, c->stream>>>(c->dm, c->g, c->d_bl, c->d_site_cat, d_ids.p, cutoff, c->R, d_out.p);
  });
  c->launches += 1;
  if (e == cudaSuccess) e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaMemcpyAsync(out, d_out.p, (size_t)n_ids * c->L, cudaMemcpyDeviceToHost, c->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
  d_ids.release(); d_out.release();
  if (e != cudaSuccess) return fail(PP_ECUDA, std::string("pp_map_states: ") + cudaGetErrorString(e));
  return PP_OK;
}

extern "C" int pp_check_like(pp_ctx *c, double *out) {
  if (!c || !out) return fail(PP_EINVAL, "pp_check_like: bad arguments");
  if (!c->precomputed) return fail(PP_ESTATE, "pp_check_like: call pp_precompute first");
  const char any = c->S == 4 ? 'N' : 'X';
  const int n = c->n_nodes;
  std::vector<uint8_t> q((size_t)n * c->L, (uint8_t)any);
  std::vector<double> sc(n), ll(n), pend(n, 0.1), dist(n);
  std::vector<int32_t> edge(n);
  for (int i = 0; i < n; ++i) { edge[i] = i; dist[i] = c->bl[i] / 2.0; }
  int rc;
  if ((rc = pp_prefilter_scores(c, 0.1, 1, q.data(), sc.data(), 0))) return rc;
  if ((rc = pp_three_tax_log_like(c, n, q.data(), edge.data(), pend.data(), dist.data(), ll.data()))) return rc;
  for (int i = 0; i < n; ++i) { out[2 * i] = ll[i]; out[2 * i + 1] = sc[i]; }
  return PP_OK;
}
