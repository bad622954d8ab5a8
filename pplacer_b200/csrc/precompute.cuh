// Reference-GLV precompute on device: Felsenstein pruning over the edges in level order.
//   Like_stree.calc_distal_and_evolv_dist   pplacer_src/like_stree.ml:43-69
//   Like_stree.calc_proximal                pplacer_src/like_stree.ml:74-100
//   pull exponents at -10                   pplacer_src/pplacer_run.ml:333-335
//   Glv_arr.prep_supernodes                 pplacer_src/glv_arr.ml:46-56
// The reference recurses serially over the tree; here every level of the tree is one launch with
// one thread per (node, site).  Arithmetic per site follows the reference order (children
// multiplied in list order, exponents added, pull, then evolve).
#pragma once
#include "device_common.cuh"

namespace ppb {

constexpr int kPreThreads = 128;
constexpr int kMinAllowedTwoexp = -50;  // like_stree.ml:8

// Transition matrices for every node: P[node][r][S*S] for branch length scale*bl[node].
template <int S>
__global__ void pmat_kernel(const DevModel *__restrict__ m, const double *__restrict__ bl, int n_nodes, double scale,
                            int Rt, double *__restrict__ P) {
  extern __shared__ double sm[];
  double *ex = sm;            // Rt*S
  double *tmp = sm + Rt * S;  // Rt*S*S
  const int node = blockIdx.x;
  if (node >= n_nodes) return;
  block_build_P<S>(m, Rt, bl[node] * scale, tmp, S * S, ex);
  double *dst = P + (int64_t)node * Rt * S * S;
  for (int t = threadIdx.x; t < Rt * S * S; t += blockDim.x) dst[t] = tmp[t];
}

// Glv.perhaps_pull_exponent (gmix_model.ml:163-187, gcat_model.ml:195-212) for one site whose
// values live in global memory planes base[k*Lp]: returns the exponent pulled (0 if none) after
// rescaling in place.  Per-rate max -> frexp exponent -> max over rates, as the reference.
template <int S>
__device__ __forceinline__ int pull_exponent_site(double *base, int64_t Lp, int Rg, int threshold) {
  int max_twoexp = INT_MIN;
  for (int r = 0; r < Rg; ++r) {
    double mx = base[(int64_t)(r * S) * Lp];
#pragma unroll
    for (int s = 1; s < S; ++s) { const double v = base[(int64_t)(r * S + s) * Lp]; if (v > mx) mx = v; }
    int ex;
    (void)frexp(mx, &ex);
    if (ex > max_twoexp) max_twoexp = ex;
  }
  if (max_twoexp < threshold) {
    for (int k = 0; k < Rg * S; ++k) base[(int64_t)k * Lp] = scalbn(base[(int64_t)k * Lp], -max_twoexp);
    return max_twoexp;
  }
  return 0;
}

struct TreeDev {
  const int32_t *child_ptr, *child_idx, *parent, *leaf_row;
  const uint8_t *ref_codes;  // [n_leaves][Lp] alphabet codes of the reference alignment
  const int32_t *site_cat;   // Gcat: [Lp]
};

// One tree level of calc_distal_and_evolv_dist.  grid = (site tiles, nodes in level).
template <int S, bool GCAT>
__global__ void __launch_bounds__(kPreThreads)
distal_level_kernel(const DevModel *__restrict__ m, GlvArrays g, TreeDev t, const int32_t *__restrict__ level_nodes,
                    const double *__restrict__ Pfull, int Rt) {
  extern __shared__ double sm[];
  constexpr int RS = p_rstride<S, GCAT>();
  double *P = sm;  // Rt * RS
  const int node = level_nodes[blockIdx.y];
  const double *Pn = Pfull + (int64_t)node * Rt * S * S;
  for (int i = threadIdx.x; i < Rt * S * S; i += blockDim.x) P[(i / (S * S)) * RS + (i % (S * S))] = Pn[i];
  __syncthreads();
  const int site = blockIdx.x * blockDim.x + threadIdx.x;
  if (site >= g.L) return;
  const int Rg = m->Rg;
  const int64_t Lp = g.Lp;
  double *d = glv_plane_w(g, 0, node, 0) + site;
  const int c0 = t.child_ptr[node], c1 = t.child_ptr[node + 1];
  int e = 0;
  if (c0 == c1) {
    // leaf: Glv.prep_constant_rate_glv_from_lv_arr (gmix_model.ml:149-156)
    const uint8_t code = t.ref_codes[(int64_t)t.leaf_row[node] * Lp + site] & 0x7f;
    const uint32_t mask = m->mask_of_code[code];
    for (int r = 0; r < Rg; ++r)
#pragma unroll
      for (int s = 0; s < S; ++s) d[(int64_t)(r * S + s) * Lp] = (mask >> s & 1u) ? 1.0 : 0.0;
  } else {
    // Glv.listwise_prod (gmix_model.ml:224-233): (x*y)*rest..., exponents add
    for (int k = 0; k < Rg * S; ++k) {
      double v = glv_plane(g, 2, t.child_idx[c0], k)[site];
      for (int c = c0 + 1; c < c1; ++c) v *= glv_plane(g, 2, t.child_idx[c], k)[site];
      d[(int64_t)k * Lp] = v;
    }
    for (int c = c0; c < c1; ++c) e += g.e[2][(int64_t)t.child_idx[c] * Lp + site];
    e += pull_exponent_site<S>(d, Lp, Rg, kMinAllowedTwoexp);
  }
  g.e[0][(int64_t)node * Lp + site] = e;
  // evolv_dist = evolve(distal, bl[node]) into the snodes array used as scratch (like_stree.ml:107)
  double *ev = glv_plane_w(g, 2, node, 0) + site;
  for (int r = 0; r < Rg; ++r) {
    const double *Pr = P + (GCAT ? t.site_cat[site] : r) * RS;
    double src[S];
#pragma unroll
    for (int s = 0; s < S; ++s) src[s] = d[(int64_t)(r * S + s) * Lp];
#pragma unroll
    for (int i = 0; i < S; ++i) {
      double acc = 0.0;
#pragma unroll
      for (int j = 0; j < S; ++j) acc = fma(Pr[i * S + j], src[j], acc);
      ev[(int64_t)(r * S + i) * Lp] = acc;
    }
  }
  g.e[2][(int64_t)node * Lp + site] = e;
}

// One tree level of calc_proximal: threads own (child, site); the parent's evolved proximal is
// recomputed per child (cheaper than a round trip through HBM).  grid = (site tiles, children).
template <int S, bool GCAT>
__global__ void __launch_bounds__(kPreThreads)
proximal_level_kernel(const DevModel *__restrict__ m, GlvArrays g, TreeDev t, const int32_t *__restrict__ level_nodes,
                      const double *__restrict__ Pfull, int Rt) {
  extern __shared__ double sm[];
  constexpr int RS = p_rstride<S, GCAT>();
  double *P = sm;
  const int child = level_nodes[blockIdx.y];
  const int parent = t.parent[child];
  const double *Pn = Pfull + (int64_t)parent * Rt * S * S;
  for (int i = threadIdx.x; i < Rt * S * S; i += blockDim.x) P[(i / (S * S)) * RS + (i % (S * S))] = Pn[i];
  __syncthreads();
  const int site = blockIdx.x * blockDim.x + threadIdx.x;
  if (site >= g.L) return;
  const int Rg = m->Rg;
  const int64_t Lp = g.Lp;
  const double *pp = glv_plane(g, 1, parent, 0) + site;
  double *out = glv_plane_w(g, 1, child, 0) + site;
  const int c0 = t.child_ptr[parent], c1 = t.child_ptr[parent + 1];
  for (int r = 0; r < Rg; ++r) {
    const double *Pr = P + (GCAT ? t.site_cat[site] : r) * RS;
    double src[S];
#pragma unroll
    for (int s = 0; s < S; ++s) src[s] = pp[(int64_t)(r * S + s) * Lp];
#pragma unroll
    for (int i = 0; i < S; ++i) {
      double acc = 0.0;
#pragma unroll
      for (int j = 0; j < S; ++j) acc = fma(Pr[i * S + j], src[j], acc);
      // listwise_prod (evolved_prox :: siblings' evolv_dist in list order), like_stree.ml:88-91
      for (int c = c0; c < c1; ++c) {
        const int sib = t.child_idx[c];
        if (sib != child) acc *= glv_plane(g, 2, sib, r * S + i)[site];
      }
      out[(int64_t)(r * S + i) * Lp] = acc;
    }
  }
  int e = g.e[1][(int64_t)parent * Lp + site];
  for (int c = c0; c < c1; ++c) {
    const int sib = t.child_idx[c];
    if (sib != child) e += g.e[2][(int64_t)sib * Lp + site];
  }
  e += pull_exponent_site<S>(out, Lp, Rg, kMinAllowedTwoexp);
  g.e[1][(int64_t)child * Lp + site] = e;
}

// Gcat_model.refine (gcat_model.ml:387-419): one attempt of Like_stree.site_log_like_arr (like_stree.ml:113-131) with
// every site in category `cat`.  After the distal pass the evolved distal vectors of the root's children sit in the
// snodes array (evolv_dist); per site  log(sum_s statd_s * x_s * y_s * z_s) + ln2 * (e_x + e_y + e_z)
// (site_log_like_arr3, gcat_model.ml:355-375) with z = the product of the remaining children (listwise_prod) or the
// unit vector for a bifurcating root, then the FastTree prior  + 2 log(rate) - 3 rate, keeping the FIRST best
// category (strict >).  A non-positive site likelihood raises in the reference: counted in *n_bad.
template <int S>
__global__ void __launch_bounds__(kPreThreads)
refine_attempt_kernel(const DevModel *__restrict__ m, GlvArrays g, TreeDev t, int root, int cat, double prior_term,
                      double *__restrict__ best_ll, int32_t *__restrict__ best_cat, double *__restrict__ site_ll_out,
                      int *__restrict__ n_bad) {
  const int site = blockIdx.x * blockDim.x + threadIdx.x;
  if (site >= g.L) return;
  const int c0 = t.child_ptr[root], c1 = t.child_ptr[root + 1];
  const int i = t.child_idx[c0], j = t.child_idx[c0 + 1];
  double like = 0.0;
  int e = g.e[2][(int64_t)i * g.Lp + site] + g.e[2][(int64_t)j * g.Lp + site];
  double z[S];
#pragma unroll
  for (int s = 0; s < S; ++s) z[s] = 1.0;
  for (int c = c0 + 2; c < c1; ++c) {
    const int k = t.child_idx[c];
#pragma unroll
    for (int s = 0; s < S; ++s) z[s] = (c == c0 + 2) ? glv_plane(g, 2, k, s)[site] : z[s] * glv_plane(g, 2, k, s)[site];
    e += g.e[2][(int64_t)k * g.Lp + site];
  }
#pragma unroll
  for (int s = 0; s < S; ++s) like += m->statd[s] * glv_plane(g, 2, i, s)[site] * glv_plane(g, 2, j, s)[site] * z[s];
  if (!(like > 0.0)) atomicAdd(n_bad, 1);
  const double ll = log(like) + kLn2 * (double)e;
  if (site_ll_out) site_ll_out[(int64_t)cat * g.L + site] = ll;
  const double x = ll + prior_term;
  if (x > best_ll[site]) { best_ll[site] = x; best_cat[site] = cat; }
}
__global__ void fill_i32_kernel(int32_t *p, int32_t v, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) p[i] = v;
}
__global__ void fill_f64_kernel(double *p, double v, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) p[i] = v;
}

// Root proximal = unit (like_stree.ml:98-99).
__global__ void set_unit_kernel(GlvArrays g, int which, int node) {
  const int64_t n = (int64_t)g.K * g.Lp;
  double *a = g.a[which] + (int64_t)node * g.a_stride;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) a[i] = 1.0;
  int32_t *e = g.e[which] + (int64_t)node * g.Lp;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < g.Lp; i += (int64_t)gridDim.x * blockDim.x) e[i] = 0;
}

// pull(-10) on darr and parr in place (pplacer_run.ml:333-335) fused with prep_supernodes
// (glv_arr.ml:46-56): snodes[i] = statd * evolve(darr[i], bl_i/2) * evolve(parr[i], bl_i/2).
// grid = (site tiles, n_nodes).
template <int S, bool GCAT>
__global__ void __launch_bounds__(kPreThreads)
supernode_kernel(const DevModel *__restrict__ m, GlvArrays g, TreeDev t, const double *__restrict__ Phalf, int Rt, int node0) {
  extern __shared__ double sm[];
  constexpr int RS = p_rstride<S, GCAT>();
  double *P = sm;
  const int node = node0 + blockIdx.y;  // grid.y is capped at 65535: the launch is chunked over nodes
  const double *Pn = Phalf + (int64_t)node * Rt * S * S;
  for (int i = threadIdx.x; i < Rt * S * S; i += blockDim.x) P[(i / (S * S)) * RS + (i % (S * S))] = Pn[i];
  __syncthreads();
  const int site = blockIdx.x * blockDim.x + threadIdx.x;
  if (site >= g.L) return;
  const int Rg = m->Rg;
  const int64_t Lp = g.Lp;
  double *d = glv_plane_w(g, 0, node, 0) + site;
  double *p = glv_plane_w(g, 1, node, 0) + site;
  double *sn = glv_plane_w(g, 2, node, 0) + site;
  const int64_t eo = (int64_t)node * Lp + site;
  const int ed = g.e[0][eo] + pull_exponent_site<S>(d, Lp, Rg, -10);
  const int ep = g.e[1][eo] + pull_exponent_site<S>(p, Lp, Rg, -10);
  g.e[0][eo] = ed;
  g.e[1][eo] = ep;
  for (int r = 0; r < Rg; ++r) {
    const double *Pr = P + (GCAT ? t.site_cat[site] : r) * RS;
    double sd[S], sp[S];
#pragma unroll
    for (int s = 0; s < S; ++s) { sd[s] = d[(int64_t)(r * S + s) * Lp]; sp[s] = p[(int64_t)(r * S + s) * Lp]; }
#pragma unroll
    for (int i = 0; i < S; ++i) {
      double a = 0.0, b = 0.0;
#pragma unroll
      for (int j = 0; j < S; ++j) { a = fma(Pr[i * S + j], sd[j], a); b = fma(Pr[i * S + j], sp[j], b); }
      sn[(int64_t)(r * S + i) * Lp] = m->statd[i] * a * b;  // linear_c.c:443-463: statd*a*b
    }
  }
  g.e[2][eo] = ed + ep;
}

}  // namespace ppb
