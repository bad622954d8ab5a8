// Fig ranking on the device: Fig._enum_by_score (pplacer_src/fig.ml:120-138) consumed by play_ball (core.ml:283-318).
//
// The reference scores ONE representative edge per fig, sorts the figs by that score (stable, descending) and expands
// them lazily: the next GROUP is the next fig plus every following fig whose representative score lies within
// strike_box of its own; the group's edges that were not yielded before are scored, sorted (score descending, edge
// ascending) and handed to baseball one by one.  On the device the laziness is kept per WAVE of baseball (baseball.cuh):
// before a wave is planned, fig_need_kernel flags every unfinished query whose ranked-but-unplayed locations are fewer
// than the wave needs; the flagged queries get their next group (fig_expand_kernel: group range, union of the figs'
// edge lists through a per-query bitmap of yielded edges; fig_score_kernel: Glv.masked_logdot of the new
// (query, edge) pairs, one warp each; two library segmented sorts put the group in the reference's order - edge
// ascending, then stable by score descending - and fig_append_kernel adds it to the query's ranking).  No edge is scored that the reference would not have scored.
#pragma once
#include "baseball.cuh"
#include "device_common.cuh"

namespace ppb {

struct FigDev {
  const int32_t *ptr, *idx;  // CSR edge lists of the m figs (host order: Fig._figs_of_gtree)
  int m;
};

struct FigState {
  const int32_t *order;   // [Q][m] figs sorted by representative score (stable, descending)
  const double *rep_sc;   // [Q][m] representative scores (fig index order)
  int32_t *fig_next;      // [Q] next fig in the order
  int32_t *n_ranked;      // [Q] entries of rank_loc available
  int32_t *n_new;         // [Q] size of the group being appended
  int32_t *need;          // [Q] flag
  uint32_t *yielded;      // [Q][words] bitmap of edges already yielded (Fig.uniquifier, fig.ml:109-113)
  int words;
};

__global__ void fig_init_kernel(FigState f, int Q) {
  const int64_t n = (int64_t)Q * f.words;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) f.yielded[i] = 0u;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < Q; i += (int64_t)gridDim.x * blockDim.x) {
    f.fig_next[i] = 0; f.n_ranked[i] = 0; f.n_new[i] = 0; f.need[i] = 0;
  }
}

// which unfinished queries must expand before the next wave can be planned (same `need` as ball_plan_kernel)
__global__ void fig_need_kernel(BallState b, BallPrefs bp, FigState f, int m, int Q, int *__restrict__ count) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= Q) return;
  f.need[q] = 0;
  if (b.done[q]) return;
  long long need = (long long)bp.t_max_strikes - b.n_strikes[q];
  const long long by_pitch = (long long)bp.t_max_pitches + 1 - b.n_results[q];
  if (by_pitch < need) need = by_pitch;
  if (need < 1) need = 1;
  if (f.n_ranked[q] - b.next_rank[q] < need && f.fig_next[q] < m) { f.need[q] = 1; atomicAdd(count, 1); }
}

// one block per query: the next group of figs -> new (query, edge) pairs appended (unsorted) to rank_loc
__global__ void __launch_bounds__(128)
fig_expand_kernel(FigState f, FigDev fd, double strike_box, int32_t *__restrict__ rank_loc, int n_nodes, int Q,
                  int *__restrict__ n_pairs) {
  const int q = blockIdx.x;
  if (q >= Q || !f.need[q]) return;
  __shared__ int s_g0, s_g1, s_cnt;
  const int32_t *ord = f.order + (int64_t)q * fd.m;
  const double *sc = f.rep_sc + (int64_t)q * fd.m;
  if (threadIdx.x == 0) {
    int g = f.fig_next[q];
    const double s0 = sc[ord[g]];
    s_g0 = g++;
    while (g < fd.m && fabs(s0 - sc[ord[g]]) < strike_box) ++g;  // approx_equal ~epsilon:strike_box (ppatteries.ml:96)
    s_g1 = g;
    f.fig_next[q] = g;
    s_cnt = 0;
  }
  __syncthreads();
  uint32_t *bits = f.yielded + (int64_t)q * f.words;
  int32_t *dst = rank_loc + (int64_t)q * n_nodes + f.n_ranked[q];
  for (int g = s_g0; g < s_g1; ++g) {
    const int fi = ord[g];
    for (int k = fd.ptr[fi] + threadIdx.x; k < fd.ptr[fi + 1]; k += blockDim.x) {
      const int e = fd.idx[k];
      const uint32_t bit = 1u << (e & 31);
      if (!(atomicOr(bits + (e >> 5), bit) & bit)) dst[atomicAdd(&s_cnt, 1)] = e;
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) { f.n_new[q] = s_cnt; atomicAdd(n_pairs, s_cnt); }
}
// the new members of every flagged query as one flat list of (query, edge) pairs, a segment per query
__global__ void __launch_bounds__(128)
fig_pairs_kernel(FigState f, const int32_t *__restrict__ rank_loc, int n_nodes, int Q, int32_t *__restrict__ pair_q,
                 int32_t *__restrict__ pair_e, int32_t *__restrict__ seg_begin, int32_t *__restrict__ seg_end, int *__restrict__ cursor) {
  const int q = blockIdx.x;
  if (q >= Q) return;
  __shared__ int s_base;
  const int n = f.need[q] ? f.n_new[q] : 0;
  if (threadIdx.x == 0) { s_base = n > 0 ? atomicAdd(cursor, n) : 0; seg_begin[q] = s_base; seg_end[q] = s_base + n; }
  __syncthreads();
  const int32_t *src = rank_loc + (int64_t)q * n_nodes + f.n_ranked[q];
  for (int i = threadIdx.x; i < n; i += blockDim.x) { pair_q[s_base + i] = q; pair_e[s_base + i] = src[i]; }
}

// Glv.masked_logdot of the evolved query against one supernode (gmix_model.ml:194-198 / gcat_model.ml:220-224;
// ten_masked_logdot_c linear_c.c:465-539, mat_masked_logdot_c :251-296): one warp per (query, edge) pair, lanes stride
// the query's informative sites, result into scores[q][edge].
template <int S, bool GCAT>
__global__ void __launch_bounds__(128)
fig_score_kernel(const DevModel *__restrict__ m, GlvArrays g, const double *__restrict__ Tq, const int32_t *__restrict__ site_cat,
                 const int32_t *__restrict__ qsite, const int32_t *__restrict__ nq, const int32_t *__restrict__ pair_q,
                 const int32_t *__restrict__ pair_e, int n_pairs, double *__restrict__ scores) {
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (w >= n_pairs) return;
  const int q = pair_q[w], edge = pair_e[w], n = nq[q];
  const int Rg = m->Rg, K = m->K, nc = m->n_codes;
  const int32_t *qs = qsite + (int64_t)q * g.Lp;
  const double *sn = g.a[2] + (int64_t)edge * g.a_stride;
  const int32_t *se = g.e[2] + (int64_t)edge * g.Lp;
  LogProd lp;
  lp.init();
  long long es = 0;
  for (int i = lane; i < n; i += 32) {
    const int pk = qs[i], site = pk & kSiteMask, code = (pk >> kSiteCodeShift) & 0x7f;
    double u = 0.0;
    if (GCAT) {
      const double *tv = Tq + ((int64_t)site_cat[site] * nc + code) * S;
#pragma unroll
      for (int s = 0; s < S; ++s) u = fma(tv[s], sn[(int64_t)s * g.Lp + site], u);
    } else {
      const double *tv = Tq + (int64_t)code * K;
      for (int r = 0; r < Rg; ++r)
#pragma unroll
        for (int s = 0; s < S; ++s) u = fma(tv[r * S + s], sn[(int64_t)(r * S + s) * g.Lp + site], u);
    }
    lp.mul(u);
    es += se[site];
  }
  double v = lp.value() + kLn2 * (double)es;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if (lane == 0) {
    if (!GCAT) v -= (double)((float)n) * log((double)((float)m->R));  // linear_c.c:536
    scores[(int64_t)q * g.n_nodes + edge] = v;
  }
}

// keys of the new members for the library's stable segmented sort (score descending, nan last); the members of a
// segment arrive in ascending edge order (first sort), so ties keep the reference's order (IntSet.elements, fig.ml:133)
__global__ void fig_keys_kernel(const int32_t *__restrict__ pair_q, const int32_t *__restrict__ pair_e, int n_pairs,
                                const double *__restrict__ scores, int n_nodes, unsigned long long *__restrict__ keys) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_pairs) return;
  const double s = scores[(int64_t)pair_q[i] * n_nodes + pair_e[i]];
  keys[i] = rank_key_bits(s, s != s);
}
// the sorted group goes to the end of the query's ranking
__global__ void __launch_bounds__(128)
fig_append_kernel(FigState f, const int32_t *__restrict__ seg_begin, const int32_t *__restrict__ sorted_e,
                  int32_t *__restrict__ rank_loc, int n_nodes, int Q) {
  const int q = blockIdx.x;
  if (q >= Q || !f.need[q]) return;
  const int n = f.n_new[q];
  int32_t *dst = rank_loc + (int64_t)q * n_nodes + f.n_ranked[q];
  const int32_t *src = sorted_e + seg_begin[q];
  for (int i = threadIdx.x; i < n; i += blockDim.x) dst[i] = src[i];
  __syncthreads();
  if (threadIdx.x == 0) { f.n_ranked[q] += n; f.n_new[q] = 0; }
}

// representative columns of the edge-major supernode arrays: out[site][k][j] = in[site][k][rep[j]], pads = fill
template <class T>
__global__ void gather_rep_columns_kernel(const T *__restrict__ in, T *__restrict__ out, const int32_t *__restrict__ rep, int m,
                                          int64_t rows, int64_t Np_in, int64_t Np_out, T fill) {
  const int64_t total = rows * Np_out;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = i / Np_out;
    const int j = (int)(i - row * Np_out);
    out[i] = j < m ? in[row * Np_in + rep[j]] : fill;
  }
}

}  // namespace ppb
