// Device-side bookkeeping of the second stage: baseball (Core play_ball), keep filter, weights.
//   core.ml:283-318  play_ball: walk the ranking; strikes / pitches / record
//   core.ml:336-352  sort by likelihood, keep_at_most / keep_factor
//   core.ml:354-377  refinement results, ll_normalized_prob (ppatteries.ml:175-183)
//   core.ml:379-408  --pp, Pquery.make_ml_sorted (pquery.ml:86-92)
// Every candidate starts from the same default branch lengths (core.ml:240-253), so results do not
// depend on the order in which candidates are optimised: the ranking is consumed in WAVES.  A wave
// takes, per unfinished query, exactly the number of further results that must exist before any
// stop rule can fire (min(max_strikes - strikes, max_pitches + 1 - results)), optimises all of them
// in one launch, then replays the reference's sequential rules over them in rank order.  No
// candidate is ever optimised that the reference would not have optimised.
#pragma once
#include "device_common.cuh"
#include "threetax.cuh"

namespace ppb {

struct BallState {
  // per query
  int32_t *next_rank, *n_results, *n_strikes, *done, *wave_begin, *wave_n, *qstatus;
  double *like_record;
  // per query x P result slots, in play order
  int32_t *res_loc;
  double *res_ll, *res_pend, *res_dist;
  int64_t *n_calls;
  int P;  // slots per query
};

struct BallPrefs {
  int t_max_pitches, t_max_strikes;
  double strike_box;
  int keep_at_most;
  double log_keep_factor;
};

__global__ void ball_init_kernel(BallState b, const int32_t *__restrict__ prep_status, int Q) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= Q) return;
  b.next_rank[q] = 0; b.n_results[q] = 0; b.n_strikes[q] = 0;
  b.like_record[q] = -__longlong_as_double(0x7ff0000000000000LL);
  b.qstatus[q] = prep_status[q];
  b.done[q] = prep_status[q] != 0;
  b.wave_n[q] = 0; b.wave_begin[q] = 0; b.n_calls[q] = 0;
}

// ranking: rank_loc[q][n_nodes] = edge ids sorted by score descending (stable); entries equal to
// root_id are skipped (the root is never scored: fig.ml:103, stree.ml:33-35).
__global__ void ball_plan_kernel(BallState b, BallPrefs bp, const int32_t *__restrict__ rank_loc, int n_nodes, int n_rank,
                                 int root_id, int Q, CandIn *__restrict__ cand, int *__restrict__ wave_count,
                                 int *__restrict__ exhausted, const int32_t *__restrict__ n_ranked = nullptr) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= Q || b.done[q]) { if (q < Q) b.wave_n[q] = 0; return; }
  if (n_ranked) n_rank = n_ranked[q];  // fig ranking (fig.cuh): the ranking grows group by group, per query
  long long need = (long long)bp.t_max_strikes - b.n_strikes[q];
  const long long by_pitch = (long long)bp.t_max_pitches + 1 - b.n_results[q];
  if (by_pitch < need) need = by_pitch;
  if (need < 1) need = 1;
  const int32_t *rk = rank_loc + (int64_t)q * n_nodes;
  // count how many non-root ranks remain, up to need
  int r = b.next_rank[q], cnt = 0;
  for (int rr = r; rr < n_rank && cnt < need; ++rr) cnt += (rk[rr] != root_id);
  if (cnt == 0) {  // ranking used up (with a truncated ranking the caller must redo the batch on the full one)
    if (!n_ranked && n_rank < n_nodes) atomicAdd(exhausted, 1);
    b.done[q] = 1; b.wave_n[q] = 0;
    return;
  }
  const int base = atomicAdd(wave_count, cnt);
  b.wave_begin[q] = base; b.wave_n[q] = cnt;
  int k = 0;
  while (k < cnt) {
    const int loc = rk[r++];
    if (loc == root_id) continue;
    CandIn c; c.q = q; c.loc = loc; c.pend0 = 0.0; c.dist0 = 0.0;
    cand[base + k++] = c;
  }
  b.next_rank[q] = r;
}

__global__ void ball_update_kernel(BallState b, BallPrefs bp, const CandIn *__restrict__ cand,
                                   const CandOut *__restrict__ out, int n_rank_total, int n_nodes, int Q, int *__restrict__ n_active,
                                   int *__restrict__ exhausted, const int32_t *__restrict__ n_ranked = nullptr,
                                   const int32_t *__restrict__ fig_next = nullptr, int n_figs = 0) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= Q || b.done[q]) return;
  int nres = b.n_results[q], strikes = b.n_strikes[q];
  double record = b.like_record[q];
  bool stop = false;
  const int base = b.wave_begin[q];
  for (int j = 0; j < b.wave_n[q] && !stop; ++j) {
    const CandOut o = out[base + j];
    b.n_calls[q] += o.calls;
    if (st_is_gsl(o.status)) continue;  // skipped with a warning (core.ml:291-297)
    if (o.status != ST_OK) { b.qstatus[q] = o.status; stop = true; break; }  // uncaught in the reference
    const int len_before = nres;
    if (nres < b.P) {
      const int64_t s = (int64_t)q * b.P + nres;
      b.res_loc[s] = cand[base + j].loc; b.res_ll[s] = o.ll; b.res_pend[s] = o.pend; b.res_dist[s] = o.dist;
    }
    ++nres;
    if (len_before >= bp.t_max_pitches) stop = true;
    else if (o.ll > record) record = o.ll;
    else if (o.ll < record - bp.strike_box) {
      if (strikes + 1 >= bp.t_max_strikes) stop = true; else ++strikes;
    }
  }
  b.n_results[q] = nres; b.n_strikes[q] = strikes; b.like_record[q] = record;
  if (n_ranked) {  // fig ranking: used up only when every fig has been expanded
    if (b.next_rank[q] >= n_ranked[q] && fig_next[q] >= n_figs) stop = true;
  } else if (b.next_rank[q] >= n_rank_total) {
    if (!stop && n_rank_total < n_nodes) atomicAdd(exhausted, 1);
    stop = true;
  }
  if (stop) b.done[q] = 1; else atomicAdd(n_active, 1);
}

struct KeepState {
  int32_t *keep_begin, *n_keep;  // per query
};

// core.ml:336-352: stable sort by likelihood descending, keep filter; emits refinement candidates
// starting from the first-round optimum (core.ml:357-358).  One thread per query (<= 41 results in
// the default mode).
__global__ void keep_select_kernel(BallState b, BallPrefs bp, KeepState ks, int Q, CandIn *__restrict__ cand,
                                   int *__restrict__ count, int empty_status) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= Q) return;
  ks.n_keep[q] = 0; ks.keep_begin[q] = 0;
  if (b.qstatus[q] != 0) return;
  const int n = min(b.n_results[q], b.P);
  if (n == 0) { b.qstatus[q] = empty_status; return; }
  const int64_t o = (int64_t)q * b.P;
  // insertion sort, descending, stable (List.sort is a stable merge sort; compare on floats puts nan last)
  for (int i = 1; i < n; ++i) {
    const int loc = b.res_loc[o + i];
    const double ll = b.res_ll[o + i], pe = b.res_pend[o + i], di = b.res_dist[o + i];
    int j = i - 1;
    while (j >= 0 && (b.res_ll[o + j] < ll || (b.res_ll[o + j] != b.res_ll[o + j] && ll == ll))) {
      b.res_loc[o + j + 1] = b.res_loc[o + j]; b.res_ll[o + j + 1] = b.res_ll[o + j];
      b.res_pend[o + j + 1] = b.res_pend[o + j]; b.res_dist[o + j + 1] = b.res_dist[o + j];
      --j;
    }
    b.res_loc[o + j + 1] = loc; b.res_ll[o + j + 1] = ll; b.res_pend[o + j + 1] = pe; b.res_dist[o + j + 1] = di;
  }
  const double best = b.res_ll[o];
  int nk = 0;
  for (int i = 0; i < n; ++i)
    if (i < bp.keep_at_most && b.res_ll[o + i] >= bp.log_keep_factor + best) ++nk;
  const int base = atomicAdd(count, nk);
  ks.keep_begin[q] = base; ks.n_keep[q] = nk;
  int k = 0;
  for (int i = 0; i < n; ++i)
    if (i < bp.keep_at_most && b.res_ll[o + i] >= bp.log_keep_factor + best) {
      CandIn c; c.q = q; c.loc = b.res_loc[o + i]; c.pend0 = b.res_pend[o + i]; c.dist0 = b.res_dist[o + i];
      cand[base + k++] = c;
    }
}

struct PlaceOut {
  int32_t edge, _pad;
  double log_like, pendant, distal, ml_ratio, marginal_like, post_prob;
};

// core.ml:354-377: take the refined results (or the initial ones where GSL failed twice), compute
// ml_ratio; when --pp follows, rewrite the candidate list with the refined branch lengths
// (tt_edges_from_placement, core.ml:246-251).
__global__ void refine_finish_kernel(BallState b, KeepState ks, int Q, CandIn *__restrict__ cand,
                                     const CandOut *__restrict__ out, PlaceOut *__restrict__ res, int keep_at_most,
                                     int32_t *__restrict__ n_out) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= Q) return;
  n_out[q] = 0;
  if (b.qstatus[q] != 0) return;
  const int nk = ks.n_keep[q], base = ks.keep_begin[q];
  PlaceOut *r = res + (int64_t)q * keep_at_most;
  for (int i = 0; i < nk; ++i) {
    const CandOut o = out[base + i];
    const CandIn c = cand[base + i];
    b.n_calls[q] += o.calls;
    r[i].edge = c.loc; r[i]._pad = 0;
    if (st_is_gsl(o.status)) { r[i].pendant = c.pend0; r[i].distal = c.dist0; r[i].log_like = __longlong_as_double(0x7ff8000000000000LL); }
    else {
      if (o.status != ST_OK) { b.qstatus[q] = o.status; return; }
      r[i].pendant = o.pend; r[i].distal = o.dist; r[i].log_like = o.ll;
    }
    r[i].marginal_like = r[i].post_prob = __longlong_as_double(0x7ff8000000000000LL);
  }
  // a location whose refinement failed keeps its first-round tuple, likelihood included
  // (core.ml:361-368 returns `initial`): find it back in the sorted first-round results
  for (int i = 0; i < nk; ++i)
    if (r[i].log_like != r[i].log_like) {
      const int64_t o = (int64_t)q * b.P;
      for (int j = 0; j < min(b.n_results[q], b.P); ++j)
        if (b.res_loc[o + j] == r[i].edge) { r[i].log_like = b.res_ll[o + j]; break; }
    }
  for (int i = 0; i < nk; ++i) {  // ll_normalized_prob
    double tot = 0.0;
    for (int j = 0; j < nk; ++j) tot += exp(r[j].log_like - r[i].log_like);
    r[i].ml_ratio = 1.0 / tot;
    cand[base + i].pend0 = r[i].pendant;
    cand[base + i].dist0 = r[i].distal;
  }
  n_out[q] = nk;
}

// core.ml:385-404 + pquery.ml:86-92: attach marginal likelihoods / posterior probabilities, then
// stable-sort the placements by ml_ratio descending.
__global__ void final_sort_kernel(BallState b, KeepState ks, int Q, const CandOut *__restrict__ pp_out, int calc_pp,
                                  PlaceOut *__restrict__ res, int keep_at_most, int32_t *__restrict__ n_out,
                                  int pp_failed_status) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= Q) return;
  PlaceOut *r = res + (int64_t)q * keep_at_most;
  PlaceOut zero;
  zero.edge = 0; zero._pad = 0; zero.log_like = zero.pendant = zero.distal = zero.ml_ratio = zero.marginal_like = zero.post_prob = 0.0;
  int nk = b.qstatus[q] != 0 ? 0 : n_out[q];
  const int base = ks.keep_begin[q];
  if (calc_pp) {
    for (int i = 0; i < nk; ++i) {
      const CandOut o = pp_out[base + i];
      if (o.status != ST_OK) { b.qstatus[q] = (o.status == ST_EXCEEDED_MAX_ITER) ? pp_failed_status : o.status; nk = 0; break; }
      r[i].marginal_like = o.marg;
    }
  }
  // unused record slots are zero-filled so the output bytes are a pure function of the inputs
  for (int i = nk; i < keep_at_most; ++i) r[i] = zero;
  n_out[q] = nk;
  if (calc_pp) {
    for (int i = 0; i < nk; ++i) {
      double tot = 0.0;
      for (int j = 0; j < nk; ++j) tot += exp(r[j].marginal_like - r[i].marginal_like);
      r[i].post_prob = 1.0 / tot;
    }
  }
  for (int i = 1; i < nk; ++i) {
    const PlaceOut v = r[i];
    int j = i - 1;
    while (j >= 0 && (r[j].ml_ratio < v.ml_ratio || (r[j].ml_ratio != r[j].ml_ratio && v.ml_ratio == v.ml_ratio))) { r[j + 1] = r[j]; --j; }
    r[j + 1] = v;
  }
}

__global__ void iota_segments_kernel(int32_t *__restrict__ v, int n_nodes, int64_t total) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x)
    v[i] = (int32_t)(i % n_nodes);
}
// sort keys: the root and NaN scores sort last (-inf), like OCaml's compare puts nan below everything
__global__ void rank_keys_kernel(const double *__restrict__ scores, double *__restrict__ keys, int n_nodes, int root_id,
                                 int64_t total) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const double s = scores[i];
    const bool bad = (i % n_nodes) == root_id || s != s;
    keys[i] = bad ? -__longlong_as_double(0x7ff0000000000000LL) : s;
  }
}


// Truncated ranking: the first KTOP entries of Fig.enum_by_score's stable descending order
// (fig.ml:143-148), which is all baseball ever reads unless a query runs through more than KTOP
// locations (then the caller falls back to the full segmented sort).  One block per query: an
// eight-pass MSB radix select finds the KTOP-th largest key, ties at the threshold are taken in
// ascending edge order (stability), the survivors are sorted in shared memory.
constexpr int kTopK = 256;
__device__ __forceinline__ unsigned long long rank_key_bits(double s, bool bad) {
  if (bad) s = -__longlong_as_double(0x7ff0000000000000LL);
  const unsigned long long u = (unsigned long long)__double_as_longlong(s);
  return (u >> 63) ? ~u : (u | 0x8000000000000000ULL);  // monotone: larger score <=> larger key
}
__global__ void __launch_bounds__(256)
topk_rank_kernel(const double *__restrict__ scores, int32_t *__restrict__ rank_loc, int n_nodes, int root_id) {
  __shared__ unsigned hist[256];
  __shared__ unsigned long long s_prefix;
  __shared__ int s_remaining, s_gt, s_eq_base;
  __shared__ unsigned long long skey[kTopK];
  __shared__ int sidx[kTopK];
  __shared__ int warp_cnt[8];
  const int q = blockIdx.x, tid = threadIdx.x;
  const double *sc = scores + (int64_t)q * n_nodes;
  if (tid == 0) { s_prefix = 0; s_remaining = kTopK; }
  for (int pass = 0; pass < 8; ++pass) {
    const int shift = 56 - 8 * pass;
    hist[tid] = 0;
    __syncthreads();
    const unsigned long long prefix = s_prefix;
    for (int i = tid; i < n_nodes; i += 256) {
      const double v = sc[i];
      const unsigned long long k = rank_key_bits(v, i == root_id || v != v);
      if (pass == 0 || (k >> (shift + 8)) == (prefix >> (shift + 8))) atomicAdd(&hist[(unsigned)(k >> shift) & 255u], 1u);
    }
    __syncthreads();
    if (tid == 0) {
      int need = s_remaining, cum = 0, bsel = 0;
      for (int bb = 255; bb >= 0; --bb) {
        if (cum + (int)hist[bb] >= need) { bsel = bb; break; }
        cum += (int)hist[bb];
      }
      s_prefix = prefix | ((unsigned long long)bsel << shift);
      s_remaining = need - cum;  // still to take from the selected bin
    }
    __syncthreads();
  }
  const unsigned long long T = s_prefix;  // the KTOP-th largest key
  const int need_eq = s_remaining;        // how many entries equal to T belong to the top KTOP
  if (tid == 0) { s_gt = 0; s_eq_base = 0; }
  __syncthreads();
  for (int base = 0; base < n_nodes; base += 256) {  // ascending edge order: ties keep the smallest ids
    const int i = base + tid;
    unsigned long long k = 0;
    bool gt = false, eq = false;
    if (i < n_nodes) {
      const double v = sc[i];
      k = rank_key_bits(v, i == root_id || v != v);
      gt = k > T; eq = k == T;
    }
    const unsigned bal = __ballot_sync(0xffffffffu, eq);
    if ((tid & 31) == 0) warp_cnt[tid >> 5] = __popc(bal);
    __syncthreads();
    int before = s_eq_base;
    for (int w = 0; w < (tid >> 5); ++w) before += warp_cnt[w];
    const int my_eq_rank = before + __popc(bal & ((1u << (tid & 31)) - 1u));
    if (gt) { const int slot = atomicAdd(&s_gt, 1); skey[slot] = k; sidx[slot] = i; }
    if (eq && my_eq_rank < need_eq) { const int slot = kTopK - 1 - my_eq_rank; skey[slot] = k; sidx[slot] = i; }
    __syncthreads();
    if (tid == 0) { int t = 0; for (int w = 0; w < 8; ++w) t += warp_cnt[w]; s_eq_base += t; }
    __syncthreads();
  }
  // bitonic sort of the KTOP (key, id) pairs: key descending, id ascending
  unsigned long long mk = skey[tid];
  int mi = sidx[tid];
  for (int size = 2; size <= kTopK; size <<= 1)
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      __syncthreads();
      skey[tid] = mk; sidx[tid] = mi;
      __syncthreads();
      const int partner = tid ^ stride;
      const unsigned long long pk = skey[partner];
      const int pi = sidx[partner];
      const bool up = (tid & size) == 0;               // this block of the network sorts "forward"
      const bool mine_first = mk > pk || (mk == pk && mi < pi);  // (key desc, id asc) order
      const bool keep_mine = ((tid < partner) == up) ? mine_first : !mine_first;
      if (!keep_mine) { mk = pk; mi = pi; }
    }
  rank_loc[(int64_t)q * n_nodes + tid] = mi;
}

// FP64 FMA throughput probe: 8 independent chains per thread.
__global__ void dfma_probe_kernel(double *out, int iters, double a, double b) {
  double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
  for (int i = 0; i < iters; ++i) {
    x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
    x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

}  // namespace ppb
