// Prefilter, query-lane log-table form (DNA, Gamma mixture with K = 4R = 16): the default for batches of >= 256 queries.
//
//   score(q, e) = sum over the informative sites of q of  LT[site][e][col(q, site)]  - n_q log R
//   LT[site][e][c] = log( sum_k Tq[c][k] * snode_e[k][site] ) + ln2 * e_snode[site][e],   c = A, C, G, T, N
// (Glv.masked_logdot, core.ml:187-194; ten_masked_logdot_c linear_c.c:465-539 sums log(util[site]) in exactly this way.)
//
// Why this shape (ncu r4e / r4g): the edge-lane table kernels spend ~40 warp instructions per (query, site, 64 edges) -
// code extraction, a warp-uniform branch per query, a switch, renormalisation - and one L1 request per (query, site).
// Here a LANE owns a QUERY and a warp owns 32 queries x 32 edges: per site the warp issues 32 shared-memory loads (lane
// l reads column col(q_l, site) of edge e; at most five distinct addresses in five distinct banks per request, so a
// request is a broadcast) and 32 DADDs for 1024 (query, edge, site) units, with NO branch: an uninformative site or
// a pad lane selects a sixth column that holds 0.  -inf (zero site likelihood) and NaN propagate through the additions
// as they do through the reference's sum of logs.  The two- and three-base ambiguity codes (rare) also select the zero
// column; prefilter_qlane_fix_kernel then adds their exact terms for the queries that have any.
//
// The table tile of a block ([site][32 edges][6 columns], 1.5 KB per site, contiguous in HBM per (edge tile, site range))
// is streamed by the TMA engine (cp.async.bulk + mbarrier ring, one producer warp); the queries are sorted by the window
// of sites they cover, so a warp walks the union of 32 similar windows.
#pragma once
#include "prefilter_staged.cuh"

namespace ppb {

constexpr int kQlEdges = 32, kQlCols = 6;

// LTq[et][site][e][6]: columns A, C, G, T, N, 0  (et = edge tile of 32 edges)
template <int K>
__global__ void __launch_bounds__(128)
qlane_table_kernel(const SnT sn, const TqFive tp, double *__restrict__ LT, int site0) {
  const int64_t Np = sn.Np;
  const int e = blockIdx.x * 128 + threadIdx.x;
  const int site = site0 + blockIdx.y;
  if (e >= Np) return;
  const double *col = sn.a + ((int64_t)site * K) * Np + e;
  double v[K];
#pragma unroll
  for (int k = 0; k < K; ++k) v[k] = __ldg(col + (int64_t)k * Np);
  const double ex = kLn2 * (double)__ldg(sn.e + (int64_t)site * Np + e);
  double *dst = LT + (((int64_t)(e / kQlEdges) * sn.L + site) * kQlEdges + (e % kQlEdges)) * kQlCols;
#pragma unroll
  for (int j = 0; j < 5; ++j) {
    double lo = tp.v[j][0] * v[0], hi = tp.v[j][K / 2] * v[K / 2];
#pragma unroll
    for (int k = 1; k < K / 2; ++k) { lo = fma(tp.v[j][k], v[k], lo); hi = fma(tp.v[j][K / 2 + k], v[K / 2 + k], hi); }
    dst[j] = log(lo + hi) + ex;
  }
  dst[5] = 0.0;
}

// column bytes of the sorted queries, four sites per word: qcol[group][site / 4][lane] (byte j = site 4*(site/4) + j).
// 0-4 = table column, 5 = adds nothing (uninformative, pad lane, or an ambiguity code left to the fix kernel).
__global__ void __launch_bounds__(256)
qlane_pack_kernel(const uint8_t *__restrict__ codes, const int32_t *__restrict__ order, uint32_t *__restrict__ qcol, int Q, int L,
                  int64_t Lp, int n_groups, int32_t *__restrict__ slow_flag) {
  const int s4 = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31, grp = blockIdx.y;
  const int n4 = (L + 3) / 4;
  if (s4 >= n4) return;
  const int slot = grp * 32 + lane;
  uint32_t w = 0x05050505u;
  if (slot < Q) {
    const int q = order[slot];
    const uint8_t *row = codes + (int64_t)q * Lp;
    w = 0;
    bool slow = false;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int site = s4 * 4 + j;
      unsigned c = 5;
      if (site < L) {
        const unsigned code = row[site];
        if (!(code & 0x80u)) {
          c = (unsigned)(kLogColOfCode >> (4 * (code & 15u))) & 15u;
          if (c == 15u) { c = 5; slow = true; }
        }
      }
      w |= c << (8 * j);
    }
    if (slow) slow_flag[q] = 1;
  }
  qcol[((int64_t)grp * n4 + s4) * 32 + lane] = w;
}

template <int W, int SPS, int NS, int NC = kQlCols>
struct QlaneGeom {
  static constexpr int SITE_BYTES = kQlEdges * NC * 8;  // 1536 (DNA, six columns); 5632 (proteins, 22 columns)
  static constexpr int STAGE_BYTES = SPS * SITE_BYTES;
  static constexpr int OFF_BAR = NS * STAGE_BYTES;
  static constexpr int OFF_TR = OFF_BAR + 2 * NS * 8 + 16;    // per-warp [32][33] transpose tile for the coalesced store
  static __host__ __device__ constexpr size_t smem_bytes() { return (size_t)OFF_TR + (size_t)W * 32 * 33 * 8; }
  static_assert(SPS % 4 == 0, "the column words hold four sites");
};

template <int W, int SPS, int NS, int MINB, int NC = kQlCols, bool GCAT = false>
__global__ void __launch_bounds__((W + 1) * 32, MINB)
prefilter_qlane_kernel(const PreArgs A, const double *__restrict__ LT, const uint32_t *__restrict__ qcol) {
  using G = QlaneGeom<W, SPS, NS, NC>;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw + G::OFF_BAR);
  uint64_t *empty = full + NS;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int L = A.sn.L;
  const int g0 = blockIdx.x * W;  // query groups (32 queries each) of this block
  int blo = 0x7fffffff, bhi = 0;
  for (int w = 0; w < W; ++w)
    if (g0 + w < A.nqt) { blo = min(blo, __ldg(A.tile_lo + g0 + w)); bhi = max(bhi, __ldg(A.tile_hi + g0 + w)); }
  const int z0 = blockIdx.z * A.sites_per_z, z1 = min(L, z0 + A.sites_per_z);
  blo = max(blo, z0) & ~3;  // stages start at a multiple of four sites (the column words)
  bhi = min(bhi, z1);
  const int n_stage = bhi > blo ? (bhi - blo + SPS - 1) / SPS : 0;
  if (threadIdx.x == 0) {
    for (int s = 0; s < NS; ++s) { stg::mbar_init(full + s, 1); stg::mbar_init(empty + s, W); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int et = blockIdx.y;
  const double *tile = LT + (int64_t)et * L * (kQlEdges * NC);

  if (warp == W) {  // ---- producer: one bulk copy per stage (the tile is contiguous over sites) ----
    if (lane == 0)
      for (int st = 0; st < n_stage; ++st) {
        const int slot = st % NS;
        if (st >= NS) stg::mbar_wait(empty + slot, ((st / NS) - 1) & 1);
        const int sb = blo + st * SPS;
        const uint32_t bytes = (uint32_t)min(SPS, L - sb) * G::SITE_BYTES;
        stg::mbar_expect_tx(full + slot, bytes);
        stg::bulk_g2s(smem_raw + (size_t)slot * G::STAGE_BYTES, tile + (int64_t)sb * (kQlEdges * NC), bytes, full + slot);
      }
    return;
  }

  // ---- consumers: lane = query ----
  const int grp = g0 + warp;
  const bool live = grp < A.nqt;
  const int wlo = live ? max(__ldg(A.tile_lo + grp), z0) : 0, whi = live ? min(__ldg(A.tile_hi + grp), z1) : 0;
  const int n4 = (L + 3) / 4;
  const uint32_t *qc = qcol + ((int64_t)(live ? grp : 0) * n4) * 32 + lane;
  double acc[kQlEdges];
#pragma unroll
  for (int e = 0; e < kQlEdges; ++e) acc[e] = 0.0;
  for (int st = 0; st < n_stage; ++st) {
    const int slot = st % NS;
    const int sb = blo + st * SPS;
    // this stage's column bytes (SPS / 4 words per lane), requested before the wait
    uint32_t cw[SPS / 4];
#pragma unroll
    for (int j = 0; j < SPS / 4; ++j) {
      const int s4 = (sb >> 2) + j;
      cw[j] = (live && s4 < n4) ? __ldg(qc + (int64_t)s4 * 32) : 0x01010101u * (unsigned)(NC - 1);  // the zero column is the last one in both layouts
    }
    stg::mbar_wait(full + slot, (st / NS) & 1);
    const double *stage = reinterpret_cast<const double *>(smem_raw + (size_t)slot * G::STAGE_BYTES);
    if (sb + SPS > wlo && sb < whi) {  // warp-uniform: the stage overlaps this warp's windows
#pragma unroll
      for (int j = 0; j < SPS / 4; ++j) {
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          const int site = sb + 4 * j + b;
          if (site < wlo || site >= whi) continue;  // warp-uniform: outside this z-slice / the warp's windows (or past the alignment)
          const unsigned c = (cw[j] >> (8 * b)) & 0xffu;
          const double *row = stage + (size_t)(4 * j + b) * (kQlEdges * NC) + c;
#pragma unroll
          for (int e = 0; e < kQlEdges; ++e) acc[e] += row[e * NC];
        }
      }
    }
    __syncwarp();
    if (lane == 0) stg::mbar_arrive(empty + slot);
  }
  if (!live) return;
  // ---- coalesced store through a per-warp transpose tile: row = query, 32 consecutive edges ----
  double *tr = reinterpret_cast<double *>(smem_raw + G::OFF_TR) + (size_t)warp * 32 * 33;
#pragma unroll
  for (int e = 0; e < kQlEdges; ++e) tr[lane * 33 + e] = acc[e];
  __syncwarp();
  double *outz = A.out + (int64_t)blockIdx.z * A.Q * A.sn.n_nodes;
  const int e = et * kQlEdges + lane;
  for (int r = 0; r < 32; ++r) {
    const int slot = grp * 32 + r;
    if (slot >= A.Q) break;
    const int q = __ldg(A.order + slot);
    if (e < A.sn.n_nodes) {
      double v = tr[r * 33 + lane];
      if (blockIdx.z == 0 && !GCAT) v -= (double)((float)__ldg(A.nq + q)) * log((double)((float)A.R));  // linear_c.c:536 (no such term in mat_masked_logdot_c)
      if (A.zs == 1 && e == A.root_id) v = __longlong_as_double(0x7ff8000000000000LL);
      outz[(int64_t)q * A.sn.n_nodes + e] = v;
    }
  }
}

// the exact terms of the two- and three-base ambiguity codes, for the queries flagged by qlane_pack_kernel: one thread
// per (flagged query, edge) walks the query's site list in order (deterministic sum).  grid = (Q, edge tiles of 128).
__global__ void __launch_bounds__(128)
prefilter_qlane_fix_kernel(const PreArgs A, const int32_t *__restrict__ qsite, int64_t Lp, const int32_t *__restrict__ slow_flag,
                           double *__restrict__ scores) {
  const int q = blockIdx.x;
  if (!slow_flag[q]) return;
  const int e = blockIdx.y * blockDim.x + threadIdx.x;
  if (e >= A.sn.n_nodes || e == A.root_id) return;
  const int n = A.nq[q];
  const int32_t *qs = qsite + (int64_t)q * Lp;
  double add = 0.0;
  for (int i = 0; i < n; ++i) {
    const int pk = qs[i];
    const unsigned code = ((unsigned)pk >> kSiteCodeShift) & 15u;
    if (((unsigned)(kLogColOfCode >> (4 * code)) & 15u) == 15u) add += slow_log_term(A, pk & kSiteMask, code, e);
  }
  scores[(int64_t)q * A.sn.n_nodes + e] += add;
}


// ---------------------------------------------------------------------------------------------------------------
// Pair-of-sites table (DNA Gamma-4, r5y).  ncu r5b and the r5x layout experiment: the query-lane form is bound by the
// BYTES the shared-memory pipe delivers to the lanes - 128 B per clock per SM, one 8-byte word per (query, edge, site)
// unit - not by requests or bank conflicts.  So two adjacent sites share one word:
//     LT2[et][pair][e][c1 * 6 + c2] = LT[2 pair][e][c1] + LT[2 pair + 1][e][c2]        (36 columns, column 35 = 0)
// and the same kernel (prefilter_qlane_kernel<..., NC = 36>) walks PAIRS: half the loads and half the additions per
// unit for three times the table (C4: 2.2 GB).  The pairwise sums reassociate the reference's sum of logs; -inf and
// NaN still propagate through plain additions.
// ---------------------------------------------------------------------------------------------------------------
constexpr int kQlPairCols = 36;

// one thread per (pair, edge): the five log columns of both sites, then the 36 sums
template <int K>
__global__ void __launch_bounds__(128)
qlane_pair_table_kernel(const SnT sn, const TqFive tp, double *__restrict__ LT2, int pair0) {
  const int64_t Np = sn.Np;
  const int e = blockIdx.x * 128 + threadIdx.x;
  const int pair = pair0 + blockIdx.y;
  if (e >= Np) return;
  double lg[2][6];
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int site = 2 * pair + h;
#pragma unroll
    for (int j = 0; j < 6; ++j) lg[h][j] = 0.0;
    if (site < sn.L) {
      const double *col = sn.a + ((int64_t)site * K) * Np + e;
      double v[K];
#pragma unroll
      for (int k = 0; k < K; ++k) v[k] = __ldg(col + (int64_t)k * Np);
      const double ex = kLn2 * (double)__ldg(sn.e + (int64_t)site * Np + e);
#pragma unroll
      for (int j = 0; j < 5; ++j) {
        double lo = tp.v[j][0] * v[0], hi = tp.v[j][K / 2] * v[K / 2];
#pragma unroll
        for (int k = 1; k < K / 2; ++k) { lo = fma(tp.v[j][k], v[k], lo); hi = fma(tp.v[j][K / 2 + k], v[K / 2 + k], hi); }
        lg[h][j] = log(lo + hi) + ex;
      }
    }
  }
  const int n_pairs = (sn.L + 1) / 2;
  double *dst = LT2 + (((int64_t)(e / kQlEdges) * n_pairs + pair) * kQlEdges + (e % kQlEdges)) * kQlPairCols;
#pragma unroll
  for (int c1 = 0; c1 < 6; ++c1)
#pragma unroll
    for (int c2 = 0; c2 < 6; ++c2) dst[c1 * 6 + c2] = lg[0][c1] + lg[1][c2];
}

// column bytes per PAIR, four pairs per word: byte = c1 * 6 + c2 (c = 5: uninformative, pad, ambiguity code -> fix kernel)
__global__ void __launch_bounds__(256)
qlane_pair_pack_kernel(const uint8_t *__restrict__ codes, const int32_t *__restrict__ order, uint32_t *__restrict__ qcol, int Q, int L,
                       int64_t Lp, int n_groups, int32_t *__restrict__ slow_flag) {
  const int p4 = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31, grp = blockIdx.y;
  const int n_pairs = (L + 1) / 2, n4 = (n_pairs + 3) / 4;
  if (p4 >= n4) return;
  const int slot = grp * 32 + lane;
  uint32_t w = 0x01010101u * (kQlPairCols - 1);
  if (slot < Q) {
    const int q = order[slot];
    const uint8_t *row = codes + (int64_t)q * Lp;
    w = 0;
    bool slow = false;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      unsigned cc[2];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int site = (p4 * 4 + j) * 2 + h;
        unsigned c = 5;
        if (site < L) {
          const unsigned code = row[site];
          if (!(code & 0x80u)) {
            c = (unsigned)(kLogColOfCode >> (4 * (code & 15u))) & 15u;
            if (c == 15u) { c = 5; slow = true; }
          }
        }
        cc[h] = c;
      }
      w |= (cc[0] * 6 + cc[1]) << (8 * j);
    }
    if (slow) slow_flag[q] = 1;
  }
  qcol[((int64_t)grp * n4 + p4) * 32 + lane] = w;
}

// tile windows in pair units
__global__ void qlane_pair_ranges_kernel(const int32_t *__restrict__ lo, const int32_t *__restrict__ hi, int n, int32_t *__restrict__ lo2,
                                         int32_t *__restrict__ hi2) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { lo2[i] = lo[i] >> 1; hi2[i] = (hi[i] + 1) >> 1; }
}

// ---------------------------------------------------------------------------------------------------------------
// The same form for any alphabet and model class (proteins, DNA with Price-CAT or a Gamma mixture other than four
// rates): NC = S + 2 columns - one per state, one for the "any" code (X / N: informative in the reference,
// alignment.ml:235), and a zero column (uninformative sites, pad lanes, the other multi-state codes).
// mat_masked_logdot_c (linear_c.c:251-296) / ten_masked_logdot_c (:465-539).
// ---------------------------------------------------------------------------------------------------------------
struct QlColMap {
  uint8_t col[kMaxCodesDev];    // table column of an alphabet code (255 = two- / three-state code: exact term in the fix kernel)
  uint8_t code_of_state[32];    // the alphabet code whose mask is exactly state s; [S] = the "any" code
};

// LTq[et][site][e][NC]; one thread per (site, edge), the S column sums run as S independent chains over (rate, state)
template <int S, bool GCAT, int NC>
__global__ void __launch_bounds__(128)
qlane_table_generic_kernel(const SnT sn, const double *__restrict__ Tq, const int32_t *__restrict__ site_cat, int nc, int Rg,
                           const QlColMap map, double *__restrict__ LT, int site0) {
  const int64_t Np = sn.Np;
  const int e = blockIdx.x * 128 + threadIdx.x;
  const int site = site0 + blockIdx.y;
  if (e >= Np) return;
  const int K = Rg * S;
  const double *col = sn.a + ((int64_t)site * K) * Np + e;
  const int cat = GCAT ? __ldg(site_cat + site) : 0;
  constexpr int NCOL = S + 1;  // the states and the "any" code
  double dot[NCOL];
#pragma unroll
  for (int c = 0; c < NCOL; ++c) dot[c] = 0.0;
  for (int r = 0; r < Rg; ++r) {
    double v[S];
#pragma unroll
    for (int s = 0; s < S; ++s) v[s] = __ldg(col + (int64_t)(r * S + s) * Np);
#pragma unroll
    for (int c = 0; c < NCOL; ++c) {
      const double *tq = GCAT ? Tq + ((int64_t)cat * nc + map.code_of_state[c]) * S : Tq + (int64_t)map.code_of_state[c] * K + r * S;
#pragma unroll
      for (int s = 0; s < S; ++s) dot[c] = fma(__ldg(tq + s), v[s], dot[c]);
    }
  }
  const double ex = kLn2 * (double)__ldg(sn.e + (int64_t)site * Np + e);
  double *dst = LT + (((int64_t)(e / kQlEdges) * sn.L + site) * kQlEdges + (e % kQlEdges)) * NC;
#pragma unroll
  for (int c = 0; c < NCOL; ++c) dst[c] = log(dot[c]) + ex;
#pragma unroll
  for (int c = NCOL; c < NC; ++c) dst[c] = 0.0;
}

__global__ void __launch_bounds__(256)
qlane_pack_generic_kernel(const uint8_t *__restrict__ codes, const int32_t *__restrict__ order, uint32_t *__restrict__ qcol, int Q, int L,
                          int64_t Lp, int n_groups, const QlColMap map, unsigned zero_col, int32_t *__restrict__ slow_flag) {
  const int s4 = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31, grp = blockIdx.y;
  const int n4 = (L + 3) / 4;
  if (s4 >= n4) return;
  const int slot = grp * 32 + lane;
  uint32_t w = 0x01010101u * zero_col;
  if (slot < Q) {
    const int q = order[slot];
    const uint8_t *row = codes + (int64_t)q * Lp;
    w = 0;
    bool slow = false;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int site = s4 * 4 + j;
      unsigned c = zero_col;
      if (site < L) {
        const unsigned code = row[site];
        if (!(code & 0x80u)) {
          c = map.col[code & (kMaxCodesDev - 1)];
          if (c == 255u) { c = zero_col; slow = true; }
        }
      }
      w |= c << (8 * j);
    }
    if (slow) slow_flag[q] = 1;
  }
  qcol[((int64_t)grp * n4 + s4) * 32 + lane] = w;
}

// exact terms of the multi-state codes for the flagged queries: one thread per (query, edge), sites in order
template <int S, bool GCAT>
__global__ void __launch_bounds__(128)
prefilter_qlane_fix_generic_kernel(const PreArgs A, const int32_t *__restrict__ qsite, int64_t Lp, const QlColMap map, int Rg,
                                   const int32_t *__restrict__ slow_flag, double *__restrict__ scores) {
  const int q = blockIdx.x;
  if (!slow_flag[q]) return;
  const int e = blockIdx.y * blockDim.x + threadIdx.x;
  if (e >= A.sn.n_nodes || e == A.root_id) return;
  const int n = A.nq[q];
  const int32_t *qs = qsite + (int64_t)q * Lp;
  const int K = Rg * S;
  double add = 0.0;
  for (int i = 0; i < n; ++i) {
    const int pk = qs[i];
    const unsigned code = ((unsigned)pk >> kSiteCodeShift) & 0x7fu;
    if (map.col[code & (kMaxCodesDev - 1)] != 255u) continue;
    const int site = pk & kSiteMask;
    const double *tq = GCAT ? A.Tq + ((int64_t)A.site_cat[site] * A.nc + code) * S : A.Tq + (int64_t)code * K;
    const double *col = A.sn.a + ((int64_t)site * K) * A.sn.Np + e;
    double dot = 0.0;
    for (int k = 0; k < K; ++k) dot = fma(tq[k], col[(int64_t)k * A.sn.Np], dot);
    add += log(dot) + kLn2 * (double)A.sn.e[(int64_t)site * A.sn.Np + e];
  }
  scores[(int64_t)q * A.sn.n_nodes + e] += add;
}

}  // namespace ppb
