// The baseball prefilter (first-stage scoring of every edge for every query).
//   core.ml:187-194   score loc = Glv.masked_logdot full_query_evolv snodes.(loc) mask_vec
//   gmix_model.ml:194-198 -> ten_masked_logdot_c  linear_c.c:465-539  (+ masked_total_twoexp, gstar_support.ml:52-59)
//   gcat_model.ml:220-224 -> mat_masked_logdot_c  linear_c.c:251-296
//     score(q,e) = sum_{site in mask_q} log( sum_{r,s} Qev[r,site,s] * snode_e[r,site,s] )
//                  - n_q log R + ln2 * sum_{site in mask_q} e_snode[site]       (query exponents are 0)
//
// Work decomposition: LANES WALK EDGES, a warp owns QT queries, and the site loop is
// warp-uniform.  The supernodes are therefore kept in a prefilter layout snT[site][k][edge] (edge
// fastest): for one (site, k) a warp reads 32*TE consecutive edges with one coalesced LDG.64/128 and
// keeps the K values of its TE edges in registers while all QT queries of the warp consume them.
// Every quantity that depends on the query (its character at the site, hence the evolved-query
// column) is warp-uniform, so it costs no per-lane shared-memory bandwidth: in the DNA Gamma-mixture
// case the evolved-query table for the four bases lives in the kernel-parameter constant bank and is
// an immediate operand of the DFMAs.  Each thread owns its (edge, query) running products over ALL
// the sites it is given: there is no cross-lane reduction at all.
//
// Three kernels share this skeleton:
//   prefilter_direct_kernel   the FMA form: K = Rg*S DFMAs per (query, edge, informative site).
//   mtable_kernel             M[site][c][edge] = sum_k Tq[c][k] * snT[site][k][edge] for the 4 bases,
//   prefilter_table_kernel    ... after which a unit is one select + one DMUL (DNA, Gamma mixture).
#pragma once
#include "device_common.cuh"

namespace ppb {

// supernodes in prefilter layout
struct SnT {
  const double *a;   // [L][K][Np]
  const int32_t *e;  // [L][Np]
  int64_t Np;        // n_nodes rounded up to 64; pad edges hold 1.0 / exponent 0
  int L, K, n_nodes;
};

// evolved-query columns of the four bases for mtable_kernel, passed BY VALUE so they sit in the
// parameter constant bank: v[j][k] = Tq[code = 1<<j][k]   (DNA Gamma mixture, K = 4R <= 16)
struct TqPrimary {
  double v[4][16];
};

struct PreArgs {
  SnT sn;
  const uint8_t *qc;        // packed query codes [n_qtiles][L][QT]: code | 0x80 if uninformative, 0xFF pad
  const double *Tq;         // evolved-query table, global (Gmix [code][K]; Gcat [cat][code][S])
  const int32_t *site_cat;  // Gcat
  const int32_t *nq;        // [Q]
  double *out;              // zs == 1: scores[Q][n_nodes]; else partial[zs][Q][n_nodes]
  const int32_t *bad_edge;  // table path: edges whose table left the safe range (or null)
  const int32_t *order;     // [Q] tile slot -> query: queries sorted by the window of sites they cover (window_keys_kernel)
  const int32_t *tile_lo;   // [n_qtiles] first / one-past-last site at which ANY query of the tile is informative
  const int32_t *tile_hi;
  int Q, nqt, root_id, R, nc, zs, sites_per_z;
};

// Reads cover a WINDOW of the alignment (an amplicon, a shotgun read): a tile of queries with similar windows walks
// only the union of its windows instead of all L columns, and nearly every site it visits is informative for every
// query of the tile.  Sort key: (first informative site / 64, last informative site); queries without informative
// sites go last.  Scores do not depend on the tiling (every (query, edge) product is private to one thread and the
// renormalisations are exact powers of two), so placements are independent of the order (tested).
__global__ void window_keys_kernel(const int32_t *__restrict__ qsite, const int32_t *__restrict__ nq, int Q, int64_t Lp,
                                   unsigned long long *__restrict__ keys, int32_t *__restrict__ vals) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= Q) return;
  const int n = nq[q];
  unsigned long long key = ~0ull;
  if (n > 0) {
    const unsigned first = (unsigned)(qsite[(int64_t)q * Lp] & kSiteMask), last = (unsigned)(qsite[(int64_t)q * Lp + n - 1] & kSiteMask);
    key = ((unsigned long long)(first >> 6) << 24) | last;
  }
  keys[q] = key;
  vals[q] = q;
}
template <int QT>
__global__ void tile_range_kernel(const int32_t *__restrict__ qsite, const int32_t *__restrict__ nq, const int32_t *__restrict__ order,
                                  int Q, int nqt, int64_t Lp, int32_t *__restrict__ lo, int32_t *__restrict__ hi) {
  const int qt = blockIdx.x * blockDim.x + threadIdx.x;
  if (qt >= nqt) return;
  int a = 0x7fffffff, b = 0;
  for (int i = 0; i < QT; ++i) {
    const int slot = qt * QT + i;
    if (slot >= Q) break;
    const int q = order[slot], n = nq[q];
    if (n <= 0) continue;
    a = min(a, qsite[(int64_t)q * Lp] & kSiteMask);
    b = max(b, (qsite[(int64_t)q * Lp + n - 1] & kSiteMask) + 1);
  }
  lo[qt] = min(a, b);
  hi[qt] = b;
}

// ---- layout kernels ---------------------------------------------------------------------------
// [n_nodes][K][Lp] (site fastest) -> [L][K][Np] (edge fastest); pad edges get `fill`.
template <class T>
__global__ void __launch_bounds__(256)
transpose_to_edge_major_kernel(const T *__restrict__ in, T *__restrict__ out, int n_nodes, int K, int L, int64_t Lp,
                               int64_t Np, T fill) {
  __shared__ T tile[32][33];
  const int k = blockIdx.z;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int node = blockIdx.y * 32 + ty + 8 * j, site = blockIdx.x * 32 + tx;
    tile[ty + 8 * j][tx] = (node < n_nodes && site < L) ? in[((int64_t)node * K + k) * Lp + site] : fill;
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int site = blockIdx.x * 32 + ty + 8 * j, node = blockIdx.y * 32 + tx;
    if (site < L && node < Np) out[((int64_t)site * K + k) * Np + node] = tile[tx][ty + 8 * j];
  }
}

// codes[Q][Lp] -> qc[n_qtiles][L][QT]
template <int QT>
__global__ void __launch_bounds__(256)
pack_codes_kernel(const uint8_t *__restrict__ codes, const int32_t *__restrict__ order, uint8_t *__restrict__ qc, int Q, int L, int64_t Lp) {
  const int site = blockIdx.x * blockDim.x + threadIdx.x;
  const int qt = blockIdx.y;
  if (site >= L) return;
  uint32_t w[QT / 4];
#pragma unroll
  for (int i = 0; i < QT / 4; ++i) {
    uint32_t v = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int slot = qt * QT + i * 4 + j;
      const uint32_t c = slot < Q ? codes[(int64_t)order[slot] * Lp + site] : 0xFFu;
      v |= c << (8 * j);
    }
    w[i] = v;
  }
  uint32_t *dst = reinterpret_cast<uint32_t *>(qc + ((int64_t)qt * L + site) * QT);
#pragma unroll
  for (int i = 0; i < QT / 4; ++i) dst[i] = w[i];
}

// ---- shared pieces ------------------------------------------------------------------------------
template <int QT>
struct CodeWords {
  uint32_t w[QT / 4];
  __device__ __forceinline__ bool all_uninformative() const {
    uint32_t a = 0x80808080u;
#pragma unroll
    for (int i = 0; i < QT / 4; ++i) a &= w[i];
    return a == 0x80808080u;
  }
  __device__ __forceinline__ unsigned code(int b) const { return (w[b >> 2] >> (8 * (b & 3))) & 0xffu; }
};

template <int QT>
__device__ __forceinline__ CodeWords<QT> load_codes(const uint8_t *p) {
  CodeWords<QT> c;
  if constexpr (QT == 4) {
    c.w[0] = __ldg(reinterpret_cast<const uint32_t *>(p));
  } else if constexpr (QT == 8) {
    const uint2 v = __ldg(reinterpret_cast<const uint2 *>(p));
    c.w[0] = v.x; c.w[1] = v.y;
  } else {
#pragma unroll
    for (int i = 0; i < QT / 16; ++i) {
      const uint4 v = __ldg(reinterpret_cast<const uint4 *>(p) + i);
      c.w[4 * i] = v.x; c.w[4 * i + 1] = v.y; c.w[4 * i + 2] = v.z; c.w[4 * i + 3] = v.w;
    }
  }
  return c;
}

// Running products with DEFERRED renormalisation.  Per-site factors are accepted only inside
// [2^-120, 2^120): then a mantissa in [1,2) times up to kRenormEvery such factors cannot leave the
// normal range, and the exponent is folded into an integer every kRenormEvery sites (3 integer
// instructions per accumulator).  A factor outside the range (zero, negative, NaN, inf, denormal,
// absurdly small) only raises `bad`; the affected (query, edge) scores are then recomputed by
// exact_partial_score() with one log() per site exactly as linear_c.c:532 does, so -inf / NaN
// propagate as in the reference.
constexpr int kRenormEvery = 8;
constexpr unsigned kSafeLoHi = (1023u - 120u) << 20;  // high word of 2^-120
constexpr unsigned kSafeSpan = 240u << 20;
__device__ __forceinline__ unsigned range_key(double u) { return (unsigned)__double2hiint(u) - kSafeLoHi; }
__device__ __forceinline__ void renorm(double &m, int &ex) {
  const int hi = __double2hiint(m);
  ex += hi >> 20;  // biased: the caller subtracts 1023 per renormalisation at the end
  m = __hiloint2double((hi & 0x000fffff) | 0x3ff00000, __double2loint(m));
}

// Exact (slow) partial score of one (query, edge) over sites [s0, s1): sum of log() per site.
template <int S, bool GCAT>
__device__ __noinline__ double exact_partial_score(const PreArgs &A, int qt, int QT, int b, int e, int s0, int s1) {
  const int K = A.sn.K;
  double acc = 0.0;
  long long es = 0;
  for (int site = s0; site < s1; ++site) {
    const unsigned code = A.qc[((int64_t)qt * A.sn.L + site) * QT + b];
    if (code & 0x80u) continue;
    const double *tv = GCAT ? A.Tq + ((int64_t)A.site_cat[site] * A.nc + code) * S : A.Tq + (int64_t)code * K;
    const double *col = A.sn.a + ((int64_t)site * K) * A.sn.Np + e;
    double u = 0.0;
    for (int k = 0; k < K; ++k) u = fma(tv[k], col[(int64_t)k * A.sn.Np], u);
    acc += log(u);
    es += A.sn.e[(int64_t)site * A.sn.Np + e];
  }
  return acc + kLn2 * (double)es;
}

// ---- the direct (FMA) form ----------------------------------------------------------------------
// grid = (ceil(n_qtiles / W), Np / (32*TE), zs); block = W warps.
//   warp  -> query tile (QT queries), lane -> TE consecutive edges, blockIdx.z -> site range.
// TQSRC: where the warp-uniform evolved-query column of a (query, site) comes from
//   TQ_CONST  __constant__ c_tq (Gamma mixtures: n_codes x K doubles, indexed loads from the constant
//             cache - no shared-memory / L1 bandwidth is spent on the query side at all)
//   TQ_SMEM   shared-memory copy of the same table (kept for comparison)
//   TQ_GLOBAL global memory through L1 (Gcat: the table is R times larger, 77 KB for 20 categories
//             of amino acids, and only one category row is live per site)
// NRC > 0: K = NRC*S is known at compile time and one register chunk holds all K planes of an edge
// (DNA Gamma-4: NRC 4, K 16; any Gcat: NRC 1).  NRC == 0: any number of rates, one rate (S planes) at
// a time with the per-query sums carried across the rate loop (amino-acid Gamma mixtures).
enum : int { TQ_CONST = 0, TQ_SMEM = 1, TQ_GLOBAL = 2 };
constexpr int kConstTqDoubles = 4096;
__constant__ double c_tq[kConstTqDoubles];

template <int S, bool GCAT, int NRC, int TE, int QT, int W, int TQSRC>
__global__ void __launch_bounds__(W * 32)
prefilter_direct_kernel(const PreArgs A) {
  extern __shared__ double tq_s[];
  constexpr bool SINGLE = NRC > 0;
  constexpr int KC = SINGLE ? NRC * S : S;
  const int K = SINGLE ? KC : A.sn.K;
  const int nch = SINGLE ? 1 : K / S;
  static_assert(!GCAT || NRC == 1, "a Gcat GLV has one plane set per site");
  static_assert(GCAT == (TQSRC == TQ_GLOBAL), "Gcat tables are read from global memory, Gmix tables are not");
  if constexpr (TQSRC == TQ_SMEM) {
    const int tq_len = A.nc * K;
    for (int i = threadIdx.x; i < tq_len; i += W * 32) tq_s[i] = A.Tq[i];
    __syncthreads();
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int qt = blockIdx.x * W + warp;
  if (qt >= A.nqt) return;
  const int L = A.sn.L;
  const int64_t Np = A.sn.Np;
  const int e0 = (blockIdx.y * 32 + lane) * TE;
  const int s0 = max(blockIdx.z * A.sites_per_z, __ldg(A.tile_lo + qt)), s1 = min(min(L, (blockIdx.z + 1) * A.sites_per_z), __ldg(A.tile_hi + qt));

  double mm[QT][TE];
  int ex[QT][TE];
#pragma unroll
  for (int b = 0; b < QT; ++b)
#pragma unroll
    for (int a = 0; a < TE; ++a) { mm[b][a] = 1.0; ex[b][a] = 0; }
  unsigned bad = 0;
  int nren = 0, cnt = 0;
  const uint8_t *qcp = A.qc + ((int64_t)qt * L) * QT;

  for (int site = s0; site < s1; ++site) {
    const CodeWords<QT> cw = load_codes<QT>(qcp + (int64_t)site * QT);
    if (cw.all_uninformative()) continue;
    const double *row = A.sn.a + ((int64_t)site * K) * Np + e0;
    int es[TE];
    if constexpr (TE == 2) {
      const int2 v = __ldg(reinterpret_cast<const int2 *>(A.sn.e + (int64_t)site * Np + e0));
      es[0] = v.x; es[1] = v.y;
    } else {
#pragma unroll
      for (int a = 0; a < TE; ++a) es[a] = __ldg(A.sn.e + (int64_t)site * Np + e0 + a);
    }
    const int cat = GCAT ? A.site_cat[site] : 0;
    double u[SINGLE ? 1 : QT][TE];
    for (int ch = 0; ch < nch; ++ch) {
      double sn[KC][TE];
#pragma unroll
      for (int k = 0; k < KC; ++k) {
        if constexpr (TE == 2) {
          const double2 v = __ldg(reinterpret_cast<const double2 *>(row + (int64_t)(ch * KC + k) * Np));
          sn[k][0] = v.x; sn[k][1] = v.y;
        } else {
#pragma unroll
          for (int a = 0; a < TE; ++a) sn[k][a] = __ldg(row + (int64_t)(ch * KC + k) * Np + a);
        }
      }
#pragma unroll
      for (int b = 0; b < QT; ++b) {
        const unsigned code = cw.code(b);
        if (code & 0x80u) continue;
        double *ub = u[SINGLE ? 0 : b];
        // two half-length chains per edge keep the FP64 pipe busy at low occupancy
        double lo[TE], hi[TE];
        auto dot = [&](auto tv) {
#pragma unroll
          for (int a = 0; a < TE; ++a) { lo[a] = tv(0) * sn[0][a]; hi[a] = tv(KC / 2) * sn[KC / 2][a]; }
#pragma unroll
          for (int k = 1; k < KC / 2; ++k)
#pragma unroll
            for (int a = 0; a < TE; ++a) { lo[a] = fma(tv(k), sn[k][a], lo[a]); hi[a] = fma(tv(KC / 2 + k), sn[KC / 2 + k][a], hi[a]); }
        };
        if constexpr (TQSRC == TQ_GLOBAL) {
          const double *tv = A.Tq + ((int64_t)cat * A.nc + code) * S;
          dot([&](int k) { return __ldg(tv + k); });
        } else if constexpr (TQSRC == TQ_SMEM) {
          const double *tv = tq_s + (int64_t)code * K + ch * KC;
          dot([&](int k) { return tv[k]; });
        } else {
          const int base = code * K + ch * KC;
          dot([&](int k) { return c_tq[base + k]; });
        }
#pragma unroll
        for (int a = 0; a < TE; ++a) {
          const double v = lo[a] + hi[a];
          if (SINGLE || ch == 0) ub[a] = v; else ub[a] += v;
        }
        if (SINGLE) {
#pragma unroll
          for (int a = 0; a < TE; ++a) {
            bad = max(bad, range_key(ub[a]));
            mm[b][a] *= ub[a];
            ex[b][a] += es[a];
          }
        }
      }
    }
    if (!SINGLE) {
#pragma unroll
      for (int b = 0; b < QT; ++b) {
        if (cw.code(b) & 0x80u) continue;
#pragma unroll
        for (int a = 0; a < TE; ++a) {
          bad = max(bad, range_key(u[SINGLE ? 0 : b][a]));
          mm[b][a] *= u[SINGLE ? 0 : b][a];
          ex[b][a] += es[a];
        }
      }
    }
    if ((++cnt & (kRenormEvery - 1)) == 0) {
      ++nren;
#pragma unroll
      for (int b = 0; b < QT; ++b)
#pragma unroll
        for (int a = 0; a < TE; ++a) renorm(mm[b][a], ex[b][a]);
    }
  }
  ++nren;
  const bool is_bad = bad >= kSafeSpan;
  double *outz = A.out + (int64_t)blockIdx.z * A.Q * A.sn.n_nodes;
#pragma unroll
  for (int b = 0; b < QT; ++b) {
    if (qt * QT + b >= A.Q) continue;
    const int q = __ldg(A.order + qt * QT + b);
    const double rate_term = (!GCAT && blockIdx.z == 0) ? (double)((float)A.nq[q]) * log((double)((float)A.R)) : 0.0;  // linear_c.c:536
#pragma unroll
    for (int a = 0; a < TE; ++a) {
      const int e = e0 + a;
      if (e >= A.sn.n_nodes) continue;
      renorm(mm[b][a], ex[b][a]);
      double v = log(mm[b][a]) + kLn2 * (double)(ex[b][a] - 1023 * nren);
      if (is_bad) v = exact_partial_score<S, GCAT>(A, qt, QT, b, e, s0, s1);
      v -= rate_term;
      if (A.zs == 1 && e == A.root_id) v = __longlong_as_double(0x7ff8000000000000LL);
      outz[(int64_t)q * A.sn.n_nodes + e] = v;
    }
  }
}

// sum the per-site-range partial scores in a fixed order (deterministic), NaN at the root
__global__ void combine_partials_kernel(const double *__restrict__ partial, double *__restrict__ scores, int zs, int n_nodes,
                                        int root_id, int64_t total) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    double v = 0.0;
    for (int z = 0; z < zs; ++z) v += partial[(int64_t)z * total + i];
    scores[i] = (int)(i % n_nodes) == root_id ? __longlong_as_double(0x7ff8000000000000LL) : v;
  }
}

// ---- the table form (DNA Gamma mixture) ----------------------------------------------------------
// The evolved query takes one of four values per (rate, state) at an unambiguous site, so
//   M[site][j][e] = sum_k Tq[1<<j][k] * snT[site][k][e]      (j = A, C, G, T)
// is computed ONCE per (edge, site) (4K FMAs, one streaming pass over the supernodes) and a
// (query, edge, site) unit becomes: pick M[.][code] (sum the columns of an ambiguity code), one DMUL
// into the running product, one integer add for the exponent.  Also records edges whose table
// leaves the safe range of the deferred renormalisation.
template <int K>
__global__ void __launch_bounds__(128)
mtable_kernel(const SnT sn, const TqPrimary tp, double *__restrict__ MT, int32_t *__restrict__ MTe, int site0) {
  const int64_t Np = sn.Np;
  const int e = blockIdx.x * 128 + threadIdx.x;
  const int site = site0 + blockIdx.y;
  if (e >= Np) return;
  const double *col = sn.a + ((int64_t)site * K) * Np + e;
  double v[K];
#pragma unroll
  for (int k = 0; k < K; ++k) v[k] = __ldg(col + (int64_t)k * Np);
  double M[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    double lo = tp.v[j][0] * v[0], hi = tp.v[j][K / 2] * v[K / 2];
#pragma unroll
    for (int k = 1; k < K / 2; ++k) { lo = fma(tp.v[j][k], v[k], lo); hi = fma(tp.v[j][K / 2 + k], v[K / 2 + k], hi); }
    M[j] = lo + hi;
  }
  // Per-(site, edge) power-of-two rescaling: the largest of the four entries goes to [1,2) and its
  // exponent joins the supernode's own (exact, so the scores keep their bits).  Long trees leave site
  // maxima hundreds of binary orders below 1 (the reference stops rescaling a site once a rate category
  // has underflowed, gmix_model.ml:172-176); without this every such edge would fall out of the fast
  // range.  Bit 0 of the stored word flags a (site, edge) whose entries still leave [2^-120, 2^120).
  const double mx = fmax(fmax(M[0], M[1]), fmax(M[2], M[3]));
  const int hi = __double2hiint(mx);
  const int ef = (hi >> 20) & 0x7ff;
  int shift = 0;
  unsigned flag = 0;
  if (hi >= 0 && ef != 0 && ef != 0x7ff) {
    shift = max(-1000, min(1000, ef - 1023));
    const double sc = __hiloint2double((1023 - shift) << 20, 0);  // 2^-shift, exact
#pragma unroll
    for (int j = 0; j < 4; ++j) M[j] *= sc;
  } else
    flag = 1;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    if (range_key(M[j]) >= kSafeSpan) flag = 1;
    MT[((int64_t)site * 4 + j) * Np + e] = M[j];
  }
  MTe[(int64_t)site * Np + e] = ((__ldg(sn.e + (int64_t)site * Np + e) + shift) << 1) | (int)flag;
}

// grid / block exactly as prefilter_direct_kernel; MT[L][4][Np].
template <int TE, int QT, int W, int MINB = 1>
__global__ void __launch_bounds__(W * 32, MINB)
prefilter_table_kernel(const PreArgs A, const double *__restrict__ MT, const int32_t *__restrict__ MTe) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int qt = blockIdx.x * W + warp;
  if (qt >= A.nqt) return;
  const int L = A.sn.L;
  const int64_t Np = A.sn.Np;
  const int e0 = (blockIdx.y * 32 + lane) * TE;
  const int s0 = max(blockIdx.z * A.sites_per_z, __ldg(A.tile_lo + qt)), s1 = min(min(L, (blockIdx.z + 1) * A.sites_per_z), __ldg(A.tile_hi + qt));
  double mm[QT][TE];
  int ex[QT][TE];
#pragma unroll
  for (int b = 0; b < QT; ++b)
#pragma unroll
    for (int a = 0; a < TE; ++a) { mm[b][a] = 1.0; ex[b][a] = 0; }
  int nren = 0, cnt = 0;
  unsigned dead = 0;       // bit b*TE+a: (query b, edge a) met a negative / NaN / inf / subnormal factor -> exact path
  unsigned dead_zero = 0;  // ... met a factor that is exactly 0 -> log-likelihood -inf (linear_c.c:532: log 0), unless `dead` too
  static_assert(QT * TE <= 32, "one dead bit per (query, edge) of the thread");
  const uint8_t *qcp = A.qc + ((int64_t)qt * L) * QT;
  for (int site = s0; site < s1; ++site) {
    const CodeWords<QT> cw = load_codes<QT>(qcp + (int64_t)site * QT);
    if (cw.all_uninformative()) continue;
    const double *row = MT + ((int64_t)site * 4) * Np + e0;
    double M[4][TE];
    int es[TE];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if constexpr (TE == 2) {
        const double2 v = __ldg(reinterpret_cast<const double2 *>(row + (int64_t)j * Np));
        M[j][0] = v.x; M[j][1] = v.y;
      } else {
#pragma unroll
        for (int a = 0; a < TE; ++a) M[j][a] = __ldg(row + (int64_t)j * Np + a);
      }
    }
    int flagged = 0;  // bit a: this (site, edge) has entries outside the deferred-renormalisation range
    if constexpr (TE == 2) {
      const int2 v = __ldg(reinterpret_cast<const int2 *>(MTe + (int64_t)site * Np + e0));
      es[0] = v.x >> 1; es[1] = v.y >> 1;
      flagged = (v.x & 1) | ((v.y & 1) << 1);
    } else {
#pragma unroll
      for (int a = 0; a < TE; ++a) {
        const int v = __ldg(MTe + (int64_t)site * Np + e0 + a);
        es[a] = v >> 1;
        flagged |= (v & 1) << a;
      }
    }
    if (!flagged) {  // the common case: exactly the unguarded loop
#pragma unroll
      for (int b = 0; b < QT; ++b) {
        const unsigned code = cw.code(b);
        if (code & 0x80u) continue;
        double u[TE];
        switch (code) {  // the DNA code IS the state mask (A=1, C=2, G=4, T=8)
          case 1:
#pragma unroll
            for (int a = 0; a < TE; ++a) u[a] = M[0][a];
            break;
          case 2:
#pragma unroll
            for (int a = 0; a < TE; ++a) u[a] = M[1][a];
            break;
          case 4:
#pragma unroll
            for (int a = 0; a < TE; ++a) u[a] = M[2][a];
            break;
          case 8:
#pragma unroll
            for (int a = 0; a < TE; ++a) u[a] = M[3][a];
            break;
          default:  // ambiguity code: sum of the compatible columns, state order (linear_c.c:29-90 on a 0/1 vector)
#pragma unroll
            for (int a = 0; a < TE; ++a) {
              double s = 0.0;
#pragma unroll
              for (int j = 0; j < 4; ++j)
                if (code >> j & 1u) s += M[j][a];
              u[a] = s;
            }
            break;
        }
#pragma unroll
        for (int a = 0; a < TE; ++a) { mm[b][a] *= u[a]; ex[b][a] += es[a]; }
      }
    } else {
#pragma unroll
      for (int b = 0; b < QT; ++b) {
        const unsigned code = cw.code(b);
        if (code & 0x80u) continue;
        double u[TE];
        switch (code) {  // the DNA code IS the state mask (A=1, C=2, G=4, T=8)
          case 1:
#pragma unroll
            for (int a = 0; a < TE; ++a) u[a] = M[0][a];
            break;
          case 2:
#pragma unroll
            for (int a = 0; a < TE; ++a) u[a] = M[1][a];
            break;
          case 4:
#pragma unroll
            for (int a = 0; a < TE; ++a) u[a] = M[2][a];
            break;
          case 8:
#pragma unroll
            for (int a = 0; a < TE; ++a) u[a] = M[3][a];
            break;
          default:  // ambiguity code: sum of the compatible columns, state order (linear_c.c:29-90 on a 0/1 vector)
#pragma unroll
            for (int a = 0; a < TE; ++a) {
              double s = 0.0;
#pragma unroll
              for (int j = 0; j < 4; ++j)
                if (code >> j & 1u) s += M[j][a];
              u[a] = s;
            }
            break;
        }
#pragma unroll
        for (int a = 0; a < TE; ++a)
          if ((flagged >> a) & 1) {  // pull the factor's own exponent out at once; non-normal factors kill the pair
            const int hi = __double2hiint(u[a]);
            const int ef = (hi >> 20) & 0x7ff;
            if (u[a] == 0.0) { dead_zero |= 1u << (b * TE + a); u[a] = 1.0; }
            else if (hi < 0 || ef == 0 || ef == 0x7ff) { dead |= 1u << (b * TE + a); u[a] = 1.0; }
            else { ex[b][a] += ef - 1023; u[a] = __hiloint2double((hi & 0x000fffff) | 0x3ff00000, __double2loint(u[a])); }
          }
#pragma unroll
        for (int a = 0; a < TE; ++a) { mm[b][a] *= u[a]; ex[b][a] += es[a]; }
      }
    }
    if ((++cnt & (kRenormEvery - 1)) == 0) {
      ++nren;
#pragma unroll
      for (int b = 0; b < QT; ++b)
#pragma unroll
        for (int a = 0; a < TE; ++a) renorm(mm[b][a], ex[b][a]);
    }
  }
  ++nren;
  double *outz = A.out + (int64_t)blockIdx.z * A.Q * A.sn.n_nodes;
#pragma unroll
  for (int b = 0; b < QT; ++b) {
    if (qt * QT + b >= A.Q) continue;
    const int q = __ldg(A.order + qt * QT + b);
    const double rate_term = (blockIdx.z == 0) ? (double)((float)A.nq[q]) * log((double)((float)A.R)) : 0.0;  // linear_c.c:536
#pragma unroll
    for (int a = 0; a < TE; ++a) {
      const int e = e0 + a;
      if (e >= A.sn.n_nodes) continue;
      renorm(mm[b][a], ex[b][a]);
      double v = log(mm[b][a]) + kLn2 * (double)(ex[b][a] - 1023 * nren);
      if ((dead >> (b * TE + a)) & 1u) v = exact_partial_score<4, false>(A, qt, QT, b, e, s0, s1);  // -inf / NaN as linear_c.c:532
      else if ((dead_zero >> (b * TE + a)) & 1u) v = -__longlong_as_double(0x7ff0000000000000LL);
      v -= rate_term;
      if (A.zs == 1 && e == A.root_id) v = __longlong_as_double(0x7ff8000000000000LL);
      outz[(int64_t)q * A.sn.n_nodes + e] = v;
    }
  }
}

// ---- the log-table form (DNA, Gamma mixture, K = 16): the default ----------------------------------
// The reference sums log(util[site]) over the query's sites (linear_c.c:532).  util depends on the query only through
// its character at the site, so
//     LT[site][c][e] = log( sum_k Tq[c][k] * snT[site][k][e] ) + ln2 * e_snode[site][e],   c = A, C, G, T, N
// is computed once per (reference, pendant length) and a (query, edge, site) unit is ONE load and ONE add - in the
// reference's own order of summation.  -inf (a zero site likelihood) and NaN propagate through the additions exactly
// as they do through the reference's sum of logs: no range guards, no renormalisation, no fallback path.  The ten
// two- and three-base ambiguity codes (rare) take the slow term: the K-term dot product and a log on the spot.
struct TqFive {
  double v[5][16];  // evolved-query columns of A, C, G, T and N (code 15)
};
constexpr unsigned long long kLogColOfCode = 0x4ffffff3fff2f10full;  // nibble[code] = table column, 15 = slow term

template <int K>
__global__ void __launch_bounds__(128)
logtable_kernel(const SnT sn, const TqFive tp, double *__restrict__ LT, int site0) {
  const int64_t Np = sn.Np;
  const int e = blockIdx.x * 128 + threadIdx.x;
  const int site = site0 + blockIdx.y;
  if (e >= Np) return;
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
    LT[((int64_t)site * 5 + j) * Np + e] = log(lo + hi) + ex;
  }
}

__device__ __noinline__ double slow_log_term(const PreArgs &A, int site, unsigned code, int e) {
  const int K = A.sn.K;
  const double *tv = A.Tq + (int64_t)code * K;
  const double *col = A.sn.a + ((int64_t)site * K) * A.sn.Np + e;
  double lo = tv[0] * col[0], hi = tv[K / 2] * col[(int64_t)(K / 2) * A.sn.Np];
  for (int k = 1; k < K / 2; ++k) { lo = fma(tv[k], col[(int64_t)k * A.sn.Np], lo); hi = fma(tv[K / 2 + k], col[(int64_t)(K / 2 + k) * A.sn.Np], hi); }
  return log(lo + hi) + kLn2 * (double)A.sn.e[(int64_t)site * A.sn.Np + e];
}

// grid / block as prefilter_direct_kernel (TE = 2 edges per lane); LT[L][5][Np].
template <int QT, int W, int MINB>
__global__ void __launch_bounds__(W * 32, MINB)
prefilter_logtable_kernel(const PreArgs A, const double *__restrict__ LT) {
  constexpr int TE = 2;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int qt = blockIdx.x * W + warp;
  if (qt >= A.nqt) return;
  const int L = A.sn.L;
  const int64_t Np = A.sn.Np;
  const int e0 = (blockIdx.y * 32 + lane) * TE;
  const int s0 = max(blockIdx.z * A.sites_per_z, __ldg(A.tile_lo + qt)), s1 = min(min(L, (blockIdx.z + 1) * A.sites_per_z), __ldg(A.tile_hi + qt));
  double acc[QT][TE];
#pragma unroll
  for (int b = 0; b < QT; ++b)
#pragma unroll
    for (int a = 0; a < TE; ++a) acc[b][a] = 0.0;
  const uint8_t *qcp = A.qc + ((int64_t)qt * L) * QT;
  const double *row = LT + ((int64_t)s0 * 5) * Np + e0;
  for (int site = s0; site < s1; ++site, row += 5 * Np) {
    const CodeWords<QT> cw = load_codes<QT>(qcp + (int64_t)site * QT);
    if (cw.all_uninformative()) continue;
#pragma unroll
    for (int b = 0; b < QT; ++b) {
      const unsigned code = cw.code(b);
      if (code & 0x80u) continue;  // warp-uniform
      const unsigned colc = (unsigned)(kLogColOfCode >> (4 * code)) & 15u;
      if (colc != 15u) {
        const double2 v = __ldg(reinterpret_cast<const double2 *>(row + (int64_t)colc * Np));
        acc[b][0] += v.x; acc[b][1] += v.y;
      } else {
#pragma unroll
        for (int a = 0; a < TE; ++a) acc[b][a] += slow_log_term(A, site, code, e0 + a);
      }
    }
  }
  double *outz = A.out + (int64_t)blockIdx.z * A.Q * A.sn.n_nodes;
#pragma unroll
  for (int b = 0; b < QT; ++b) {
    if (qt * QT + b >= A.Q) continue;
    const int q = __ldg(A.order + qt * QT + b);
    const double rate_term = (blockIdx.z == 0) ? (double)((float)A.nq[q]) * log((double)((float)A.R)) : 0.0;  // linear_c.c:536
#pragma unroll
    for (int a = 0; a < TE; ++a) {
      const int e = e0 + a;
      if (e >= A.sn.n_nodes) continue;
      double v = acc[b][a] - rate_term;
      if (A.zs == 1 && e == A.root_id) v = __longlong_as_double(0x7ff8000000000000LL);
      outz[(int64_t)q * A.sn.n_nodes + e] = v;
    }
  }
}

}  // namespace ppb
