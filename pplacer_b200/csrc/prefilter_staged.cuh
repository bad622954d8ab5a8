// Prefilter, staged form (DNA, Gamma mixture with K = 4R = 16): Glv.masked_logdot of every query against every
// supernode (core.ml:187-194; ten_masked_logdot_c, linear_c.c:465-539) with the supernode tiles moved by the TMA
// engine (cp.async.bulk + mbarrier, a ring of stages in shared memory) instead of per-lane global loads, and the
// K-contraction either on the FP64 CUDA cores (MMA = false: the same FMA order as prefilter_direct_kernel) or on the
// FP64 tensor cores (MMA = true: DMMA m8n8k4 - per site an 8-query x 8-edge x 16 tile is four DMMAs).
//
// Why (ncu r4b, prefilter_direct_kernel, 36.0 ms = 36.8 % of the DFMA peak): 12 warps per SM at 157 registers, every
// site iteration begins with 17 dependent global loads (long_scoreboard 25 % of the stall samples on the first use of
// the loaded tile), each warp re-reads the tile of its block mates (31 k warps x 8 KB x 1440 sites = 360 GB from L2),
// and two thirds of the executed instructions are not DFMAs (64-bit address arithmetic, constant loads, branches).
//
// Tiling: a block of W consumer warps + 1 producer warp owns (W x 8 queries) x 64 edges.  The queries arrive sorted by
// the window of sites they cover (window_keys_kernel), so the block walks the union of its windows only.  Per stage the
// producer copies SPS sites of the edge tile: 16 rows of 64 doubles per site (row stride 68 doubles in shared memory:
// conflict-free for the DMMA fragment pattern lane -> (k = lane & 3, edge = lane >> 2) and for the 16-byte row loads
// of the FMA form) + the 64 supernode exponents.  Everything a consumer reads inside the site loop is in shared memory
// or registers; the query codes are fetched one stage ahead.
#pragma once
#include "prefilter.cuh"

namespace ppb {

namespace stg {
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {}
}
// TMA engine: contiguous bytes global -> shared, completion counted on the mbarrier (UBLKCP in SASS)
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src),
               "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
}  // namespace stg

template <int W, int SPS, int NS>
struct StagedGeom {
  static constexpr int K = 16, ET = 64, ROW = 68, QT = 8;
  static constexpr int SITE_DOUBLES = K * ROW;                        // one site of the edge tile
  static constexpr int STAGE_BYTES = SPS * (SITE_DOUBLES * 8 + ET * 4);
  static constexpr int TQ_STRIDE = 17;                                // evolved-query table rows: 16 doubles + 1 pad
  static constexpr int OFF_TQ = NS * STAGE_BYTES;                     // [n_codes][17] doubles (MMA form)
  static __host__ __device__ constexpr size_t smem_bytes(int n_codes) { return (size_t)OFF_TQ + (size_t)n_codes * TQ_STRIDE * 8 + 2 * NS * 8 + 16; }
  static constexpr uint32_t TX_BYTES = SPS * (K * ET * 8 + ET * 4);   // bytes the TMA engine delivers per full stage
};

template <int W, int SPS, int NS, bool MMA, int MINB>
__global__ void __launch_bounds__((W + 1) * 32, MINB)
prefilter_staged_kernel(const PreArgs A) {
  using G = StagedGeom<W, SPS, NS>;
  constexpr int K = G::K, ET = G::ET, ROW = G::ROW, QT = G::QT, TE = 2;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double *tq_s = reinterpret_cast<double *>(smem_raw + G::OFF_TQ);
  uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw + G::OFF_TQ + (((size_t)A.nc * G::TQ_STRIDE * 8 + 15) & ~(size_t)15));
  uint64_t *empty = full + NS;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int L = A.sn.L;
  const int64_t Np = A.sn.Np;
  const int qt0 = blockIdx.x * W;
  // the block's site range: union of its tiles' windows, cut to this z-slice
  int blo = 0x7fffffff, bhi = 0;
  for (int w = 0; w < W; ++w)
    if (qt0 + w < A.nqt) { blo = min(blo, __ldg(A.tile_lo + qt0 + w)); bhi = max(bhi, __ldg(A.tile_hi + qt0 + w)); }
  const int z0 = blockIdx.z * A.sites_per_z, z1 = min(L, z0 + A.sites_per_z);
  blo = max(blo, z0); bhi = min(bhi, z1);
  const int n_stage = bhi > blo ? (bhi - blo + SPS - 1) / SPS : 0;
  if (threadIdx.x == 0) {
    for (int s = 0; s < NS; ++s) { stg::mbar_init(full + s, 1); stg::mbar_init(empty + s, W); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = threadIdx.x; i < A.nc * K; i += (W + 1) * 32) tq_s[(i >> 4) * G::TQ_STRIDE + (i & 15)] = A.Tq[i];
  __syncthreads();
  const int e_base = blockIdx.y * ET;

  if (warp == W) {
    // ---------------- producer warp: the TMA engine fills the ring ----------------
    for (int st = 0; st < n_stage; ++st) {
      const int slot = st % NS;
      if (st >= NS) stg::mbar_wait(empty + slot, ((st / NS) - 1) & 1);
      const int sb = blo + st * SPS;
      const int ns = min(SPS, L - sb);  // sites that exist (the tail stage of the alignment may be short)
      unsigned char *dst = smem_raw + (size_t)slot * G::STAGE_BYTES;
      if (lane == 0) stg::mbar_expect_tx(full + slot, (uint32_t)ns * (K * ET * 8 + ET * 4));
      __syncwarp();
      for (int i = lane; i < ns * (K + 1); i += 32) {
        const int s = i / (K + 1), r = i - s * (K + 1);
        if (r < K)
          stg::bulk_g2s(dst + ((size_t)s * G::SITE_DOUBLES + (size_t)r * ROW) * 8, A.sn.a + ((int64_t)(sb + s) * K + r) * Np + e_base, ET * 8, full + slot);
        else
          stg::bulk_g2s(dst + (size_t)SPS * G::SITE_DOUBLES * 8 + (size_t)s * ET * 4, A.sn.e + (int64_t)(sb + s) * Np + e_base, ET * 4, full + slot);
      }
    }
    return;
  }

  // ---------------- consumer warps ----------------
  const int qt = qt0 + warp;
  const bool live = qt < A.nqt;
  const int wlo = live ? max(__ldg(A.tile_lo + qt), z0) : 0, whi = live ? min(__ldg(A.tile_hi + qt), z1) : 0;
  const uint8_t *qcp = A.qc + ((int64_t)(live ? qt : 0) * L) * QT;
  // accumulators.  FMA form: lane owns edges 2*lane, 2*lane+1 for the tile's 8 queries.
  // MMA form: lane owns query lane>>2 and edges 8t + 2*(lane&3) + {0,1}, t = 0..7.
  double mm[QT][TE];
  int ex[QT][TE];
#pragma unroll
  for (int b = 0; b < QT; ++b)
#pragma unroll
    for (int a = 0; a < TE; ++a) { mm[b][a] = 1.0; ex[b][a] = 0; }
  unsigned bad = 0;
  int nren = 0, cnt = 0;
  const int g = lane >> 2, tq = lane & 3;
  auto fetch_codes = [&](int sb) -> uint2 {  // lane i < SPS: the 8 code bytes of site sb + i (zero-extended to "all uninformative")
    uint2 v = make_uint2(0x80808080u, 0x80808080u);
    const int site = sb + lane;
    if (live && lane < SPS && site >= wlo && site < whi) v = __ldg(reinterpret_cast<const uint2 *>(qcp + (int64_t)site * QT));
    return v;
  };
  uint2 cw_next = n_stage > 0 ? fetch_codes(blo) : make_uint2(0x80808080u, 0x80808080u);
  for (int st = 0; st < n_stage; ++st) {
    const int slot = st % NS;
    const int sb = blo + st * SPS;
    const uint2 cw_cur = cw_next;
    if (st + 1 < n_stage) cw_next = fetch_codes(sb + SPS);
    stg::mbar_wait(full + slot, (st / NS) & 1);
    const double *stage = reinterpret_cast<const double *>(smem_raw + (size_t)slot * G::STAGE_BYTES);
    const int *stage_e = reinterpret_cast<const int *>(smem_raw + (size_t)slot * G::STAGE_BYTES + (size_t)SPS * G::SITE_DOUBLES * 8);
#pragma unroll 1
    for (int s = 0; s < SPS; ++s) {
      CodeWords<QT> cw;
      cw.w[0] = __shfl_sync(0xffffffffu, cw_cur.x, s);
      cw.w[1] = __shfl_sync(0xffffffffu, cw_cur.y, s);
      if (cw.all_uninformative()) continue;  // warp-uniform
      const double *sv = stage + (size_t)s * G::SITE_DOUBLES;
      if constexpr (!MMA) {
        double sn[K][TE];
#pragma unroll
        for (int k = 0; k < K; ++k) {
          const double2 v = *reinterpret_cast<const double2 *>(sv + k * ROW + 2 * lane);
          sn[k][0] = v.x; sn[k][1] = v.y;
        }
        const int2 ev = *reinterpret_cast<const int2 *>(stage_e + s * ET + 2 * lane);
        const int es[TE] = {ev.x, ev.y};
#pragma unroll
        for (int b = 0; b < QT; ++b) {
          const unsigned code = cw.code(b);
          if (code & 0x80u) continue;
          const double *tv = tq_s + code * G::TQ_STRIDE;  // warp-uniform: broadcast loads
          double lo[TE], hi[TE];
#pragma unroll
          for (int a = 0; a < TE; ++a) { lo[a] = tv[0] * sn[0][a]; hi[a] = tv[K / 2] * sn[K / 2][a]; }
#pragma unroll
          for (int k = 1; k < K / 2; ++k)
#pragma unroll
            for (int a = 0; a < TE; ++a) { lo[a] = fma(tv[k], sn[k][a], lo[a]); hi[a] = fma(tv[K / 2 + k], sn[K / 2 + k][a], hi[a]); }
#pragma unroll
          for (int a = 0; a < TE; ++a) {
            const double u = lo[a] + hi[a];
            bad = max(bad, range_key(u));
            mm[b][a] *= u;
            ex[b][a] += es[a];
          }
        }
      } else {
        // A fragments: this lane's query (row g) against k = 4j + tq; a masked query reads row 0 and its products are dropped
        const unsigned code = cw.code(g);
        const bool inf = !(code & 0x80u);
        const double *tv = tq_s + (inf ? code : 0u) * G::TQ_STRIDE + tq;
        double af[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) af[j] = tv[4 * j];
#pragma unroll
        for (int t = 0; t < 8; ++t) {
          double d0 = 0.0, d1 = 0.0;
#pragma unroll
          for (int j = 0; j < 4; ++j) stg::dmma884(d0, d1, af[j], sv[(4 * j + tq) * ROW + 8 * t + g]);
          const int2 ev = *reinterpret_cast<const int2 *>(stage_e + s * ET + 8 * t + 2 * tq);
          d0 = inf ? d0 : 1.0;
          d1 = inf ? d1 : 1.0;
          bad = max(bad, max(range_key(d0), range_key(d1)));
          mm[t][0] *= d0; mm[t][1] *= d1;
          ex[t][0] += inf ? ev.x : 0; ex[t][1] += inf ? ev.y : 0;
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
    __syncwarp();
    if (lane == 0) stg::mbar_arrive(empty + slot);
  }
  if (!live) return;
  ++nren;
  double *outz = A.out + (int64_t)blockIdx.z * A.Q * A.sn.n_nodes;
  if constexpr (!MMA) {
    const bool is_bad = bad >= kSafeSpan;
#pragma unroll
    for (int b = 0; b < QT; ++b) {
      if (qt * QT + b >= A.Q) continue;
      const int q = __ldg(A.order + qt * QT + b);
      const double rate_term = blockIdx.z == 0 ? (double)((float)A.nq[q]) * log((double)((float)A.R)) : 0.0;  // linear_c.c:536
#pragma unroll
      for (int a = 0; a < TE; ++a) {
        const int e = e_base + 2 * lane + a;
        if (e >= A.sn.n_nodes) continue;
        renorm(mm[b][a], ex[b][a]);
        double v = log(mm[b][a]) + kLn2 * (double)(ex[b][a] - 1023 * nren);
        if (is_bad) v = exact_partial_score<4, false>(A, qt, QT, b, e, max(wlo, z0), min(whi, z1));
        v -= rate_term;
        if (A.zs == 1 && e == A.root_id) v = __longlong_as_double(0x7ff8000000000000LL);
        outz[(int64_t)q * A.sn.n_nodes + e] = v;
      }
    }
  } else {
    const bool is_bad = bad >= kSafeSpan;
    if (qt * QT + g < A.Q) {
      const int q = __ldg(A.order + qt * QT + g);
      const double rate_term = blockIdx.z == 0 ? (double)((float)A.nq[q]) * log((double)((float)A.R)) : 0.0;
#pragma unroll
      for (int t = 0; t < 8; ++t)
#pragma unroll
        for (int a = 0; a < TE; ++a) {
          const int e = e_base + 8 * t + 2 * tq + a;
          if (e >= A.sn.n_nodes) continue;
          renorm(mm[t][a], ex[t][a]);
          double v = log(mm[t][a]) + kLn2 * (double)(ex[t][a] - 1023 * nren);
          if (is_bad) v = exact_partial_score<4, false>(A, qt, QT, g, e, max(wlo, z0), min(whi, z1));
          v -= rate_term;
          if (A.zs == 1 && e == A.root_id) v = __longlong_as_double(0x7ff8000000000000LL);
          outz[(int64_t)q * A.sn.n_nodes + e] = v;
        }
    }
  }
}

}  // namespace ppb
