// Fast 'spline' Limber kernels (sm_100a): ln k node grids shared across tracer pairs, 64-node chunks.
//
// Reference path: integ_cls_limber_spline (src/ccl_cls.c:189-225) -> cl_integrand (:166-187)
// -> ccl_integ_spline (src/ccl_utils.c:274-345).  The reference evaluates the integrand
// k P(k,a) D1 D2 independently for every (pair, ell).  Its node grid depends only on the pair's
// (chi_min, chi_max) (get_k_interval, :64-100), so pairs with equal ranges share every node:
// a(chi), k P(k, a) and each collection's D(node) are evaluated ONCE per grid and the per-pair
// remainder is one product plus the Akima integral.  Node values and the
// integrand's arithmetic are exactly those of the general kernel (kernels.cu, spline_task).
//
// Work decomposition: one CTA per (cosmology, ell); its warps pull that cosmology's grid tasks from
// a shared-memory counter.  A warp streams a grid in chunks of CH = 32 * NPL nodes:
//   phase 1 (lane = node, NPL sub-chunks of 32): a(chi), k P(k,a) and every collection's D into
//     shared memory.  The task's sub-tracers arrive ordered by spline grid (limber_group_kernel), so
//     splines on identical grids -- the lensing kernels of all source bins, the bias / alignment
//     transfers -- share ONE interval search per node and then read only their 32-byte coefficient
//     records;
//   phase 2 (lane = NPL consecutive nodes): each pair's integrand and its Akima integral in streaming
//     form; the per-pair cost (two vector loads, two shuffles, NPL ballots) is paid once per CH nodes.
// Anything the fast path does not cover (a node outside the a(chi) or P(k,a) tables, ...) is pushed
// to a redo list for the general kernel.
#include <cstddef>
#include <cstdlib>

#include "fast_common.cuh"

namespace cclb {

namespace {

#ifndef CCLB_FW
#define CCLB_FW 5
#endif
#ifndef CCLB_NPL
#define CCLB_NPL 2
#endif
#ifndef CCLB_MINB
#define CCLB_MINB 4
#endif
#ifndef CCLB_TR_UNROLL
#define CCLB_TR_UNROLL 1
#endif
constexpr int FW = CCLB_FW;     // warps per CTA
constexpr int NPL = CCLB_NPL;   // nodes per lane in phase 2
constexpr int CH = 32 * NPL;    // nodes per chunk
constexpr int NCG = CCLB_NCG;   // collections per grid task
constexpr int NPG = CCLB_NPG;   // pairs per grid task
static_assert(NPL == 2 || NPL == 4, "phase 2 is written for 2 or 4 nodes per lane");
static_assert(CCLB_NPG <= 32 && CCLB_NCG <= 16, "carry bits are one per pair; collection slots are 4 bits");
constexpr int NSG = 2 * CCLB_NCG;  // sub-tracers per grid task
constexpr int TRU = CCLB_TR_UNROLL;  // unroll factor of the sub-tracer loop
constexpr int GB = 56;          // (ell, grid) tasks whose limits are precomputed per CTA pass

// ln k limits and node spacing of one (ell, grid) task (get_k_interval + ccl_linear_spacing)
struct __align__(16) GridPar {
  double lkmin, lkmax, dx, inv_dx, ratio, iratio;
  int nk, grid, ls, pad;
};
// the last interval ends on lkmax itself (ccl_linear_spacing overwrites the end node)
__device__ __forceinline__ double last_interval(const GridPar& g) { return g.lkmax - (g.lkmin + g.dx * (g.nk - 2)); }

// Per-CTA shared state; followed in shared memory by FastTr tr[ntr_max] and the per-warp blocks.
struct FastCta {
  Spline1D achi;
  F2D psp;
  GridPar gp[GB];
  int next_task, pad;
};


struct __align__(16) FastWarp {
  FastEnt ents[NSG];       // the task's sub-tracers, ordered by spline grid (first: FastEnt offsets are positive)
  double dflat[NCG * CH];  // D of the grid's collection slot c at chunk node j: [c * CH + j]
  double kp[CH];           // k P(k, a) at chunk node j
  double acc[NPG][16];     // partial integrals: lanes l and l + 16 share an entry
  double fc[NPG][4];       // integrand at the last 4 nodes of the previous chunk, per pair
  double fe[NPG][8];       // f_0..f_2 and f_{nk-5}..f_{nk-1} per pair (end-node terms)
  int2 po[NPG];            // offsets of the pair's two collections in dflat
};

// Layout of the CTA's dynamic shared memory; tr[] really holds FastPlan::ntr_max entries.
struct __align__(16) FastSmem {
  FastCta cta;
  FastWarp warps[FW];
  FastTr tr[1];
};

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace

// ------------------------------------------------------------------------------------------
// Grouping: one CTA per cosmology builds its grid tasks: pairs with equal (chi_min, chi_max), their
// collections, and the flattened sub-tracer list ordered by spline grid.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) limber_group_kernel(const LimberProblem P, const FastPlan G) {
  extern __shared__ double gsm[];
  __shared__ int s_kgrp[FAST_MAX_TRACERS], s_agrp[FAST_MAX_TRACERS];
  const int ic = blockIdx.x;
  const int ncoll = P.ncoll, npair = P.npair;
  double* cmin = gsm;
  double* cmax = cmin + ncoll;
  double* pmin = cmax + ncoll;
  double* pmax = pmin + npair;
  int* leader = reinterpret_cast<int*>(pmax + npair);
  const Cosmo* cs = P.cosmos + ic;
  const int ntr = cs->ntracers;
  for (int c = threadIdx.x; c < ncoll; c += blockDim.x) {
    const int2 cr = P.colls[c];
    double lo = 1E15, hi = -1E15;  // get_k_interval (src/ccl_cls.c:70-90)
    for (int t = cr.x; t < cr.x + cr.y; ++t) {
      lo = fmin(lo, cs->tracers[t].chi_min);
      hi = fmax(hi, cs->tracers[t].chi_max);
    }
    cmin[c] = lo;
    cmax[c] = hi;
  }
  __syncthreads();
  for (int p = threadIdx.x; p < npair; p += blockDim.x) {
    const int2 pr = P.pairs[p];
    pmin[p] = fmax(cmin[pr.x], cmin[pr.y]);
    pmax[p] = fmin(cmax[pr.x], cmax[pr.y]);
  }
  __syncthreads();
  for (int p = threadIdx.x; p < npair; p += blockDim.x) {
    int ld = p;
    const double a = pmin[p], b = pmax[p];
    for (int q = 0; q < p; ++q)
      if (pmin[q] == a && pmax[q] == b) { ld = q; break; }
    leader[p] = ld;
  }
  // Spline grids: kgrp[t] / agrp[t] = first sub-tracer whose kernel / a-only transfer has exactly the
  // same nodes (every interval record {x_i, x_{i+1}} equal).  agrp = -1: no transfer lookup.
  for (int t = 0; t < ntr; ++t) {
    const Tracer& tr = cs->tracers[t];
    int kg = t, ag = -1;
    for (int u = 0; u < t; ++u) {
      if (s_kgrp[u] != u) continue;
      const Spline1D& a = tr.kernel;
      const Spline1D& b = cs->tracers[u].kernel;
      bool same = (a.n == b.n) && (a.x0 == b.x0) && (a.xn == b.xn);
      if (same)
        for (int i = threadIdx.x; i < a.n - 1; i += blockDim.x) {
          const double2 p = a.xr[i], q = b.xr[i];
          if (p.x != q.x || p.y != q.y) same = false;
        }
      if (__syncthreads_and(same)) { kg = u; break; }
    }
    const bool fold = tr.has_transfer && tr.const_transfer;
    if (tr.has_transfer && !fold) {
      ag = t;
      for (int u = 0; u < t; ++u) {
        if (s_agrp[u] != u) continue;
        const Spline1D& a = tr.transfer.fa;
        const Spline1D& b = cs->tracers[u].transfer.fa;
        bool same = (a.n == b.n) && (a.x0 == b.x0) && (a.xn == b.xn);
        if (same)
          for (int i = threadIdx.x; i < a.n - 1; i += blockDim.x) {
            const double2 p = a.xr[i], q = b.xr[i];
            if (p.x != q.x || p.y != q.y) same = false;
          }
        if (__syncthreads_and(same)) { ag = u; break; }
      }
    }
    __syncthreads();
    if (threadIdx.x == 0) { s_kgrp[t] = kg; s_agrp[t] = ag; }
    __syncthreads();
  }
  if (threadIdx.x != 0) return;
  GridDesc* grids = G.grids + (size_t)ic * G.gstride;
  GridPair* gp = G.gpairs + (size_t)ic * G.gstride;
  int* gc = G.gcolls + (size_t)ic * G.cstride;
  unsigned* gs = G.gsubs + (size_t)ic * G.sstride;
  int ng = 0, npw = 0, ncw = 0;
  auto close_task = [&](GridDesc& cur) {
    // flatten the task's collections into one sub-tracer list, ordered by kernel grid (stable)
    unsigned* out = gs + (size_t)ng * NSG;
    int key[NSG];
    int n = 0, rsd = 0;
    for (int s = 0; s < cur.ncoll; ++s) {
      const int2 cr = P.colls[gc[cur.coll_begin + s]];
      for (int j = 0; j < cr.y; ++j) {
        const int t = cr.x + j;
        if (n < NSG) {
          int at = n;
          while (at > 0 && key[at - 1] > s_kgrp[t]) { key[at] = key[at - 1]; out[at] = out[at - 1]; --at; }
          key[at] = s_kgrp[t];
          out[at] = (unsigned)t | ((unsigned)s << 8);
        }
        ++n;
        rsd |= (cs->tracers[t].der_bessel >= 1);
      }
    }
    cur.nsub = n;  // n > NSG: the fast kernel leaves the task to the general kernel
    cur.has_rsd = rsd;
    cur.sub_begin = ng * NSG;
    unsigned seen = 0u;
    int last_k = -1, last_a = -2;
    for (int e = 0; e < min(n, NSG); ++e) {
      unsigned ent = out[e];
      const int t = (int)(ent & 0xffu), s = (int)((ent >> 8) & 0xfu);
      if (!((seen >> s) & 1u)) { ent |= SUB_FIRST; seen |= 1u << s; }
      if (s_kgrp[t] != last_k) { ent |= SUB_NEWK; last_k = s_kgrp[t]; }
      if (s_agrp[t] >= 0 && s_agrp[t] != last_a) { ent |= SUB_NEWA; last_a = s_agrp[t]; }
      out[e] = ent;
    }
    grids[ng++] = cur;
    ncw += cur.ncoll;
    npw += cur.npair;
  };
  for (int p = 0; p < npair; ++p) {
    if (leader[p] != p) continue;
    GridDesc cur{pmin[p], pmax[p], ncw, 0, npw, 0, 0, 0, 0, 0};
    int cur_nsub = 0;
    for (int q = p; q < npair; ++q) {
      if (leader[q] != p) continue;
      const int2 pr = P.pairs[q];
      int sa = -1, sb = -1;
      for (int i = 0; i < cur.ncoll; ++i) {
        if (gc[cur.coll_begin + i] == pr.x) sa = i;
        if (gc[cur.coll_begin + i] == pr.y) sb = i;
      }
      const bool new_b = (sb < 0) && (pr.y != pr.x);
      const int need = (sa < 0) + new_b;
      const int need_sub = (sa < 0 ? P.colls[pr.x].y : 0) + (new_b ? P.colls[pr.y].y : 0);
      if (cur.npair > 0 && (cur.ncoll + need > NCG || cur.npair + 1 > NPG || cur_nsub + need_sub > NSG)) {
        // close this task, open another on the same grid
        close_task(cur);
        cur = GridDesc{pmin[p], pmax[p], ncw, 0, npw, 0, 0, 0, 0, 0};
        cur_nsub = 0;
        sa = sb = -1;
      }
      if (sa < 0) { sa = cur.ncoll; gc[cur.coll_begin + cur.ncoll++] = pr.x; cur_nsub += P.colls[pr.x].y; }
      if (pr.y == pr.x) sb = sa;
      else if (sb < 0) { sb = cur.ncoll; gc[cur.coll_begin + cur.ncoll++] = pr.y; cur_nsub += P.colls[pr.y].y; }
      gp[cur.pair_begin + cur.npair++] = GridPair{q, (short)sa, (short)sb};
    }
    close_task(cur);
  }
  G.ngrids[ic] = ng;
}

// ------------------------------------------------------------------------------------------
// Phase 2 of one chunk: per-pair integrand and Akima integral.
//
// With m_j the slopes and t_j the Akima node derivatives, GSL's piecewise cubic integrates to
//   sum_i h_i (f_i + f_{i+1}) / 2  +  (1/12) sum_{i: NE_i != 0} h_i^2 (t_i - tL_{i+1}),
// tL_{i+1} = t_{i+1} unless NE_{i+1} == 0 (then m_i).  The first sum is a node-weighted sum (weights
// dx inside, half intervals at the two ends, the last interval with its own width).  Per node c the
// second sum contributes h^2 ([!z_c] t_c - [!z_{c-1}] (z_c ? m_{c-1} : t_c)) with z_c = (NE_c == 0): on
// the uniform part of the grid that vanishes unless z_c != z_{c-1}.  The chunk loop therefore adds
// weighted node values and tests for tie transitions (NE_c == 0 <=> m_{c+1} == m_c and m_{c-1} ==
// m_{c-2}, i.e. E_{c+2} && E_c with E_s = (d_s == d_{s-1}), d_s = f_s - f_{s-1}); only a chunk holding
// a transition evaluates the per-node form, for the interior nodes 1..nk-3.  The three end-node
// terms (nodes 0, nk-2, nk-1) are added once per pair after the loop from the saved values fe.
//
// Lane `lane` holds the chunk nodes n = NPL*lane + r (r < NPL) and owns the correction term of node
// c = base + n - 2.  The left neighbours of a lane's first node come from two shuffles; the previous
// chunk's last four values (W.fc) supply lane 0, and three carried bits per pair (E at the last two
// nodes, z at the last owned node) connect the tie flags across chunks.
// ------------------------------------------------------------------------------------------
template <bool EDGE>
__device__ __forceinline__ void pair_phase(FastWarp& W, const GridPar& gp, int np, int base, int lane,
                                           unsigned& cE1, unsigned& cE2, unsigned& cZ) {
  const unsigned FULL = 0xffffffffu;
  const int n0 = NPL * lane;
  const int nk = gp.nk;
  const double dx = gp.dx;
  double kpv[NPL], wn[NPL];
  unsigned vmask[NPL];
  if (NPL == 2) {
    const double2 k2 = *reinterpret_cast<const double2*>(&W.kp[n0]);
    kpv[0] = k2.x; kpv[1] = k2.y;
  } else {
    const double2 k2 = *reinterpret_cast<const double2*>(&W.kp[n0]);
    const double2 k3 = *reinterpret_cast<const double2*>(&W.kp[n0 + 2]);
    kpv[0] = k2.x; kpv[1] = k2.y; kpv[NPL - 2] = k3.x; kpv[NPL - 1] = k3.y;
  }
#pragma unroll
  for (int r = 0; r < NPL; ++r) {
    if (EDGE) {
      const int s = base + n0 + r;
      const double hl = last_interval(gp);
      double w = dx;
      if (s == 0) w = 0.5 * dx;
      if (s == nk - 2) w = 0.5 * dx + 0.5 * hl;
      if (s == nk - 1) w = 0.5 * hl;
      if (s >= nk) w = 0.0;
      wn[r] = w;
      vmask[r] = __ballot_sync(FULL, s >= 3 && s <= nk - 1);  // owned correction node c = s - 2 in 1..nk-3
    } else {
      wn[r] = dx;
      vmask[r] = FULL;
    }
  }
  const bool first = EDGE && (base == 0);
#pragma unroll 1
  for (int q = 0; q < np; ++q) {
    const int2 po = W.po[q];
    // cl_integrand (src/ccl_cls.c:166-187): k P d1 d2 (nodes past the end hold zeros)
    double f[NPL], d[NPL];
    {
      const double2 a0 = *reinterpret_cast<const double2*>(&W.dflat[po.x + n0]);
      const double2 b0 = *reinterpret_cast<const double2*>(&W.dflat[po.y + n0]);
      f[0] = kpv[0] * a0.x * b0.x;
      f[1] = kpv[1] * a0.y * b0.y;
      if (NPL == 4) {
        const double2 a1 = *reinterpret_cast<const double2*>(&W.dflat[po.x + n0 + 2]);
        const double2 b1 = *reinterpret_cast<const double2*>(&W.dflat[po.y + n0 + 2]);
        f[NPL - 2] = kpv[NPL - 2] * a1.x * b1.x;
        f[NPL - 1] = kpv[NPL - 1] * a1.y * b1.y;
      }
    }
    double pf = __shfl_up_sync(FULL, f[NPL - 1], 1);
#pragma unroll
    for (int r = 1; r < NPL; ++r) d[r] = f[r] - f[r - 1];
    double pd = __shfl_up_sync(FULL, d[NPL - 1], 1);
    if (lane == 0) {  // the previous chunk's last values: f at nodes base-2, base-1
      const double2 c23 = *reinterpret_cast<const double2*>(&W.fc[q][2]);
      pf = c23.y;
      pd = c23.y - c23.x;
    }
    d[0] = f[0] - pf;
    unsigned m[NPL], zm[NPL], tr[NPL];
    m[0] = __ballot_sync(FULL, d[0] == pd);
#pragma unroll
    for (int r = 1; r < NPL; ++r) m[r] = __ballot_sync(FULL, d[r] == d[r - 1]);
    if (first) {
      // boundary slopes repeat the first difference of slopes (gsl akima_init): z_0 <=> m_1 == m_0 and
      // z_1 <=> m_2 == m_1 == m_0, i.e. E_0 := true, E_1 := E_2
      const unsigned e2 = (NPL == 2) ? ((m[0] >> 1) & 1u) : (m[NPL - 2] & 1u);
      m[0] |= 1u;
      m[1] = (m[1] & ~1u) | e2;
    }
    // carried flags of this pair sit in bit 0; the new ones enter at bit 31 and the words rotate right once per pair
    const unsigned qE1 = cE1 & 1u, qE2 = cE2 & 1u, qZ = cZ & 1u;
    if (NPL == 2) {
      zm[0] = m[0] & ((m[0] << 1) | qE2);
      zm[1] = m[1] & ((m[1] << 1) | qE1);
    } else {
      zm[NPL - 2] = m[NPL - 2] & m[0];
      zm[NPL - 1] = m[NPL - 1] & m[1];
      zm[0] = m[0] & ((m[NPL - 2] << 1) | qE2);
      zm[1] = m[1] & ((m[NPL - 1] << 1) | qE1);
    }
    tr[0] = (zm[0] ^ ((zm[NPL - 1] << 1) | qZ)) & vmask[0];
    unsigned any = tr[0];
#pragma unroll
    for (int r = 1; r < NPL; ++r) {
      tr[r] = (zm[r] ^ zm[r - 1]) & vmask[r];
      any |= tr[r];
    }
    double acc = 0.0;
#pragma unroll
    for (int r = 0; r < NPL; ++r) acc = fma(wn[r], f[r], acc);
    if (EDGE) {  // f_0..f_2 and f_{nk-5}..f_{nk-1} feed the end-node terms
#pragma unroll
      for (int r = 0; r < NPL; ++r) {
        const int s = base + n0 + r;
        if (s < 3) W.fe[q][s] = f[r];
        const int e = s - (nk - 5);
        if (e >= 0 && s < nk) W.fe[q][3 + e] = f[r];
      }
    }
    if (any != 0u) {
      // ---- a tie transition in this chunk: per-node form for the flagged nodes ----
      const double2 c01 = *reinterpret_cast<const double2*>(&W.fc[q][0]);  // f at nodes base-4, base-3
      const double2 c23 = *reinterpret_cast<const double2*>(&W.fc[q][2]);  // f at nodes base-2, base-1
      const double cd1 = c23.y - c23.x, cd2 = c23.x - c01.y, cd3 = c01.y - c01.x;  // d at nodes base-1, -2, -3
      const double inv_dx = gp.inv_dx;
      double w4[NPL + 3];  // d at chunk nodes n0-3 .. n0+NPL-1
      if (NPL == 2) {
        double p0 = __shfl_up_sync(FULL, d[0], 1), pp1 = __shfl_up_sync(FULL, d[1], 2);
        if (lane == 0) { p0 = cd2; pp1 = cd3; }
        if (lane == 1) pp1 = cd1;
        w4[0] = pp1; w4[1] = p0;
      } else {
        double p1 = __shfl_up_sync(FULL, d[1], 1), p2 = __shfl_up_sync(FULL, d[NPL - 2], 1);
        if (lane == 0) { p1 = cd3; p2 = cd2; }
        w4[0] = p1; w4[1] = p2;
      }
      w4[2] = pd;
#pragma unroll
      for (int r = 0; r < NPL; ++r) w4[3 + r] = d[r];
#pragma unroll
      for (int r = 0; r < NPL; ++r) {
        if (((tr[r] >> lane) & 1u) != 0u) {
          const int cn = base + n0 + r - 2;
          double mm2 = w4[r] * inv_dx, mm1 = w4[r + 1] * inv_dx, m0 = w4[r + 2] * inv_dx, mp1 = w4[r + 3] * inv_dx;
          if (EDGE) {
            if (cn == nk - 3) mp1 = w4[r + 3] / last_interval(gp);
            if (cn == 1) mm2 = 2.0 * mm1 - m0;
          }
          const bool z = ((zm[r] >> lane) & 1u) != 0u;
          bool zprev;
          if (r > 0) zprev = ((zm[r > 0 ? r - 1 : 0] >> lane) & 1u) != 0u;
          else zprev = lane ? (((zm[NPL - 1] >> (lane - 1)) & 1u) != 0u) : (qZ != 0u);
          const double d12 = fabs(mm1 - mm2);
          const double NE = fabs(mp1 - m0) + d12;
          double tc = 0.0;
          if (!z && NE != 0.0) {
            const double alpha = d12 / NE;
            tc = (1.0 - alpha) * mm1 + alpha * m0;
          }
          double cc = z ? 0.0 : tc;
          if (!zprev) cc -= (z ? mm1 : tc);
          acc += cc * (dx * dx * (1.0 / 12.0));
        }
      }
    }
    acc += __shfl_down_sync(FULL, acc, 16);
    if (lane < 16) W.acc[q][lane] += acc;
    __syncwarp();  // every lane has read W.fc
    if (NPL == 2) {
      if (lane >= 30) *reinterpret_cast<double2*>(&W.fc[q][2 * (lane - 30)]) = make_double2(f[0], f[1]);
    } else {
      if (lane == 31) {
        *reinterpret_cast<double2*>(&W.fc[q][0]) = make_double2(f[0], f[1]);
        *reinterpret_cast<double2*>(&W.fc[q][2]) = make_double2(f[NPL - 2], f[NPL - 1]);
      }
    }
    cE1 = (cE1 >> 1) | (m[NPL - 1] & 0x80000000u);
    cE2 = (cE2 >> 1) | (m[NPL - 2] & 0x80000000u);
    cZ = (cZ >> 1) | (zm[NPL - 1] & 0x80000000u);
  }
  // after np pairs the flags occupy bits 32-np..31: pair 0 back to bit 0 for the next chunk
  cE1 >>= (32 - np);
  cE2 >>= (32 - np);
  cZ >>= (32 - np);
}

// ------------------------------------------------------------------------------------------
// The fast integrator.
// ------------------------------------------------------------------------------------------
// RSD = true adds the der_bessel 1/2 sub-tracers (transfer_limber_single, src/ccl_cls.c:121-145): the
// second evaluation point chi' = (l+3/2)/k, a' = a(chi') and the ratio P(k,a')/P(k,a) are computed once per
// node and shared by every such sub-tracer of the grid.
template <bool RSD>
__global__ void __launch_bounds__(FW * 32, CCLB_MINB)
limber_spline_fast_kernel(const LimberProblem P, const FastPlan G) {
  // one typed view of the dynamic shared memory: every access stays a direct shared-space access
  extern __shared__ __align__(16) FastSmem fsm[];
  FastCta& S = fsm[0].cta;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  FastTr* const s_tr = fsm[0].tr;
  FastWarp& W = fsm[0].warps[warp];
  const int nlt = (P.nl + LPC - 1) / LPC;
  const int il0 = (blockIdx.x % nlt) * LPC;
  const int ic = blockIdx.x / nlt;
  const int nls = min(LPC, P.nl - il0);
  const Cosmo* cs = P.cosmos + ic;

  // ---- per-CTA staging of the descriptors this (cosmology, ell tile) needs ----
  for (int t = threadIdx.x; t < cs->ntracers; t += blockDim.x) {
    const Tracer& tr = cs->tracers[t];
    s_tr[t].kernel = pack_s1(tr.kernel);
    s_tr[t].fa = pack_s1(tr.transfer.fa);
    // a constant a-only transfer is folded into the prefactor: (w * t) * pref == w * (t * pref) up to
    // one rounding, and its spline evaluates to exactly the constant
    const bool fold = tr.has_transfer && tr.const_transfer;
    s_tr[t].has_fa = tr.has_transfer && !fold;
    s_tr[t].der_bessel = tr.der_bessel;
    for (int ls = 0; ls < nls; ++ls)
      s_tr[t].pref[ls] = tracer_prefactor(tr, P.ells[il0 + ls]) * (fold ? tr.const_transfer_value : 1.0);
  }
  if (threadIdx.x == 32) S.achi = cs->achi;
  if (threadIdx.x == 64) S.psp = *cs->psp;

  const GridDesc* grids = G.grids + (size_t)ic * G.gstride;
  const GridPair* gpairs = G.gpairs + (size_t)ic * G.gstride;
  const unsigned* gsubs = G.gsubs + (size_t)ic * G.sstride;
  const unsigned FULL = 0xffffffffu;
  const int ngrids = G.ngrids[ic];
  const int ntasks = ngrids * nls;  // task = (ell of the tile, grid)

  for (int t0 = 0; t0 < ntasks; t0 += GB) {
    __syncthreads();  // staging done / previous pass drained
    const int nb = min(GB, ntasks - t0);
    if (threadIdx.x < nb) {
      // get_k_interval (src/ccl_cls.c:91-99) and ccl_linear_spacing (src/ccl_utils.c:17-37), once
      // per (grid, ell) instead of once per lane
      const int task = t0 + threadIdx.x;
      GridPar q;
      q.ls = task / ngrids;
      q.grid = task - q.ls * ngrids;
      const GridDesc gd = grids[q.grid];
      const double lp1h = P.ells[il0 + q.ls] + 0.5;
      double chi_min = gd.chi_min;
      if (chi_min <= 0) chi_min = 0.5 * lp1h / P.par.K_MAX;
      q.lkmax = log(fmin(P.par.K_MAX, 2 * lp1h / chi_min));
      q.lkmin = log(fmax(P.par.K_MIN, lp1h / gd.chi_max));
      q.nk = (int)(fmax((q.lkmax - q.lkmin) / P.par.DLOGK_INTEGRATION + 0.5, 1.0)) + 1;
      q.dx = (q.lkmax - q.lkmin) / (q.nk - 1.);
      q.inv_dx = 1.0 / q.dx;
      q.ratio = exp(32.0 * q.dx);
      q.iratio = 1.0 / q.ratio;
      q.pad = 0;
      S.gp[threadIdx.x] = q;
    }
    if (threadIdx.x == 0) S.next_task = 0;
    __syncthreads();

    for (;;) {
      int gl = 0;
      if (lane == 0) gl = atomicAdd(&S.next_task, 1);
      gl = __shfl_sync(FULL, gl, 0);
      if (gl >= nb) break;
      const GridDesc gd = grids[S.gp[gl].grid];
      const int ls = S.gp[gl].ls;
      const int il = il0 + ls;
      const double lp1h = P.ells[il] + 0.5;
      const int ncoll = gd.ncoll, np = gd.npair, nsub = gd.nsub;
      __syncwarp();
      if (lane < np) {
        const GridPair pr = gpairs[gd.pair_begin + lane];
        W.po[lane] = make_int2(pr.a * CH, pr.b * CH);
      }
      FastEnt en;
      en.flags = 0u;
      if (lane < min(nsub, NSG)) {
        const unsigned ent = gsubs[gd.sub_begin + lane];
        const FastTr& tr = s_tr[ent & 0xffu];
        en.kcr = tr.kernel.cr;
        en.pref = tr.pref[ls];
        en.acr = tr.has_fa ? tr.fa.cr : nullptr;
        en.flags = (ent & 0xffffu) | ((RSD && tr.der_bessel >= 1) ? SUB_RSD : 0u) | (tr.has_fa ? SUB_HASA : 0u);
        en.addr = (unsigned)__cvta_generic_to_shared(&W.dflat[((ent >> 8) & 0xfu) * CH]);
      }
      {  // entries of the spline-grid group a NEWK entry opens: distance to the next NEWK entry (or the end)
        const int ne = min(nsub, NSG);
        const unsigned heads = __ballot_sync(FULL, lane < ne && (en.flags & SUB_NEWK)) | (1u << ne);
        if (lane < ne) {
          const unsigned later = heads >> (lane + 1);
          en.flags |= (unsigned)__ffs(later) << 16;
          W.ents[lane] = en;
        }
      }
      const bool task_rsd = RSD && gd.has_rsd;
      const int nk = S.gp[gl].nk;
      const double lkmin = S.gp[gl].lkmin, lkmax = S.gp[gl].lkmax;
      if (nk < 5) {  // gsl_spline_alloc(akima, n < 5) fails: NaN + CCL_ERROR_INTEG
        if (lane < np) {
          const int ip = gpairs[gd.pair_begin + lane].out;
          const size_t s = (size_t)ic * P.npair + ip;
          P.cl[s * P.nl + il] = nan("");
          P.status[s] = ST_ERR_INTEG;
          atomicCAS(P.status_detail + s, 0, (int)ST_ERR_MEMORY);
        }
        continue;
      }
      // outside the P(k,a) table in ln k, or more sub-tracers than the list holds: general kernel
      bool redo = !(lkmin >= S.psp.lkmin && lkmax <= S.psp.lkmax) || nsub > NSG;
      const double dx = S.gp[gl].dx;
      // k and chi advance by constant ratios from one 32-node sub-chunk to the next (one exp per lane per grid)
      double k = exp(lkmin + dx * lane);
      double chi = lp1h / k;
      unsigned cE1 = 0u, cE2 = 0u, cZ = 0u;  // bit q: carried tie flags of pair q (see pair_phase); np >= 1
      for (int i = lane; i < NPG * 4; i += 32) (&W.fc[0][0])[i] = 0.0;  // f at nodes -4..-1: finite placeholders
#pragma unroll 1
      for (int q = 0; q < np; ++q)
        if (lane < 16) W.acc[q][lane] = 0.0;
      __syncwarp();

#pragma unroll 1
      for (int base = 0; base < nk && !redo; base += CH) {
        // ---------------- phase 1: node quantities (lane = node base + 32 u + lane) ----------------
        bool bad = false;
#pragma unroll 1
        for (int u = 0; u < NPL; ++u) {
          const int j = 32 * u + lane;
          const int s = base + j;
          double kp = 0.0;
          if (s < nk) {
            const double lk = (s == 0) ? lkmin : ((s == nk - 1) ? S.gp[gl].lkmax : lkmin + dx * s);
            // ccl_scale_factor_of_chi (src/ccl_background.c:1251-1277); chi > 0 here
            double a = 1.0;
            if (!(chi < 1.e-8)) {
              if (chi >= S.achi.x0 && chi <= S.achi.xn) {
                const Rec1 r = find_rec(S.achi, chi);
                a = cubic(make_double4(r.y, r.b, r.c, r.d), chi - r.x0);
              } else
                bad = true;  // GSL_EDOM: the general kernel reports it
            }
            if (!(a >= S.psp.amin && a <= S.psp.amax)) bad = true;  // growth extrapolation: general kernel
            double pk = 0.0;
            if (!bad) pk = fast_exp(bicubic_eval(S.psp, lk, a, 0));
            kp = k * pk;
            double chi_lp = 0.0, a_lp = 1.0, sqell = 0.0;
            if (RSD && task_rsd) {
              const double lp3h = lp1h + 1.0;
              chi_lp = lp3h / k;
              if (!(chi_lp < 1.e-8)) {
                if (chi_lp >= S.achi.x0 && chi_lp <= S.achi.xn) {
                  const Rec1 r = find_rec(S.achi, chi_lp);
                  a_lp = cubic(make_double4(r.y, r.b, r.c, r.d), chi_lp - r.x0);
                } else
                  bad = true;
              }
              if (!(a_lp >= S.psp.amin && a_lp <= S.psp.amax)) bad = true;
              if (!bad) sqell = sqrt(lp1h * fabs(fast_exp(bicubic_eval(S.psp, lk, a_lp, 0)) / pk) / lp3h);
            }
            // transfer_limber_wrap (src/ccl_cls.c:150-164) for every collection of the grid, as one loop
            // over the task's sub-tracer list.  Entries on the same spline grid share the interval search
            // (ki, kdel for the kernels in chi; ai, adel for the a-only transfers); each then reads its own
            // 32-byte coefficient record.  Outside a kernel's support the lookup runs on the first / last
            // interval and the term is dropped (ccl_f1d_t_eval: the end nodes count as outside,
            // src/ccl_f1d.c:137-161).
            // The list is walked group by group: the entries of a spline-grid group share one interval search in chi
            // (done at the head of the group), each then reads its own 32-byte coefficient record and adds into the
            // D row of its collection (absolute shared-memory address kept in the entry).
            int ai = 0;
            double adel = 0.0;
            const FastEnt* ep = W.ents;
            const unsigned j8 = (unsigned)j * 8u;
            int e = 0;
#pragma unroll 1
            while (e < nsub) {
              const FastEnt hd = *ep;
              const FastTr& trk = s_tr[hd.flags & 0xffu];
              double x0k;
              const int ki = find_interval(trk.kernel, chi, x0k);
              const double kdel = chi - x0k;
              const bool inside = (chi > trk.kernel.x0) && (chi < trk.kernel.xn);
              const int gend = e + (int)((hd.flags >> 16) & 63u);
#pragma unroll 1
              do {
                const FastEnt en = *ep;
                const unsigned ent = en.flags;
                double wt = cubic(ldg4(en.kcr + ki), kdel);
                if (ent & SUB_HASA) {
                  if (__builtin_expect((ent & SUB_NEWA) != 0u, 0)) {
                    const FastTr& tr = s_tr[ent & 0xffu];
                    const double a_ev = fmin(fmax(a, tr.fa.x0), tr.fa.xn);
                    double x0;
                    ai = find_interval(tr.fa, a_ev, x0);
                    adel = a_ev - x0;
                  }
                  wt *= cubic(ldg4(en.acr + ai), adel);
                }
                double val = inside ? wt * en.pref : 0.0;
                if (RSD && (ent & SUB_RSD)) {
                  const FastTr& tr = s_tr[ent & 0xffu];
                  // w t at chi is `inside ? wt : 0`; the second point has its own support test
                  const double l = lp1h - 0.5, lp3h = lp1h + 1.0;
                  const double w_t = inside ? wt : 0.0;
                  const bool inside_p = (chi_lp > tr.kernel.x0) && (chi_lp < tr.kernel.xn);
                  double x0;
                  const int ip = find_interval(tr.kernel, chi_lp, x0);
                  double wt_p = cubic(ldg4(tr.kernel.cr + ip), chi_lp - x0);
                  if (tr.has_fa) {
                    const double a_ev = fmin(fmax(a_lp, tr.fa.x0), tr.fa.xn);
                    const int ia = find_interval(tr.fa, a_ev, x0);
                    wt_p *= cubic(ldg4(tr.fa.cr + ia), a_ev - x0);
                  }
                  if (!inside_p) wt_p = 0.0;
                  double dd;
                  if (tr.der_bessel == 1) dd = l * w_t / lp1h - sqell * wt_p;
                  else dd = sqell * 2 * wt_p / lp3h - (0.25 + 2 * l) * w_t / (lp1h * lp1h);
                  val = dd * en.pref;
                }
                const unsigned da = en.addr + j8;
                if (!(ent & SUB_FIRST)) val += lds_f64(da);
                sts_f64(da, val);
                ++ep;
                ++e;
              } while (e < gend);
            }
            k *= S.gp[gl].ratio;
            chi *= S.gp[gl].iratio;
          } else {
            for (int c = 0; c < ncoll; ++c) W.dflat[c * CH + j] = 0.0;  // nodes past the end: f = 0
          }
          W.kp[j] = kp;
        }
        if (__any_sync(FULL, bad)) { redo = true; break; }
        __syncwarp();
        // ---------------- phase 2: per-pair integrand and Akima integral ----------------
        if ((base == 0) || (base + CH - 1 >= nk - 5))
          pair_phase<true>(W, S.gp[gl], np, base, lane, cE1, cE2, cZ);
        else
          pair_phase<false>(W, S.gp[gl], np, base, lane, cE1, cE2, cZ);
        __syncwarp();
      }

      if (redo) {
        if (lane < np) {
          const unsigned long long task =
              ((unsigned long long)ic * P.npair + gpairs[gd.pair_begin + lane].out) * P.nl + il;
          const unsigned long long at = atomicAdd(G.redo, 1ull);
          if (at < G.redo_cap) G.redo[1 + at] = task;
        }
        continue;
      }
      // ---- end-node terms, one lane per pair (nodes 0, nk-2, nk-1; gsl akima_init boundary slopes) ----
      double endc = 0.0;
      int out = 0;
      if (lane < np) {
        out = gpairs[gd.pair_begin + lane].out;
        const double* e = W.fe[lane];
        const double hl = last_interval(S.gp[gl]), inv_hl = 1.0 / hl, inv_dx = S.gp[gl].inv_dx;
        auto blend = [](double d12, double NE, double ml, double mr) {
          const double alpha = d12 / NE;
          return (1.0 - alpha) * ml + alpha * mr;
        };
        {  // node 0: slopes m_{-2}, m_{-1}, m_0, m_1
          const double m0 = (e[1] - e[0]) * inv_dx, m1 = (e[2] - e[1]) * inv_dx;
          const double mm1 = 2.0 * m0 - m1, mm2 = 3.0 * m0 - 2.0 * m1;
          const double d12 = fabs(mm1 - mm2), NE = fabs(m1 - m0) + d12;
          if (NE != 0.0) endc += dx * dx * blend(d12, NE, mm1, m0);
        }
        const double g0 = e[3], g1 = e[4], g2 = e[5], g3 = e[6], g4 = e[7];
        const double ma = (g1 - g0) * inv_dx, mb = (g2 - g1) * inv_dx, mc = (g3 - g2) * inv_dx;
        const double md = (g4 - g3) * inv_hl;  // last interval: its end node was overwritten with lkmax
        const double me = 2.0 * md - mc, mf = 3.0 * md - 2.0 * mc;
        const bool z3 = (fabs(md - mc) + fabs(mb - ma)) == 0.0;  // node nk-3
        {  // node nk-2: h_c = hl, h_{c-1} = dx
          const double d12 = fabs(mc - mb), NE = fabs(me - md) + d12;
          const bool z2 = (NE == 0.0);
          const double t2 = z2 ? 0.0 : blend(d12, NE, mc, md);
          if (!z2) endc += hl * hl * t2;
          if (!z3) endc -= dx * dx * (z2 ? mc : t2);
          // node nk-1: h_{c-1} = hl
          const double e12 = fabs(md - mc), NE1 = fabs(mf - me) + e12;
          const bool z1 = (NE1 == 0.0);
          const double t1 = z1 ? 0.0 : blend(e12, NE1, md, me);
          if (!z2) endc -= hl * hl * (z1 ? md : t1);
        }
        endc = endc * (1.0 / 12.0);
      }
#pragma unroll 1
      for (int q = 0; q < np; ++q) {
        const double v = warp_sum(lane < 16 ? W.acc[q][lane] : 0.0) + __shfl_sync(FULL, endc, q);
        const int o = __shfl_sync(FULL, out, q);
        if (lane == 0) P.cl[((size_t)ic * P.npair + o) * P.nl + il] = v / lp1h;  // src/ccl_cls.c:340
      }
    }
  }
}

size_t limber_fast_plan_bytes(int ncosmo, int npair, unsigned long long redo_cap) {
  auto up = [](size_t b) { return (b + 255) / 256 * 256; };
  const size_t nc = (size_t)ncosmo, np = (size_t)npair;
  return up(nc * np * sizeof(GridDesc)) + up(nc * np * sizeof(GridPair)) + up(nc * 2 * np * sizeof(int)) +
         up(nc * sizeof(int)) + up(nc * np * 2 * 16 * sizeof(unsigned)) +
         up((redo_cap + 1) * sizeof(unsigned long long));
}

FastPlan limber_fast_plan_carve(void* base, int ncosmo, int npair, unsigned long long redo_cap) {
  auto up = [](size_t b) { return (b + 255) / 256 * 256; };
  const size_t nc = (size_t)ncosmo, np = (size_t)npair;
  char* p = reinterpret_cast<char*>(base);
  FastPlan G;
  G.grids = reinterpret_cast<GridDesc*>(p); p += up(nc * np * sizeof(GridDesc));
  G.gpairs = reinterpret_cast<GridPair*>(p); p += up(nc * np * sizeof(GridPair));
  G.gcolls = reinterpret_cast<int*>(p); p += up(nc * 2 * np * sizeof(int));
  G.ngrids = reinterpret_cast<int*>(p); p += up(nc * sizeof(int));
  G.gsubs = reinterpret_cast<unsigned*>(p); p += up(nc * np * 2 * 16 * sizeof(unsigned));
  G.redo = reinterpret_cast<unsigned long long*>(p);
  G.gstride = npair;
  G.cstride = 2 * npair;
  G.sstride = npair * 2 * 16;
  G.redo_cap = redo_cap;
  G.ntr_max = FAST_MAX_TRACERS;
  G.has_rsd = 1;
  G.regroup = 1;
  return G;
}

// grid tasks of every cosmology (shared by the 'spline' and 'qag_quad' grid-sharing integrators)
void launch_limber_group(const LimberProblem& prob, const FastPlan& plan, cudaStream_t stream) {
  cudaMemsetAsync(plan.redo, 0, sizeof(unsigned long long), stream);
  if (!plan.regroup) return;
  const size_t gsm = sizeof(double) * 2 * ((size_t)prob.ncoll + prob.npair) + sizeof(int) * (size_t)prob.npair;
  if (gsm > 40 * 1024)
    cudaFuncSetAttribute(limber_group_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gsm);
  limber_group_kernel<<<prob.ncosmo, 128, gsm, stream>>>(prob, plan);
}

cudaError_t launch_limber_fast(const LimberProblem& prob, const FastPlan& plan, cudaStream_t stream) {
  if (prob.ncosmo == 0 || prob.npair == 0 || prob.nl == 0) return cudaSuccess;
  if ((long long)prob.ncosmo * prob.nl > 0x7fffffffLL) return cudaErrorInvalidValue;
  launch_limber_group(prob, plan, stream);
  const size_t fsm = offsetof(FastSmem, tr) + (size_t)plan.ntr_max * sizeof(FastTr);
  static size_t attr_set[64] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  if (attr_set[dev & 63] < fsm) {
    cudaFuncSetAttribute(limber_spline_fast_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fsm);
    cudaFuncSetAttribute(limber_spline_fast_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fsm);
    attr_set[dev & 63] = fsm;
  }
  const long long nlt = (prob.nl + LPC - 1) / LPC;
  const unsigned nblk = (unsigned)((long long)prob.ncosmo * nlt);
  if (plan.has_rsd) limber_spline_fast_kernel<true><<<nblk, FW * 32, fsm, stream>>>(prob, plan);
  else limber_spline_fast_kernel<false><<<nblk, FW * 32, fsm, stream>>>(prob, plan);
  launch_limber_redo(prob, plan, stream);
  return cudaGetLastError();
}

}  // namespace cclb
