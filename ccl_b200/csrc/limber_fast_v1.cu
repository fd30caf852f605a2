// Fast 'spline' Limber kernels (sm_100a): ln k node grids shared across tracer pairs.
//
// Reference path: integ_cls_limber_spline (src/ccl_cls.c:189-225) -> cl_integrand (:166-187)
// -> ccl_integ_spline (src/ccl_utils.c:274-345).  The reference evaluates the integrand
// k P(k,a) D1 D2 independently for every (pair, ell).  Its node grid depends only on the pair's
// (chi_min, chi_max) (get_k_interval, :64-100), so pairs with equal ranges share every node:
// a(chi), k P(k, a) and each collection's D(node) are evaluated ONCE per grid and the per-pair
// remainder is one product plus the Akima integral.  Node values and the
// integrand's arithmetic are exactly those of the general kernel (kernels.cu, spline_task).
//
// Work decomposition: one CTA per (cosmology, ell); its 8 warps pull that cosmology's grids from
// a shared-memory counter.  A warp streams a grid in chunks of 32 nodes (lane = node): phase 1
// fills a 36-column shared-memory window (4 carried + 32 new) with k P and the collections' D,
// phase 2 forms each pair's integrand over the window and advances its Akima integral, which
// needs only f_{j-2..j+2} per node.  Anything the fast path does not cover (a node outside the
// a(chi) or P(k,a) tables, ...) is pushed to a redo list for the general kernel.
#include "device_eval.cuh"

namespace cclb {

namespace {

#ifndef CCLB_FW
#define CCLB_FW 8
#endif
#ifndef CCLB_NCG
#define CCLB_NCG 10
#endif
#ifndef CCLB_NPG
#define CCLB_NPG 10
#endif
#ifndef CCLB_LPC
#define CCLB_LPC 1
#endif
#ifndef CCLB_MINB
#define CCLB_MINB 3
#endif
constexpr int FW = CCLB_FW;     // warps per CTA
constexpr int NCG = CCLB_NCG;   // collections per grid task
constexpr int NPG = CCLB_NPG;   // pairs per grid task
static_assert(CCLB_NPG <= 16, "pair flags are packed two bits per pair");
constexpr int LPC = CCLB_LPC;   // ells per CTA (descriptor staging is amortised over them)
constexpr int NSG = 2 * CCLB_NCG;  // sub-tracers per grid task
constexpr int GB = 64;          // (ell, grid) tasks whose limits are precomputed per CTA pass

// Compact lookup descriptor of one 1-D spline, laid out so that a lookup reads it with three 16-byte
// shared-memory loads: {x0, xn}, {inv_bin, rec}, {lut, nlut - 1, n - 2}.
struct __align__(16) FastS1 {
  double x0, xn;
  double inv_bin;
  const double2* xr;
  const double4* cr;
  const int* lut;
  int nlutm1, last;
  int pad[2];
};

struct __align__(16) FastTr {
  FastS1 kernel;
  FastS1 fa;
  double pref[LPC];  // f_ell (/(l+1/2)^2 when der_bessel == -1) for each ell of this CTA
  int has_fa, der_bessel;
  int pad[2];
};

__device__ __forceinline__ FastS1 pack_s1(const Spline1D& s) {
  FastS1 f;
  f.x0 = s.x0; f.xn = s.xn; f.inv_bin = s.inv_bin; f.xr = s.xr; f.cr = s.cr; f.lut = s.lut; f.nlutm1 = s.nlut - 1; f.last = s.n - 2;
  return f;
}

// find_rec through the LUT only (built for every grid): the entry is the interval of a point just
// left of the bin, so the walk goes right only.
__device__ __forceinline__ Rec1 find_rec_lut(const FastS1& s, double v) {
  int j = (int)((v - s.x0) * s.inv_bin);
  j = min(max(j, 0), s.nlutm1);
  int i = __ldg(s.lut + j);
  const int last = s.last;
  auto load = [&](int q) {
    const double2 a = __ldg(s.xr + q);
    const double4 c = ldg4(s.cr + q);
    Rec1 o;
    o.x0 = a.x; o.x1 = a.y; o.y = c.x; o.b = c.y; o.c = c.z; o.d = c.w;
    return o;
  };
  Rec1 r = load(i);
  int steps = 0;
  while (v >= r.x1 && i < last) {
    if (++steps > 4) {  // strongly non-uniform grid: finish by bisection on the interval starts
      int lo = i, hi = last + 1;
      while (hi > lo + 1) {
        const int m = (hi + lo) >> 1;
        if (__ldg(&s.xr[m].x) > v) hi = m; else lo = m;
      }
      i = lo;
      r = load(i);
      break;
    }
    r = load(++i);
  }
  return r;
}

// ln k limits and node spacing of one (ell, grid) task (get_k_interval + ccl_linear_spacing)
struct GridPar {
  double lkmin, lkmax, dx, hl, inv_dx, inv_hl, ratio, iratio;
  int nk, grid, ls, pad;
};

// Per-CTA shared state; followed in shared memory by FastTr tr[ntr_max] and the per-warp blocks.
struct FastCta {
  Spline1D achi;
  F2D psp;
  int2 colls[FAST_MAX_COLLS];
  GridPar gp[GB];
  int next_task, pad;
};

__host__ __device__ inline size_t fast_tr_offset() { return (sizeof(FastCta) + 15) & ~size_t(15); }
__host__ __device__ inline size_t fast_warp_offset(int ntr_max) {
  return (fast_tr_offset() + (size_t)ntr_max * sizeof(FastTr) + 15) & ~size_t(15);
}

// sub-tracer list entry of a grid task: tracer index | collection slot << 8 | first << 16 | last << 17
constexpr unsigned SUB_FIRST = 1u << 16, SUB_LAST = 1u << 17;

struct FastWarp {
  double dflat[NCG * 32]; // D of the grid's collection c at this lane's node: [c * 32 + lane]
  double sf[2][36];       // integrand of the current pair: 4 carried + 32 new nodes (double-buffered)
  double fc[NPG][4];      // last 4 integrand values of the previous chunk, per pair
  double fe[NPG][8];      // f_0..f_2 and f_{nk-5}..f_{nk-1} per pair (end-node terms)
  int2 po[NPG];           // offsets of the pair's two collections in dflat
  double acc[NPG][32];    // per-lane partial integrals
  GridPair prs[NPG];
  unsigned subs[NSG];
};

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ double rec_value(const Rec1& r, double v) {
  const double d = v - r.x0;
  return r.y + d * (r.b + d * (r.c + d * r.d));
}

}  // namespace

// ------------------------------------------------------------------------------------------
// Grouping: one CTA per cosmology builds its grid descriptors.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) limber_group_v1_kernel(const LimberProblem P, const FastPlan G) {
  extern __shared__ double gsm[];
  const int ic = blockIdx.x;
  const int ncoll = P.ncoll, npair = P.npair;
  double* cmin = gsm;
  double* cmax = cmin + ncoll;
  double* pmin = cmax + ncoll;
  double* pmax = pmin + npair;
  int* leader = reinterpret_cast<int*>(pmax + npair);
  const Cosmo* cs = P.cosmos + ic;
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
  __syncthreads();
  if (threadIdx.x != 0) return;
  GridDesc* grids = G.grids + (size_t)ic * G.gstride;
  GridPair* gp = G.gpairs + (size_t)ic * G.gstride;
  int* gc = G.gcolls + (size_t)ic * G.cstride;
  int ng = 0, npw = 0, ncw = 0;
  for (int p = 0; p < npair; ++p) {
    if (leader[p] != p) continue;
    GridDesc cur{pmin[p], pmax[p], ncw, 0, npw, 0};
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
        grids[ng++] = cur;
        ncw += cur.ncoll;
        npw += cur.npair;
        cur = GridDesc{pmin[p], pmax[p], ncw, 0, npw, 0};
        cur_nsub = 0;
        sa = sb = -1;
      }
      if (sa < 0) { sa = cur.ncoll; gc[cur.coll_begin + cur.ncoll++] = pr.x; cur_nsub += P.colls[pr.x].y; }
      if (pr.y == pr.x) sb = sa;
      else if (sb < 0) { sb = cur.ncoll; gc[cur.coll_begin + cur.ncoll++] = pr.y; cur_nsub += P.colls[pr.y].y; }
      gp[cur.pair_begin + cur.npair++] = GridPair{q, (short)sa, (short)sb};
    }
    grids[ng++] = cur;
    ncw += cur.ncoll;
    npw += cur.npair;
  }
  G.ngrids[ic] = ng;
}

// ------------------------------------------------------------------------------------------
// Phase 2 of one chunk: per-pair integrand and Akima integral.
//
// With m_j the slopes and t_j the Akima node derivatives, GSL's piecewise cubic integrates to
//   sum_i h_i (f_i + f_{i+1}) / 2  +  (1/12) sum_{i: NE_i != 0} h_i^2 (t_i - tL_{i+1}),
// tL_{i+1} = t_{i+1} unless NE_{i+1} == 0 (then m_i).  Per node c the second sum contributes
// h^2 ([!z_c] t_c - [!z_{c-1}] (z_c ? m_{c-1} : t_c)) with z_c = (NE_c == 0): on the uniform part of the
// grid that vanishes unless z_c != z_{c-1}.  The chunk loop therefore adds trapezoids and tests for
// tie transitions (NE_c == 0 <=> m_{c+1} == m_c and m_{c-1} == m_{c-2}, i.e. equal consecutive first
// differences); only a chunk holding a transition evaluates the per-node form, for the interior
// nodes 1..nk-3.  The three end-node terms (nodes 0, nk-2, nk-1) are added once per pair after the
// loop from the saved values fe.  Lane `lane` holds node s = base + lane and owns the trapezoid of
// interval (s-1, s) and the correction term of node c = s - 2.
// ------------------------------------------------------------------------------------------
template <bool EDGE>
__device__ __forceinline__ void pair_phase(FastWarp& W, int np, int s, int base, int nk, int lane, double kp,
                                           double dx, double inv_dx, double inv_hl, unsigned& zdirty) {
  const unsigned FULL = 0xffffffffu;
  const bool valid = s < nk;
  const double wtrap = (s >= 1 && valid) ? 0.5 * dx : 0.0;  // the last interval is fixed up after the loop
  const int cn = s - 2;
  const bool cvalid = EDGE ? ((cn >= 0) && (cn <= nk - 3)) : true;
#pragma unroll 1
  for (int q = 0; q < np; ++q) {
    const int2 po = W.po[q];
    double* sf = W.sf[q & 1];
    // cl_integrand (src/ccl_cls.c:166-187): k P d1 d2, exactly 0 when either transfer vanishes
    // (lanes past the last node read stale shared memory: select, do not multiply by zero)
    const double f4 = valid ? kp * W.dflat[po.x + lane] * W.dflat[po.y + lane] : 0.0;
    sf[4 + lane] = f4;
    if (lane < 4) sf[lane] = W.fc[q][lane];
    if (EDGE) {  // f_0..f_2 and f_{nk-5}..f_{nk-1} feed the end-node terms
      if (s < 3) W.fe[q][s] = f4;
      const int r = s - (nk - 5);
      if (r >= 0 && valid) W.fe[q][3 + r] = f4;
    }
    __syncwarp();
    const double f0 = sf[lane], f1 = sf[lane + 1], f2 = sf[lane + 2], f3 = sf[lane + 3];
    if (lane >= 28) W.fc[q][lane - 28] = f4;
    const double da = f1 - f0, db = f2 - f1, dc = f3 - f2, dd = f4 - f3;
    bool z = (dd == dc) && (db == da);  // tie flag of node c
    if (EDGE) {
      // boundary slopes repeat the first difference of slopes: z_0 <=> m_1 == m_0; z_1 <=> m_2 == m_1 == m_0
      if (cn == 0) z = (dd == dc);
      if (cn == 1) z = (dd == dc) && (dc == db);
      z = z && cvalid;
    }
    const unsigned zm = __ballot_sync(FULL, z);
    const unsigned carry = (zdirty >> q) & 1u;
    unsigned trans = (zm ^ ((zm << 1) | carry));
    if (EDGE) {
      unsigned m = (base == 0) ? ~7u : ~0u;  // node 0 has no left neighbour term; nodes 1..nk-3 only
      if (nk - base < 32) m &= (1u << (nk - base)) - 1u;
      trans &= m;
    }
    double add = wtrap * (f3 + f4);
    if (trans != 0u) {
      // ---- a tie transition in this chunk: per-node form for its interior nodes ----
      double mm2 = da * inv_dx, mm1 = db * inv_dx, m0 = dc * inv_dx, mp1 = dd * inv_dx;
      if (EDGE) {
        if (cn == nk - 3) mp1 = dd * inv_hl;
        if (cn == 1) mm2 = 2.0 * mm1 - m0;
        if (cn == 0) { mm1 = 2.0 * m0 - mp1; mm2 = 3.0 * m0 - 2.0 * mp1; }
      }
      const bool zprev = lane ? ((zm >> (lane - 1)) & 1u) : (carry != 0u);
      if (((trans >> lane) & 1u) != 0u) {
        const double d12 = fabs(mm1 - mm2);
        const double NE = fabs(mp1 - m0) + d12;
        double tc = 0.0;
        if (!z && NE != 0.0) {
          const double alpha = d12 / NE;
          tc = (1.0 - alpha) * mm1 + alpha * m0;
        }
        double cc = z ? 0.0 : tc;
        if (!zprev) cc -= (z ? mm1 : tc);
        add += cc * (dx * dx * (1.0 / 12.0));
      }
    }
    W.acc[q][lane] += add;
    const unsigned top = (zm >> 31) & 1u;  // tie flag of this chunk's last node, carried to the next chunk
    if ((carry | top) != 0u) zdirty = (zdirty & ~(1u << q)) | (top << q);
  }
}

// Interior chunks (every lane holds a node in 4..nk-6 of the uniform part of the grid): the same sums
// without the shared-memory window.  The left neighbour's value and first difference come from two
// warp shuffles; the tie flags follow from ONE ballot of E_s = (d_s == d_{s-1}), d_s = f_s - f_{s-1}:
// z_c = E_{c+2} && E_c.  The previous chunk's last four values (W.fc, also what the edge form reads)
// supply lane 0's neighbour and the two carried E bits.
__device__ __forceinline__ void pair_phase_interior(FastWarp& W, int np, int lane, double kp, double dx,
                                                    double inv_dx, unsigned& zdirty) {
  const unsigned FULL = 0xffffffffu;
  const double wtrap = 0.5 * dx;
#pragma unroll 1
  for (int q = 0; q < np; ++q) {
    const int2 po = W.po[q];
    const double f4 = kp * W.dflat[po.x + lane] * W.dflat[po.y + lane];
    const double2 ca = *reinterpret_cast<const double2*>(&W.fc[q][0]);  // f at nodes base-4, base-3
    const double2 cb = *reinterpret_cast<const double2*>(&W.fc[q][2]);  // f at nodes base-2, base-1
    const double cd1 = cb.y - cb.x, cd2 = cb.x - ca.y, cd3 = ca.y - ca.x;  // d at base-1, base-2, base-3
    double f3 = __shfl_up_sync(FULL, f4, 1);
    if (lane == 0) f3 = cb.y;
    const double dd = f4 - f3;
    double dc = __shfl_up_sync(FULL, dd, 1);
    if (lane == 0) dc = cd1;
    const unsigned em = __ballot_sync(FULL, dd == dc);
    const unsigned ec = (cd2 == cd3 ? 1u : 0u) | (cd1 == cd2 ? 2u : 0u);  // E at nodes base-2, base-1
    const unsigned zm = em & ((em << 2) | ec);  // bit lane: tie flag of node c = base + lane - 2
    __syncwarp();  // every lane has read W.fc
    if (lane >= 28) W.fc[q][lane - 28] = f4;
    const unsigned carry = (zdirty >> q) & 1u;
    const unsigned trans = zm ^ ((zm << 1) | carry);
    double add = wtrap * (f3 + f4);
    if (trans != 0u) {
      // ---- a tie transition in this chunk: per-node form (see pair_phase) ----
      double db = __shfl_up_sync(FULL, dd, 2), da = __shfl_up_sync(FULL, dd, 3);
      if (lane < 2) db = (lane == 1) ? cd1 : cd2;
      if (lane < 3) da = (lane == 2) ? cd1 : ((lane == 1) ? cd2 : cd3);
      if (((trans >> lane) & 1u) != 0u) {
        const double mm2 = da * inv_dx, mm1 = db * inv_dx, m0 = dc * inv_dx, mp1 = dd * inv_dx;
        const bool z = ((zm >> lane) & 1u) != 0u;
        const bool zprev = lane ? (((zm >> (lane - 1)) & 1u) != 0u) : (carry != 0u);
        const double d12 = fabs(mm1 - mm2);
        const double NE = fabs(mp1 - m0) + d12;
        double tc = 0.0;
        if (!z && NE != 0.0) {
          const double alpha = d12 / NE;
          tc = (1.0 - alpha) * mm1 + alpha * m0;
        }
        double cc = z ? 0.0 : tc;
        if (!zprev) cc -= (z ? mm1 : tc);
        add += cc * (dx * dx * (1.0 / 12.0));
      }
    }
    W.acc[q][lane] += add;
    const unsigned top = (zm >> 31) & 1u;
    if ((carry | top) != 0u) zdirty = (zdirty & ~(1u << q)) | (top << q);
  }
}

// ------------------------------------------------------------------------------------------
// The fast integrator.
// ------------------------------------------------------------------------------------------
// RSD = true adds the der_bessel 1/2 sub-tracers (transfer_limber_single, src/ccl_cls.c:121-145): the
// second evaluation point chi' = (l+3/2)/k, a' = a(chi') and the ratio P(k,a')/P(k,a) are computed once per
// node and shared by every such sub-tracer of the grid.
template <bool RSD>
__global__ void __launch_bounds__(FW * 32, CCLB_MINB)
limber_spline_fastv1_kernel(const LimberProblem P, const FastPlan G) {
  extern __shared__ __align__(16) unsigned char smraw[];
  FastCta& S = *reinterpret_cast<FastCta*>(smraw);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  FastTr* const s_tr = reinterpret_cast<FastTr*>(smraw + fast_tr_offset());
  FastWarp& W = reinterpret_cast<FastWarp*>(smraw + fast_warp_offset(G.ntr_max))[warp];
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
  for (int c = threadIdx.x; c < P.ncoll; c += blockDim.x) S.colls[c] = P.colls[c];
  if (threadIdx.x == 32) S.achi = cs->achi;
  if (threadIdx.x == 64) S.psp = *cs->psp;

  const GridDesc* grids = G.grids + (size_t)ic * G.gstride;
  const GridPair* gpairs = G.gpairs + (size_t)ic * G.gstride;
  const int* gcolls = G.gcolls + (size_t)ic * G.cstride;
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
      q.hl = q.lkmax - (q.lkmin + q.dx * (q.nk - 2));  // last interval (end node overwritten)
      q.inv_dx = 1.0 / q.dx;
      q.inv_hl = 1.0 / q.hl;
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
      const int ncoll = gd.ncoll, np = gd.npair;
      __syncwarp();
      if (lane < np) {
        const GridPair pr = gpairs[gd.pair_begin + lane];
        W.prs[lane] = pr;
        W.po[lane] = make_int2(pr.a * 32, pr.b * 32);
      }
      // flatten the grid's collections into one sub-tracer list (lane c expands collection c)
      int2 cr = make_int2(0, 0);
      if (lane < ncoll) cr = S.colls[gcolls[gd.coll_begin + lane]];
      int off = cr.y;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(FULL, off, o);
        if (lane >= o) off += v;
      }
      bool task_rsd = false;
      if (RSD) {
        bool mine = false;
        for (int j = 0; j < cr.y; ++j) mine = mine || (s_tr[cr.x + j].der_bessel >= 1);
        task_rsd = __any_sync(FULL, mine);
      }
      const int nsub = __shfl_sync(FULL, off, 31);
      off -= cr.y;
      if (nsub <= NSG)
        for (int j = 0; j < cr.y; ++j)
          W.subs[off + j] = (unsigned)(cr.x + j) | ((unsigned)lane << 8) | (j == 0 ? SUB_FIRST : 0u) |
                            (j == cr.y - 1 ? SUB_LAST : 0u);
      const int nk = S.gp[gl].nk;
      const double lkmin = S.gp[gl].lkmin, lkmax = S.gp[gl].lkmax;
      __syncwarp();
      if (nk < 5) {  // gsl_spline_alloc(akima, n < 5) fails: NaN + CCL_ERROR_INTEG
        if (lane < np) {
          const int ip = W.prs[lane].out;
          const size_t s = (size_t)ic * P.npair + ip;
          P.cl[s * P.nl + il] = nan("");
          P.status[s] = ST_ERR_INTEG;
          atomicCAS(P.status_detail + s, 0, (int)ST_ERR_MEMORY);
        }
        continue;
      }
      // outside the P(k,a) table in ln k, or more sub-tracers than the list holds: general kernel
      bool redo = !(lkmin >= S.psp.lkmin && lkmax <= S.psp.lkmax) || nsub > NSG;
      const double dx = S.gp[gl].dx, hl = S.gp[gl].hl, inv_dx = S.gp[gl].inv_dx, inv_hl = S.gp[gl].inv_hl;
      // k and chi advance by constant ratios from chunk to chunk (one exp per lane per grid)
      double k = exp(lkmin + dx * lane);
      double chi = lp1h / k;
      unsigned zdirty = 0u;  // bit q: the last node of pair q's previous chunk was an Akima tie
      for (int i = lane; i < NPG * 4; i += 32) (&W.fc[0][0])[i] = 0.0;  // f at nodes -4..-1: finite placeholders
#pragma unroll 1
      for (int q = 0; q < np; ++q) W.acc[q][lane] = 0.0;

#pragma unroll 1
      for (int base = 0; base < nk && !redo; base += 32) {
        // ---------------- phase 1: node quantities (lane = node base + lane) ----------------
        const int s = base + lane;
        bool bad = false;
        double kp = 0.0;
        if (s < nk) {
          const double lk = (s == 0) ? lkmin : ((s == nk - 1) ? lkmax : lkmin + dx * s);
          // ccl_scale_factor_of_chi (src/ccl_background.c:1251-1277); chi > 0 here
          double a = 1.0;
          if (!(chi < 1.e-8)) {
            if (chi >= S.achi.x0 && chi <= S.achi.xn) a = rec_value(find_rec(S.achi, chi), chi);
            else bad = true;  // GSL_EDOM: the general kernel reports it
          }
          if (!(a >= S.psp.amin && a <= S.psp.amax)) bad = true;  // growth extrapolation: general kernel
          double pk = 0.0;
          if (!bad) pk = exp(bicubic_eval(S.psp, lk, a, 0));
          kp = k * pk;
          double chi_lp = 0.0, a_lp = 1.0, sqell = 0.0;
          if (RSD && task_rsd) {
            const double lp3h = lp1h + 1.0;
            chi_lp = lp3h / k;
            if (!(chi_lp < 1.e-8)) {
              if (chi_lp >= S.achi.x0 && chi_lp <= S.achi.xn) a_lp = rec_value(find_rec(S.achi, chi_lp), chi_lp);
              else bad = true;
            }
            if (!(a_lp >= S.psp.amin && a_lp <= S.psp.amax)) bad = true;
            if (!bad) sqell = sqrt(lp1h * fabs(exp(bicubic_eval(S.psp, lk, a_lp, 0)) / pk) / lp3h);
          }
          // transfer_limber_wrap (src/ccl_cls.c:150-164) for every collection of the grid, as one
          // branch-free loop over the flattened sub-tracer list so that independent table lookups
          // overlap.  Outside a kernel's support the lookup runs on the clamped abscissa and the
          // term is dropped (ccl_f1d_t_eval: the end nodes count as outside, src/ccl_f1d.c:137-161).
          double D = 0.0;
#pragma unroll 2
          for (int e = 0; e < nsub; ++e) {
            const unsigned ent = W.subs[e];
            const FastTr& tr = s_tr[ent & 0xffu];
            const bool inside = (chi > tr.kernel.x0) && (chi < tr.kernel.xn);
            double wt = rec_value(find_rec_lut(tr.kernel, chi), chi);  // index clamped inside; dropped if outside
            if (tr.has_fa) {
              const double a_ev = fmin(fmax(a, tr.fa.x0), tr.fa.xn);
              wt *= rec_value(find_rec_lut(tr.fa, a_ev), a_ev);
            }
            double val = inside ? wt * tr.pref[ls] : 0.0;
            if (RSD && tr.der_bessel >= 1) {
              // w t at chi is `inside ? wt : 0`; the second point has its own support test
              const double l = lp1h - 0.5, lp3h = lp1h + 1.0;
              const double w_t = inside ? wt : 0.0;
              const bool inside_p = (chi_lp > tr.kernel.x0) && (chi_lp < tr.kernel.xn);
              double wt_p = rec_value(find_rec_lut(tr.kernel, chi_lp), chi_lp);
              if (tr.has_fa) {
                const double a_ev = fmin(fmax(a_lp, tr.fa.x0), tr.fa.xn);
                wt_p *= rec_value(find_rec_lut(tr.fa, a_ev), a_ev);
              }
              if (!inside_p) wt_p = 0.0;
              double dd;
              if (tr.der_bessel == 1) dd = l * w_t / lp1h - sqell * wt_p;
              else dd = sqell * 2 * wt_p / lp3h - (0.25 + 2 * l) * w_t / (lp1h * lp1h);
              val = dd * tr.pref[ls];
            }
            D = (ent & SUB_FIRST) ? val : D + val;
            if (ent & SUB_LAST) W.dflat[((ent >> 8) & 0xffu) * 32 + lane] = D;
          }
          k *= S.gp[gl].ratio;
          chi *= S.gp[gl].iratio;
        }
        if (__any_sync(FULL, bad)) { redo = true; break; }
        // ---------------- phase 2: per-pair integrand and Akima integral ----------------
        if ((base == 0) || (base + 31 >= nk - 5))
          pair_phase<true>(W, np, s, base, nk, lane, kp, dx, inv_dx, inv_hl, zdirty);
        else
          pair_phase_interior(W, np, lane, kp, dx, inv_dx, zdirty);
        __syncwarp();
      }

      if (redo) {
        if (lane < np) {
          const unsigned long long task =
              ((unsigned long long)ic * P.npair + W.prs[lane].out) * P.nl + il;
          const unsigned long long at = atomicAdd(G.redo, 1ull);
          if (at < G.redo_cap) G.redo[1 + at] = task;
        }
        continue;
      }
      // ---- end-node terms, one lane per pair (nodes 0, nk-2, nk-1; gsl akima_init boundary slopes) ----
      double endc = 0.0;
      if (lane < np) {
        const double* e = W.fe[lane];
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
        endc = endc * (1.0 / 12.0) + 0.5 * (hl - dx) * (g3 + g4);
      }
#pragma unroll 1
      for (int q = 0; q < np; ++q) {
        const double v = warp_sum(W.acc[q][lane]) + __shfl_sync(FULL, endc, q);
        if (lane == 0) P.cl[((size_t)ic * P.npair + W.prs[q].out) * P.nl + il] = v / lp1h;  // src/ccl_cls.c:340
      }
    }
  }
}

cudaError_t launch_limber_fast_v1(const LimberProblem& prob, const FastPlan& plan, cudaStream_t stream) {
  if (prob.ncosmo == 0 || prob.npair == 0 || prob.nl == 0) return cudaSuccess;
  if ((long long)prob.ncosmo * prob.nl > 0x7fffffffLL) return cudaErrorInvalidValue;
  cudaMemsetAsync(plan.redo, 0, sizeof(unsigned long long), stream);
  const size_t gsm = sizeof(double) * 2 * ((size_t)prob.ncoll + prob.npair) + sizeof(int) * (size_t)prob.npair;
  if (gsm > 48 * 1024)
    cudaFuncSetAttribute(limber_group_v1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gsm);
  limber_group_v1_kernel<<<prob.ncosmo, 128, gsm, stream>>>(prob, plan);
  const size_t fsm = fast_warp_offset(plan.ntr_max) + FW * sizeof(FastWarp);
  static size_t attr_set[64] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  if (attr_set[dev & 63] < fsm) {
    cudaFuncSetAttribute(limber_spline_fastv1_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fsm);
    cudaFuncSetAttribute(limber_spline_fastv1_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fsm);
    attr_set[dev & 63] = fsm;
  }
  const long long nlt = (prob.nl + LPC - 1) / LPC;
  const unsigned nblk = (unsigned)((long long)prob.ncosmo * nlt);
  if (plan.has_rsd) limber_spline_fastv1_kernel<true><<<nblk, FW * 32, fsm, stream>>>(prob, plan);
  else limber_spline_fastv1_kernel<false><<<nblk, FW * 32, fsm, stream>>>(prob, plan);
  launch_limber_redo(prob, plan, stream);
  return cudaGetLastError();
}

}  // namespace cclb
