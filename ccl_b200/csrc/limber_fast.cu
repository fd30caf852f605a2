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

constexpr int FW = 8;     // warps per CTA
constexpr int NCG = 12;   // collections per grid task
constexpr int NPG = 12;   // pairs per grid task
constexpr int WIN = 36;   // window columns

struct FastTr {
  Spline1D kernel;
  Spline1D fa;
  double pref;  // f_ell (/(l+1/2)^2 when der_bessel == -1)
  int has_fa, pad;
};

struct FastCta {
  Spline1D achi;
  F2D psp;
  FastTr tr[FAST_MAX_TRACERS];
  int2 colls[FAST_MAX_COLLS];
  int next_grid, ngrids;
};

struct FastWarp {
  double w[1 + NCG][WIN];  // row 0: k P(k,a); row 1+c: D of the grid's collection c
  double sf[2][WIN + 4];   // integrand of the current pair over the window (double-buffered)
  double acc[NPG][32];     // per-lane partial integrals
  GridPair prs[NPG];
  int cl[NCG];
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
__global__ void __launch_bounds__(128) limber_group_kernel(const LimberProblem P, const FastPlan G) {
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
    for (int q = p; q < npair; ++q) {
      if (leader[q] != p) continue;
      const int2 pr = P.pairs[q];
      int sa = -1, sb = -1;
      for (int i = 0; i < cur.ncoll; ++i) {
        if (gc[cur.coll_begin + i] == pr.x) sa = i;
        if (gc[cur.coll_begin + i] == pr.y) sb = i;
      }
      const int need = (sa < 0) + ((sb < 0) && (pr.y != pr.x));
      if (cur.ncoll + need > NCG || cur.npair + 1 > NPG) {  // close this task, open another on the same grid
        grids[ng++] = cur;
        ncw += cur.ncoll;
        npw += cur.npair;
        cur = GridDesc{pmin[p], pmax[p], ncw, 0, npw, 0};
        sa = sb = -1;
      }
      if (sa < 0) { sa = cur.ncoll; gc[cur.coll_begin + cur.ncoll++] = pr.x; }
      if (pr.y == pr.x) sb = sa;
      else if (sb < 0) { sb = cur.ncoll; gc[cur.coll_begin + cur.ncoll++] = pr.y; }
      gp[cur.pair_begin + cur.npair++] = GridPair{q, (short)sa, (short)sb};
    }
    grids[ng++] = cur;
    ncw += cur.ncoll;
    npw += cur.npair;
  }
  G.ngrids[ic] = ng;
}

// ------------------------------------------------------------------------------------------
// The fast integrator.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(FW * 32, 3) limber_spline_fast_kernel(const LimberProblem P, const FastPlan G) {
  extern __shared__ __align__(16) unsigned char smraw[];
  FastCta& S = *reinterpret_cast<FastCta*>(smraw);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  FastWarp& W = reinterpret_cast<FastWarp*>(smraw + ((sizeof(FastCta) + 15) & ~size_t(15)))[warp];
  const int il = blockIdx.x % P.nl;
  const int ic = blockIdx.x / P.nl;
  const Cosmo* cs = P.cosmos + ic;
  const double l = P.ells[il];
  const double lp1h = l + 0.5;

  // ---- per-CTA staging of the descriptors this (cosmology, ell) needs ----
  for (int t = threadIdx.x; t < cs->ntracers; t += blockDim.x) {
    const Tracer& tr = cs->tracers[t];
    S.tr[t].kernel = tr.kernel;
    S.tr[t].fa = tr.transfer.fa;
    S.tr[t].has_fa = tr.has_transfer;
    S.tr[t].pref = tracer_prefactor(tr, l);
  }
  for (int c = threadIdx.x; c < P.ncoll; c += blockDim.x) S.colls[c] = P.colls[c];
  if (threadIdx.x == 32) S.achi = cs->achi;
  if (threadIdx.x == 64) S.psp = *cs->psp;
  if (threadIdx.x == 0) {
    S.next_grid = 0;
    S.ngrids = G.ngrids[ic];
  }
  __syncthreads();

  const GridDesc* grids = G.grids + (size_t)ic * G.gstride;
  const GridPair* gpairs = G.gpairs + (size_t)ic * G.gstride;
  const int* gcolls = G.gcolls + (size_t)ic * G.cstride;
  const unsigned FULL = 0xffffffffu;

  for (;;) {
    int g = 0;
    if (lane == 0) g = atomicAdd(&S.next_grid, 1);
    g = __shfl_sync(FULL, g, 0);
    if (g >= S.ngrids) break;
    const GridDesc gd = grids[g];
    const int ncoll = gd.ncoll, np = gd.npair;
    __syncwarp();
    if (lane < np) W.prs[lane] = gpairs[gd.pair_begin + lane];
    if (lane < ncoll) W.cl[lane] = gcolls[gd.coll_begin + lane];
    // get_k_interval (src/ccl_cls.c:91-99)
    double chi_min = gd.chi_min;
    if (chi_min <= 0) chi_min = 0.5 * lp1h / P.par.K_MAX;
    const double lkmax = log(fmin(P.par.K_MAX, 2 * lp1h / chi_min));
    const double lkmin = log(fmax(P.par.K_MIN, lp1h / gd.chi_max));
    const int nk = (int)(fmax((lkmax - lkmin) / P.par.DLOGK_INTEGRATION + 0.5, 1.0)) + 1;
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
    bool redo = !(lkmin >= S.psp.lkmin && lkmax <= S.psp.lkmax);
    // ccl_linear_spacing (src/ccl_utils.c:17-37); k and chi advance by constant ratios per chunk
    const double dx = (lkmax - lkmin) / (nk - 1.);
    const double hl = lkmax - (lkmin + dx * (nk - 2));  // last interval (end node overwritten)
    const double inv_dx = 1.0 / dx;
    double k = exp(lkmin + dx * lane);
    double chi = lp1h / k;
    const double ratio = exp(32.0 * dx);
    const double iratio = 1.0 / ratio;
    unsigned zcarry = 0;  // bit q: Akima tie flag of the last node of the previous chunk, pair q
#pragma unroll 1
    for (int q = 0; q < np; ++q) W.acc[q][lane] = 0.0;

#pragma unroll 1
    for (int base = 0; base < nk + 2 && !redo; base += 32) {
      if (base > 0 && lane < 4) {
#pragma unroll 1
        for (int r = 0; r <= ncoll; ++r) W.w[r][lane] = W.w[r][32 + lane];
      }
      __syncwarp();
      // ---------------- phase 1: node quantities ----------------
      const int s = base + lane;
      bool bad = false;
      if (s < nk) {
        const double lk = (s == 0) ? lkmin : ((s == nk - 1) ? lkmax : lkmin + dx * s);
        // ccl_scale_factor_of_chi (src/ccl_background.c:1251-1277); chi > 0 here
        double a = 1.0;
        if (!(chi < 1.e-8)) {
          if (chi >= S.achi.x0 && chi <= S.achi.xn) a = rec_value(find_rec(S.achi, chi), chi);
          else bad = true;  // GSL_EDOM: the general kernel reports it
        }
        if (!(a >= S.psp.amin && a <= S.psp.amax)) bad = true;  // growth extrapolation: general kernel
        double kp = 0.0;
        if (!bad) kp = k * exp(bicubic_eval(S.psp, lk, a, 0));
        W.w[0][4 + lane] = kp;
#pragma unroll 1
        for (int c = 0; c < ncoll; ++c) {
          const int2 cr = S.colls[W.cl[c]];
          double D = 0.0;
#pragma unroll 1
          for (int t = cr.x; t < cr.x + cr.y; ++t) {
            const FastTr& tr = S.tr[t];
            // ccl_f1d_t_eval: the end nodes count as outside (src/ccl_f1d.c:137-161)
            if (chi > tr.kernel.x0 && chi < tr.kernel.xn) {
              double wt = rec_value(find_rec(tr.kernel, chi), chi);
              if (tr.has_fa) {
                const double a_ev = fmin(fmax(a, tr.fa.x0), tr.fa.xn);
                wt *= rec_value(find_rec(tr.fa, a_ev), a_ev);
              }
              D += wt * tr.pref;
            }
          }
          W.w[1 + c][4 + lane] = D;
        }
        k *= ratio;
        chi *= iratio;
      }
      if (__any_sync(FULL, bad)) { redo = true; break; }
      __syncwarp();
      // ---------------- phase 2: per-pair integrand and Akima integral ----------------
      const int cn = base + lane - 2;  // node integrated by this lane (window column lane + 2)
      const bool cvalid = (cn >= 0) && (cn < nk);
      const bool edge = (cn < 2) || (cn > nk - 4);
#pragma unroll 1
      for (int q = 0; q < np; ++q) {
        const GridPair pr = W.prs[q];
        const double* da = W.w[1 + pr.a];
        const double* db = W.w[1 + pr.b];
        double* sf = W.sf[q & 1];
        {
          // cl_integrand (src/ccl_cls.c:166-187): 0 when either transfer vanishes
          const double d1 = da[4 + lane], d2 = db[4 + lane];
          sf[4 + lane] = (d1 == 0.0 || d2 == 0.0) ? 0.0 : W.w[0][4 + lane] * d1 * d2;
          if (lane < 4) {
            const double e1 = da[lane], e2 = db[lane];
            sf[lane] = (e1 == 0.0 || e2 == 0.0) ? 0.0 : W.w[0][lane] * e1 * e2;
          }
        }
        __syncwarp();
        const double f0 = sf[lane], f1 = sf[lane + 1], f2 = sf[lane + 2], f3 = sf[lane + 3], f4 = sf[lane + 4];
        // slopes m_{c-2}, m_{c-1}, m_c, m_{c+1} (gsl akima_init boundary rules at the ends)
        double mm2 = (f1 - f0) * inv_dx, mm1 = (f2 - f1) * inv_dx, m0 = (f3 - f2) * inv_dx,
               mp1 = (f4 - f3) * inv_dx;
        double h = dx, hprev = dx;
        if (edge) {
          if (cn == nk - 3) mp1 = (f4 - f3) / hl;
          if (cn == nk - 2) { m0 = (f3 - f2) / hl; h = hl; mp1 = 2.0 * m0 - mm1; }
          if (cn == nk - 1) {
            mm1 = (f2 - f1) / hl;
            hprev = hl;
            m0 = 2.0 * mm1 - mm2;
            mp1 = 3.0 * mm1 - 2.0 * mm2;
          }
          if (cn == 1) mm2 = 2.0 * mm1 - m0;
          if (cn == 0) { mm1 = 2.0 * m0 - mp1; mm2 = 3.0 * m0 - 2.0 * mp1; }
        }
        const double d12 = fabs(mm1 - mm2);
        const double NE = fabs(mp1 - m0) + d12;
        const bool z = (NE == 0.0);
        const unsigned zm = __ballot_sync(FULL, cvalid && z);
        const bool zprev = lane ? ((zm >> (lane - 1)) & 1u) : ((zcarry >> q) & 1u);
        zcarry = (zcarry & ~(1u << q)) | (((zm >> 31) & 1u) << q);
        if (cvalid) {
          double add = 0.0;
          if (cn < nk - 1) add = 0.5 * h * (f2 + f3);  // trapezoid part of interval cn
          if (edge || z || zprev) {
            // h_c^2 b_c - h_{c-1}^2 tL_c: vanishes at regular interior nodes (uniform h, no ties)
            double tc = 0.0;
            if (!z) {
              const double alpha = d12 / NE;
              tc = (1.0 - alpha) * mm1 + alpha * m0;
            }
            double cc = 0.0;
            if (cn < nk - 1 && !z) cc = h * h * tc;
            if (cn >= 1 && !zprev) cc -= hprev * hprev * (z ? mm1 : tc);
            add += cc * (1.0 / 12.0);
          }
          W.acc[q][lane] += add;
        }
      }
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
#pragma unroll 1
    for (int q = 0; q < np; ++q) {
      const double v = warp_sum(W.acc[q][lane]);
      if (lane == 0) P.cl[((size_t)ic * P.npair + W.prs[q].out) * P.nl + il] = v / lp1h;  // src/ccl_cls.c:340
    }
  }
}

size_t limber_fast_plan_bytes(int ncosmo, int npair, unsigned long long redo_cap) {
  auto up = [](size_t b) { return (b + 255) / 256 * 256; };
  const size_t nc = (size_t)ncosmo, np = (size_t)npair;
  return up(nc * np * sizeof(GridDesc)) + up(nc * np * sizeof(GridPair)) + up(nc * 2 * np * sizeof(int)) +
         up(nc * sizeof(int)) + up((redo_cap + 1) * sizeof(unsigned long long));
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
  G.redo = reinterpret_cast<unsigned long long*>(p);
  G.gstride = npair;
  G.cstride = 2 * npair;
  G.redo_cap = redo_cap;
  return G;
}

cudaError_t launch_limber_fast(const LimberProblem& prob, const FastPlan& plan, cudaStream_t stream) {
  if (prob.ncosmo == 0 || prob.npair == 0 || prob.nl == 0) return cudaSuccess;
  if ((long long)prob.ncosmo * prob.nl > 0x7fffffffLL) return cudaErrorInvalidValue;
  cudaMemsetAsync(plan.redo, 0, sizeof(unsigned long long), stream);
  const size_t gsm = sizeof(double) * 2 * ((size_t)prob.ncoll + prob.npair) + sizeof(int) * (size_t)prob.npair;
  if (gsm > 48 * 1024)
    cudaFuncSetAttribute(limber_group_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gsm);
  limber_group_kernel<<<prob.ncosmo, 128, gsm, stream>>>(prob, plan);
  const size_t fsm = ((sizeof(FastCta) + 15) & ~size_t(15)) + FW * sizeof(FastWarp);
  static bool attr_set[64] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  if (!attr_set[dev & 63]) {
    cudaFuncSetAttribute(limber_spline_fast_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fsm);
    attr_set[dev & 63] = true;
  }
  limber_spline_fast_kernel<<<(unsigned)((long long)prob.ncosmo * prob.nl), FW * 32, fsm, stream>>>(prob, plan);
  launch_limber_redo(prob, plan, stream);
  return cudaGetLastError();
}

}  // namespace cclb
