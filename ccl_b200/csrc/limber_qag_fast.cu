// Fast 'qag_quad' Limber kernel (sm_100a): gsl_integration_qag panels shared across the pairs of a node grid.
//
// Reference path: integ_cls_limber_qag_quad (src/ccl_cls.c:227-263) -> gsl_integration_qag(F, lkmin, lkmax, 0,
// epsrel, limit, GSL_INTEG_GAUSS41) with F = cl_integrand (:166-187), once per (pair, ell).  QAG is adaptive, so
// every C_ell bisects its own intervals; but its panels are always dyadic sub-intervals of [lkmin, lkmax], and
// [lkmin, lkmax] depends only on the pair's (chi_min, chi_max) (get_k_interval, :64-100).  Pairs with equal
// ranges (the grid tasks limber_group_kernel builds for the 'spline' integrator) therefore request the SAME
// panels again and again: the first panel and its two halves always, deeper ones whenever the integrands peak in
// the same place.  One warp runs the QAG state machines of a task's pairs one after the other over a per-warp
// PANEL CACHE, filled lazily at two levels: the node part of a panel (k, chi, a(chi), k P(k,a) at its 41
// Gauss-Kronrod abscissae) is evaluated the first time ANY pair of the task asks for the panel, the transfer D of a
// collection on that panel the first time a pair USING that collection asks for it -- with the lean lookups of the
// 'spline' kernel (32-byte coefficient records, descriptors in shared memory).  A pair that finds everything
// cached only forms k P D1 D2 and the four GK sums.  Both children of a bisection are evaluated together
// (82 abscissae over three lane passes).  Nothing is ever evaluated that the reference would not evaluate.  The bookkeeping (interval list, error sums, round-off counters, stop tests) is GSL's, in the
// form of the general kernel (kernels.cu, limber_qag_kernel), which also serves as the redo path for whatever
// this kernel does not cover: nodes outside the a(chi) / P(k,a) tables, der_bessel 1/2 sub-tracers, oversize tasks.
#include <cfloat>
#include <cstddef>

#include "fast_common.cuh"
#include "gk41_table.h"

namespace cclb {


namespace {

#ifndef CCLB_QFW
#define CCLB_QFW 12
#endif
constexpr int QFW = CCLB_QFW;      // warps per CTA
constexpr int NCG = CCLB_NCG;     // collections per grid task
constexpr int NPG = CCLB_NPG;     // pairs per grid task
constexpr int NSG = 2 * CCLB_NCG; // sub-tracers per grid task
constexpr int NPC = 24;           // cached panels per warp (the last two slots are scratch once the cache is full)
constexpr int PN = 41;            // abscissae per panel
constexpr int PS = 48;            // row stride of a cached panel (doubles)
constexpr int PROWS = NCG + 3;    // k P, chi, a rows + one D row per collection
constexpr int QLB = 10;           // ells per CTA work item
#ifndef CCLB_QCTAS
#define CCLB_QCTAS 2
#endif
constexpr int QCTAS_PER_SM = CCLB_QCTAS;

__constant__ double q_xgk[21];
__constant__ double q_wgk[21];
__constant__ double q_wg[10];

struct QCta {
  Spline1D achi;
  F2D psp;
  int der_angles[FAST_MAX_TRACERS];
  int next_task, item;
};

struct __align__(16) QWarp {
  FastEnt ents[NSG];    // the task's sub-tracers, ordered by collection slot
  double2 keys[NPC];    // (a, b) of the cached panels
  int2 po[NPG];         // collection slots of the task's pairs
  int2 cs[NCG];         // {first entry, entries} of every collection slot
  unsigned dmask[NPC];  // bit c: the D row of collection slot c is valid in this panel
  unsigned raw[NSG];    // the task's list entries as limber_group_kernel wrote them
};

// Layout of the CTA's dynamic shared memory; tr[] really holds FastPlan::ntr_max entries.
struct __align__(16) QSmem {
  QCta cta;
  QWarp warps[QFW];
  FastTr tr[1];
};

__device__ __forceinline__ double qwarp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// rescale_error (gsl integration/qk.c)
__device__ __forceinline__ double q_rescale_error(double err, double result_abs, double result_asc) {
  err = fabs(err);
  if (result_asc != 0 && err != 0) {
    // pow(x, 1.5) as x sqrt(x): both are correctly rounded to within an ulp, and the value only ranks intervals
    const double x = 200 * err / result_asc;
    const double scale = x * sqrt(x);
    err = (scale < 1) ? result_asc * scale : result_asc;
  }
  if (result_abs > DBL_MIN / (50 * DBL_EPSILON)) {
    const double min_err = 50 * DBL_EPSILON * result_abs;
    if (min_err > err) err = min_err;
  }
  return err;
}

// subinterval_too_small (gsl integration/util.c)
__device__ __forceinline__ bool q_subinterval_too_small(double a1, double a2, double b2) {
  const double tmp = (1 + 100 * DBL_EPSILON) * (fabs(a2) + 1000 * DBL_MIN);
  return fabs(a1) <= tmp && fabs(b2) <= tmp;
}

// gsl_integration_qk with the 41-point rule on ONE or TWO cached panels at once: lanes 0..15 work on panel 0,
// lanes 16..31 on panel 1.  A half-warp lane h holds the symmetric pair of abscissae (h, h + 20) and, for
// h < 5, also the item h + 16 (pairs 16..19 and the centre, item 20); the four sums are reduced by a 4-step
// butterfly inside each half-warp (both panels for the price of one), and the halves exchange their results.
// panel: [PROWS][PS] doubles, row 0 = k P, row 3 + c = D of collection slot c.  Same sums as the general
// kernel's qk41_warp up to the association of the additions.
__device__ __forceinline__ void panel_sums2(const double* panel0, const double* panel1, int npan, int ca, int cb,
                                            const double* qa, const double* qb, int lane, double* result,
                                            double* abserr, double* resabs, double* resasc) {
  const unsigned FULL = 0xffffffffu;
  const int half = lane >> 4, h = lane & 15;
  const bool active = half < npan;
  const double* panel = half ? panel1 : panel0;
  const double* kp = panel;
  const double* da = panel + (3 + ca) * PS;
  const double* db = panel + (3 + cb) * PS;
  const double a = (half && npan > 1) ? qa[1] : qa[0], b = (half && npan > 1) ? qb[1] : qb[0];
  const double half_length = 0.5 * (b - a);
  const double abs_half_length = fabs(half_length);
  double rk = 0, rg = 0, ra = 0;
  double f1[2] = {0, 0}, f2[2] = {0, 0};
  bool centre[2] = {false, false}, used[2] = {false, false};
  if (active) {
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int i = h + 16 * u;  // item: 0..19 symmetric pairs, 20 the centre
      if (i > 20) continue;
      used[u] = true;
      if (i < 20) {
        // cl_integrand (src/ccl_cls.c:166-187): k P d1 d2
        f1[u] = __ldcg(kp + i) * __ldcg(da + i) * __ldcg(db + i);
        f2[u] = __ldcg(kp + i + 20) * __ldcg(da + i + 20) * __ldcg(db + i + 20);
        const double fsum = f1[u] + f2[u];
        rk += q_wgk[i] * fsum;
        ra += q_wgk[i] * (fabs(f1[u]) + fabs(f2[u]));
        if (i & 1) rg += q_wg[i >> 1] * fsum;
      } else {
        centre[u] = true;
        f1[u] = __ldcg(kp + 40) * __ldcg(da + 40) * __ldcg(db + 40);
        rk += f1[u] * q_wgk[20];
        ra += fabs(f1[u] * q_wgk[20]);
      }
    }
  }
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) {
    rk += __shfl_xor_sync(FULL, rk, o);
    rg += __shfl_xor_sync(FULL, rg, o);
    ra += __shfl_xor_sync(FULL, ra, o);
  }
  const double mean = rk * 0.5;
  double rs = 0;
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    if (!used[u]) continue;
    const int i = h + 16 * u;
    if (centre[u]) rs += q_wgk[20] * fabs(f1[u] - mean);
    else rs += q_wgk[i] * (fabs(f1[u] - mean) + fabs(f2[u] - mean));
  }
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) rs += __shfl_xor_sync(FULL, rs, o);
  const double err = (rk - rg) * half_length;
  rk *= half_length;
  ra *= abs_half_length;
  rs *= abs_half_length;
  const double ae = q_rescale_error(err, ra, rs);
  // every lane leaves with both panels' results
  const double ork = __shfl_xor_sync(FULL, rk, 16), ora = __shfl_xor_sync(FULL, ra, 16);
  const double ors = __shfl_xor_sync(FULL, rs, 16), oae = __shfl_xor_sync(FULL, ae, 16);
  result[0] = half ? ork : rk; resabs[0] = half ? ora : ra; resasc[0] = half ? ors : rs; abserr[0] = half ? oae : ae;
  result[1] = half ? rk : ork; resabs[1] = half ? ra : ora; resasc[1] = half ? rs : ors; abserr[1] = half ? ae : oae;
}

}  // namespace

__global__ void __launch_bounds__(QFW * 32, QCTAS_PER_SM)
limber_qag_fast_kernel(const LimberProblem P, const FastPlan G, double* ws, double* cache,
                       unsigned long long* next_item, unsigned long long* cq_list, unsigned long long cq_cap) {
  extern __shared__ __align__(16) QSmem qsm[];
  QCta& S = qsm[0].cta;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  FastTr* const s_tr = qsm[0].tr;
  QWarp& W = qsm[0].warps[warp];
  const unsigned FULL = 0xffffffffu;
  const int limit = P.par.N_ITERATION;
  const size_t gw = (size_t)blockIdx.x * QFW + warp;
  double* alist = ws + gw * 4 * (size_t)limit;
  double* blist = alist + limit;
  double* rlist = blist + limit;
  double* elist = rlist + limit;
  double* pcache = cache + gw * (size_t)NPC * PROWS * PS;
  // work item of a CTA: one cosmology x a block of QLB ells (its descriptors are staged once, its (ell, grid)
  // tasks are pulled by the warps from a shared counter: one barrier per few hundred tasks)
  const int nlb = (P.nl + QLB - 1) / QLB;
  const long long nitems = (long long)P.ncosmo * nlb;

  for (;;) {
    __syncthreads();  // the previous item is drained
    if (threadIdx.x == 0) {
      S.item = (int)min((unsigned long long)nitems, atomicAdd(next_item, 1ull));
      S.next_task = 0;
    }
    __syncthreads();
    const long long item = S.item;
    if (item >= nitems) break;
    const int il0 = (int)(item % nlb) * QLB;
    const int nls = min(QLB, P.nl - il0);
    const int ic = (int)(item / nlb);
    const Cosmo* cs = P.cosmos + ic;

    // ---- per-item staging of the descriptors this cosmology needs ----
    for (int t = threadIdx.x; t < cs->ntracers; t += blockDim.x) {
      const Tracer& tr = cs->tracers[t];
      s_tr[t].kernel = pack_s1(tr.kernel);
      s_tr[t].fa = pack_s1(tr.transfer.fa);
      const bool fold = tr.has_transfer && tr.const_transfer;  // see limber_fast.cu
      s_tr[t].has_fa = tr.has_transfer && !fold;
      s_tr[t].der_bessel = tr.der_bessel;
      s_tr[t].pref[0] = fold ? tr.const_transfer_value : 1.0;  // the ell prefactor is applied per task
      S.der_angles[t] = tr.der_angles;
    }
    if (threadIdx.x == 32) S.achi = cs->achi;
    if (threadIdx.x == 64) S.psp = *cs->psp;
    const GridDesc* grids = G.grids + (size_t)ic * G.gstride;
    const GridPair* gpairs = G.gpairs + (size_t)ic * G.gstride;
    const unsigned* gsubs = G.gsubs + (size_t)ic * G.sstride;
    const int ngrids = G.ngrids[ic];
    const int ntasks = ngrids * nls;
    __syncthreads();

    for (;;) {
      int task = 0;
      if (lane == 0) task = atomicAdd(&S.next_task, 1);
      task = __shfl_sync(FULL, task, 0);
      if (task >= ntasks) break;
      const int il = il0 + task / ngrids;
      const GridDesc gd = grids[task % ngrids];
      const double ell = P.ells[il];
      const double lp1h = ell + 0.5;
      const int ncoll = gd.ncoll, np = gd.npair, nsub = gd.nsub;
      // get_k_interval (src/ccl_cls.c:91-99)
      double chi_min = gd.chi_min;
      if (chi_min <= 0) chi_min = 0.5 * lp1h / P.par.K_MAX;
      const double lkmax = log(fmin(P.par.K_MAX, 2 * lp1h / chi_min));
      const double lkmin = log(fmax(P.par.K_MIN, lp1h / gd.chi_max));
      __syncwarp();
      if (lane < np) {
        const GridPair pr = gpairs[gd.pair_begin + lane];
        W.po[lane] = make_int2(pr.a, pr.b);
      }
      const int nent = min(nsub, NSG);
      if (lane < nent) W.raw[lane] = gsubs[gd.sub_begin + lane];
      __syncwarp();
      if (lane < nent) {
        // order by collection slot (stable): collections are evaluated one at a time, on demand
        const unsigned ent = W.raw[lane];
        const unsigned slot = (ent >> 8) & 0xfu;
        int rank = 0;
        for (int j = 0; j < nent; ++j) {
          const unsigned sj = (W.raw[j] >> 8) & 0xfu;
          rank += (sj < slot) || (sj == slot && j < lane);
        }
        const int t = (int)(ent & 0xffu);
        const FastTr& tr = s_tr[t];
        // ccl_cl_tracer_t_get_f_ell (src/ccl_tracers.c:703-726), /(l+1/2)^2 for der_bessel == -1 (src/ccl_cls.c:118)
        double pref = tracer_f_ell(S.der_angles[t], ell);
        if (tr.der_bessel == -1) pref /= (lp1h * lp1h);
        FastEnt en;
        en.kcr = tr.kernel.cr;
        en.pref = pref * tr.pref[0];
        en.acr = tr.has_fa ? tr.fa.cr : nullptr;
        en.flags = ent & 0xfffu;
        en.addr = (unsigned)(reinterpret_cast<const char*>(&tr) - reinterpret_cast<const char*>(&W.ents[rank]));
        W.ents[rank] = en;
      }
      if (lane < ncoll) {
        int start = 0, cnt = 0;
        for (int j = 0; j < nent; ++j) {
          const unsigned sj = (W.raw[j] >> 8) & 0xfu;
          start += sj < (unsigned)lane;
          cnt += sj == (unsigned)lane;
        }
        W.cs[lane] = make_int2(start, cnt);
      }
      __syncwarp();
      // outside the P(k,a) table in ln k, der_bessel 1/2 sub-tracers, oversize list: general kernel
      bool redo = !(lkmin >= S.psp.lkmin && lkmax <= S.psp.lkmax) || nsub > NSG || gd.has_rsd;
      int ncached = 0;
      unsigned long long nodes_done = 0;

#pragma unroll 1
      for (int q = 0; q < np && !redo; ++q) {
        const int2 po = W.po[q];
        const unsigned pmask = (1u << po.x) | (1u << po.y);  // the two collections of this pair
        const int ipair = gpairs[gd.pair_begin + q].out;
        const double epsabs = 0.0, epsrel = P.par.LIMBER_EPSREL;
        int gsl = 0;
        unsigned long long nev = 0;
        double result = 0.0;
        bool running = true;
        if (epsabs <= 0 && (epsrel < 50 * DBL_EPSILON || epsrel < 0.5e-28)) {
          gsl = ST_GSL_EBADTOL;
          nev = 41;
          running = false;
        } else if (P.par.force_cquad) {
          gsl = ST_GSL_EROUND;  // tests of the fallback branch: skip QAG
          running = false;
        }
        // gsl_integration_qag as ONE loop: step 0 requests the whole interval, every later step the two halves of
        // the interval with the largest error estimate (one call site each for the cache look-up, the node part,
        // the transfers and the GK sums keeps the kernel inside the instruction cache)
        int size = 0, iteration = 0, roundoff_type1 = 0, roundoff_type2 = 0, error_type = 0, imax = 0;
        double area = 0.0, errsum = 0.0, tolerance = 0.0, r_i = 0.0, e_i = 0.0;
        double qa[2] = {lkmin, 0.0}, qb[2] = {lkmax, 0.0};
        int npan = 1;
        while (running) {
          // ---- panel cache: hit -> its slot; miss -> a fresh slot, or one of the two scratch slots when full ----
          int sl[2] = {0, 0};
          double pa[2], pb[2];
          int ps[2], nmiss = 0;
#pragma unroll 1
          for (int p = 0; p < npan; ++p) {
            bool hit = false;
            if (lane < ncached) {
              const double2 key = W.keys[lane];
              hit = (key.x == qa[p]) && (key.y == qb[p]);
            }
            const unsigned m = __ballot_sync(FULL, hit);
            int slot;
            if (m) slot = __ffs(m) - 1;
            else {
              slot = NPC - 2 + p;
              if (ncached < NPC - 2) {
                slot = ncached++;
                if (lane == 0) W.keys[slot] = make_double2(qa[p], qb[p]);
              }
              pa[nmiss] = qa[p]; pb[nmiss] = qb[p]; ps[nmiss] = slot;
              ++nmiss;
            }
            sl[p] = slot;
            __syncwarp();
          }
          // ---- node part of the missing panels: k P, chi, a at their 41 abscissae each ----
          if (nmiss) {
            bool bad = false;
            for (int e = lane; e < PN * nmiss; e += 32) {
              const int p = e >= PN ? 1 : 0, j = e - p * PN;
              const double center = 0.5 * (pa[p] + pb[p]);
              const double half_length = 0.5 * (pb[p] - pa[p]);
              double lk;  // gsl_integration_qk abscissae
              if (j == 40) lk = center;
              else if (j < 20) lk = center - half_length * q_xgk[j];
              else lk = center + half_length * q_xgk[j - 20];
              double* row = pcache + (size_t)ps[p] * PROWS * PS + j;
              const double k = fast_exp(lk);
              const double chi = lp1h / k;
              // ccl_scale_factor_of_chi (src/ccl_background.c:1251-1277); chi > 0 here
              double a = 1.0;
              if (!(chi < 1.e-8)) {
                if (chi >= S.achi.x0 && chi <= S.achi.xn) {
                  const Rec1 r = find_rec(S.achi, chi);
                  a = cubic(make_double4(r.y, r.b, r.c, r.d), chi - r.x0);
                } else
                  bad = true;  // the general kernel reports it
              }
              if (!(a >= S.psp.amin && a <= S.psp.amax)) bad = true;  // growth extrapolation: general kernel
              double pk = 0.0;
              if (!bad) pk = fast_exp(bicubic_eval(S.psp, lk, a, 0));
              __stcg(row, k * pk);
              __stcg(row + PS, chi);
              __stcg(row + 2 * PS, a);
            }
            if (lane < nmiss) W.dmask[ps[lane]] = 0u;
            nodes_done += (unsigned long long)PN * nmiss;
            if (__any_sync(FULL, bad)) { redo = true; break; }
            __syncwarp();
          }
          // ---- transfers of this pair's collections where a panel does not hold them yet ----
          // transfer_limber_wrap (src/ccl_cls.c:150-164) over the sub-tracers of those collections only.  Outside a
          // kernel's support the lookup runs on the first / last interval and the term is dropped
          // (ccl_f1d_t_eval: the end nodes count as outside, src/ccl_f1d.c:137-161).
          unsigned need[2];
          need[0] = pmask & ~W.dmask[sl[0]];
          need[1] = npan > 1 ? (pmask & ~W.dmask[sl[1]]) : 0u;
          if (need[0] | need[1]) {
            for (int e = lane; e < PN * npan; e += 32) {
              const int p = e >= PN ? 1 : 0, j = e - p * PN;
              unsigned todo = need[p];
              if (!todo) continue;
              double* row = pcache + (size_t)sl[p] * PROWS * PS + j;
              const double chi = __ldcg(row + PS), a = __ldcg(row + 2 * PS);
              while (todo) {
                const int c = __ffs(todo) - 1;
                todo &= todo - 1;
                const int2 cr = W.cs[c];
                double D = 0.0;
                const FastEnt* ep = W.ents + cr.x;
#pragma unroll 1
                for (int s = 0; s < cr.y; ++s, ++ep) {
                  const FastEnt en = *ep;
                  const FastTr& tr = *reinterpret_cast<const FastTr*>(reinterpret_cast<const char*>(ep) + (int)en.addr);
                  double x0;
                  const int ki = find_interval(tr.kernel, chi, x0);
                  const bool inside = (chi > tr.kernel.x0) && (chi < tr.kernel.xn);
                  double wt = cubic(ldg4(en.kcr + ki), chi - x0);
                  if (en.acr != nullptr) {
                    const double a_ev = fmin(fmax(a, tr.fa.x0), tr.fa.xn);
                    const int ai = find_interval(tr.fa, a_ev, x0);
                    wt *= cubic(ldg4(en.acr + ai), a_ev - x0);
                  }
                  D += inside ? wt * en.pref : 0.0;
                }
                __stcg(row + (3 + c) * PS, D);
              }
            }
            if (lane < npan && need[lane]) W.dmask[sl[lane]] |= need[lane];
            __syncwarp();
          }
          // ---- GK41 sums of the requested panels ----
          double ar[2] = {0, 0}, er[2] = {0, 0}, rab[2] = {0, 0}, ras[2] = {0, 0};
          panel_sums2(pcache + (size_t)sl[0] * PROWS * PS, pcache + (size_t)sl[1] * PROWS * PS, npan, po.x, po.y, qa, qb, lane,
                      ar, er, rab, ras);
          nev += (unsigned long long)PN * npan;
          // ---- gsl_integration_qag bookkeeping (integration/qag.c) ----
          if (iteration == 0) {
            tolerance = fmax(epsabs, epsrel * fabs(ar[0]));
            const double round_off = 50 * DBL_EPSILON * rab[0];
            result = ar[0];
            if (er[0] <= round_off && er[0] > tolerance) { gsl = ST_GSL_EROUND; break; }
            if ((er[0] <= tolerance && er[0] != ras[0]) || er[0] == 0.0) { gsl = 0; break; }
            if (limit == 1) { gsl = ST_GSL_EMAXITER; break; }
            __syncwarp();
            if (lane == 0) { alist[0] = lkmin; blist[0] = lkmax; rlist[0] = ar[0]; elist[0] = er[0]; }
            __syncwarp();
            size = 1;
            area = ar[0];
            errsum = er[0];
            iteration = 1;
          } else {
            const double a1 = qa[0], b1 = qb[0], a2 = qa[1], b2 = qb[1];
            const double area12 = ar[0] + ar[1];
            const double error12 = er[0] + er[1];
            errsum += (error12 - e_i);
            area += area12 - r_i;
            if (ras[0] != er[0] && ras[1] != er[1]) {
              const double delta = r_i - area12;
              if (fabs(delta) <= 1.0e-5 * fabs(area12) && error12 >= 0.99 * e_i) roundoff_type1++;
              if (iteration >= 10 && error12 > e_i) roundoff_type2++;
            }
            tolerance = fmax(epsabs, epsrel * fabs(area));
            if (errsum > tolerance) {
              if (roundoff_type1 >= 6 || roundoff_type2 >= 20) error_type = 2;
              if (q_subinterval_too_small(a1, a2, b2)) error_type = 3;
            }
            // update(): larger-error half stays at imax, the other is appended
            __syncwarp();
            if (lane == 0) {
              if (er[1] > er[0]) {
                alist[imax] = a2; rlist[imax] = ar[1]; elist[imax] = er[1];
                alist[size] = a1; blist[size] = b1; rlist[size] = ar[0]; elist[size] = er[0];
              } else {
                blist[imax] = b1; rlist[imax] = ar[0]; elist[imax] = er[0];
                alist[size] = a2; blist[size] = b2; rlist[size] = ar[1]; elist[size] = er[1];
              }
            }
            __syncwarp();
            size++;
            iteration++;
            if (!(iteration < limit && !error_type && errsum > tolerance)) {
              double sum = 0.0;
              for (int i = lane; i < size; i += 32) sum += rlist[i];
              result = qwarp_sum(sum);
              if (errsum <= tolerance) gsl = 0;
              else if (error_type == 2) gsl = ST_GSL_EROUND;
              else if (error_type == 3) gsl = ST_GSL_ESING;
              else if (iteration == limit) gsl = ST_GSL_EMAXITER;
              else gsl = ST_GSL_EFAILED;
              break;
            }
          }
          // retrieve(): interval with the largest error estimate; its two halves are the next request
          // error estimates are non-negative doubles: their bit patterns order like unsigned integers, so the
          // maximum of a 32-entry block is two integer warp reductions (high words, then low words) and a ballot
          double emax = -1.0;
          imax = 0;
          for (int i0 = 0; i0 < size; i0 += 32) {
            const int i = i0 + lane;
            const double e = (i < size) ? elist[i] : -1.0;
            const bool live = (i < size) && (e >= 0.0);
            const unsigned hi = live ? (unsigned)__double2hiint(e) : 0u, lo = live ? (unsigned)__double2loint(e) : 0u;
            const unsigned mhi = __reduce_max_sync(FULL, hi);
            const unsigned mlo = __reduce_max_sync(FULL, (live && hi == mhi) ? lo : 0u);
            const unsigned who = __ballot_sync(FULL, live && hi == mhi && lo == mlo);
            if (who) {
              const double eb = __hiloint2double((int)mhi, (int)mlo);
              if (eb > emax) { emax = eb; imax = i0 + __ffs(who) - 1; }   // ties: the lowest index, as before
            }
          }
          const double a_i = alist[imax], b_i = blist[imax];
          r_i = rlist[imax];
          e_i = elist[imax];
          qa[0] = a_i; qb[0] = 0.5 * (a_i + b_i); qa[1] = qb[0]; qb[1] = b_i;
          npan = 2;
        }
        if (redo) break;
        if (P.counters && lane == 0) atomicAdd(P.counters, nev);
        const unsigned long long task_id = ((unsigned long long)ic * P.npair + ipair) * P.nl + il;
        // GSL_EROUND switches the reference to CQUAD (src/ccl_cls.c:245-259): handed to limber_cquad_kernel
        if (gsl == ST_GSL_EROUND && cq_list) {
          unsigned long long at = cq_cap;
          if (lane == 0) {
            at = atomicAdd(&cq_list[0], 1ull);
            if (at < cq_cap) cq_list[2 + at] = task_id;
          }
          at = __shfl_sync(FULL, at, 0);
          if (at < cq_cap) continue;
        }
        if (lane == 0) {
          if (gsl == 0) {
            P.cl[task_id] = result / lp1h;  // src/ccl_cls.c:340
          } else {
            P.cl[task_id] = nan("");         // src/ccl_cls.c:343-345
            const size_t sidx = (size_t)ic * P.npair + ipair;
            P.status[sidx] = ST_ERR_INTEG;
            atomicCAS(P.status_detail + sidx, 0, gsl);
          }
        }
      }
      if (redo) {
        // every pair of the task goes to the general kernel (pairs already finished are simply recomputed)
        if (lane < np) {
          const unsigned long long task_id =
              ((unsigned long long)ic * P.npair + gpairs[gd.pair_begin + lane].out) * P.nl + il;
          const unsigned long long at = atomicAdd(G.redo, 1ull);
          if (at < G.redo_cap) G.redo[1 + at] = task_id;
        }
      }
      if (P.counters && lane == 0 && nodes_done) atomicAdd(P.counters + 1, nodes_done);
    }
  }
}

// workspace carved after the general kernel's regions: [interval lists | panel caches | item counter]
static size_t qf_up(size_t b) { return (b + 255) / 256 * 256; }
static int qf_blocks() {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  return sms * QCTAS_PER_SM;
}
size_t limber_qag_fast_workspace_bytes(int n_iteration) {
  const size_t warps = (size_t)qf_blocks() * QFW;
  return qf_up(warps * 4 * (size_t)n_iteration * sizeof(double)) + qf_up(warps * NPC * PROWS * PS * sizeof(double)) + 256;
}

// `ws` holds limber_qag_workspace_bytes() (general kernel + CQUAD) followed by limber_qag_fast_workspace_bytes()
cudaError_t launch_limber_qag_fast(const LimberProblem& prob, const FastPlan& plan, double* ws, size_t ws_bytes,
                                   cudaStream_t stream) {
  if (prob.ncosmo == 0 || prob.npair == 0 || prob.nl == 0) return cudaSuccess;
  static bool uploaded[64] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  if (!uploaded[dev & 63]) {
    cudaMemcpyToSymbol(q_xgk, GK41_XGK, sizeof(double) * 21);
    cudaMemcpyToSymbol(q_wgk, GK41_WGK, sizeof(double) * 21);
    cudaMemcpyToSymbol(q_wg, GK41_WG, sizeof(double) * 10);
    uploaded[dev & 63] = true;
  }
  const int nit = prob.par.N_ITERATION;
  const size_t general = limber_qag_workspace_bytes(nit);
  if (general + limber_qag_fast_workspace_bytes(nit) > ws_bytes) return cudaErrorInvalidValue;
  const int nblk = qf_blocks();
  const size_t warps = (size_t)nblk * QFW;
  char* base = reinterpret_cast<char*>(ws) + general;
  double* lists = reinterpret_cast<double*>(base);
  double* cache = reinterpret_cast<double*>(base + qf_up(warps * 4 * (size_t)nit * sizeof(double)));
  unsigned long long* next_item = reinterpret_cast<unsigned long long*>(
      reinterpret_cast<char*>(cache) + qf_up(warps * NPC * PROWS * PS * sizeof(double)));
  launch_limber_group(prob, plan, stream);
  cudaMemsetAsync(next_item, 0, sizeof(unsigned long long), stream);
  // the general launcher zeroes the CQUAD list; the fast kernel appends to the same list first
  unsigned long long* cq_list = nullptr;
  unsigned long long cq_cap = 0;
  limber_qag_cq_list(ws, nit, &cq_list, &cq_cap, stream);
  const size_t smem = offsetof(QSmem, tr) + (size_t)plan.ntr_max * sizeof(FastTr);
  static size_t attr_set[64] = {0};
  if (attr_set[dev & 63] < smem) {
    cudaFuncSetAttribute(limber_qag_fast_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    attr_set[dev & 63] = smem;
  }
  limber_qag_fast_kernel<<<nblk, QFW * 32, smem, stream>>>(prob, plan, lists, cache, next_item, cq_list, cq_cap);
  // redo list -> general QAG kernel, then the CQUAD fallback over everything both kernels handed over
  launch_limber_qag_list(prob, ws, general, plan.redo, plan.redo_cap, stream);
  return cudaGetLastError();
}

}  // namespace cclb
