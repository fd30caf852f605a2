// Fast 'qag_quad' Limber kernel (sm_100a): gsl_integration_qag panels shared across the pairs of a node grid.
//
// Reference path: integ_cls_limber_qag_quad (src/ccl_cls.c:227-263) -> gsl_integration_qag(F, lkmin, lkmax, 0,
// epsrel, limit, GSL_INTEG_GAUSS41) with F = cl_integrand (:166-187), once per (pair, ell).  QAG is adaptive, so
// every C_ell bisects its own intervals; but its panels are always dyadic sub-intervals of [lkmin, lkmax], and
// [lkmin, lkmax] depends only on the pair's (chi_min, chi_max) (get_k_interval, :64-100).  Pairs with equal
// ranges (the grid tasks limber_group_kernel builds for the 'spline' integrator) therefore request the SAME
// panels again and again: the first panel and its two halves always, deeper ones whenever the integrands peak in
// the same place.  One warp runs the QAG state machines of a task's pairs one after the other over a per-warp
// PANEL CACHE: a panel's 41 Gauss-Kronrod abscissae are evaluated once -- k, chi, a(chi), k P(k,a) and the
// transfer D of EVERY collection of the task, with the lean lookups of the 'spline' kernel (sub-tracers on
// identical grids share one interval search) -- and every pair that asks for the panel again only forms
// k P D1 D2 and the four GK sums.  Both children of a bisection are evaluated together (82 abscissae over three
// lane passes).  The bookkeeping (interval list, error sums, round-off counters, stop tests) is GSL's, in the
// form of the general kernel (kernels.cu, limber_qag_kernel), which also serves as the redo path for whatever
// this kernel does not cover: nodes outside the a(chi) / P(k,a) tables, der_bessel 1/2 sub-tracers, oversize tasks.
#include <cfloat>
#include <cstddef>

#include "fast_common.cuh"
#include "gk41_table.h"

namespace cclb {


namespace {

constexpr int QFW = 8;            // warps per CTA
constexpr int NCG = CCLB_NCG;     // collections per grid task
constexpr int NPG = CCLB_NPG;     // pairs per grid task
constexpr int NSG = 2 * CCLB_NCG; // sub-tracers per grid task
constexpr int NPC = 24;           // cached panels per warp (the last slot doubles as scratch when the cache is full)
constexpr int PN = 41;            // abscissae per panel
constexpr int PS = 48;            // row stride of a cached panel (doubles)
constexpr int PROWS = NCG + 1;    // k P row + one D row per collection
constexpr int QGB = 64;           // grid tasks whose limits are precomputed per CTA pass
constexpr int QCTAS_PER_SM = 2;

__constant__ double q_xgk[21];
__constant__ double q_wgk[21];
__constant__ double q_wg[10];

struct QPar {
  double lkmin, lkmax;
  int grid, pad;
};

struct QCta {
  Spline1D achi;
  F2D psp;
  QPar gp[QGB];
  int next_task, item;
};

struct __align__(16) QWarp {
  FastEnt ents[NSG];  // the task's sub-tracers (doff = collection slot)
  double2 keys[NPC];  // (a, b) of the cached panels
  int2 po[NPG];       // collection slots of the task's pairs
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
    const double scale = pow((200 * err / result_asc), 1.5);
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

// gsl_integration_qk with the 41-point rule on a cached panel: lanes 0..19 hold the symmetric pairs of
// abscissae (e, e + 20), lane 20 the centre; the four sums are shuffle-reduced exactly like the general
// kernel's qk41_warp.  panel: [PROWS][PS] doubles, row 0 = k P, row 1 + c = D of collection slot c.
__device__ __forceinline__ void panel_sums(const double* panel, int ca, int cb, double a, double b, int lane,
                                           double& result, double& abserr, double& resabs, double& resasc) {
  const double* kp = panel;
  const double* da = panel + (1 + ca) * PS;
  const double* db = panel + (1 + cb) * PS;
  const double half_length = 0.5 * (b - a);
  const double abs_half_length = fabs(half_length);
  double rk = 0, rg = 0, ra = 0, f1 = 0, f2 = 0;
  if (lane < 20) {
    // cl_integrand (src/ccl_cls.c:166-187): k P d1 d2
    f1 = __ldcg(kp + lane) * __ldcg(da + lane) * __ldcg(db + lane);
    f2 = __ldcg(kp + lane + 20) * __ldcg(da + lane + 20) * __ldcg(db + lane + 20);
    const double fsum = f1 + f2;
    rk = q_wgk[lane] * fsum;
    ra = q_wgk[lane] * (fabs(f1) + fabs(f2));
    if (lane & 1) rg = q_wg[lane >> 1] * fsum;
  } else if (lane == 20) {
    f1 = __ldcg(kp + 40) * __ldcg(da + 40) * __ldcg(db + 40);
    rk = f1 * q_wgk[20];
    ra = fabs(rk);
  }
  rk = qwarp_sum(rk);
  rg = qwarp_sum(rg);
  ra = qwarp_sum(ra);
  const double mean = rk * 0.5;
  double rs = 0;
  if (lane < 20) rs = q_wgk[lane] * (fabs(f1 - mean) + fabs(f2 - mean));
  else if (lane == 20) rs = q_wgk[20] * fabs(f1 - mean);
  rs = qwarp_sum(rs);
  const double err = (rk - rg) * half_length;
  rk *= half_length;
  ra *= abs_half_length;
  rs *= abs_half_length;
  result = rk;
  resabs = ra;
  resasc = rs;
  abserr = q_rescale_error(err, ra, rs);
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
  const long long nitems = (long long)P.ncosmo * P.nl;

  for (;;) {
    __syncthreads();  // the previous item is drained
    if (threadIdx.x == 0) S.item = (int)min((unsigned long long)nitems, atomicAdd(next_item, 1ull));
    __syncthreads();
    const long long item = S.item;
    if (item >= nitems) break;
    const int il = (int)(item % P.nl);
    const int ic = (int)(item / P.nl);
    const Cosmo* cs = P.cosmos + ic;
    const double ell = P.ells[il];
    const double lp1h = ell + 0.5;

    // ---- per-item staging of the descriptors this (cosmology, ell) needs ----
    for (int t = threadIdx.x; t < cs->ntracers; t += blockDim.x) {
      const Tracer& tr = cs->tracers[t];
      s_tr[t].kernel = pack_s1(tr.kernel);
      s_tr[t].fa = pack_s1(tr.transfer.fa);
      const bool fold = tr.has_transfer && tr.const_transfer;  // see limber_fast.cu
      s_tr[t].has_fa = tr.has_transfer && !fold;
      s_tr[t].der_bessel = tr.der_bessel;
      s_tr[t].pref[0] = tracer_prefactor(tr, ell) * (fold ? tr.const_transfer_value : 1.0);
    }
    if (threadIdx.x == 32) S.achi = cs->achi;
    if (threadIdx.x == 64) S.psp = *cs->psp;
    const GridDesc* grids = G.grids + (size_t)ic * G.gstride;
    const GridPair* gpairs = G.gpairs + (size_t)ic * G.gstride;
    const unsigned* gsubs = G.gsubs + (size_t)ic * G.sstride;
    const int ngrids = G.ngrids[ic];

    for (int t0 = 0; t0 < ngrids; t0 += QGB) {
      __syncthreads();
      const int nb = min(QGB, ngrids - t0);
      if (threadIdx.x < nb) {  // get_k_interval (src/ccl_cls.c:91-99) once per grid
        QPar q;
        q.grid = t0 + threadIdx.x;
        const GridDesc gd = grids[q.grid];
        double chi_min = gd.chi_min;
        if (chi_min <= 0) chi_min = 0.5 * lp1h / P.par.K_MAX;
        q.lkmax = log(fmin(P.par.K_MAX, 2 * lp1h / chi_min));
        q.lkmin = log(fmax(P.par.K_MIN, lp1h / gd.chi_max));
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
        const int ncoll = gd.ncoll, np = gd.npair, nsub = gd.nsub;
        const double lkmin = S.gp[gl].lkmin, lkmax = S.gp[gl].lkmax;
        __syncwarp();
        if (lane < np) {
          const GridPair pr = gpairs[gd.pair_begin + lane];
          W.po[lane] = make_int2(pr.a, pr.b);
        }
        if (lane < min(nsub, NSG)) {
          const unsigned ent = gsubs[gd.sub_begin + lane];
          const FastTr& tr = s_tr[ent & 0xffu];
          FastEnt en;
          en.kcr = tr.kernel.cr;
          en.pref = tr.pref[0];
          en.acr = tr.has_fa ? tr.fa.cr : nullptr;
          en.flags = ent & 0xffffu;
          en.troff = (int)(reinterpret_cast<const char*>(&tr) - reinterpret_cast<const char*>(&W.ents[lane]));
          W.ents[lane] = en;
        }
        __syncwarp();
        // outside the P(k,a) table in ln k, der_bessel 1/2 sub-tracers, oversize list: general kernel
        bool redo = !(lkmin >= S.psp.lkmin && lkmax <= S.psp.lkmax) || nsub > NSG || gd.has_rsd;
        int ncached = 0;

        // Evaluate the abscissae of nmiss (1 or 2) panels into their cache slots.  Returns false when a node
        // leaves the tables the fast path reads (the task is then redone by the general kernel).
        auto eval_panels = [&](int nmiss, const double* pa, const double* pb, const int* slot) -> bool {
          bool bad = false;
          for (int e = lane; e < PN * nmiss; e += 32) {
            const int p = e >= PN ? 1 : 0, q = e - p * PN;
            const double center = 0.5 * (pa[p] + pb[p]);
            const double half_length = 0.5 * (pb[p] - pa[p]);
            double lk;  // gsl_integration_qk abscissae
            if (q == 40) lk = center;
            else if (q < 20) lk = center - half_length * q_xgk[q];
            else lk = center + half_length * q_xgk[q - 20];
            double* row = pcache + (size_t)slot[p] * PROWS * PS + q;
            const double k = exp(lk);
            const double chi = lp1h / k;
            // ccl_scale_factor_of_chi (src/ccl_background.c:1251-1277); chi > 0 here
            double a = 1.0;
            if (!(chi < 1.e-8)) {
              if (chi >= S.achi.x0 && chi <= S.achi.xn) {
                const Rec1 r = find_rec(S.achi, chi);
                a = cubic(make_double4(r.y, r.b, r.c, r.d), chi - r.x0);
              } else
                bad = true;
            }
            if (!(a >= S.psp.amin && a <= S.psp.amax)) bad = true;
            double pk = 0.0;
            if (!bad) pk = fast_exp(bicubic_eval(S.psp, lk, a, 0));
            __stcg(row, k * pk);
            int ki = 0, ai = 0;
            double kdel = 0.0, adel = 0.0;
            bool inside = false;
            const FastEnt* ep = W.ents;
#pragma unroll 1
            for (int s = 0; s < nsub; ++s, ++ep) {
              const FastEnt en = *ep;
              const unsigned ent = en.flags;
              if (__builtin_expect((ent & SUB_NEWK) != 0u, 0)) {
                const FastTr& tr = *reinterpret_cast<const FastTr*>(reinterpret_cast<const char*>(ep) + en.troff);
                double x0;
                ki = find_interval(tr.kernel, chi, x0);
                kdel = chi - x0;
                inside = (chi > tr.kernel.x0) && (chi < tr.kernel.xn);
              }
              double wt = cubic(ldg4(en.kcr + ki), kdel);
              if (en.acr != nullptr) {
                if (__builtin_expect((ent & SUB_NEWA) != 0u, 0)) {
                  const FastTr& tr = *reinterpret_cast<const FastTr*>(reinterpret_cast<const char*>(ep) + en.troff);
                  const double a_ev = fmin(fmax(a, tr.fa.x0), tr.fa.xn);
                  double x0;
                  ai = find_interval(tr.fa, a_ev, x0);
                  adel = a_ev - x0;
                }
                wt *= cubic(ldg4(en.acr + ai), adel);
              }
              double val = inside ? wt * en.pref : 0.0;
              double* dp = row + (1 + ((ent >> 8) & 0xfu)) * PS;
              if (!(ent & SUB_FIRST)) val += __ldcg(dp);
              __stcg(dp, val);
            }
          }
          bad = __any_sync(FULL, bad);
          __syncwarp();
          return !bad;
        };

        // Cache slot of panel (a, b): hit -> its slot; miss -> a fresh slot (or the scratch slot when full),
        // flagged in `miss`.
        auto find_panel = [&](double a, double b, bool& miss) -> int {
          bool hit = false;
          if (lane < ncached) {
            const double2 key = W.keys[lane];
            hit = (key.x == a) && (key.y == b);
          }
          const unsigned m = __ballot_sync(FULL, hit);
          if (m) { miss = false; return __ffs(m) - 1; }
          miss = true;
          int slot = NPC - 1;
          if (ncached < NPC - 1) {
            slot = ncached++;
            if (lane == 0) W.keys[slot] = make_double2(a, b);
          }
          __syncwarp();
          return slot;
        };

#pragma unroll 1
        for (int q = 0; q < np && !redo; ++q) {
          const int2 po = W.po[q];
          const int ipair = gpairs[gd.pair_begin + q].out;
          const double epsabs = 0.0, epsrel = P.par.LIMBER_EPSREL;
          int gsl = 0;
          unsigned long long nev = 41;
          double result = 0.0;
          if (epsabs <= 0 && (epsrel < 50 * DBL_EPSILON || epsrel < 0.5e-28)) {
            gsl = ST_GSL_EBADTOL;
          } else if (P.par.force_cquad) {
            gsl = ST_GSL_EROUND;  // tests of the fallback branch: skip QAG
            nev = 0;
          } else {
            bool miss;
            const int s0 = find_panel(lkmin, lkmax, miss);
            if (miss && !eval_panels(1, &lkmin, &lkmax, &s0)) { redo = true; break; }
            double r0, e0, ab0, as0;
            panel_sums(pcache + (size_t)s0 * PROWS * PS, po.x, po.y, lkmin, lkmax, lane, r0, e0, ab0, as0);
            double tolerance = fmax(epsabs, epsrel * fabs(r0));
            const double round_off = 50 * DBL_EPSILON * ab0;
            result = r0;
            if (e0 <= round_off && e0 > tolerance) gsl = ST_GSL_EROUND;
            else if ((e0 <= tolerance && e0 != as0) || e0 == 0.0) gsl = 0;
            else if (limit == 1) gsl = ST_GSL_EMAXITER;
            else {
              __syncwarp();
              if (lane == 0) { alist[0] = lkmin; blist[0] = lkmax; rlist[0] = r0; elist[0] = e0; }
              __syncwarp();
              int size = 1;
              double area = r0, errsum = e0;
              int iteration = 1, roundoff_type1 = 0, roundoff_type2 = 0, error_type = 0;
              do {
                // retrieve(): interval with the largest error estimate
                double emax = -1.0;
                int imax = 0;
                for (int i = lane; i < size; i += 32) {
                  const double e = elist[i];
                  if (e > emax) { emax = e; imax = i; }
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                  const double eo = __shfl_xor_sync(FULL, emax, o);
                  const int io = __shfl_xor_sync(FULL, imax, o);
                  if (eo > emax || (eo == emax && io < imax)) { emax = eo; imax = io; }
                }
                const double a_i = alist[imax], b_i = blist[imax], r_i = rlist[imax], e_i = elist[imax];
                const double a1 = a_i, b1 = 0.5 * (a_i + b_i), a2 = b1, b2 = b_i;
                // both children together: hits are read back, misses are evaluated in one pass
                bool m1, m2;
                const int sl1 = find_panel(a1, b1, m1);
                int sl2 = find_panel(a2, b2, m2);
                if (m1 && m2 && sl1 == sl2) {
                  // cache full: the scratch slot cannot hold both children; evaluate and reduce one after the other
                  const int one = sl1;
                  if (!eval_panels(1, &a1, &b1, &one)) { redo = true; break; }
                }
                double ar[2], er[2], rab[2], ras[2];
                if (m1 && m2 && sl1 == sl2) {
                  panel_sums(pcache + (size_t)sl1 * PROWS * PS, po.x, po.y, a1, b1, lane, ar[0], er[0], rab[0], ras[0]);
                  if (!eval_panels(1, &a2, &b2, &sl2)) { redo = true; break; }
                  panel_sums(pcache + (size_t)sl2 * PROWS * PS, po.x, po.y, a2, b2, lane, ar[1], er[1], rab[1], ras[1]);
                } else {
                  const double pa[2] = {m1 ? a1 : a2, a2}, pb[2] = {m1 ? b1 : b2, b2};
                  const int ps[2] = {m1 ? sl1 : sl2, sl2};
                  const int nmiss = (m1 ? 1 : 0) + (m2 ? 1 : 0);
                  if (nmiss && !eval_panels(nmiss, pa, pb, ps)) { redo = true; break; }
                  panel_sums(pcache + (size_t)sl1 * PROWS * PS, po.x, po.y, a1, b1, lane, ar[0], er[0], rab[0], ras[0]);
                  panel_sums(pcache + (size_t)sl2 * PROWS * PS, po.x, po.y, a2, b2, lane, ar[1], er[1], rab[1], ras[1]);
                }
                nev += 82;
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
              } while (iteration < limit && !error_type && errsum > tolerance);
              if (redo) break;
              double sum = 0.0;
              for (int i = lane; i < size; i += 32) sum += rlist[i];
              result = qwarp_sum(sum);
              if (errsum <= tolerance) gsl = 0;
              else if (error_type == 2) gsl = ST_GSL_EROUND;
              else if (error_type == 3) gsl = ST_GSL_ESING;
              else if (iteration == limit) gsl = ST_GSL_EMAXITER;
              else gsl = ST_GSL_EFAILED;
            }
          }
          if (P.counters && lane == 0) atomicAdd(P.counters, nev);
          const unsigned long long task = ((unsigned long long)ic * P.npair + ipair) * P.nl + il;
          // GSL_EROUND switches the reference to CQUAD (src/ccl_cls.c:245-259): handed to limber_cquad_kernel
          if (gsl == ST_GSL_EROUND && cq_list) {
            unsigned long long at = cq_cap;
            if (lane == 0) {
              at = atomicAdd(&cq_list[0], 1ull);
              if (at < cq_cap) cq_list[2 + at] = task;
            }
            at = __shfl_sync(FULL, at, 0);
            if (at < cq_cap) continue;
          }
          if (lane == 0) {
            const size_t o = (size_t)task;
            if (gsl == 0) {
              P.cl[o] = result / lp1h;  // src/ccl_cls.c:340
            } else {
              P.cl[o] = nan("");         // src/ccl_cls.c:343-345
              const size_t sidx = (size_t)ic * P.npair + ipair;
              P.status[sidx] = ST_ERR_INTEG;
              atomicCAS(P.status_detail + sidx, 0, gsl);
            }
          }
          (void)ncoll;
        }
        if (redo) {
          // every pair of the task goes to the general kernel (pairs already finished are simply recomputed)
          if (lane < np) {
            const unsigned long long task =
                ((unsigned long long)ic * P.npair + gpairs[gd.pair_begin + lane].out) * P.nl + il;
            const unsigned long long at = atomicAdd(G.redo, 1ull);
            if (at < G.redo_cap) G.redo[1 + at] = task;
          }
        }
        if (P.counters && lane == 0 && ncached) atomicAdd(P.counters + 1, (unsigned long long)ncached * PN);
      }
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
