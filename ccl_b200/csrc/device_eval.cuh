// Device evaluators for the tables the Limber integrand reads.
//
// Each function restates one reference routine on flattened device tables; the reference
// call it replaces is cited above it.  All arithmetic is IEEE fp64.  Interval selection
// follows gsl_interp_bsearch exactly (x_i <= x < x_{i+1}, last interval closed), so a
// node that lands on a table node picks the same polynomial piece as the reference.
#pragma once
#include "tables.h"

namespace cclb {

// One 32-byte table node with a single 256-bit read-only load (LDG.E.256 on sm_100a); every double4
// table is carved on a 32-byte boundary (BatchPlan::derive, Arena).
__device__ __forceinline__ double4 ldg4(const double4* p) {
#ifdef CCLB_NO_LDG256
  const double2* q = reinterpret_cast<const double2*>(p);
  const double2 a = __ldg(q), b = __ldg(q + 1);
  return make_double4(a.x, a.y, b.x, b.y);
#else
  double4 r;
  asm("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(r.x), "=d"(r.y), "=d"(r.z), "=d"(r.w) : "l"(p));
  return r;
#endif
}

// Interval i of a 1-D spline: end points (LDG.128) and coefficients (LDG.256), issued together.
__device__ __forceinline__ Rec1 load_rec(const Spline1D& s, int i) {
  const double2 a = __ldg(s.xr + i);
  const double4 c = ldg4(s.cr + i);
  Rec1 r;
  r.x0 = a.x; r.x1 = a.y; r.y = c.x; r.b = c.y; r.c = c.z; r.d = c.w;
  return r;
}

// Record of the interval gsl_interp_bsearch(x, v, 0, n-1) would return, for x[0] <= v <= x[n-1].
__device__ __forceinline__ Rec1 find_rec(const Spline1D& s, double v) {
  int i;
  if (s.uniform) {
    i = (int)((v - s.x0) * s.inv_step);
  } else {
    int j = (int)((v - s.x0) * s.inv_bin);
    j = min(max(j, 0), s.nlut - 1);
    i = __ldg(s.lut + j);
  }
  const int last = s.n - 2;
  i = min(max(i, 0), last);
  Rec1 r = load_rec(s, i);
  // the LUT entry is the interval of a point just left of the bin (lut_kernel): never past the target
  if (s.uniform)
    while (v < r.x0 && i > 0) r = load_rec(s, --i);
  int steps = 0;
  while (v >= r.x1 && i < last) {
    if (++steps > 4) {  // strongly non-uniform grid: finish by bisection on the interval starts
      int lo = i, hi = last + 1;
      while (hi > lo + 1) {
        const int m = (hi + lo) >> 1;  // m <= last
        if (__ldg(&s.xr[m].x) > v) hi = m; else lo = m;
      }
      i = lo;
      r = load_rec(s, i);
      break;
    }
    r = load_rec(s, ++i);
  }
  return r;
}

__device__ __forceinline__ int find_axis(const Axis& s, double v, AxisRec& out) {
  int i;
  if (s.uniform) {
    i = (int)((v - s.x0) * s.inv_step);
  } else {
    int j = (int)((v - s.x0) * s.inv_bin);
    j = min(max(j, 0), s.nlut - 1);
    i = __ldg(s.lut + j);
  }
  const int last = s.n - 2;
  i = min(max(i, 0), last);
  const double4* p = reinterpret_cast<const double4*>(s.rec);
  double4 r = ldg4(p + i);
  while (v < r.x && i > 0) r = ldg4(p + (--i));
  while (v >= r.y && i < last) r = ldg4(p + (++i));
  out.x0 = r.x; out.x1 = r.y; out.inv_dx = r.z;
  return i;
}

// gsl_spline_eval_e on an akima/cspline object; EDOM outside the node range.
// which = 0 value, 1 first derivative, 2 second derivative (cspline_eval_deriv / _deriv2).
__device__ __forceinline__ double spline_eval(const Spline1D& s, double v, int which, int& st) {
  if (!(v >= s.x0 && v <= s.xn)) { st |= ST_GSL_EDOM; return nan(""); }
  const Rec1 r = find_rec(s, v);
  const double delx = v - r.x0;
  if (which == 0) return r.y + delx * (r.b + delx * (r.c + delx * r.d));
  if (which == 1) return r.b + delx * (2.0 * r.c + 3.0 * r.d * delx);
  return 2.0 * r.c + 6.0 * r.d * delx;
}

// ccl_scale_factor_of_chi (src/ccl_background.c:1251-1277)
__device__ __forceinline__ double scale_factor_of_chi(const Cosmo& cs, double chi, int& st) {
  if ((chi < 1.e-8) && (chi >= 0.)) return 1.;
  if (chi < 0.) { st = ST_ERR_COMPUTECHI; return nan(""); }
  if (!cs.has_achi) { st = ST_ERR_DISTANCES_INIT; return nan(""); }
  return spline_eval(cs.achi, chi, 0, st);
}

// ccl_growth_factor (src/ccl_background.c:1290-1325)
__device__ __forceinline__ double growth_factor(const Cosmo& cs, double a, int& st) {
  if (a == 1.) return 1.;
  if (a > 1.) { st = ST_ERR_COMPUTECHI; return nan(""); }
  if (!cs.has_growth) { st = ST_ERR_GROWTH_INIT; return nan(""); }
  int s2 = 0;
  const double D = spline_eval(cs.growth, a, 0, s2);
  if (s2) { st |= s2; return nan(""); }
  return D;
}

// gsl_spline2d_eval_e / _deriv_x_e / _deriv_xx_e / _deriv_y_e (which = 0 / 1 / 2 / 3) with gsl_interp2d_bicubic
// (interp2d/bicubic.c).  The bicubic Hermite patch is evaluated in tensor Hermite-basis
// form (same polynomial as GSL's 16-coefficient expansion, ~45 fp64 instructions).
__device__ __forceinline__ double bicubic_eval(const F2D& f, double x, double y, int which) {
  AxisRec ax, ay;
  const int xi = find_axis(f.gk, x, ax);
  const int yi = find_axis(f.ga, y, ay);
  const int nk = f.gk.n;
  const double4* r0 = f.tab + (size_t)yi * nk + xi;
  const double4* r1 = r0 + nk;
  const double4 c00 = ldg4(r0), c10 = ldg4(r0 + 1);  // (xmin,ymin) (xmax,ymin)
  const double4 c01 = ldg4(r1), c11 = ldg4(r1 + 1);  // (xmin,ymax) (xmax,ymax)
  const double dx = ax.x1 - ax.x0, dy = ay.x1 - ay.x0;
  const double t = (x - ax.x0) * ax.inv_dx, u = (y - ay.x0) * ay.inv_dx;
  const double um = 1.0 - u;
  const double u2 = u * u, um2 = um * um;
  double u00 = (1.0 + 2.0 * u) * um2;  // value weight, y-min edge
  double u01 = u2 * (3.0 - 2.0 * u);   // value weight, y-max edge
  double u10 = u * um2 * dy;           // slope weight, y-min edge
  double u11 = -u2 * um * dy;          // slope weight, y-max edge
  if (which == 3) {                    // d/dy (gsl bicubic_deriv_y): derivative of the y weights, x weights as for the value
    u00 = 6.0 * u * (u - 1.0) * ay.inv_dx;
    u01 = -u00;
    u10 = (3.0 * u * u - 4.0 * u + 1.0);
    u11 = (3.0 * u * u - 2.0 * u);
  }
  double t00, t01, t10, t11;
  const double tm = 1.0 - t;
  if (which == 0 || which == 3) {
    const double t2 = t * t, tm2 = tm * tm;
    t00 = (1.0 + 2.0 * t) * tm2;
    t01 = t2 * (3.0 - 2.0 * t);
    t10 = t * tm2 * dx;
    t11 = -t2 * tm * dx;
  } else if (which == 1) {
    t00 = 6.0 * t * (t - 1.0) * ax.inv_dx;
    t01 = -t00;
    t10 = (3.0 * t * t - 4.0 * t + 1.0);
    t11 = (3.0 * t * t - 2.0 * t);
  } else {
    t00 = (12.0 * t - 6.0) * ax.inv_dx * ax.inv_dx;
    t01 = -t00;
    t10 = (6.0 * t - 4.0) * ax.inv_dx;
    t11 = (6.0 * t - 2.0) * ax.inv_dx;
  }
  // y-direction Hermite interpolation of value (p) and x-slope (q) on both x edges
  const double p0 = u00 * c00.x + u01 * c01.x + u10 * c00.z + u11 * c01.z;
  const double p1 = u00 * c10.x + u01 * c11.x + u10 * c10.z + u11 * c11.z;
  const double q0 = u00 * c00.y + u01 * c01.y + u10 * c00.w + u11 * c01.w;
  const double q1 = u00 * c10.y + u01 * c11.y + u10 * c10.w + u11 * c11.w;
  return t00 * p0 + t01 * p1 + t10 * q0 + t11 * q1;
}

// ccl_f2d_t_eval (src/ccl_f2d.c:203-373), every branch.
static __device__ __noinline__ double f2d_eval(const F2D& f, double lk, double a, const Cosmo* cs, int& st) {
  bool is_hiz = false, is_loz = false;
  double a_ev = a;
  if (!f.is_a_constant) {
    is_hiz = a < f.amin;
    is_loz = a > f.amax;
    if (is_loz) {
      if (f.extrap_linear_growth == GROWTH_NONE) { st = ST_ERR_SPLINE_EV; return nan(""); }
      a_ev = f.amax;
    } else if (is_hiz) {
      if (f.extrap_linear_growth == GROWTH_NONE) { st = ST_ERR_SPLINE_EV; return nan(""); }
      a_ev = f.amin;
    }
  }
  bool is_hik = false, is_lok = false;
  double lk_ev = lk;
  if (!f.is_k_constant) {
    is_hik = lk > f.lkmax;
    is_lok = lk < f.lkmin;
    if (is_hik) lk_ev = f.lkmax;
    else if (is_lok) lk_ev = f.lkmin;
  }
  double pre = 0.0;
  int sp = 0;
  if (f.is_factorizable) {
    double fk = f.is_log ? 0. : 1., fa = f.is_log ? 0. : 1.;
    if (f.has_fk) fk = spline_eval(f.fk, lk_ev, 0, sp);
    if (f.has_fa) fa = spline_eval(f.fa, a_ev, 0, sp);
    pre = f.is_log ? fk + fa : fk * fa;
  } else {
    if (!f.has_fka) pre = f.is_log ? 0. : 1.;
    else {
      // gsl_interp2d_eval_e domain check (NaN inputs fail it)
      if (!(lk_ev >= f.gk.x0 && lk_ev <= f.gk.xn && a_ev >= f.ga.x0 && a_ev <= f.ga.xn)) sp = ST_GSL_EDOM;
      else pre = bicubic_eval(f, lk_ev, a_ev, 0);
    }
  }
  if (sp) { st = ST_ERR_SPLINE_EV; return nan(""); }
  double post = pre;
  if (is_hik || is_lok) {
    const int order = is_hik ? f.extrap_order_hik : f.extrap_order_lok;
    if (order > 0) {
      const double dlk = lk - lk_ev;
      double pd = f.is_factorizable ? spline_eval(f.fk, lk_ev, 1, sp) : bicubic_eval(f, lk_ev, a_ev, 1);
      if (sp) { st = ST_ERR_SPLINE_EV; return nan(""); }
      post += pd * dlk;
      if (order > 1) {
        pd = f.is_factorizable ? spline_eval(f.fk, lk_ev, 2, sp) : bicubic_eval(f, lk_ev, a_ev, 2);
        if (sp) { st = ST_ERR_SPLINE_EV; return nan(""); }
        post += pd * dlk * dlk * 0.5;
      }
    }
  }
  if (f.is_log) post = exp(post);
  if (is_hiz) {
    double gz;
    if (f.extrap_linear_growth == GROWTH_CCL) {
      if (cs == nullptr || !cs->has_growth) { st = ST_ERR_GROWTH_INIT; return nan(""); }
      gz = growth_factor(*cs, a, st) / growth_factor(*cs, a_ev, st);
    } else
      gz = f.growth_factor_0;
    post *= pow(gz, (double)f.growth_exponent);
  }
  return post;
}

// ccl_f2d_t_dlogf_dlk_eval (src/ccl_f2d.c:375-531): d ln f / d ln k
static __device__ __noinline__ double f2d_dlogf_dlk_eval(const F2D& f, double lk, double a, const Cosmo* cs, int& st) {
  if (f.is_k_constant) return 0.0;
  double inv_pk0 = 1.0;
  if (!f.is_log) {
    inv_pk0 = f2d_eval(f, lk, a, cs, st);
    if (inv_pk0 == 0) return 0.0;
    inv_pk0 = 1.0 / inv_pk0;
  }
  bool is_hiz = false, is_loz = false;
  double a_ev = a;
  if (!f.is_a_constant) {
    is_hiz = a < f.amin;
    is_loz = a > f.amax;
    if (is_loz || is_hiz) {
      if (f.extrap_linear_growth == GROWTH_NONE) { st = ST_ERR_SPLINE_EV; return nan(""); }
      a_ev = is_loz ? f.amax : f.amin;
    }
  }
  const bool is_hik = lk > f.lkmax, is_lok = lk < f.lkmin;
  const double lk_ev = is_hik ? f.lkmax : (is_lok ? f.lkmin : lk);
  double pre = 0.0;
  int sp = 0;
  if (f.is_factorizable) {
    double fk = f.is_log ? 0. : 1., fa = f.is_log ? 0. : 1.;
    if (f.has_fk) fk = spline_eval(f.fk, lk_ev, 1, sp);
    if (f.has_fa) fa = spline_eval(f.fa, a_ev, 0, sp);
    pre = f.is_log ? fk + fa : fk * fa;
  } else {
    if (!f.has_fka) pre = f.is_log ? 0. : 1.;
    else if (!(lk_ev >= f.gk.x0 && lk_ev <= f.gk.xn && a_ev >= f.ga.x0 && a_ev <= f.ga.xn)) sp = ST_GSL_EDOM;
    else pre = bicubic_eval(f, lk_ev, a_ev, 1);
  }
  if (sp) { st = ST_ERR_SPLINE_EV; return nan(""); }
  double post = pre;
  if ((is_hik && f.extrap_order_hik > 1) || (is_lok && f.extrap_order_lok > 1)) {
    const double pd = f.is_factorizable ? spline_eval(f.fk, lk_ev, 2, sp) : bicubic_eval(f, lk_ev, a_ev, 2);
    if (sp) { st = ST_ERR_SPLINE_EV; return nan(""); }
    post += pd * (lk - lk_ev);
  }
  if (!f.is_log) {
    post *= inv_pk0;
    if (is_hiz) {
      double gz;
      if (f.extrap_linear_growth == GROWTH_CCL) {
        if (cs == nullptr || !cs->has_growth) { st = ST_ERR_GROWTH_INIT; return nan(""); }
        gz = growth_factor(*cs, a, st) / growth_factor(*cs, a_ev, st);
      } else
        gz = f.growth_factor_0;
      post *= pow(gz, (double)f.growth_exponent);
    }
  }
  return post;
}

// P(k,a) with the in-range bicubic case inlined (what the Limber path hits on default tables:
// ln k never leaves the table, a < amin only for z > 99), everything else through f2d_eval.
__device__ __forceinline__ double psp_eval(const F2D& f, double lk, double a, const Cosmo* cs, int& st) {
  if (f.has_fka && lk >= f.lkmin && lk <= f.lkmax && a >= f.amin && a <= f.amax) {
    const double z = bicubic_eval(f, lk, a, 0);
    return f.is_log ? exp(z) : z;
  }
  return f2d_eval(f, lk, a, cs, st);
}

// ccl_cl_tracer_t_get_f_ell (src/ccl_tracers.c:703-726)
__device__ __forceinline__ double tracer_f_ell(int der_angles, double ell) {
  if (der_angles == 1) return ell * (ell + 1.);
  if (der_angles == 2) {
    if (ell <= 1) return 0;
    if (ell <= 10) return sqrt((ell + 2) * (ell + 1) * ell * (ell - 1));
    const double lp1h = ell + 0.5;
    const double lp1h2 = lp1h * lp1h;
    if (ell <= 1000) return lp1h2 * (1 - 1.25 / lp1h2);
    return lp1h2;
  }
  return 1;
}

// ccl_cl_tracer_t_get_kernel (src/ccl_tracers.c:728-737) + ccl_f1d_t_eval constant branch
__device__ __forceinline__ double tracer_kernel(const Tracer& tr, double chi) {
  if (!tr.has_kernel) return 1;
  if (chi <= tr.kernel.x0) return 0;  // y0 = 0, end node itself counts as outside
  if (chi >= tr.kernel.xn) return 0;  // yf = 0 (a NaN chi falls through to the spline: NaN)
  int st = 0;
  return spline_eval(tr.kernel, chi, 0, st);
}

// ccl_cl_tracer_t_get_transfer (src/ccl_tracers.c:739-749); the a-only factorised transfer
// (galaxy bias, IA amplitude, growth rate) is the common case and is inlined.
__device__ __forceinline__ double tracer_transfer(const Tracer& tr, double lk, double a, int& st) {
  if (!tr.has_transfer) return 1;
  // all samples of an a-only transfer equal (a constant galaxy bias): its natural cubic spline is that
  // constant exactly (every c_i, b_i, d_i is 0), inside and, through the clamp below, outside the table
  if (tr.const_transfer) return tr.const_transfer_value;
  const F2D& f = tr.transfer;
  if (f.is_factorizable && !f.has_fk && f.has_fa && !f.is_log) {
    // constant in k, clamped in a (growth_factor_0 = 1, exponent 0: src/ccl_tracers.c:670-684)
    const double a_ev = fmin(fmax(a, f.amin), f.amax);
    int sp = 0;
    const double v = spline_eval(f.fa, a_ev, 0, sp);
    if (sp) { st = ST_ERR_SPLINE_EV; return nan(""); }
    return v;
  }
  return f2d_eval(f, lk, a, nullptr, st);
}

struct TaskCtx {
  const Cosmo* cs;
  const Tracer* t1;
  const Tracer* t2;
  int n1, n2;
  bool same;  // both collections are the same tracer list
  double l;
};

constexpr int MAX_SUB = 8;  // sub-tracers per collection whose prefactors are hoisted

__device__ __forceinline__ double tracer_prefactor(const Tracer& tr, double l) {
  double p = tracer_f_ell(tr.der_angles, l);
  if (tr.der_bessel == -1) {
    const double lp1h = l + 0.5;
    p /= (lp1h * lp1h);
  }
  return p;
}

// transfer_limber_single + transfer_limber_wrap (src/ccl_cls.c:102-164).
// pref[itr] = f_ell (/ (l+1/2)^2 when der_bessel == -1), hoisted out of the node loop.
// der_bessel 1 / 2 terms of transfer_limber_single (src/ccl_cls.c:121-145): the second evaluation point
// chi' = (l + 3/2) / k.  Kept out of line: it is the larger half of the integrand's code and only
// redshift-space-distortion sub-tracers reach it (the adaptive integrators are instruction-fetch bound).
static __device__ __noinline__ double rsd_term(const TaskCtx& c, const Tracer& tr, double lk, double k, double a,
                                               double w, double t, int& st) {
  const double l = c.l;
  const double lp1h = l + 0.5;
  const double lp3h = l + 1.5;
  const double chi_lp = lp3h / k;
  const double a_lp = scale_factor_of_chi(*c.cs, chi_lp, st);
  const double pk_ratio =
      fabs(psp_eval(*c.cs->psp, lk, a_lp, c.cs, st) / psp_eval(*c.cs->psp, lk, a, c.cs, st));
  const double w_p = tracer_kernel(tr, chi_lp);
  const double t_p = tracer_transfer(tr, lk, a_lp, st);
  const double sqell = sqrt(lp1h * pk_ratio / lp3h);
  if (tr.der_bessel == 1) return l * w * t / lp1h - sqell * w_p * t_p;
  return sqell * 2 * w_p * t_p / lp3h - (0.25 + 2 * l) * w * t / (lp1h * lp1h);
}

__device__ __forceinline__ double transfer_limber_wrap(const TaskCtx& c, const Tracer* trs, int n,
                                                       const double* pref, double lk, double k, double chi,
                                                       double a, int& st) {
  double transfer = 0;
  const double l = c.l;
  for (int itr = 0; itr < n; ++itr) {
    const Tracer& tr = trs[itr];
    double dd;
    const double w = tracer_kernel(tr, chi);
    const double t = tracer_transfer(tr, lk, a, st);
    if (tr.der_bessel < 1) dd = w * t;
    else dd = rsd_term(c, tr, lk, k, a, w, t, st);
    transfer += dd * (itr < MAX_SUB ? pref[itr] : tracer_prefactor(tr, l));
    if (st != 0) return -1;
  }
  return transfer;
}

// cl_integrand (src/ccl_cls.c:166-187) with k = exp(lk) and chi = (l+1/2)/k supplied by the caller
__device__ __forceinline__ double cl_integrand_kchi(const TaskCtx& c, const double* pref1, const double* pref2,
                                                    double lk, double k, double chi, int& st) {
  const double a = scale_factor_of_chi(*c.cs, chi, st);
  const double d1 = transfer_limber_wrap(c, c.t1, c.n1, pref1, lk, k, chi, a, st);
  if (d1 == 0) return 0;
  const double d2 = c.same ? d1 : transfer_limber_wrap(c, c.t2, c.n2, pref2, lk, k, chi, a, st);
  if (d2 == 0) return 0;
  const double pk = psp_eval(*c.cs->psp, lk, a, c.cs, st);
  return k * pk * d1 * d2;
}

__device__ __forceinline__ double cl_integrand(const TaskCtx& c, const double* pref1, const double* pref2,
                                               double lk, int& st) {
  const double k = exp(lk);
  const double chi = (c.l + 0.5) / k;
  return cl_integrand_kchi(c, pref1, pref2, lk, k, chi, st);
}

// get_k_interval (src/ccl_cls.c:64-100)
__device__ __forceinline__ void get_k_interval(const TaskCtx& c, const LimberParams& par, double& lkmin,
                                               double& lkmax) {
  double chi_min1 = 1E15, chi_max1 = -1E15;
  for (int i = 0; i < c.n1; ++i) {
    chi_min1 = fmin(chi_min1, c.t1[i].chi_min);
    chi_max1 = fmax(chi_max1, c.t1[i].chi_max);
  }
  double chi_min2 = 1E15, chi_max2 = -1E15;
  for (int i = 0; i < c.n2; ++i) {
    chi_min2 = fmin(chi_min2, c.t2[i].chi_min);
    chi_max2 = fmax(chi_max2, c.t2[i].chi_max);
  }
  double chi_min = fmax(chi_min1, chi_min2);
  const double chi_max = fmin(chi_max1, chi_max2);
  if (chi_min <= 0) chi_min = 0.5 * (c.l + 0.5) / par.K_MAX;
  lkmax = log(fmin(par.K_MAX, 2 * (c.l + 0.5) / chi_min));
  lkmin = log(fmax(par.K_MIN, (c.l + 0.5) / chi_max));
}

}  // namespace cclb
