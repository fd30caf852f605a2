// Device builders of the tables the Limber path reads (north-star item 5): background
// distances and growth, analytic linear P(k) with sigma8 normalisation, tracer radial kernels.
// They restate the reference's table builders (src/ccl_background.c, ccl_power.c, ccl_bbks.c,
// ccl_eh.c, ccl_tracers.c) for a whole batch of cosmologies at once, with the same algorithms as the
// host builders of this package (ccl_b200/background.py, power.py, tracers.py), against which they
// are tested: fixed high-order Gauss-Legendre panels where the reference uses adaptive CQUAD/QAG
// (tighter than its tolerances), the reference's own Newton chain and stopping rule for a(chi), and
// GSL-compatible Akima / natural-cubic-spline arithmetic wherever a table is read through a spline.
#include <cmath>
#include <cstdio>

#include "../../include/ccl_b200.h"
#include "gl_tables.h"
#include "tables.h"

namespace cclb {

namespace {

constexpr double CLIGHT_HMPC = 2997.92458;  // src/ccl_core.c:152

// parameter vector of one cosmology (derived like pyccl/cosmology.py:398-496, m_nu = 0)
enum : int { P_OC = 0, P_OB, P_OL, P_OK, P_OG, P_ONR, P_W0, P_WA, P_H, P_OM, P_NS, P_S8, P_TCMB, P_KSIGN, P_SQRTK, P_RES };
static_assert(P_RES + 1 == CCL_B200_NPAR, "parameter vector layout");

// h_over_h0 (src/ccl_background.c:18-48) without massive neutrinos
__device__ __forceinline__ double h_over_h0(const double* p, double a) {
  return sqrt((p[P_OC] + p[P_OB] + p[P_OL] * pow(a, -3 * (p[P_W0] + p[P_WA])) * exp(3 * p[P_WA] * (a - 1)) +
               p[P_OK] * a + (p[P_OG] + p[P_ONR]) / a) /
              (a * a * a));
}

// chi_integrand / h (src/ccl_background.c:141-147, 313-315)
__device__ __forceinline__ double chi_integrand(const double* p, double a) {
  return CLIGHT_HMPC / (a * a * h_over_h0(p, a)) / p[P_H];
}

// 12-point Gauss-Legendre integral of the chi integrand over [lo, hi]
__device__ __forceinline__ double chi_panel(const double* p, double lo, double hi) {
  const double mid = 0.5 * (hi + lo), half = 0.5 * (hi - lo);
  double s = 0.0;
#pragma unroll 1
  for (int q = 0; q < 12; ++q) s += GL12_W[q] * chi_integrand(p, mid + half * GL12_X[q]);
  return half * s;
}

}  // namespace

// E(a) and chi(a) on the a-grid: one CTA per cosmology (ccl_cosmology_compute_distances,
// src/ccl_background.c:417-496; chi(a) per node by quadrature, accumulated from a = 1).
__global__ void bg_distance_kernel(int na, const double* __restrict__ a_grid, const double* __restrict__ params,
                                   double* __restrict__ E, double* __restrict__ chi) {
  extern __shared__ double sm[];
  double* seg = sm;
  const int ic = blockIdx.x;
  const double* p = params + (size_t)ic * CCL_B200_NPAR;
  for (int i = threadIdx.x; i < na; i += blockDim.x) {
    E[(size_t)ic * na + i] = h_over_h0(p, a_grid[i]);
    if (i < na - 1) seg[i] = chi_panel(p, a_grid[i], a_grid[i + 1]);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double acc = 0.0;
    chi[(size_t)ic * na + na - 1] = 0.0;
    for (int i = na - 2; i >= 0; --i) {
      acc += seg[i];
      chi[(size_t)ic * na + i] = acc;
    }
  }
}

// a(chi) on a uniform chi grid of step <= dchi: Newton iterations stopped when |da| < epsrel |a| (a_of_chi,
// src/ccl_background.c:370-410, 498-571).  The reference walks the grid seeding every node with the previous root;
// here one CTA per cosmology cuts the grid into ACHI_HW runs, one per half-warp: the first node of a run is seeded
// by linear inverse interpolation of the chi(a) table, the others by the previous root as in the reference (Newton
// converges quadratically: the accepted root does not depend on the seed beyond its last bits).  A half-warp
// spreads the 12 quadrature nodes of every chi(a) evaluation over its lanes.
constexpr int ACHI_HW = 16;  // half-warps (runs of consecutive chi nodes) per cosmology

__global__ void __launch_bounds__(16 * ACHI_HW)
bg_achi_kernel(int ncosmo, int na, const double* __restrict__ a_grid, const double* __restrict__ params,
               const double* __restrict__ chi, double dchi, double epsrel, int nmax, int stride,
               int* __restrict__ n_out, double* __restrict__ achi_chi, double* __restrict__ achi_a) {
  const int sub = threadIdx.x & 15;                      // lane within the half-warp
  const int hw = threadIdx.x >> 4;                       // run index
  const int ic = blockIdx.x;
  const unsigned hmask = 0xffffu << (threadIdx.x & 16);  // the half-warp's lanes
  const double* p = params + (size_t)ic * CCL_B200_NPAR;
  const double* ch = chi + (size_t)ic * na;
  double* gx = achi_chi + (size_t)ic * stride;
  double* ga = achi_a + (size_t)ic * stride;
  const double chi0 = ch[na - 1], chif = ch[0];
  const double a0 = a_grid[na - 1], af = a_grid[0];
  int n = (int)((chif - chi0) / dchi);
  if (n > stride) n = -n;  // reported to the host: the caller's stride is too small
  if (threadIdx.x == 0) n_out[ic] = n;
  if (n < 5) return;
  const double dx = (chif - chi0) / (n - 1.);
  // 12-point panel with one node per lane, summed in node order by lane 0's shuffle chain so that the
  // value does not depend on the lane layout
  auto panel = [&](double lo, double hi) -> double {
    const double mid = 0.5 * (hi + lo), half = 0.5 * (hi - lo);
    const double term = (sub < 12) ? GL12_W[sub] * chi_integrand(p, mid + half * GL12_X[sub]) : 0.0;
    double s = 0.0;
#pragma unroll
    for (int q = 0; q < 12; ++q) s += __shfl_sync(hmask, term, q, 16);
    return half * s;
  };
  int j = na - 1;  // a_grid[j-1] <= cur < a_grid[j], hunted as chi grows
  auto chi_of = [&](double a) -> double {
    if (a >= 1.0) return (a > 1.0) ? -panel(1.0, a) : 0.0;
    while (j > 1 && a < a_grid[j - 1]) --j;
    while (j < na - 1 && a >= a_grid[j]) ++j;
    return ch[j] + panel(a, a_grid[j]);
  };
  if (threadIdx.x == 0) { gx[0] = chi0; ga[0] = a0; gx[n - 1] = chif; ga[n - 1] = af; }
  // interior nodes 1..n-2 in ACHI_HW runs
  const int per = (n - 2 + ACHI_HW - 1) / ACHI_HW;
  const int i_lo = 1 + hw * per, i_hi = min(i_lo + per, n - 1);
  if (i_lo >= i_hi) return;
  double cur = a0;
  if (hw > 0) {  // seed of the run: the chi(a) table inverted linearly around chi(i_lo - 1)
    const double t = chi0 + dx * (i_lo - 1);
    int lo = 0, hi = na - 1;  // ch decreases with the index: ch[lo] >= t >= ch[hi]
    while (hi > lo + 1) {
      const int m = (hi + lo) >> 1;
      if (ch[m] >= t) lo = m; else hi = m;
    }
    const double w = (ch[lo] - t) / (ch[lo] - ch[hi]);
    cur = a_grid[lo] + w * (a_grid[hi] - a_grid[lo]);
    j = hi;
  }
  for (int i = i_lo; i < i_hi; ++i) {
    const double target = chi0 + dx * i;
    double f = target - chi_of(cur);
    double df = chi_integrand(p, cur);
    int it = 0;
    for (;;) {
      ++it;
      const double prev = cur;
      cur = cur - f / df;
      f = target - chi_of(cur);
      df = chi_integrand(p, cur);
      if (fabs(cur - prev) < epsrel * fabs(cur) || cur == prev || it > nmax) break;
    }
    if (sub == 0) { gx[i] = target; ga[i] = cur; }
  }
}

// Linear growth: y0' = y1 / (a^3 E), y1' = 1.5 E a Omega_m(a) y0 from a_i with y0 = a_i, y1 = a_i^3 E(a_i)
// (src/ccl_background.c:153-165, 209-284), integrated in ln a with classical RK4 at a step far below
// the reference's RKCK tolerance (1e-6); D = y0 / y0(1), f = y1 / (a^2 E y0).
//
// One CTA per cosmology.  The system is linear, y' = A(ln a) y, and so is an RK4 step: the steps between two
// a-nodes (ns = ceil(span / hmax) steps of span / ns, as a serial integration would take them) are cut into
// segments of at most GSEG steps, every thread propagates the two basis vectors through its segments -- a 2x2
// transfer matrix each, three evaluations of E(a) per step -- and one thread chains the matrices node by node.
// The serial chain shrinks from ~7000 RK4 steps (2300 of them between a_init = 1e-6 and the first node) to GSEG
// steps plus ~650 2x2 products; the result is the serial one up to the rounding of the linear combinations.
constexpr int GSEG = 16;
constexpr int GROWTH_THREADS = 256;
// segments of the whole integration: at most one partial segment per node plus the full ones
inline int growth_seg_cap(int na, double a_init, double hmax) {
  return na + (int)(log(1.0 / a_init) / hmax / GSEG) + 8;
}

__global__ void __launch_bounds__(GROWTH_THREADS)
bg_growth_kernel(int ncosmo, int na, const double* __restrict__ a_grid, const double* __restrict__ params, double a_init,
                 double hmax, int seg_cap, double* __restrict__ growth, double* __restrict__ fgrowth,
                 double* __restrict__ growth0) {
  extern __shared__ double4 growth_sm[];
  double4* s_M = growth_sm;                                       // [seg_cap] transfer matrix {m00, m01, m10, m11}
  double2* s_y = reinterpret_cast<double2*>(s_M + seg_cap);       // [na] (y0, y1) at node i
  int* s_first = reinterpret_cast<int*>(s_y + na);                // [na + 1] first segment of node i
  const int ic = blockIdx.x;
  const double* p = params + (size_t)ic * CCL_B200_NPAR;
  const double lna_init = log(a_init);
  // interval of node i: from the previous node (or a_init) to a_grid[i]; ns steps of h = span / ns
  auto interval = [&](int i, double& lna0, double& h, int& ns) {
    const double a = a_grid[i];
    const double prev = (i > 0 && a_grid[i - 1] >= a_init) ? log(a_grid[i - 1]) : lna_init;
    const double span = log(a) - prev;
    lna0 = prev;
    ns = (a >= a_init && span > 0) ? (int)ceil(span / hmax) : 0;
    h = ns ? span / ns : 0.0;
  };
  for (int i = threadIdx.x; i < na; i += blockDim.x) {
    double lna0, h;
    int ns;
    interval(i, lna0, h, ns);
    s_first[i + 1] = (ns + GSEG - 1) / GSEG;
  }
  if (threadIdx.x == 0) s_first[0] = 0;
  __syncthreads();
  if (threadIdx.x == 0)
    for (int i = 0; i < na; ++i) s_first[i + 1] += s_first[i];
  __syncthreads();
  const int nseg = s_first[na];
  // coefficients of y' = (c0 y1, c1 y0) at ln a
  auto coef = [&](double lna, double& c0, double& c1) {
    const double a = exp(lna);
    const double hn = h_over_h0(p, a);
    const double om = (p[P_OC] + p[P_OB]) / (a * a * a) / hn / hn;
    c0 = a / (a * a * a * hn);
    c1 = a * 1.5 * hn * a * om;
  };
  for (int sg = threadIdx.x; sg < nseg && sg < seg_cap; sg += blockDim.x) {
    int lo = 0, hi = na;  // node i with s_first[i] <= sg < s_first[i + 1]
    while (hi > lo + 1) {
      const int m = (hi + lo) >> 1;
      if (s_first[m] > sg) hi = m; else lo = m;
    }
    double lna0, h;
    int ns;
    interval(lo, lna0, h, ns);
    const int s0 = (sg - s_first[lo]) * GSEG, s1 = min(s0 + GSEG, ns);
    double lna = lna0 + s0 * h;
    double u0 = 1.0, u1 = 0.0, v0 = 0.0, v1 = 1.0;  // images of the basis vectors (1, 0) and (0, 1)
    double ca0, ca1;
    coef(lna, ca0, ca1);
    for (int st = s0; st < s1; ++st) {
      double cm0, cm1, cb0, cb1;
      coef(lna + 0.5 * h, cm0, cm1);
      coef(lna + h, cb0, cb1);
      auto step = [&](double& y0, double& y1) {
        const double k10 = ca0 * y1, k11 = ca1 * y0;
        const double k20 = cm0 * (y1 + 0.5 * h * k11), k21 = cm1 * (y0 + 0.5 * h * k10);
        const double k30 = cm0 * (y1 + 0.5 * h * k21), k31 = cm1 * (y0 + 0.5 * h * k20);
        const double k40 = cb0 * (y1 + h * k31), k41 = cb1 * (y0 + h * k30);
        y0 += h / 6.0 * (k10 + 2 * k20 + 2 * k30 + k40);
        y1 += h / 6.0 * (k11 + 2 * k21 + 2 * k31 + k41);
      };
      step(u0, u1);
      step(v0, v1);
      lna += h;
      ca0 = cb0; ca1 = cb1;
    }
    s_M[sg] = make_double4(u0, v0, u1, v1);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double y0 = a_init, y1 = a_init * a_init * a_init * h_over_h0(p, a_init);
    int sg = 0;
    for (int i = 0; i < na; ++i) {
      const int end = min(s_first[i + 1], seg_cap);
      for (; sg < end; ++sg) {
        const double4 M = s_M[sg];
        const double n0 = M.x * y0 + M.y * y1, n1 = M.z * y0 + M.w * y1;
        y0 = n0; y1 = n1;
      }
      s_y[i] = make_double2(y0, y1);
    }
  }
  __syncthreads();
  double* D = growth + (size_t)ic * na;
  double* F = fgrowth + (size_t)ic * na;
  // the a-grid ends at a = 1 (A_SPLINE_MAX is fixed to 1): normalise D(1) = 1
  const double g0 = s_y[na - 1].x;
  if (threadIdx.x == 0) growth0[ic] = g0;
  for (int i = threadIdx.x; i < na; i += blockDim.x) {
    const double a = a_grid[i];
    if (a < a_init) {  // below the starting point the reference sets D = a, f = 1
      D[i] = a;
      F[i] = 1.0;
    } else {
      const double2 y = s_y[i];
      D[i] = y.x / g0;
      F[i] = y.y / (a * a * h_over_h0(p, a) * y.x);
    }
  }
}

// ------------------------------------------------------------------------------------------
// GSL-compatible spline arithmetic evaluated straight from the nodes (no coefficient table): used
// where a freshly built table is read a handful of times per cosmology.
// ------------------------------------------------------------------------------------------
// gsl_interp_bsearch: i with x[i] <= v < x[i+1] (last interval closed)
__device__ __forceinline__ int bsearch_nodes(const double* x, int n, double v) {
  int lo = 0, hi = n - 1;
  while (hi > lo + 1) {
    const int m = (hi + lo) >> 1;
    if (x[m] > v) hi = m; else lo = m;
  }
  return lo;
}

// slope m_j of the akima construction, j in [-2, n] (akima_init boundary rules)
__device__ __forceinline__ double akima_slope(const double* x, const double* y, int n, int j) {
  auto raw = [&](int q) { return (y[q + 1] - y[q]) / (x[q + 1] - x[q]); };
  if (j >= 0 && j <= n - 2) return raw(j);
  if (j == -1) return 2.0 * raw(0) - raw(1);
  if (j == -2) return 3.0 * raw(0) - 2.0 * raw(1);
  if (j == n - 1) return 2.0 * raw(n - 2) - raw(n - 3);
  return 3.0 * raw(n - 2) - 2.0 * raw(n - 3);
}

// coefficients (b, c, d) of interval i (akima_calc)
__device__ __forceinline__ void akima_interval(const double* x, const double* y, int n, int i, double& b,
                                               double& c, double& d) {
  const double mm2 = akima_slope(x, y, n, i - 2), mm1 = akima_slope(x, y, n, i - 1), m0 = akima_slope(x, y, n, i),
               mp1 = akima_slope(x, y, n, i + 1), mp2 = akima_slope(x, y, n, i + 2);
  const double NE = fabs(mp1 - m0) + fabs(mm1 - mm2);
  if (NE == 0.0) { b = m0; c = 0.0; d = 0.0; return; }
  const double h = x[i + 1] - x[i];
  const double NE_next = fabs(mp2 - mp1) + fabs(m0 - mm1);
  const double alpha = fabs(mm1 - mm2) / NE;
  double tL;
  if (NE_next == 0.0) tL = m0;
  else {
    const double alpha1 = fabs(m0 - mm1) / NE_next;
    tL = (1.0 - alpha1) * m0 + alpha1 * mp1;
  }
  b = (1.0 - alpha) * mm1 + alpha * m0;
  c = (3.0 * m0 - 2.0 * b - tL) / h;
  d = (b + tL - 2.0 * m0) / (h * h);
}

__device__ __forceinline__ double akima_eval_direct(const double* x, const double* y, int n, double v) {
  const int i = bsearch_nodes(x, n, v);
  double b, c, d;
  akima_interval(x, y, n, i, b, c, d);
  const double dx = v - x[i];
  return y[i] + dx * (b + dx * (c + dx * d));
}

// ------------------------------------------------------------------------------------------
// Linear matter power: ccl_compute_linpower_analytic (src/ccl_power.c:28-146) with the BBKS
// (src/ccl_bbks.c:10-26) or Eisenstein-Hu (src/ccl_eh.c:9-204) transfer function and the sigma8
// normalisation (ccl_sigmaR, src/ccl_power.c:390-454).  One CTA per cosmology.
// ------------------------------------------------------------------------------------------
namespace {

struct EHConst {
  double h, th2p7, keq, rsound, kSilk, alphac, betac, alphab, betab, bnode, rsound_approx, b_frac, OMh2, Om;
};

__device__ EHConst eh_constants(const double* p) {
  EHConst e;
  const double h = p[P_H];
  const double OMh2 = p[P_OM] * h * h, OBh2 = p[P_OB] * h * h;
  const double th2p7 = p[P_TCMB] / 2.7;
  const double zeq = 2.5E4 * OMh2 / pow(th2p7, 4);
  const double keq = 0.0746 * OMh2 / (h * th2p7 * th2p7);
  const double b1 = 0.313 * pow(OMh2, -0.419) * (1 + 0.607 * pow(OMh2, 0.674));
  const double b2 = 0.238 * pow(OMh2, 0.223);
  const double zdrag = 1291 * pow(OMh2, 0.251) * (1 + b1 * pow(OBh2, b2)) / (1 + 0.659 * pow(OMh2, 0.828));
  const double Req = 31.5 * OBh2 * 1000. / (zeq * pow(th2p7, 4));
  const double Rd = 31.5 * OBh2 * 1000. / ((1 + zdrag) * pow(th2p7, 4));
  e.rsound = 2 / (3 * keq) * sqrt(6 / Req) * log((sqrt(1 + Rd) + sqrt(Rd + Req)) / (1 + sqrt(Req)));
  e.kSilk = 1.6 * pow(OBh2, 0.52) * pow(OMh2, 0.73) * (1 + pow(10.4 * OMh2, -0.95)) / h;
  const double a1 = pow(46.9 * OMh2, 0.670) * (1 + pow(32.1 * OMh2, -0.532));
  const double a2 = pow(12.0 * OMh2, 0.424) * (1 + pow(45.0 * OMh2, -0.582));
  const double b_frac = OBh2 / OMh2;
  e.alphac = pow(a1, -b_frac) * pow(a2, -(b_frac * b_frac * b_frac));
  const double bb1 = 0.944 / (1 + pow(458 * OMh2, -0.708));
  const double bb2 = pow(0.395 * OMh2, -0.0266);
  e.betac = 1 / (1 + bb1 * (pow(1 - b_frac, bb2) - 1));
  const double y = zeq / (1 + zdrag);
  const double sqy = sqrt(1 + y);
  const double gy = y * (-6 * sqy + (2 + 3 * y) * log((sqy + 1) / (sqy - 1)));
  e.alphab = 2.07 * keq * e.rsound * pow(1 + Rd, -0.75) * gy;
  e.betab = 0.5 + b_frac + (3 - 2 * b_frac) * sqrt((17.2 * OMh2) * (17.2 * OMh2) + 1);
  e.bnode = 8.41 * pow(OMh2, 0.435);
  e.rsound_approx = h * 44.5 * log(9.83 / OMh2) / sqrt(1 + 10 * pow(OBh2, 0.75));
  e.h = h; e.th2p7 = th2p7; e.keq = keq; e.b_frac = b_frac; e.OMh2 = OMh2; e.Om = p[P_OM];
  return e;
}

__device__ __forceinline__ double eh_tk0(const EHConst& e, double kk, double a, double b) {
  const double q = kk / (13.41 * e.keq);
  const double c = 14.2 / a + 386. / (1 + 69.9 * pow(q, 1.08));
  const double l = log(M_E + 1.8 * b * q);
  return l / (l + c * q * q);
}

// k^ns T(k)^2, k in 1/Mpc
__device__ double analytic_pk(const double* p, const EHConst& e, int model, double k) {
  if (model == 0) {  // BBKS
    const double tfac = p[P_TCMB] / 2.7;
    const double q = tfac * tfac * k / (p[P_OM] * p[P_H] * p[P_H] * exp(-p[P_OB] * (1.0 + sqrt(2. * p[P_H]) / p[P_OM])));
    const double lg = log(1. + 2.34 * q) / (2.34 * q);
    const double t2 = lg * lg / sqrt(1. + 3.89 * q + (16.1 * q) * (16.1 * q) + (5.46 * q) * (5.46 * q) * (5.46 * q) +
                                     (6.71 * q) * (6.71 * q) * (6.71 * q) * (6.71 * q));
    return pow(k, p[P_NS]) * t2;
  }
  const double kh = k / e.h;
  double tk;
  if (model == 1) {
    const double f = 1 / (1 + pow(kh * e.rsound / 5.4, 4));
    const double tc = f * eh_tk0(e, kh, 1, e.betac) + (1 - f) * eh_tk0(e, kh, e.alphac, e.betac);
    const double x = kh * e.rsound;
    const double x_bessel = x * pow(1 + e.bnode * e.bnode * e.bnode / (x * x * x), -1. / 3.);
    const double part1 = eh_tk0(e, kh, 1, 1) / (1 + (x / 5.2) * (x / 5.2));
    const double part2 = e.alphab / (1 + pow(e.betab / x, 3)) * exp(-pow(kh / e.kSilk, 1.4));
    const double ax2 = x_bessel * x_bessel;
    const double jl = (ax2 < 1e-4) ? 1 - ax2 * (1 - ax2 / 20.) / 6. : sin(x_bessel) / x_bessel;
    const double tb = jl * (part1 + part2);
    tk = e.b_frac * tb + (1 - e.b_frac) * tc;
  } else {
    const double alpha_gamma = 1 - 0.328 * log(431 * e.OMh2) * e.b_frac + 0.38 * log(22.3 * e.OMh2) * e.b_frac * e.b_frac;
    const double gamma_eff = e.Om * e.h * (alpha_gamma + (1 - alpha_gamma) / (1 + pow(0.43 * kh * e.rsound_approx, 4)));
    const double q = kh * e.th2p7 * e.th2p7 / gamma_eff;
    const double l0 = log(2 * M_E + 1.8 * q);
    const double c0 = 14.2 + 731 / (1 + 62.5 * q);
    tk = l0 / (l0 + c0 * q * q);
  }
  return pow(k, p[P_NS]) * tk * tk;
}

// w_tophat (src/ccl_power.c:390-404)
__device__ __forceinline__ double w_tophat(double kR) {
  if (kR < 0.1) {
    const double kR2 = kR * kR;
    return 1. + kR2 * (-1.0 / 10.0 + kR2 * (1.0 / 280.0 + kR2 * (-1.0 / 15120.0 + kR2 * (1.0 / 1330560.0 +
                                                                                     kR2 * (-1.0 / 172972800.0)))));
  }
  return 3. * (sin(kR) - kR * cos(kR)) / (kR * kR * kR);
}

}  // namespace

__global__ void __launch_bounds__(256)
linpower_kernel(int nk, const double* __restrict__ k_grid, int na_pk, const double* __restrict__ a_pk,
                int na_bg, const double* __restrict__ a_bg, const double* __restrict__ growth,
                const double* __restrict__ params, int model, double lg_kmin, double lg_kmax, int npanel,
                double* __restrict__ logp, double* __restrict__ sigma8_unnorm) {
  extern __shared__ double sm[];
  double* lk = sm;          // [nk]
  double* row = lk + nk;    // log P at a = 1
  double* cc = row + nk;    // natural-spline c at the nodes
  double* tmp = cc + nk;    // forward-sweep scratch
  __shared__ double red[256];
  const int ic = blockIdx.x;
  const double* p = params + (size_t)ic * CCL_B200_NPAR;
  const EHConst e = eh_constants(p);
  for (int i = threadIdx.x; i < nk; i += blockDim.x) {
    lk[i] = log(k_grid[i]);
    row[i] = log(analytic_pk(p, e, model, k_grid[i]));
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    // natural cubic spline of the a = 1 row (gsl cspline_init): tridiagonal solve for c_1..c_{n-2}
    const int N = nk - 2;
    cc[0] = 0.0;
    cc[nk - 1] = 0.0;
    auto hh = [&](int i) { return lk[i + 1] - lk[i]; };
    auto g = [&](int i) { return 3.0 * ((row[i + 2] - row[i + 1]) / hh(i + 1) - (row[i + 1] - row[i]) / hh(i)); };
    // Thomas algorithm: tmp = modified upper diagonal, cc[1..] = modified rhs
    double beta = 2.0 * (hh(0) + hh(1));
    tmp[0] = hh(1) / beta;
    cc[1] = g(0) / beta;
    for (int i = 1; i < N; ++i) {
      beta = 2.0 * (hh(i) + hh(i + 1)) - hh(i) * tmp[i - 1];
      tmp[i] = hh(i + 1) / beta;
      cc[i + 1] = (g(i) - hh(i) * cc[i]) / beta;
    }
    for (int i = N - 2; i >= 0; --i) cc[i + 1] -= tmp[i] * cc[i + 2];
  }
  __syncthreads();
  // sigma_R at R = 8/h: composite 16-point Gauss-Legendre in log10 k over the spline of the table row
  const double R = 8.0 / p[P_H];
  const double step = (lg_kmax - lg_kmin) / npanel;
  const double inv_dlk = (nk - 1) / (lk[nk - 1] - lk[0]);
  double acc = 0.0;
  for (int t = threadIdx.x; t < npanel * 16; t += blockDim.x) {
    const int j = t >> 4, q = t & 15;
    const double e0 = lg_kmin + j * step, e1 = (j == npanel - 1) ? lg_kmax : lg_kmin + (j + 1) * step;
    const double mid = 0.5 * (e1 + e0), half = 0.5 * (e1 - e0);
    const double x = mid + half * GL16_X[q];
    const double k = pow(10., x);
    const double lkq = fmin(fmax(x * log(10.), lk[0]), lk[nk - 1]);
    int i = (int)((lkq - lk[0]) * inv_dlk);
    i = min(max(i, 0), nk - 2);
    while (i > 0 && lkq < lk[i]) --i;
    while (i < nk - 2 && lkq >= lk[i + 1]) ++i;
    const double h = lk[i + 1] - lk[i], dy = row[i + 1] - row[i];
    const double b = dy / h - h * (cc[i + 1] + 2.0 * cc[i]) / 3.0;
    const double d = (cc[i + 1] - cc[i]) / (3.0 * h);
    const double dx = lkq - lk[i];
    const double pk = exp(row[i] + dx * (b + dx * (cc[i] + dx * d)));
    const double w = w_tophat(k * R);
    acc += half * GL16_W[q] * pk * k * k * k * w * w;
  }
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  const double sigma8 = sqrt(red[0] * log(10.) / (2 * M_PI * M_PI));
  if (threadIdx.x == 0) sigma8_unnorm[ic] = sigma8;
  const double shift = 2 * (log(p[P_S8]) - log(sigma8));
  // log P(k, a) = log(k^ns T^2) + 2 log D(a) + 2 log(sigma8_in / sigma8) (src/ccl_power.c:99-130)
  const double* D = growth + (size_t)ic * na_bg;
  for (int ja = 0; ja < na_pk; ++ja) {
    const double a = a_pk[ja];
    const double Da = (a == 1.0) ? 1.0 : akima_eval_direct(a_bg, D, na_bg, a);
    const double g2 = 2. * log(Da);
    for (int i = threadIdx.x; i < nk; i += blockDim.x)
      logp[((size_t)ic * na_pk + ja) * nk + i] = row[i] + g2 + shift;
  }
}

void launch_build_linpower(int ncosmo, int nk, const double* d_k, int na_pk, const double* d_apk, int na_bg,
                           const double* d_abg, const double* d_growth, const double* d_params, int model,
                           double lg_kmin, double lg_kmax, double* d_logp, double* d_sigma8, cudaStream_t stream) {
  if (ncosmo <= 0) return;
  const size_t smem = sizeof(double) * 4 * (size_t)nk;
  if (smem > 48 * 1024)
    cudaFuncSetAttribute(linpower_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  linpower_kernel<<<ncosmo, 256, smem, stream>>>(nk, d_k, na_pk, d_apk, na_bg, d_abg, d_growth, d_params, model,
                                                 lg_kmin, lg_kmax, 600, d_logp, d_sigma8);
}

// ------------------------------------------------------------------------------------------
// Tracer radial kernels from n(z): number counts p(z) H(z) (ccl_get_number_counts_kernel,
// src/ccl_tracers.c:121-160) and the lensing / magnification kernel by Akima-spline integration
// with the reference's trapezoid edge term (integrate_lensing_kernel_spline, :347-454; node
// placement ccl_get_chis_lensing_kernel, :470-479).  One CTA per (cosmology, n(z)); one thread per
// lensing-kernel node walks the n(z) grid with a sliding window of Akima slopes.
// ------------------------------------------------------------------------------------------
namespace {

__device__ __forceinline__ double sinn(const double* p, double chi) {  // ccl_sinn, src/ccl_background.c:107-131
  const double ks = p[P_KSIGN], sk = p[P_SQRTK];
  if (ks == -1.0) return sinh(sk * chi) / sk;
  if (ks == 1.0) return sin(sk * chi) / sk;
  return chi;
}

}  // namespace

__global__ void __launch_bounds__(256)
tracer_kernel_builder(int na, const double* __restrict__ a_bg, const double* __restrict__ chi_bg,
                      const int* __restrict__ n_achi, int stride, const double* __restrict__ achi_chi,
                      const double* __restrict__ achi_a, const double* __restrict__ params, int nz,
                      const double* __restrict__ z, int ntr, const double* __restrict__ nz_arr,
                      const double* __restrict__ sz_arr, const double* __restrict__ inv_norm,
                      const double* __restrict__ lens_scale, int nchi,
                      double* __restrict__ chi_z_out, double* __restrict__ w_nc, double* __restrict__ chi_l_out,
                      double* __restrict__ w_l) {
  extern __shared__ double sm[];
  double* chz = sm;        // chi at the n(z) nodes
  double* nq = chz + nz;   // n(z) (1 - 2.5 s(z))
  double* snz = nq + nz;   // 1 / sinn(chi(z)) (0 where chi = 0: that node takes the n q branch)
  double* idz = snz + nz;  // 1 / (z[j+1] - z[j]): the reciprocals replace four divisions per step of the walk below
  const int ic = blockIdx.x, it = blockIdx.y;
  const double* p = params + (size_t)ic * CCL_B200_NPAR;
  const double* chi_a = chi_bg + (size_t)ic * na;
  const double* nzv = nz_arr + (size_t)it * nz;
  const double inorm = inv_norm[it];
  for (int j = threadIdx.x; j < nz; j += blockDim.x) {
    const double a = 1. / (1 + z[j]);
    // ccl_comoving_radial_distance: Akima spline of chi(a), 0 at a = 1
    const double c = (a == 1.0) ? 0.0 : akima_eval_direct(a_bg, chi_a, na, a);
    chz[j] = c;
    snz[j] = (c == 0.0) ? 0.0 : 1.0 / sinn(p, c);
    idz[j] = (j < nz - 1) ? 1.0 / (z[j + 1] - z[j]) : 0.0;
    nq[j] = nzv[j] * (sz_arr ? 1 - 2.5 * sz_arr[(size_t)it * nz + j] : 1.0);
    if (it == 0) chi_z_out[(size_t)ic * nz + j] = c;
    // H(z) p(z): h E(a) / c * n(z) / norm
    w_nc[((size_t)ic * ntr + it) * nz + j] = p[P_H] * h_over_h0(p, a) / CLIGHT_HMPC * nzv[j] * inorm;
  }
  __syncthreads();
  const int ik = threadIdx.x;
  if (ik >= nchi) return;
  const double z_k = (z[nz - 1] / nchi) * ik + 1E-15;
  const double ak0 = 1. / (1 + z_k);
  const double chi_end = (ak0 == 1.0) ? 0.0 : akima_eval_direct(a_bg, chi_a, na, ak0);
  if (lens_scale && lens_scale[it] == 0.0) {  // this distribution's lensing kernel is not used (pure number counts)
    if (it == 0) chi_l_out[(size_t)ic * nchi + ik] = chi_end;
    w_l[((size_t)ic * ntr + it) * nchi + ik] = 0.0;
    return;
  }
  // ccl_scale_factor_of_chi on the freshly built a(chi) table
  const double a_k = (chi_end < 1e-8) ? 1.0
                                      : akima_eval_direct(achi_chi + (size_t)ic * stride, achi_a + (size_t)ic * stride,
                                                          n_achi[ic], chi_end);
  const double z_end = 1. / a_k - 1;
  // integrand on the n(z) grid: 0 below chi_end, n q sinn(chi - chi_end) / sinn(chi) above
  auto yv = [&](int j) -> double {
    const double c = chz[j];
    if (c < chi_end) return 0.0;
    if (c == 0.0) return nq[j];
    return nq[j] * sinn(p, c - chi_end) * snz[j];
  };
  int i_end = 0;  // chi(z) is increasing: number of nodes below chi_end
  {
    int lo = 0, hi = nz;
    while (lo < hi) {
      const int m = (lo + hi) >> 1;
      if (chz[m] < chi_end) lo = m + 1; else hi = m;
    }
    i_end = lo;
  }
  auto raw = [&](int q) { return (yv(q + 1) - yv(q)) * idz[q]; };
  auto slope = [&](int j) -> double {
    if (j >= 0 && j <= nz - 2) return raw(j);
    if (j == -1) return 2.0 * raw(0) - raw(1);
    if (j == -2) return 3.0 * raw(0) - 2.0 * raw(1);
    if (j == nz - 1) return 2.0 * raw(nz - 2) - raw(nz - 3);
    return 3.0 * raw(nz - 2) - 2.0 * raw(nz - 3);
  };
  double res = 0.0;
  if (i_end <= nz - 2) {
    // gsl_spline_eval_integ from the node z[i_end] to z[nz-1]: whole intervals of the Akima spline
    double mm2 = slope(i_end - 2), mm1 = slope(i_end - 1), m0 = slope(i_end), mp1 = slope(i_end + 1),
           mp2 = slope(i_end + 2);
    // integrand values y_i .. y_{i+3}: one new evaluation per interval as the window slides
    double q0 = yv(i_end), q1 = yv(min(i_end + 1, nz - 1)), q2 = yv(min(i_end + 2, nz - 1)),
           q3 = yv(min(i_end + 3, nz - 1));
    for (int i = i_end; i <= nz - 2; ++i) {
      const double h = z[i + 1] - z[i], ih = idz[i];
      const double NE = fabs(mp1 - m0) + fabs(mm1 - mm2);
      double b, c, d;
      if (NE == 0.0) { b = m0; c = 0.0; d = 0.0; }
      else {
        const double NE_next = fabs(mp2 - mp1) + fabs(m0 - mm1);
        const double alpha = fabs(mm1 - mm2) / NE;
        double tL;
        if (NE_next == 0.0) tL = m0;
        else {
          const double alpha1 = fabs(m0 - mm1) / NE_next;
          tL = (1.0 - alpha1) * m0 + alpha1 * mp1;
        }
        b = (1.0 - alpha) * mm1 + alpha * m0;
        c = (3.0 * m0 - 2.0 * b - tL) * ih;
        d = (b + tL - 2.0 * m0) * (ih * ih);
      }
      // (x2 - x1) (y + b r12/2 + c (r1^2 + r2^2 + r1 r2)/3 + d r12 (r1^2 + r2^2)/4) with r1 = 0, r2 = h
      res += h * (q0 + 0.5 * b * h + (1. / 3.) * c * (h * h) + 0.25 * d * h * (h * h));
      // slide: slopes m_{i-1} .. m_{i+3}
      mm2 = mm1; mm1 = m0; m0 = mp1; mp1 = mp2;
      const int jn = i + 3;  // index of the new slope
      if (jn <= nz - 2) {
        const double ynew = yv(jn + 1);
        mp2 = (ynew - q3) * idz[jn];
        q0 = q1; q1 = q2; q2 = q3; q3 = ynew;
      } else {
        if (jn == nz - 1) mp2 = 2.0 * mp1 - m0;        // 2 m_{n-2} - m_{n-3}
        else if (jn == nz) mp2 = 3.0 * m0 - 2.0 * mm1;  // 3 m_{n-2} - 2 m_{n-3}
        q0 = q1; q1 = q2; q2 = q3;
      }
    }
  }
  const double y_end = (i_end < nz) ? yv(i_end) : 0.0;
  if (z_end > z[0] && i_end < nz) res += 0.5 * (z[i_end] - z_end) * y_end;
  const double hub = p[P_H] / CLIGHT_HMPC;
  const double pref = 1.5 * hub * hub * p[P_OM];  // get_lensing_prefactor, src/ccl_tracers.c:163-166
  if (it == 0) chi_l_out[(size_t)ic * nchi + ik] = chi_end;
  w_l[((size_t)ic * ntr + it) * nchi + ik] = res * pref * inorm * chi_end / a_k;
}

void launch_build_tracer_kernels(int ncosmo, int na, const double* d_abg, const double* d_chibg, const int* d_nachi,
                                 int stride, const double* d_achi_chi, const double* d_achi_a, const double* d_params,
                                 int nz, const double* d_z, int ntr, const double* d_nz, const double* d_sz,
                                 const double* d_inv_norm, const double* d_lens_scale, int nchi, double* d_chi_z,
                                 double* d_w_nc, double* d_chi_l, double* d_w_l, cudaStream_t stream) {
  if (ncosmo <= 0 || ntr <= 0) return;
  const size_t smem = sizeof(double) * 4 * (size_t)nz;
  if (smem > 48 * 1024)
    cudaFuncSetAttribute(tracer_kernel_builder, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  dim3 grid(ncosmo, ntr);
  tracer_kernel_builder<<<grid, 256, smem, stream>>>(na, d_abg, d_chibg, d_nachi, stride, d_achi_chi, d_achi_a,
                                                     d_params, nz, d_z, ntr, d_nz, d_sz, d_inv_norm, d_lens_scale, nchi,
                                                     d_chi_z, d_w_nc, d_chi_l, d_w_l);
}

// ------------------------------------------------------------------------------------------
// Halofit (Takahashi et al. 2012 with the Bird et al. 2012 neutrino terms at f_nu = 0):
// ccl_apply_halofit (src/ccl_power.c:171-232) = per scale factor the non-linear scale sigma(R) = 1
// (get_rsigma, src/ccl_halofit.c:259-315), n_eff and C from the first two derivatives (:322-554),
// then ccl_halofit_power (:828-930) on every k.  One CTA per (cosmology, a-node): the row's natural
// cubic spline is built once, the Gaussian-filtered moments are 16-point Gauss-Legendre sums over
// ln k spread over the CTA, the root is found by bracketed Newton iterations on ln R.
// ------------------------------------------------------------------------------------------
constexpr int HF_GROUP = 10;
// Quadrature of the Gaussian-filtered moments: HF_PANELS 16-point Gauss-Legendre panels per 25 e-folds of k.  The
// integrand is a natural cubic SPLINE in ln k (1220 knots by default), so the rule converges algebraically: 400 / 200
// / 100 / 64 panels reproduce R_sigma to 1e-13 / 3e-10 / 1e-8 / 2e-8 and ln P_NL to - / 2e-9 / 5e-8 / 8e-8 (host twin,
// power.py, which uses the same constants).  The reference integrates with CQUAD at INTEGRATION_SIGMAR_EPSREL = 1e-5
// and stops its root search at the same 1e-5 (src/ccl_halofit.c:239-241,296-301); 100 panels and a root tolerance of
// 1e-10 stay three orders of magnitude inside that at a quarter of the cost of the 400-panel rule.
constexpr int HF_PANELS = 100;
constexpr double HF_ROOT_TOL = 1e-10;

__global__ void __launch_bounds__(256, 3)
halofit_kernel(int nk, const double* __restrict__ lk_grid, int na, const double* __restrict__ a_grid,
               const double* __restrict__ logplin, const double* __restrict__ params, double* __restrict__ logpnl,
               double* __restrict__ rsigma_out, int* __restrict__ fail) {
  extern __shared__ double sm[];
  double* lk = sm;
  double* row = lk + nk;
  double* cc = row + nk;
  double* tmp = cc + nk;   // right-hand side of the spline system, then the b coefficients
  double* dco = tmp + nk;  // d coefficients
  double* alpha = dco + nk;  // LDL^T factors of the natural-spline system of the ln k grid: the same
  double* gamma = alpha + nk;  // for every scale factor of the group, computed once per CTA
  __shared__ double red[3][8];   // warp partials of the three moments (blockDim = 256)
  __shared__ double glx[16], glw[16];  // lane-varying node index: shared memory, not the constant cache
  if (threadIdx.x < 16) { glx[threadIdx.x] = GL16_X[threadIdx.x]; glw[threadIdx.x] = GL16_W[threadIdx.x]; }
  const int ic = blockIdx.x;
  const double* p = params + (size_t)ic * CCL_B200_NPAR;
  for (int i = threadIdx.x; i < nk; i += blockDim.x) lk[i] = lk_grid[i];
  __syncthreads();
  if (threadIdx.x == 0) {  // gsl solve_tridiag (linalg/tridiag.c): alpha_i, gamma_i of diag 2(h_i + h_{i+1}), offdiag h_{i+1}
    const int N = nk - 2;
    auto hh = [&](int i) { return lk[i + 1] - lk[i]; };
    alpha[0] = 2.0 * (hh(0) + hh(1));
    gamma[0] = hh(1) / alpha[0];
    for (int i = 1; i < N; ++i) {
      alpha[i] = 2.0 * (hh(i) + hh(i + 1)) - hh(i) * gamma[i - 1];
      gamma[i] = hh(i + 1) / alpha[i];
    }
  }
  // a CTA walks HF_GROUP consecutive scale factors; each root search starts from the linear
  // extrapolation of the previous two roots (ln R_sigma is smooth in a)
  double x = log(1.0), x_prev = 0.0, a_prev = 0.0, a_prev2 = 0.0, x_prev2 = 0.0;
  int nroots = 0;
  for (int ia = blockIdx.y * HF_GROUP; ia < min(na, (int)(blockIdx.y + 1) * HF_GROUP); ++ia) {
  const double* src = logplin + ((size_t)ic * na + ia) * nk;
  if (nroots >= 2) x = x_prev + (x_prev - x_prev2) * (a_grid[ia] - a_prev) / (a_prev - a_prev2);
  __syncthreads();
  for (int i = threadIdx.x; i < nk; i += blockDim.x) row[i] = src[i];
  __syncthreads();
  {
    // natural cubic spline of the row (gsl cspline_init): everything without a serial dependence runs
    // on all threads; thread 0 only carries the two first-order recurrences (one FMA per node each)
    const int N = nk - 2;
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
      const double h0 = lk[i + 1] - lk[i], h1 = lk[i + 2] - lk[i + 1];
      tmp[i] = 3.0 * ((row[i + 2] - row[i + 1]) / h1 - (row[i + 1] - row[i]) / h0);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      cc[0] = 0.0;
      cc[nk - 1] = 0.0;
      double z = tmp[0];
      cc[1] = z;
      for (int i = 1; i < N; ++i) {
        z = tmp[i] - gamma[i - 1] * z;
        cc[i + 1] = z;
      }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < N; i += blockDim.x) cc[i + 1] = cc[i + 1] / alpha[i];
    __syncthreads();
    if (threadIdx.x == 0)
      for (int i = N - 2; i >= 0; --i) cc[i + 1] -= gamma[i] * cc[i + 2];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < nk - 1; i += blockDim.x) {  // per-interval b and d (gsl cspline coeff_calc)
    const double h = lk[i + 1] - lk[i];
    tmp[i] = (row[i + 1] - row[i]) / h - h * (cc[i + 1] + 2.0 * cc[i]) / 3.0;
    dco[i] = (cc[i + 1] - cc[i]) / (3.0 * h);
  }
  __syncthreads();
  const double lnkmin = lk[0], lnkmax_t = lk[nk - 1];
  const double inv_dlk = (nk - 1) / (lnkmax_t - lnkmin);
  // edge derivatives for the ln k extrapolation of ccl_f2d_t_eval (order 1 below, order 2 above)
  double d1_lo, d1_hi, d2_hi;
  {
    d1_lo = tmp[0];
    const int i = nk - 2;
    const double h = lk[i + 1] - lk[i];
    d1_hi = tmp[i] + h * (2.0 * cc[i] + 3.0 * dco[i] * h);
    d2_hi = 2.0 * cc[i] + 6.0 * dco[i] * h;
  }
  auto plin = [&](double x) -> double {
    const double xi = fmin(fmax(x, lnkmin), lnkmax_t);
    int i = (int)((xi - lnkmin) * inv_dlk);
    i = min(max(i, 0), nk - 2);
    while (i > 0 && xi < lk[i]) --i;
    while (i < nk - 2 && xi >= lk[i + 1]) ++i;
    const double dx = xi - lk[i];
    double val = row[i] + dx * (tmp[i] + dx * (cc[i] + dx * dco[i]));
    const double dlk = x - xi;
    if (dlk > 0) val += d1_hi * dlk + 0.5 * d2_hi * dlk * dlk;
    else if (dlk < 0) val += d1_lo * dlk;
    return exp(val);
  };
  // moments of Delta^2(k) exp(-k^2 R^2): m0 = sigma^2, m1 = d sigma^2 / dR, m2 = d2 sigma^2 / dR2
  auto moments = [&](double R, double& m0, double& m1, double& m2) {
    const double lnkmax = fmax(lnkmax_t, log(30. / R));
    const int npanel = (int)(HF_PANELS * fmax(1.0, (lnkmax - lnkmin) / 25.0));
    const double step = (lnkmax - lnkmin) / npanel;
    double a0 = 0.0, a1 = 0.0, a2 = 0.0;
    // a thread keeps its Gauss-Legendre index q (blockDim is a multiple of 16) and visits every
    // (blockDim/16)-th panel: k = exp(panel start) * exp(offset of node q) advances by a constant ratio,
    // so the node's exp(ln k) costs one multiplication
    const int q = threadIdx.x & 15, jstride = blockDim.x >> 4;
    const double gx = glx[q], gw = glw[q];
    const double fq = exp(0.5 * step * (1.0 + gx));
    const double eratio = exp(step * jstride);
    double ej = exp(lnkmin + (threadIdx.x >> 4) * step);
    for (int j = threadIdx.x >> 4; j < npanel; j += jstride, ej *= eratio) {
      const double e0 = lnkmin + j * step, e1 = (j == npanel - 1) ? lnkmax : lnkmin + (j + 1) * step;
      const double mid = 0.5 * (e1 + e0), half = 0.5 * (e1 - e0);
      const double x = mid + half * gx;
      const double kk = ej * fq, k2 = kk * kk;
      const double base = half * gw * plin(x) * kk * k2 * (0.5 / (M_PI * M_PI)) * exp(-k2 * R * R);
      a0 += base;
      a1 += base * (-k2 * 2.0 * R);
      a2 += base * (-2.0 * k2 + 4.0 * k2 * k2 * R * R);
    }
    // warp sums by shuffles, then the eight warp partials through shared memory: two barriers per call
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      a0 += __shfl_xor_sync(0xffffffffu, a0, o);
      a1 += __shfl_xor_sync(0xffffffffu, a1, o);
      a2 += __shfl_xor_sync(0xffffffffu, a2, o);
    }
    const int w = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0) { red[0][w] = a0; red[1][w] = a1; red[2][w] = a2; }
    __syncthreads();
    m0 = 0.0; m1 = 0.0; m2 = 0.0;
    for (int q = 0; q < (int)(blockDim.x >> 5); ++q) { m0 += red[0][q]; m1 += red[1][q]; m2 += red[2][q]; }
    __syncthreads();
  };
  // ---- sigma(R) = 1 (get_rsigma, src/ccl_halofit.c:259-315, root searched in [1e-64, 1e16] Mpc there):
  //      Newton steps on ln sigma^2 as a function of ln R (slope -(n_eff + 3), nearly constant), kept
  //      inside the bracket found so far; bisection whenever a step leaves it.  The moments of the last
  //      evaluation are the ones at the returned R. ----
  double m0 = 0, m1 = 0, m2 = 0;
  const double X_MIN = log(1e-64), X_MAX = log(1e4);
  double lo = X_MIN, hi = X_MAX;      // sigma^2 > 1 at lo, < 1 at hi once seen
  bool have_lo = false, have_hi = false;
  const size_t oidx = (size_t)ic * na + ia;
  bool ok = false;
  for (int it = 0; it < 200; ++it) {
    const double R = exp(x);
    moments(R, m0, m1, m2);
    const double f = log(m0);
    if (f > 0) { lo = x; have_lo = true; } else if (f < 0) { hi = x; have_hi = true; }
    if (f == 0.0) { ok = true; break; }
    // Halley step on ln sigma^2(ln R): f' = R m1 / m0, f'' = f' + R^2 m2 / m0 - f'^2 (falls back to Newton
    // when the correction is not a modest one)
    const double f1 = R * m1 / m0;
    const double f2 = f1 + R * R * m2 / m0 - f1 * f1;
    double den = f1 - 0.5 * f * f2 / f1;
    if (!(den * f1 > 0.25 * f1 * f1)) den = f1;
    double step = -f / den;
    if (!(fabs(step) <= 8.0)) step = (f > 0) ? 8.0 : -8.0;   // also catches NaN / zero slope
    double xn = x + step;
    if (have_lo && have_hi) {
      if (!(xn > lo && xn < hi)) xn = 0.5 * (lo + hi);
    } else if (xn < X_MIN || xn > X_MAX) {
      xn = fmin(fmax(xn, X_MIN), X_MAX);
      if (xn == x) break;  // pinned at an end of the search range: no root
    }
    if (fabs(xn - x) <= HF_ROOT_TOL * (1.0 + fabs(x)) ||
        (have_lo && have_hi && (hi - lo) <= HF_ROOT_TOL * (1.0 + fabs(x)))) {
      ok = true;  // the step is below the resolution of R: keep this evaluation's moments
      break;
    }
    x = xn;
  }
  if (!ok) {
    if (threadIdx.x == 0) { rsigma_out[oidx] = -1.0; atomicExch(fail, 1); }
    for (int i = threadIdx.x; i < nk; i += blockDim.x) logpnl[oidx * nk + i] = nan("");
    x = log(1.0);
    nroots = 0;
    continue;
  }
  const double R = exp(x);
  x_prev2 = x_prev; a_prev2 = a_prev;
  x_prev = x; a_prev = a_grid[ia];
  ++nroots;
  const double sigma2 = m0;
  const double neff = -R / sigma2 * m1 - 3.0;
  const double ds = (neff + 3.0) / (-R / sigma2);
  const double Cc = -1.0 * (m2 * R * R / sigma2 + ds * R / sigma2 - ds * ds * R * R / sigma2 / sigma2);
  if (threadIdx.x == 0) rsigma_out[oidx] = R;
  // ---- ccl_halofit_power (src/ccl_halofit.c:828-930), f_nu = 0 ----
  // the coefficients of the fit (eleven powers) are formed by one thread and shared; in the loop over k the powers of
  // y = k R_sigma are exponentials of multiples of ln y = ln k + ln R_sigma, which needs no logarithm
  __shared__ double hf[12];
  if (threadIdx.x == 0) {
    const double a = a_grid[ia];
    const double hn = h_over_h0(p, a);
    const double om = (p[P_OC] + p[P_OB]) / (a * a * a) / hn / hn;
    const double ode = p[P_OL] * pow(a, -3 * (1 + p[P_W0] + p[P_WA])) * exp(3 * p[P_WA] * (a - 1)) / hn / hn;
    const double weffa = p[P_W0];
    const double n1 = neff, n2 = n1 * n1, n3 = n2 * n1, n4 = n2 * n2;
    const double an = pow(10.0, 1.5222 + 2.8553 * n1 + 2.3706 * n2 + 0.9903 * n3 + 0.2250 * n4 - 0.6038 * Cc +
                                    0.1749 * ode * (1.0 + weffa));
    const double bn = pow(10.0, -0.5642 + 0.5864 * n1 + 0.5716 * n2 - 1.5474 * Cc + 0.2279 * ode * (1.0 + weffa));
    const double cn = pow(10.0, 0.3698 + 2.0404 * n1 + 0.8161 * n2 + 0.5869 * Cc);
    const double gamman = 0.1971 - 0.0843 * n1 + 0.8460 * Cc;
    const double alphan = fabs(6.0835 + 1.3373 * n1 - 0.1959 * n2 - 5.5274 * Cc);
    const double betan = 2.0379 - 0.7354 * n1 + 0.3157 * n2 + 1.2490 * n3 + 0.3980 * n4 - 0.1682 * Cc;
    const double nun = pow(10.0, 5.2105 + 3.6902 * n1);
    double f1 = 1.0, f2 = 1.0, f3 = 1.0;
    if (fabs(1.0 - om) > 0.01) {
      const double f1a = pow(om, -0.0732), f2a = pow(om, -0.1423), f3a = pow(om, 0.0725);
      const double f1b = pow(om, -0.0307), f2b = pow(om, -0.0585), f3b = pow(om, 0.0743);
      const double fb = ode / (1.0 - om);
      f1 = fb * f1b + (1.0 - fb) * f1a;
      f2 = fb * f2b + (1.0 - fb) * f2a;
      f3 = fb * f3b + (1.0 - fb) * f3a;
    }
    hf[0] = an; hf[1] = bn; hf[2] = log(cn * f3); hf[3] = 3.0 - gamman; hf[4] = alphan; hf[5] = betan; hf[6] = nun;
    hf[7] = 3.0 * f1; hf[8] = f2;
  }
  __syncthreads();
  {
    const double an = hf[0], bn = hf[1], lcf = hf[2], e3 = hf[3], alphan = hf[4], betan = hf[5], nun = hf[6],
                 e1 = hf[7], e2 = hf[8];
    const double lnR = log(R), ln2pi2 = log(2.0 * M_PI * M_PI);
    for (int i = threadIdx.x; i < nk; i += blockDim.x) {
      const double lki = lk[i];
      const double ld2 = 3.0 * lki - ln2pi2;       // ln(k^3 / (2 pi^2))
      const double lny = lki + lnR;                // ln(k / k_sigma)
      const double y = exp(lny), y2 = y * y;
      const double fy = y / 4.0 + y2 / 8.0;
      const double DeltakL = exp(row[i] + ld2);
      const double DeltakQ = DeltakL * exp(betan * log1p(DeltakL) - fy) / (1.0 + alphan * DeltakL);
      const double DeltakHprime = an * exp(e1 * lny) / (1.0 + bn * exp(e2 * lny) + exp(e3 * (lcf + lny)));
      const double DeltakH = DeltakHprime / (1.0 + nun / y2);
      logpnl[oidx * nk + i] = log(DeltakQ + DeltakH) - ld2;
    }
  }
  }  // scale factors of the group
}

void launch_build_halofit(int ncosmo, int nk, const double* d_lk, int na, const double* d_a, const double* d_logplin,
                          const double* d_params, double* d_logpnl, double* d_rsigma, int* d_fail,
                          cudaStream_t stream) {
  if (ncosmo <= 0) return;
  const size_t smem = sizeof(double) * 7 * (size_t)nk;
  if (smem + (3 * 256 + 32) * sizeof(double) > 48 * 1024)  // dynamic + the static buffers
    cudaFuncSetAttribute(halofit_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  dim3 grid(ncosmo, (na + HF_GROUP - 1) / HF_GROUP);
  halofit_kernel<<<grid, 256, smem, stream>>>(nk, d_lk, na, d_a, d_logplin, d_params, d_logpnl, d_rsigma, d_fail);
}

// chi_min / chi_max rule of ccl_cl_tracer_t_new (src/ccl_tracers.c:626-666) for device-built kernels, and
// the end nodes of their abscissae: out[(ic*ntr + it)*4] = {chi_min, chi_max, x[0], x[n-1]}.
__global__ void tracer_range_kernel(int ncosmo, int ntr, int n, const double* __restrict__ x, int x_per_tracer,
                                    const double* __restrict__ w, double* __restrict__ out) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= ncosmo * ntr) return;
  const int ic = t / ntr;
  const double* xx = x + (size_t)(x_per_tracer ? t : ic) * n;
  const double* ww = w + (size_t)t * n;
  double w_max = fabs(ww[0]);
  for (int i = 0; i < n; ++i)
    if (fabs(ww[i]) >= w_max) w_max = fabs(ww[i]);
  w_max *= 5E-4;  // CCL_FRAC_RELEVANT
  double lo = xx[0], hi = xx[n - 1];
  for (int i = 0; i < n - 1; ++i)
    if (fabs(ww[i + 1]) >= w_max) { lo = xx[i]; break; }
  for (int i = n - 1; i >= 1; --i)
    if (fabs(ww[i - 1]) >= w_max) { hi = xx[i]; break; }
  out[(size_t)t * 4 + 0] = lo;
  out[(size_t)t * 4 + 1] = hi;
  out[(size_t)t * 4 + 2] = xx[0];
  out[(size_t)t * 4 + 3] = xx[n - 1];
}

// chi(a_eval) through the Akima spline of the chi(a) table: chi_max of kernel-less tracers
__global__ void chi_at_kernel(int ncosmo, int na, const double* __restrict__ a_bg, const double* __restrict__ chi,
                              double a_eval, double* __restrict__ out) {
  const int ic = blockIdx.x * blockDim.x + threadIdx.x;
  if (ic >= ncosmo) return;
  out[ic] = (a_eval == 1.0) ? 0.0 : akima_eval_direct(a_bg, chi + (size_t)ic * na, na, a_eval);
}

__global__ void log_kernel(int n, const double* __restrict__ x, double* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = log(x[i]);
}

__global__ void scale_rows_kernel(int ncosmo, int ntr, int n, double* __restrict__ w, const double* __restrict__ scale) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)ncosmo * ntr * n) return;
  w[i] *= scale[(i / n) % ntr];
}

void launch_scale_rows(int ncosmo, int ntr, int n, double* d_w, const double* d_scale, cudaStream_t stream) {
  const size_t tot = (size_t)ncosmo * ntr * n;
  if (tot) scale_rows_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, stream>>>(ncosmo, ntr, n, d_w, d_scale);
}

// ccl_get_kappa_kernel (src/ccl_tracers.c:554-567) on chi = linspace(0, chi(z_source), n) (pyccl/tracers.py:189-190)
// for every cosmology: one CTA per cosmology, one thread per node.
__global__ void kappa_kernel_builder(int na, const double* __restrict__ a_bg, const double* __restrict__ chi_bg,
                                     const int* __restrict__ n_achi, int stride, const double* __restrict__ achi_chi,
                                     const double* __restrict__ achi_a, const double* __restrict__ params,
                                     double z_source, int n, double* __restrict__ chi_out, double* __restrict__ w_out) {
  const int ic = blockIdx.x;
  const double* p = params + (size_t)ic * CCL_B200_NPAR;
  const double a_s = 1. / (1. + z_source);
  const double chi_s = (a_s == 1.0) ? 0.0 : akima_eval_direct(a_bg, chi_bg + (size_t)ic * na, na, a_s);
  const double hub = p[P_H] / CLIGHT_HMPC;
  const double pref = 1.5 * hub * hub * p[P_OM] / sinn(p, chi_s);  // get_lensing_prefactor / sinn(chi_source)
  const double step = chi_s / (n - 1.);
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const double chi = (i == n - 1) ? chi_s : step * i;  // numpy.linspace: start + step * i, last = stop
    const double a = (chi < 1e-8) ? 1.0
                                  : akima_eval_direct(achi_chi + (size_t)ic * stride, achi_a + (size_t)ic * stride,
                                                      n_achi[ic], chi);
    double w;
    if (chi != 0.0) w = pref * sinn(p, chi_s - chi) * chi * chi / a / sinn(p, chi);
    else w = pref * sinn(p, chi_s - chi) * chi / a;
    chi_out[(size_t)ic * n + i] = chi;
    w_out[(size_t)ic * n + i] = w;
  }
}

// a-only transfer functions that depend on the cosmology through its growth tables, on shared nodes a_ka[n]:
// kind 1 = redshift-space distortions, -f(a) (pyccl/tracers.py:895-899: transfer_a = (a, -growth_rate(a)));
// kind 2 = intrinsic alignments, -amp_j Omega_m / D(a_j) with amp_j = A_IA(z_j) 5e-14 rho_crit
// (pyccl/tracers.py:967-987).  D and f are read through their Akima splines like ccl_growth_factor /
// ccl_growth_rate (src/ccl_background.c:1290-1352).
__global__ void bg_transfer_kernel(int ncosmo, int na, const double* __restrict__ a_bg, const double* __restrict__ D,
                                   const double* __restrict__ f, const double* __restrict__ params, int kind, int n,
                                   const double* __restrict__ a_ka, const double* __restrict__ amp,
                                   double* __restrict__ out) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= ncosmo * n) return;
  const int ic = t / n, j = t - ic * n;
  const double a = a_ka[j];
  if (kind == 1) {
    out[t] = -akima_eval_direct(a_bg, f + (size_t)ic * na, na, a);
  } else {
    const double Da = (a == 1.0) ? 1.0 : akima_eval_direct(a_bg, D + (size_t)ic * na, na, a);
    out[t] = -amp[j] * params[(size_t)ic * CCL_B200_NPAR + P_OM] / Da;
  }
}

void launch_build_bg_transfer(int ncosmo, int na, const double* d_abg, const double* d_D, const double* d_f,
                              const double* d_params, int kind, int n, const double* d_a_ka, const double* d_amp,
                              double* d_out, cudaStream_t stream) {
  const int tot = ncosmo * n;
  if (tot > 0)
    bg_transfer_kernel<<<(tot + 127) / 128, 128, 0, stream>>>(ncosmo, na, d_abg, d_D, d_f, d_params, kind, n, d_a_ka,
                                                             d_amp, d_out);
}

void launch_build_kappa_kernel(int ncosmo, int na, const double* d_abg, const double* d_chibg, const int* d_nachi,
                               int stride, const double* d_achi_chi, const double* d_achi_a, const double* d_params,
                               double z_source, int n, double* d_chi, double* d_w, cudaStream_t stream) {
  if (ncosmo > 0)
    kappa_kernel_builder<<<ncosmo, 128, 0, stream>>>(na, d_abg, d_chibg, d_nachi, stride, d_achi_chi, d_achi_a, d_params,
                                                     z_source, n, d_chi, d_w);
}

void launch_tracer_ranges(int ncosmo, int ntr, int n, const double* d_x, int x_per_tracer, const double* d_w,
                          double* d_out, cudaStream_t stream) {
  const int tot = ncosmo * ntr;
  if (tot > 0) tracer_range_kernel<<<(tot + 127) / 128, 128, 0, stream>>>(ncosmo, ntr, n, d_x, x_per_tracer, d_w, d_out);
}
void launch_chi_at(int ncosmo, int na, const double* d_abg, const double* d_chi, double a_eval, double* d_out,
                   cudaStream_t stream) {
  if (ncosmo > 0) chi_at_kernel<<<(ncosmo + 127) / 128, 128, 0, stream>>>(ncosmo, na, d_abg, d_chi, a_eval, d_out);
}
void launch_log(int n, const double* d_x, double* d_out, cudaStream_t stream) {
  if (n > 0) log_kernel<<<(n + 127) / 128, 128, 0, stream>>>(n, d_x, d_out);
}

void launch_build_background(int ncosmo, int na, const double* d_a, const double* d_params, double dchi,
                             double epsrel, int nmax, double a_init, int stride, double* d_E, double* d_chi,
                             int* d_n, double* d_achi_chi, double* d_achi_a, double* d_growth, double* d_fgrowth,
                             double* d_growth0, cudaStream_t stream) {
  if (ncosmo <= 0) return;
  bg_distance_kernel<<<ncosmo, 128, sizeof(double) * na, stream>>>(na, d_a, d_params, d_E, d_chi);
  bg_achi_kernel<<<ncosmo, 16 * ACHI_HW, 0, stream>>>(ncosmo, na, d_a, d_params, d_chi, dchi, epsrel, nmax, stride, d_n,
                                                      d_achi_chi, d_achi_a);
  const double hmax = 2.0e-3;
  const int seg_cap = growth_seg_cap(na, a_init, hmax);
  const size_t gsm = (size_t)seg_cap * sizeof(double4) + (size_t)na * sizeof(double2) + ((size_t)na + 1) * sizeof(int);
  if (gsm > 48 * 1024)
    cudaFuncSetAttribute(bg_growth_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gsm);
  bg_growth_kernel<<<ncosmo, GROWTH_THREADS, gsm, stream>>>(ncosmo, na, d_a, d_params, a_init, hmax, seg_cap, d_growth,
                                                            d_fgrowth, d_growth0);
}

}  // namespace cclb
