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

// a(chi) on a uniform chi grid of step <= dchi: Newton iterations seeded by the previous node's root
// and stopped when |da| < epsrel |a| (a_of_chi, src/ccl_background.c:370-410, 498-571).  The chain is
// sequential by construction; one thread per cosmology.
__global__ void bg_achi_kernel(int ncosmo, int na, const double* __restrict__ a_grid, const double* __restrict__ params,
                               const double* __restrict__ chi, double dchi, double epsrel, int nmax, int stride,
                               int* __restrict__ n_out, double* __restrict__ achi_chi, double* __restrict__ achi_a) {
  const int ic = blockIdx.x * blockDim.x + threadIdx.x;
  if (ic >= ncosmo) return;
  const double* p = params + (size_t)ic * CCL_B200_NPAR;
  const double* ch = chi + (size_t)ic * na;
  double* gx = achi_chi + (size_t)ic * stride;
  double* ga = achi_a + (size_t)ic * stride;
  const double chi0 = ch[na - 1], chif = ch[0];
  const double a0 = a_grid[na - 1], af = a_grid[0];
  int n = (int)((chif - chi0) / dchi);
  if (n > stride) n = -n;  // reported to the host: the caller's stride is too small
  n_out[ic] = n;
  if (n < 5) return;
  const double dx = (chif - chi0) / (n - 1.);
  int j = na - 1;  // a_grid[j-1] <= cur < a_grid[j], hunted downwards as chi grows
  auto chi_of = [&](double a) -> double {
    if (a >= 1.0) return (a > 1.0) ? -chi_panel(p, 1.0, a) : 0.0;
    while (j > 1 && a < a_grid[j - 1]) --j;
    while (j < na - 1 && a >= a_grid[j]) ++j;
    return ch[j] + chi_panel(p, a, a_grid[j]);
  };
  gx[0] = chi0;
  ga[0] = a0;
  double cur = a0;
  for (int i = 1; i < n - 1; ++i) {
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
    gx[i] = target;
    ga[i] = cur;
  }
  gx[n - 1] = chif;
  ga[n - 1] = af;
}

// Linear growth: y0' = y1 / (a^3 E), y1' = 1.5 E a Omega_m(a) y0 from a_i with y0 = a_i, y1 = a_i^3 E(a_i)
// (src/ccl_background.c:153-165, 209-284), integrated in ln a with classical RK4 at a step far below
// the reference's RKCK tolerance (1e-6); D = y0 / y0(1), f = y1 / (a^2 E y0).  One thread per cosmology.
__global__ void bg_growth_kernel(int ncosmo, int na, const double* __restrict__ a_grid, const double* __restrict__ params,
                                 double a_init, double hmax, double* __restrict__ growth, double* __restrict__ fgrowth,
                                 double* __restrict__ growth0) {
  const int ic = blockIdx.x * blockDim.x + threadIdx.x;
  if (ic >= ncosmo) return;
  const double* p = params + (size_t)ic * CCL_B200_NPAR;
  auto rhs = [&](double lna, double y0, double y1, double& d0, double& d1) {
    const double a = exp(lna);
    const double hn = h_over_h0(p, a);
    const double om = (p[P_OC] + p[P_OB]) / (a * a * a) / hn / hn;
    d0 = a * y1 / (a * a * a * hn);
    d1 = a * 1.5 * hn * a * om * y0;
  };
  double lna = log(a_init);
  double y0 = a_init, y1 = a_init * a_init * a_init * h_over_h0(p, a_init);
  double* D = growth + (size_t)ic * na;
  double* F = fgrowth + (size_t)ic * na;
  for (int i = 0; i < na; ++i) {
    const double a = a_grid[i];
    if (a < a_init) {  // below the starting point the reference sets D = a, f = 1
      D[i] = a;
      F[i] = 1.0;
      continue;
    }
    const double target = log(a);
    const double span = target - lna;
    if (span > 0) {
      const int ns = (int)ceil(span / hmax);
      const double h = span / ns;
      for (int s = 0; s < ns; ++s) {
        double k10, k11, k20, k21, k30, k31, k40, k41;
        rhs(lna, y0, y1, k10, k11);
        rhs(lna + 0.5 * h, y0 + 0.5 * h * k10, y1 + 0.5 * h * k11, k20, k21);
        rhs(lna + 0.5 * h, y0 + 0.5 * h * k20, y1 + 0.5 * h * k21, k30, k31);
        rhs(lna + h, y0 + h * k30, y1 + h * k31, k40, k41);
        y0 += h / 6.0 * (k10 + 2 * k20 + 2 * k30 + k40);
        y1 += h / 6.0 * (k11 + 2 * k21 + 2 * k31 + k41);
        lna += h;
      }
      lna = target;
    }
    D[i] = y0;
    F[i] = y1 / (a * a * h_over_h0(p, a) * y0);
  }
  // the a-grid ends at a = 1 (A_SPLINE_MAX is fixed to 1): normalise D(1) = 1
  const double g0 = D[na - 1];
  growth0[ic] = g0;
  for (int i = 0; i < na; ++i)
    if (a_grid[i] >= a_init) D[i] /= g0;
}

void launch_build_background(int ncosmo, int na, const double* d_a, const double* d_params, double dchi,
                             double epsrel, int nmax, double a_init, int stride, double* d_E, double* d_chi,
                             int* d_n, double* d_achi_chi, double* d_achi_a, double* d_growth, double* d_fgrowth,
                             double* d_growth0, cudaStream_t stream) {
  if (ncosmo <= 0) return;
  bg_distance_kernel<<<ncosmo, 128, sizeof(double) * na, stream>>>(na, d_a, d_params, d_E, d_chi);
  bg_achi_kernel<<<(ncosmo + 31) / 32, 32, 0, stream>>>(ncosmo, na, d_a, d_params, d_chi, dchi, epsrel, nmax, stride,
                                                         d_n, d_achi_chi, d_achi_a);
  bg_growth_kernel<<<(ncosmo + 31) / 32, 32, 0, stream>>>(ncosmo, na, d_a, d_params, a_init, 2.0e-3, d_growth,
                                                           d_fgrowth, d_growth0);
}

}  // namespace cclb
