// gsl_integration_qag with the 41-point Gauss-Kronrod rule, one warp per integral (integration/qag.c, qk.c,
// qk41.c, util.c of GSL, restated from the QUADPACK algorithm; the same sequence as oracle/limber_oracle.c
// o_qag41).  The 41 abscissae of a panel go over the lanes (two passes), the four sums are shuffle-reduced, the
// interval lists live in a per-warp slice of global memory (4 * limit doubles) and the interval with the
// largest error is found by a lane-strided scan (retrieve() on GSL's sorted list returns the same interval:
// the first of the maximal ones).
//
// Every translation unit that includes this header owns its copy of the rule in constant memory and uploads
// it once per device with qagw_upload_tables().
#pragma once
#include <cfloat>

#include <cuda_runtime.h>

#include "akima_integ.cuh"
#include "gk41_table.h"
#include "tables.h"

namespace cclb {
namespace {

__constant__ double qw_xgk[21];
__constant__ double qw_wgk[21];
__constant__ double qw_wg[10];

inline void qagw_upload_tables() {
  int dev = 0;
  cudaGetDevice(&dev);
  static bool uploaded[64] = {false};
  if (uploaded[dev & 63]) return;
  cudaMemcpyToSymbol(qw_xgk, GK41_XGK, sizeof(double) * 21);
  cudaMemcpyToSymbol(qw_wgk, GK41_WGK, sizeof(double) * 21);
  cudaMemcpyToSymbol(qw_wg, GK41_WG, sizeof(double) * 10);
  uploaded[dev & 63] = true;
}

// first non-zero status over the lanes (lane order)
__device__ __forceinline__ int qagw_first_nonzero(int st) {
  const unsigned bad = __ballot_sync(0xffffffffu, st != 0);
  return bad ? __shfl_sync(0xffffffffu, st, __ffs(bad) - 1) : 0;
}

// rescale_error (gsl integration/qk.c)
__device__ __forceinline__ double qagw_rescale_error(double err, double result_abs, double result_asc) {
  err = fabs(err);
  if (result_asc != 0 && err != 0) {
    const double x = 200 * err / result_asc;
    const double scale = x * sqrt(x);  // pow(x, 1.5): within an ulp, and the value only ranks intervals
    err = (scale < 1) ? result_asc * scale : result_asc;
  }
  if (result_abs > DBL_MIN / (50 * DBL_EPSILON)) {
    const double min_err = 50 * DBL_EPSILON * result_abs;
    if (min_err > err) err = min_err;
  }
  return err;
}

// gsl_integration_qk with the 41-point rule on [a, b].  f(x, st) is called by every lane on its abscissae and
// may set st to a non-zero status; fv is a 41-double scratch of the warp (shared memory).
template <class F>
__device__ __forceinline__ void qagw_qk41(F& f, double a, double b, double* fv, int lane, double& result,
                                          double& abserr, double& resabs, double& resasc, int& st) {
  const double center = 0.5 * (a + b), half_length = 0.5 * (b - a);
  for (int e = lane; e < 41; e += 32) {
    double xq;
    if (e == 40) xq = center;
    else if (e < 20) xq = center - half_length * qw_xgk[e];
    else xq = center + half_length * qw_xgk[e - 20];
    int s1 = 0;
    fv[e] = f(xq, s1);
    if (s1 && !st) st = s1;
  }
  st = qagw_first_nonzero(st);
  __syncwarp();
  double rk = 0, rg = 0, ra = 0, f1 = 0, f2 = 0;
  if (lane < 20) {
    f1 = fv[lane];
    f2 = fv[lane + 20];
    const double fsum = f1 + f2;
    rk = qw_wgk[lane] * fsum;
    ra = qw_wgk[lane] * (fabs(f1) + fabs(f2));
    if (lane & 1) rg = qw_wg[lane >> 1] * fsum;
  } else if (lane == 20) {
    f1 = fv[40];
    rk = f1 * qw_wgk[20];
    ra = fabs(rk);
  }
  rk = warp_sum(rk);
  rg = warp_sum(rg);
  ra = warp_sum(ra);
  const double mean = rk * 0.5;
  double rs = 0;
  if (lane < 20) rs = qw_wgk[lane] * (fabs(f1 - mean) + fabs(f2 - mean));
  else if (lane == 20) rs = qw_wgk[20] * fabs(f1 - mean);
  rs = warp_sum(rs);
  const double err = (rk - rg) * half_length;
  result = rk * half_length;
  resabs = ra * fabs(half_length);
  resasc = rs * fabs(half_length);
  abserr = qagw_rescale_error(err, resabs, resasc);
  __syncwarp();
}

// gsl_integration_qag(f, a, b, epsabs, epsrel, limit, GSL_INTEG_GAUSS41): returns the GSL status (0, EROUND,
// ESING, EMAXITER, EBADTOL, EFAILED) in `gsl`, the integrand's own status in `st` (the loop stops on it) and the
// result on every lane.  ws: 4 * limit doubles of this warp.
template <class F>
__device__ __forceinline__ double qagw_qag41(F& f, double a, double b, double epsabs, double epsrel, int limit,
                                             double* ws, double* fv, int lane, int& gsl, int& st) {
  const unsigned FULL = 0xffffffffu;
  double* alist = ws;
  double* blist = alist + limit;
  double* rlist = blist + limit;
  double* elist = rlist + limit;
  gsl = 0;
  double result = 0.0;
  if (epsabs <= 0 && (epsrel < 50 * DBL_EPSILON || epsrel < 0.5e-28)) {
    gsl = ST_GSL_EBADTOL;
    return result;
  }
  double r0, e0, ab0, as0;
  qagw_qk41(f, a, b, fv, lane, r0, e0, ab0, as0, st);
  double tolerance = fmax(epsabs, epsrel * fabs(r0));
  const double round_off = 50 * DBL_EPSILON * ab0;
  result = r0;
  if (e0 <= round_off && e0 > tolerance) { gsl = ST_GSL_EROUND; return result; }
  if ((e0 <= tolerance && e0 != as0) || e0 == 0.0) return result;
  if (limit == 1) { gsl = ST_GSL_EMAXITER; return result; }
  __syncwarp();
  if (lane == 0) { alist[0] = a; blist[0] = b; rlist[0] = r0; elist[0] = e0; }
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
    double ar[2], er[2], rab[2], ras[2];
    qagw_qk41(f, a1, b1, fv, lane, ar[0], er[0], rab[0], ras[0], st);
    qagw_qk41(f, a2, b2, fv, lane, ar[1], er[1], rab[1], ras[1], st);
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
      const double tmp = (1 + 100 * DBL_EPSILON) * (fabs(a2) + 1000 * DBL_MIN);  // subinterval_too_small
      if (fabs(a1) <= tmp && fabs(b2) <= tmp) error_type = 3;
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
    if (st) break;
  } while (iteration < limit && !error_type && errsum > tolerance);
  double sum = 0.0;
  for (int i = lane; i < size; i += 32) sum += rlist[i];
  result = warp_sum(sum);
  if (errsum <= tolerance) gsl = 0;
  else if (error_type == 2) gsl = ST_GSL_EROUND;
  else if (error_type == 3) gsl = ST_GSL_ESING;
  else if (iteration == limit) gsl = ST_GSL_EMAXITER;
  else gsl = ST_GSL_EFAILED;
  __syncwarp();
  return result;
}

}  // namespace
}  // namespace cclb
