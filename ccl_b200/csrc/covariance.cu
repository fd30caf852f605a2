// Limber C_ell covariance (sm_100a): ccl_angular_cl_covariance (src/ccl_cls.c:377-612) over ccl_f3d_t
// (src/ccl_f3d.c).  For every (l1, l2) it integrates over chi
//     D1(l1) D2(l1) D3(l2) D4(l2) T((l1+1/2)/chi, (l2+1/2)/chi, a(chi)) ker(a) / chi^p
// either by sampling every DCHI_INTEGRATION and integrating an Akima spline of the samples
// (integ_cov_limber_spline, :413-452) or adaptively with QAG-GK41 (integ_cov_limber_qag_quad, :454-488).
// One warp owns one (l1, l2) element: lanes walk the chi samples / the Gauss-Kronrod abscissae, the sums are
// reduced with warp shuffles; the tracer machinery (kernels, transfers, f_ell) is the Limber path's
// (device_eval.cuh), with transfer_limber_wrap's `ignore_jbes_deriv = 1` rule: der_bessel 1/2 sub-tracers drop out.
// The reference's CQUAD retry after GSL_EROUND is not restated for this entry: such an element fails with
// CCL_ERROR_INTEG like any other integration failure.
#include <cfloat>

#include "akima_integ.cuh"
#include "device_eval.cuh"
#include "gk41_table.h"

namespace cclb {

namespace {

constexpr int CW = 4;  // warps per CTA

__constant__ double v_xgk[21];
__constant__ double v_wgk[21];
__constant__ double v_wg[10];

// ccl_find_a_index (src/ccl_f3d.c:46-88): interval [a_i, a_{i+1}) holding a; na-1 at and above the last node
__device__ __forceinline__ int find_a_index(const F3D& f, double a) {
  if (a >= f.a_arr[f.na - 1]) return f.na - 1;
  if (a <= f.a_arr[0]) return 0;
  int lo = 0, hi = f.na - 1;  // a_arr[lo] <= a < a_arr[hi]
  while (hi > lo + 1) {
    const int m = (hi + lo) >> 1;
    if (f.a_arr[m] > a) hi = m; else lo = m;
  }
  return lo;
}

// ccl_f3d_t_eval (src/ccl_f3d.c:273-392)
__device__ double f3d_eval(const F3D& f3d, double lk1, double lk2, double a, const Cosmo* cs, int& st) {
  double tkka_post = 0.0;
  double a_ev = a;
  const bool is_hiz = a < f3d.a_arr[0];
  const bool is_loz = a > f3d.a_arr[f3d.na - 1];
  if (is_loz || is_hiz) {
    if (f3d.extrap_linear_growth == GROWTH_NONE) { st = ST_ERR_SPLINE_EV; return nan(""); }
    a_ev = is_loz ? f3d.a_arr[f3d.na - 1] : f3d.a_arr[0];
  }
  if (f3d.is_product) {
    const double fka1 = f2d_eval(f3d.planes[0], lk1, a_ev, cs, st);
    const double fka2 = f2d_eval(f3d.planes[1], lk2, a_ev, cs, st);
    if (st != 0) return nan("");
    tkka_post = fka1 * fka2;
  } else {
    double lk1_ev = lk1, lk2_ev = lk2;
    const bool is_hik1 = lk1 > f3d.lkmax, is_lok1 = lk1 < f3d.lkmin;
    const bool extrap_k1 = (is_hik1 && (f3d.extrap_order_hik > 0)) || (is_lok1 && (f3d.extrap_order_lok > 0));
    if (is_hik1) lk1_ev = f3d.lkmax;
    else if (is_lok1) lk1_ev = f3d.lkmin;
    const bool is_hik2 = lk2 > f3d.lkmax, is_lok2 = lk2 < f3d.lkmin;
    const bool extrap_k2 = (is_hik2 && (f3d.extrap_order_hik > 0)) || (is_lok2 && (f3d.extrap_order_lok > 0));
    if (is_hik2) lk2_ev = f3d.lkmax;
    else if (is_lok2) lk2_ev = f3d.lkmin;
    const int ia = find_a_index(f3d, a_ev);
    // gsl_interp2d domain check (NaN arguments fail it)
    if (!(lk1_ev >= f3d.lkmin && lk1_ev <= f3d.lkmax && lk2_ev >= f3d.lkmin && lk2_ev <= f3d.lkmax)) {
      st = ST_ERR_SPLINE_EV;
      return nan("");
    }
    const F2D& p0 = f3d.planes[ia];
    double tkka = bicubic_eval(p0, lk1_ev, lk2_ev, 0);
    double dtkka1 = extrap_k1 ? bicubic_eval(p0, lk1_ev, lk2_ev, 1) : 0.0;
    double dtkka2 = extrap_k2 ? bicubic_eval(p0, lk1_ev, lk2_ev, 3) : 0.0;
    if (ia < f3d.na - 1) {
      const F2D& p1 = f3d.planes[ia + 1];
      const double h = (a_ev - f3d.a_arr[ia]) / (f3d.a_arr[ia + 1] - f3d.a_arr[ia]);
      tkka = tkka * (1 - h) + bicubic_eval(p1, lk1_ev, lk2_ev, 0) * h;
      if (extrap_k1) dtkka1 = dtkka1 * (1 - h) + bicubic_eval(p1, lk1_ev, lk2_ev, 1) * h;
      if (extrap_k2) dtkka2 = dtkka2 * (1 - h) + bicubic_eval(p1, lk1_ev, lk2_ev, 3) * h;
    }
    if (extrap_k1) tkka += dtkka1 * (lk1 - lk1_ev);
    if (extrap_k2) tkka += dtkka2 * (lk2 - lk2_ev);
    tkka_post = f3d.is_log ? exp(tkka) : tkka;
  }
  if (is_hiz) {
    double gz;
    if (f3d.extrap_linear_growth == GROWTH_CCL) {
      if (cs == nullptr || !cs->has_growth) { st = ST_ERR_GROWTH_INIT; return nan(""); }
      gz = growth_factor(*cs, a, st) / growth_factor(*cs, a_ev, st);
    } else
      gz = f3d.growth_factor_0;
    tkka_post *= pow(gz, (double)f3d.growth_exponent);
  }
  return tkka_post;
}

// transfer_limber_wrap(l, lk, k, chi, a, trc, cosmo, psp = NULL, ignore_jbes_deriv = 1) (src/ccl_cls.c:102-164)
__device__ double transfer_noderiv(const Tracer* trs, int n, double l, double lk, double chi, double a, int& st) {
  double transfer = 0.0;
  for (int i = 0; i < n; ++i) {
    const Tracer& tr = trs[i];
    double dd = 0.0;
    const double w = tracer_kernel(tr, chi);
    const double t = tracer_transfer(tr, lk, a, st);
    if (tr.der_bessel < 1) {
      dd = w * t;
      if (tr.der_bessel == -1) {
        const double lp1h = l + 0.5;
        dd /= (lp1h * lp1h);
      }
    }
    transfer += dd * tracer_f_ell(tr.der_angles, l);
    if (st != 0) return -1;
  }
  return transfer;
}

// cov_integrand (src/ccl_cls.c:377-411)
__device__ double cov_integrand(const CovProblem& P, double l1, double l2, double chi, int& st) {
  const double k1 = (l1 + 0.5) / chi, k2 = (l2 + 0.5) / chi;
  const double lk1 = log(k1), lk2 = log(k2);
  const double a = scale_factor_of_chi(*P.cosmo, chi, st);
  const double d1 = transfer_noderiv(P.tracers + P.coll[0].x, P.coll[0].y, l1, lk1, chi, a, st);
  if (d1 == 0) return 0;
  const double d2 = transfer_noderiv(P.tracers + P.coll[1].x, P.coll[1].y, l1, lk1, chi, a, st);
  if (d2 == 0) return 0;
  const double d3 = transfer_noderiv(P.tracers + P.coll[2].x, P.coll[2].y, l2, lk2, chi, a, st);
  if (d3 == 0) return 0;
  const double d4 = transfer_noderiv(P.tracers + P.coll[3].x, P.coll[3].y, l2, lk2, chi, a, st);
  if (d4 == 0) return 0;
  const double tkk = f3d_eval(*P.tsp, lk1, lk2, a, P.cosmo, st);
  double ker = 1.0;
  if (P.has_ker) {  // ccl_f1d_t_eval with constant extrapolation (pyccl/ccl_covs.i:95-118): the end nodes count as outside
    if (a <= P.ker.x0) ker = P.ker_y0;
    else if (a >= P.ker.xn) ker = P.ker_yf;
    else {
      int s2 = 0;
      ker = spline_eval(P.ker, a, 0, s2);
    }
  }
  return d1 * d2 * d3 * d4 * tkk * ker / pow(chi, (double)P.chipow);
}

__device__ __forceinline__ int first_nonzero(int st) {
  const unsigned bad = __ballot_sync(0xffffffffu, st != 0);
  return bad ? __shfl_sync(0xffffffffu, st, __ffs(bad) - 1) : 0;
}

__device__ __forceinline__ void cov_finish(const CovProblem& P, int i1, int i2, double result, int st, int lane) {
  if (lane != 0) return;
  const size_t o = (size_t)i1 + (size_t)P.nl1 * i2;  // cov_out[lind1 + nl1 * lind2]
  if (st == 0) P.cov[o] = result * P.prefactor;
  else {
    P.cov[o] = nan("");
    atomicCAS(P.status, 0, (int)ST_ERR_INTEG);
    atomicCAS(P.status + 1, 0, st);
  }
}

// rescale_error (gsl integration/qk.c)
__device__ __forceinline__ double c_rescale_error(double err, double result_abs, double result_asc) {
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

// gsl_integration_qk with the 41-point rule on [a, b]: the abscissae over the lanes (two passes), four shuffle-reduced sums
__device__ void cov_qk41(const CovProblem& P, double l1, double l2, double a, double b, double* fv, int lane,
                         double& result, double& abserr, double& resabs, double& resasc, int& st) {
  const double center = 0.5 * (a + b), half_length = 0.5 * (b - a);
  for (int e = lane; e < 41; e += 32) {
    double xq;
    if (e == 40) xq = center;
    else if (e < 20) xq = center - half_length * v_xgk[e];
    else xq = center + half_length * v_xgk[e - 20];
    int s1 = 0;
    fv[e] = cov_integrand(P, l1, l2, xq, s1);
    if (s1 && !st) st = s1;
  }
  st = first_nonzero(st);
  __syncwarp();
  double rk = 0, rg = 0, ra = 0, f1 = 0, f2 = 0;
  if (lane < 20) {
    f1 = fv[lane];
    f2 = fv[lane + 20];
    const double fsum = f1 + f2;
    rk = v_wgk[lane] * fsum;
    ra = v_wgk[lane] * (fabs(f1) + fabs(f2));
    if (lane & 1) rg = v_wg[lane >> 1] * fsum;
  } else if (lane == 20) {
    f1 = fv[40];
    rk = f1 * v_wgk[20];
    ra = fabs(rk);
  }
  rk = warp_sum(rk);
  rg = warp_sum(rg);
  ra = warp_sum(ra);
  const double mean = rk * 0.5;
  double rs = 0;
  if (lane < 20) rs = v_wgk[lane] * (fabs(f1 - mean) + fabs(f2 - mean));
  else if (lane == 20) rs = v_wgk[20] * fabs(f1 - mean);
  rs = warp_sum(rs);
  const double err = (rk - rg) * half_length;
  result = rk * half_length;
  resabs = ra * fabs(half_length);
  resasc = rs * fabs(half_length);
  abserr = c_rescale_error(err, resabs, resasc);
  __syncwarp();
}

}  // namespace

// integ_cov_limber_spline: sf = per-warp slice of nchi samples in global memory
__global__ void __launch_bounds__(CW * 32) cov_spline_kernel(const CovProblem P, double* scratch, int nchi) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long gw = (long long)blockIdx.x * CW + warp, nw = (long long)gridDim.x * CW;
  double* sf = scratch + (size_t)gw * nchi;
  const long long nel = (long long)P.nl1 * P.nl2;
  // ccl_linear_spacing (src/ccl_utils.c:17-37)
  const double dx = (P.chimax - P.chimin) / (nchi - 1.);
  for (long long el = gw; el < nel; el += nw) {
    const int i1 = (int)(el % P.nl1), i2 = (int)(el / P.nl1);
    const double l1 = P.l1[i1], l2 = P.l2[i2];
    int st = 0;
    if (nchi < 5) st = ST_ERR_MEMORY;  // gsl_spline_alloc(akima, n < 5)
    for (int s = lane; s < nchi && !st; s += 32) {
      const double chi = (s == 0) ? P.chimin : ((s == nchi - 1) ? P.chimax : P.chimin + dx * s);
      int s1 = 0;
      sf[s] = cov_integrand(P, l1, l2, chi, s1);
      if (s1) st = s1;
    }
    st = first_nonzero(st);
    __syncwarp();
    double acc = 0.0;
    if (!st) acc = akima_integral_uniform(sf, nchi, P.chimin, P.chimax, dx, lane);
    cov_finish(P, i1, i2, acc, st, lane);
    __syncwarp();
  }
}

// integ_cov_limber_qag_quad: gsl_integration_qag(F, chimin, chimax, 0, epsrel, limit, GSL_INTEG_GAUSS41), the
// bookkeeping of integration/qag.c; the interval list of the warp lives in its slice of the workspace
__global__ void __launch_bounds__(CW * 32) cov_qag_kernel(const CovProblem P, double* ws, int limit, double epsrel) {
  __shared__ double s_fv[CW][41];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const unsigned FULL = 0xffffffffu;
  const long long gw = (long long)blockIdx.x * CW + warp, nw = (long long)gridDim.x * CW;
  double* alist = ws + (size_t)gw * 4 * limit;
  double* blist = alist + limit;
  double* rlist = blist + limit;
  double* elist = rlist + limit;
  double* fv = s_fv[warp];
  const long long nel = (long long)P.nl1 * P.nl2;
  for (long long el = gw; el < nel; el += nw) {
    const int i1 = (int)(el % P.nl1), i2 = (int)(el / P.nl1);
    const double l1 = P.l1[i1], l2 = P.l2[i2];
    const double epsabs = 0.0;
    int st = 0, gsl = 0;
    double result = 0.0;
    if (epsabs <= 0 && (epsrel < 50 * DBL_EPSILON || epsrel < 0.5e-28)) gsl = ST_GSL_EBADTOL;
    else {
      double r0, e0, ab0, as0;
      cov_qk41(P, l1, l2, P.chimin, P.chimax, fv, lane, r0, e0, ab0, as0, st);
      double tolerance = fmax(epsabs, epsrel * fabs(r0));
      const double round_off = 50 * DBL_EPSILON * ab0;
      result = r0;
      if (e0 <= round_off && e0 > tolerance) gsl = ST_GSL_EROUND;
      else if ((e0 <= tolerance && e0 != as0) || e0 == 0.0) gsl = 0;
      else if (limit == 1) gsl = ST_GSL_EMAXITER;
      else {
        __syncwarp();
        if (lane == 0) { alist[0] = P.chimin; blist[0] = P.chimax; rlist[0] = r0; elist[0] = e0; }
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
          cov_qk41(P, l1, l2, a1, b1, fv, lane, ar[0], er[0], rab[0], ras[0], st);
          cov_qk41(P, l1, l2, a2, b2, fv, lane, ar[1], er[1], rab[1], ras[1], st);
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
      }
    }
    cov_finish(P, i1, i2, result, st ? st : gsl, lane);
    __syncwarp();
  }
}

size_t covariance_scratch_bytes(int nchi, int n_iteration) {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const size_t warps = (size_t)sms * 4 * CW;
  return warps * sizeof(double) * (size_t)std::max(nchi, 4 * n_iteration) + 256;
}

cudaError_t launch_covariance(const CovProblem& prob, int method, int nchi, int n_iteration, double epsrel,
                              double* scratch, cudaStream_t stream) {
  if (prob.nl1 == 0 || prob.nl2 == 0) return cudaSuccess;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  static bool uploaded[64] = {false};
  if (!uploaded[dev & 63]) {
    cudaMemcpyToSymbol(v_xgk, GK41_XGK, sizeof(double) * 21);
    cudaMemcpyToSymbol(v_wgk, GK41_WGK, sizeof(double) * 21);
    cudaMemcpyToSymbol(v_wg, GK41_WG, sizeof(double) * 10);
    uploaded[dev & 63] = true;
  }
  const long long nel = (long long)prob.nl1 * prob.nl2;
  const int nblk = (int)std::min<long long>((nel + CW - 1) / CW, (long long)sms * 4);
  if (method == INTEG_SPLINE) cov_spline_kernel<<<nblk, CW * 32, 0, stream>>>(prob, scratch, nchi);
  else cov_qag_kernel<<<nblk, CW * 32, 0, stream>>>(prob, scratch, n_iteration, epsrel);
  return cudaGetLastError();
}

}  // namespace cclb
