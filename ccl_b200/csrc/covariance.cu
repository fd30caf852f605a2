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
#include "qag_warp.cuh"

namespace cclb {

namespace {

constexpr int CW = 4;  // warps per CTA

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
    st = qagw_first_nonzero(st);
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
  const long long gw = (long long)blockIdx.x * CW + warp, nw = (long long)gridDim.x * CW;
  const long long nel = (long long)P.nl1 * P.nl2;
  for (long long el = gw; el < nel; el += nw) {
    const int i1 = (int)(el % P.nl1), i2 = (int)(el / P.nl1);
    const double l1 = P.l1[i1], l2 = P.l2[i2];
    auto f = [&](double chi, int& s1) { return cov_integrand(P, l1, l2, chi, s1); };
    int st = 0, gsl = 0;
    const double result = qagw_qag41(f, P.chimin, P.chimax, 0.0, epsrel, limit, ws + (size_t)gw * 4 * limit,
                                     s_fv[warp], lane, gsl, st);
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
  qagw_upload_tables();
  const long long nel = (long long)prob.nl1 * prob.nl2;
  const int nblk = (int)std::min<long long>((nel + CW - 1) / CW, (long long)sms * 4);
  if (method == INTEG_SPLINE) cov_spline_kernel<<<nblk, CW * 32, 0, stream>>>(prob, scratch, nchi);
  else cov_qag_kernel<<<nblk, CW * 32, 0, stream>>>(prob, scratch, n_iteration, epsrel);
  return cudaGetLastError();
}

}  // namespace cclb
