// C_ell -> xi(theta) (sm_100a): ccl_correlation (src/ccl_correlation.c:14-440) with its three algorithms,
//   FFTLog   ccl_tracer_corr_fftlog (:47-162) over fht() of src/ccl_fftlog.c:199-328 (q = 0, low-ringing kr),
//   Bessel   ccl_tracer_corr_bessel (:183-262): QAG-GK41 of l J_n(l theta) C(l) over [0, ELL_MAX_CORR],
//   Legendre ccl_tracer_corr_legendre (:298-400): sum_l (2l+1)/(4 pi) C_l P_l(cos theta) (P_l^2 for NG).
// The input C_ell is a ccl_f1d_t (Akima; constant below, power law above: src/ccl_f1d.c).
//
// FFTLog's two length-N FFTs (FFTW in the reference; N = N_ELL_CORR = 5000 = 2^3 5^4) are evaluated from
// their definition with a tabulated twiddle table, one CTA per output frequency: 2 x N^2/2 complex
// multiply-adds (the input is real, so only frequencies 0..N/2 are formed and the backward sum uses the
// Hermitian symmetry of b = B u / N).  That is 25 Mflop-pairs -- microseconds on 148 SMs, fewer launches than
// a mixed-radix plan and exact to the rounding of one dot product per output.
#include <algorithm>
#include <cfloat>

#include "akima_coef.cuh"
#include "qag_warp.cuh"

namespace cclb {

namespace {

constexpr double PI = 3.14159265358979323846;

// ccl_f1d_t as ccl_correlation.c builds it: y0 below x_ini; above x_end either yf or y_end (x/x_end)^der_hi
struct CorrF1D {
  const double* x;
  const double4* cr;  // [n-1] {y_i, b_i, c_i, d_i}
  int n, hi_loglog;
  double y0, yf, x_ini, x_end, y_end, der_hi;
};

// ccl_f1d_t_eval (src/ccl_f1d.c:118-191); the interval is gsl_interp_bsearch's
__device__ __forceinline__ double corr_f1d_eval(const CorrF1D& f, double x) {
  if (x <= f.x_ini) return f.y0;
  if (x >= f.x_end) return f.hi_loglog ? f.y_end * pow(x / f.x_end, f.der_hi) : f.yf;
  int lo = 0, hi = f.n - 1;
  while (hi > lo + 1) {
    const int m = (hi + lo) >> 1;
    if (f.x[m] > x) hi = m; else lo = m;
  }
  const double4 c = f.cr[lo];
  const double d = x - f.x[lo];
  return c.x + d * (c.y + d * (c.z + d * c.w));
}

// taper_cl (src/ccl_correlation.c:20-39) for one element
__device__ __forceinline__ double corr_taper(double ell, double cl, const double* lim) {
  if (ell < lim[0] || ell > lim[3]) return 0.0;
  if (ell >= lim[1] && ell <= lim[2]) return cl;
  if (ell < lim[1]) cl *= cos((ell - lim[1]) / (lim[1] - lim[0]) * PI / 2.);
  if (ell > lim[2]) cl *= cos((ell - lim[2]) / (lim[3] - lim[2]) * PI / 2.);
  return cl;
}

__global__ void corr_akima_kernel(int n, const double* x, const double* y, double4* cr) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n - 1; i += gridDim.x * blockDim.x)
    cr[i] = akima_interval_coef(x, y, n, i);
}

// out[i] = C(x_i) (tapered) times x_i^xpow; x_i = xs[i], or i when xs is null (Legendre: integer multipoles)
__global__ void corr_eval_kernel(CorrF1D f, int nx, const double* xs, double* out, int do_taper, const double* taper,
                                 int times_x) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nx; i += gridDim.x * blockDim.x) {
    const double x = xs ? xs[i] : (double)i;
    double v = corr_f1d_eval(f, x);
    if (do_taper) v = corr_taper(x, v, taper);
    out[i] = times_x ? x * v : v;  // prefac_pk = pow(k, dim/2 - q) = k for dim 2, q 0 (src/ccl_fftlog.c:270)
  }
}

// ---- gsl_sf_lngamma_complex_e (GSL specfunc/gamma.c, trig.c; published Lanczos g = 7 form), as in the oracle ----
__constant__ double c_lanczos7[9] = {
    0.99999999999980993227684700473478, 676.520368121885098567009190444019, -1259.13921672240287047156078755283,
    771.3234287776530788486528258894,   -176.61502916214059906584551354,    12.507343278686904814458936853,
    -0.13857109526572011689554707,      9.984369578019570859563e-6,         1.50563273514931155834e-7};

__device__ __forceinline__ double angle_restrict_symm(double theta) {
  const double P1 = 4 * 7.8539812564849853515625e-01, P2 = 4 * 3.7748947079307981766760e-08,
               P3 = 4 * 2.6951514290790594840552e-15, TwoPi = 2 * (P1 + P2 + P3);
  const double y = (theta >= 0 ? 1 : -1) * 2 * floor(fabs(theta) / TwoPi);
  double r = ((theta - y * P1) - y * P2) - y * P3;
  if (r > PI) r = (((r - 2 * P1) - 2 * P2) - 2 * P3);
  else if (r < -PI) r = (((r + 2 * P1) + 2 * P2) + 2 * P3);
  return r;
}

__device__ __forceinline__ void complex_log(double zr, double zi, double& lnr, double& theta) {
  const double ax = fabs(zr), ay = fabs(zi);
  const double mn = fmin(ax, ay), mx = fmax(ax, ay);
  lnr = log(mx) + 0.5 * log(1.0 + (mn / mx) * (mn / mx));
  theta = atan2(zi, zr);
}

__device__ void lngamma_lanczos_complex(double zr, double zi, double& yr, double& yi) {
  zr -= 1.0;
  double Ag_r = c_lanczos7[0], Ag_i = 0.0;
  for (int k = 1; k <= 8; k++) {
    const double R = zr + k, I = zi;
    const double a = c_lanczos7[k] / (R * R + I * I);
    Ag_r += a * R;
    Ag_i -= a * I;
  }
  double log1_r, log1_i, logAg_r, logAg_i;
  complex_log(zr + 7.5, zi, log1_r, log1_i);
  complex_log(Ag_r, Ag_i, logAg_r, logAg_i);
  yr = (zr + 0.5) * log1_r - zi * log1_i - (zr + 7.5) + 0.9189385332046727418 + logAg_r;
  yi = angle_restrict_symm(zi * log1_r + (zr + 0.5) * log1_i - zi + logAg_i);
}

__device__ void lngamma_complex(double zr, double zi, double& lnr, double& arg) {
  if (zr <= 0.5) {  // reflection to the right half plane
    double a, b, ls_r, ls_i;
    lngamma_lanczos_complex(1.0 - zr, -zi, a, b);
    const double sr = PI * zr, si = PI * zi;
    if (si > 60.0) { ls_r = -0.69314718055994530942 + si; ls_i = 0.5 * PI - sr; }
    else if (si < -60.0) { ls_r = -0.69314718055994530942 - si; ls_i = -0.5 * PI + sr; }
    else complex_log(sin(sr) * cosh(si), cos(sr) * sinh(si), ls_r, ls_i);
    ls_i = angle_restrict_symm(ls_i);
    lnr = 1.14472988584940017414 - ls_r - a;
    arg = angle_restrict_symm(-ls_i - b);
  } else
    lngamma_lanczos_complex(zr, zi, lnr, arg);
}

// goodkr (src/ccl_fftlog.c:72-86), q = 0
__device__ double fftlog_goodkr(int N, double mu, double L, double kr) {
  const double x = (mu + 1) / 2, y = PI * N / (2 * L);
  double lnr, argp;
  lngamma_complex(x, y, lnr, argp);
  const double arg = log(2 / kr) * N / L + (argp + argp) / PI;
  const double iarg = round(arg);
  if (arg != iarg) kr *= exp((arg - iarg) * L / N);
  return kr;
}

// compute_u_coefficients (src/ccl_fftlog.c:95-125, q == 0 branch) for m = 0..N/2, the twiddle table
// w_j = exp(2 pi i j / N) and the low-ringing kcrc (scal[0])
__global__ void fftlog_setup_kernel(int N, double mu, double L, double2* u, double2* w, double* scal) {
  const double kcrc = fftlog_goodkr(N, mu, L, 1.0);
  const double y = PI / L, k0r0 = kcrc * exp(-L), t = -2 * y * log(k0r0 / 2);
  const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
  if (tid == 0) scal[0] = kcrc;
  for (int m = tid; m <= N / 2; m += nth) {
    double lnr, phi;
    lngamma_complex((mu + 1) / 2, m * y, lnr, phi);
    double s, c;
    sincos(m * t + 2 * phi, &s, &c);
    if ((N % 2) == 0 && m == N / 2) s = 0.0;  // u[N/2] = creal(u[N/2])
    u[m] = make_double2(c, s);
  }
  for (int j = tid; j < N; j += nth) {
    double s, c;
    sincospi(2.0 * j / N, &s, &c);
    w[j] = make_double2(c, s);
  }
}

constexpr int FT = 256;  // threads per DFT output

__device__ __forceinline__ double2 block_sum2(double2 v, double2* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    v.x += __shfl_xor_sync(0xffffffffu, v.x, o);
    v.y += __shfl_xor_sync(0xffffffffu, v.y, o);
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) red[warp] = v;
  __syncthreads();
  double2 r = make_double2(0.0, 0.0);
  for (int q = 0; q < FT / 32; ++q) { r.x += red[q].x; r.y += red[q].y; }
  return r;
}

// b[m] = u[m] / N * sum_i a[i] exp(-2 pi i m i / N), m = 0..N/2 (fftw forward, src/ccl_fftlog.c:290-292)
__global__ void __launch_bounds__(FT) fftlog_forward_kernel(int N, const double* a, const double2* w,
                                                             const double2* u, double2* b) {
  __shared__ double2 red[FT / 32];
  const int m = blockIdx.x;
  double sr = 0.0, si = 0.0;
  for (int i = threadIdx.x; i < N; i += FT) {
    const int idx = (int)(((long long)i * m) % N);
    const double2 t = w[idx];
    const double ai = a[i];
    sr += ai * t.x;
    si -= ai * t.y;
  }
  const double2 B = block_sum2(make_double2(sr, si), red);
  if (threadIdx.x == 0) {
    const double ur = u[m].x / (double)N, ui = u[m].y / (double)N;
    b[m] = make_double2(B.x * ur - B.y * ui, B.x * ui + B.y * ur);
  }
}

// c[n] = Re sum_{m<N} b[m] exp(+2 pi i m n / N) with b[N-m] = conj(b[m]); the output array is reversed and
// scaled: th[i] = r_i = k0r0 / k_0 exp(i L / N), xi[i] = (2 pi)^-1 r_i^-1 c[N-1-i] (src/ccl_fftlog.c:274-306)
__global__ void __launch_bounds__(FT) fftlog_backward_kernel(int N, const double2* b, const double2* w,
                                                              const double* scal, double k0, double L, double* th,
                                                              double* xi) {
  __shared__ double2 red[FT / 32];
  const int n = blockIdx.x;
  const int half = (N - 1) / 2;  // paired frequencies 1..half
  double sr = 0.0;
  for (int m = 1 + threadIdx.x; m <= half; m += FT) {
    const int idx = (int)(((long long)m * n) % N);
    const double2 t = w[idx], bm = b[m];
    sr += bm.x * t.x - bm.y * t.y;
  }
  const double2 S = block_sum2(make_double2(sr, 0.0), red);
  if (threadIdx.x == 0) {
    double c = b[0].x + 2.0 * S.x;
    if ((N % 2) == 0) c += (n & 1) ? -b[N / 2].x : b[N / 2].x;
    const int i = N - 1 - n;
    const double k0r0 = scal[0] * exp(-L);
    const double r = (i == 0) ? k0r0 / k0 : (k0r0 / k0) * exp(i * L / N);
    th[i] = r;
    xi[i] = (1.0 / (2 * PI)) * (1.0 / r) * c;  // pow(2 pi, -dim/2) * pow(r, -dim/2 - q), dim 2, q 0
  }
}

// wtheta[i] = ccl_f1d_t(th_arr, wth_arr; y0 = wth_arr[0], yf = 0)(theta_i pi / 180)  (src/ccl_correlation.c:139-152)
__global__ void corr_output_kernel(int N, const double* th, const double* wth, const double4* cr, int n_theta,
                                   const double* theta, double* wtheta) {
  CorrF1D f;
  f.x = th; f.cr = cr; f.n = N; f.hi_loglog = 0;
  f.y0 = wth[0]; f.yf = 0.0; f.x_ini = th[0]; f.x_end = th[N - 1]; f.y_end = 0.0; f.der_hi = 0.0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_theta; i += gridDim.x * blockDim.x)
    wtheta[i] = corr_f1d_eval(f, theta[i] * PI / 180.);
}

__device__ __forceinline__ int bessel_order(int corr_type) {
  return corr_type == CORR_GL ? 2 : (corr_type == CORR_LM ? 4 : 0);
}

constexpr int BW = 4;  // warps per CTA of the Bessel kernel

// ccl_tracer_corr_bessel: one warp per theta
__global__ void __launch_bounds__(BW * 32) corr_bessel_kernel(CorrF1D f, CorrProblem P, double* ws) {
  __shared__ double s_fv[BW][41];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int gw = blockIdx.x * BW + warp, nw = gridDim.x * BW;
  const int nb = bessel_order(P.corr_type);
  for (int ith = gw; ith < P.n_theta; ith += nw) {
    const double th = P.theta[ith] * PI / 180;
    auto integrand = [&](double l, int&) { return l * jn(nb, l * th) * corr_f1d_eval(f, l); };
    int gsl = 0, st = 0;
    const double result = qagw_qag41(integrand, 0.0, P.ell_max_corr, 0.0, P.epsrel, P.n_iteration,
                                     ws + (size_t)gw * 4 * P.n_iteration, s_fv[warp], lane, gsl, st);
    if (lane == 0) {
      P.wtheta[ith] = result / (2 * PI);
      if (gsl) atomicOr(P.status, gsl);
    }
    __syncwarp();
  }
}

// ccl_tracer_corr_legendre.  The upward recurrences of gsl_sf_legendre_Pl_array / gsl_sf_legendre_Plm(l, 2, x)
// (legendre_poly.c) are sequential in l, so one thread per theta walks all multipoles; what does not depend on
// theta -- the recurrence coefficients (2l-1)/l, (l-1)/l (or (2l-1)/(l-2), (l+1)/(l-2) for P_l^2) and the weighted
// spectrum C_l (2l+1) (or C_l (2l+1)/(l(l+1))) -- is tabulated in parallel first (corr_legendre_tables), which takes
// the divisions out of the 60000-step chain.  The sum runs over l = 1..lmax-1 in the reference's order.
__global__ void corr_legendre_tables(int lmax, int corr_type, const double* cl_arr, double* ra, double* rb, double* clw) {
  for (int l = blockIdx.x * blockDim.x + threadIdx.x; l <= lmax; l += gridDim.x * blockDim.x) {
    const double dl = (double)l;
    if (corr_type == CORR_GG) {
      ra[l] = l ? (2 * dl - 1) / dl : 0.0;
      rb[l] = l ? (dl - 1) / dl : 0.0;
      clw[l] = cl_arr[l] * (2 * dl + 1);
    } else {
      ra[l] = l > 2 ? (2 * dl - 1) / (dl - 2) : 0.0;
      rb[l] = l > 2 ? (dl + 1) / (dl - 2) : 0.0;
      clw[l] = l >= 2 ? cl_arr[l] * ((2 * dl + 1.) / ((dl + 0.) * (dl + 1.))) : 0.0;
    }
  }
}

constexpr int LT = 1024;  // multipoles per shared-memory tile of the three tables

// One warp per 32 thetas: the lanes load a tile of the tables together (coalesced), then each lane advances its own
// recurrence through the tile from shared memory (all lanes read the same element: a broadcast).
__global__ void __launch_bounds__(32) corr_legendre_kernel(int lmax, const double* ra, const double* rb,
                                                           const double* clw, CorrProblem P) {
  __shared__ double s_ra[LT], s_rb[LT], s_cw[LT];
  const int lane = threadIdx.x;
  const int ith = blockIdx.x * 32 + lane;
  const bool live = ith < P.n_theta;
  const double cth = live ? cos(P.theta[ith] * PI / 180) : 0.0;
  const bool gg = P.corr_type == CORR_GG;
  // the first terms: l = 1 for P_l (P_0 = 1, P_1 = x); l = 2, 3 for P_l^2 (P_2^2 = 3 (1-x)(1+x), P_3^2 = 5 x P_2^2)
  double p2, p1, sum = 0.0;
  int l0;
  if (gg) {
    p2 = 1.0; p1 = cth; l0 = 2;
    if (lmax > 1) sum += clw[1] * cth;
  } else {
    const double root_factor = sqrt(1.0 - cth) * sqrt(1.0 + cth);
    double p_mm = 1.0, fact_coeff = 1.0;
    for (int i = 1; i <= 2; i++) { p_mm *= -fact_coeff * root_factor; fact_coeff += 2.0; }
    p2 = p_mm; p1 = cth * 5 * p_mm; l0 = 4;
    if (lmax > 2) sum += clw[2] * p2;
    if (lmax > 3) sum += clw[3] * p1;
  }
  for (int base = l0; base < lmax; base += LT) {
    const int n = min(LT, lmax - base);
    __syncwarp();
    for (int i = lane; i < n; i += 32) { s_ra[i] = ra[base + i]; s_rb[i] = rb[base + i]; s_cw[i] = clw[base + i]; }
    __syncwarp();
#pragma unroll 8
    for (int i = 0; i < n; ++i) {
      const double p = cth * s_ra[i] * p1 - s_rb[i] * p2;
      p2 = p1; p1 = p;
      sum += s_cw[i] * p;
    }
  }
  if (live) P.wtheta[ith] = sum / (PI * 4);
}

// ---- generalised FFTLog: ccl_fftlog_ComputeXi_general = general_fht (src/ccl_fftlog.c:330-503), batched over mu ----
__device__ __forceinline__ void cmul(double& ar, double& ai, double br, double bi) {
  const double r = ar * br - ai * bi, i = ar * bi + ai * br;
  ar = r; ai = i;
}

// u = 2^(q + sph + 2 i y - deriv) [k0r0^(-2 i y)] Gamma(xp + i y) / Gamma(xm - i y) prod_j (q + 2 i y - (j - sph)) (-1)^deriv
// (compute_u_coefficients_new_deriv :171-197; goodkr_new_deriv :134-164 without the k0r0 factor); q includes plaw
__device__ void general_coef(double mu, double q, int sph, double deriv, double y, double lnk0r0, int with_k0r0,
                             double& ur, double& ui) {
  const double xp = (mu + 1 + q - deriv - 0.5 * sph) / 2, xm = (mu + 1 - q + deriv + 0.5 * sph) / 2;
  double lnrp, phip, lnrm, phim;
  lngamma_complex(xp, y, lnrp, phip);
  lngamma_complex(xm, -y, lnrm, phim);
  const double mod2 = pow(2.0, q + 1.0 * sph - deriv);
  double ph = 2 * y * 0.69314718055994530942;
  if (with_k0r0) ph += -2 * y * lnk0r0;
  double r = mod2 * cos(ph), i = mod2 * sin(ph);
  const double g = exp(lnrp - lnrm);
  cmul(r, i, g * cos(phip - phim), g * sin(phip - phim));
  for (int j = 1; j <= (int)deriv; j++) cmul(r, i, q - (j - sph), 2 * y);
  const double sgn = ((int)deriv & 1) ? -1.0 : 1.0;  // pow(-1, deriv) for the integer orders the wrapper admits
  ur = r * sgn; ui = i * sgn;
}

struct FGProblem {
  int nmu, npk, N, sph;
  double q, deriv, plaw, L;  // q = q_in - sph (general_fht :351)
  const double* mu;          // [nmu] as passed (before + sph / 2)
  const double* k;           // [N]
  const double* f;           // [npk][N]
  double* a;                 // [npk][N] k^(-sph - q) f
  double2* w;                // [N]
  double2* u;                // [nmu][N/2 + 1]
  double2* B;                // [npk][N/2 + 1]
  double* kcrc;              // [nmu]
  double* out;               // [nmu][npk + 1][N]
};

__global__ void fg_setup_kernel(FGProblem P) {
  const int N = P.N, imu = blockIdx.x;
  const double mu = P.mu[imu] + 0.5 * P.sph;
  __shared__ double s_kcrc;
  if (threadIdx.x == 0) {
    double pr, pi_;
    general_coef(mu, P.q + P.plaw, P.sph, P.deriv, PI * N / (2 * P.L), 0.0, 0, pr, pi_);
    const double arg = atan2(pi_, pr) / PI;
    const double iarg = round(arg);
    double kr = P.mu[imu] + 1.0;
    if (arg != iarg) kr *= exp((arg - iarg) * P.L / N);
    s_kcrc = kr;
    P.kcrc[imu] = kr;
  }
  __syncthreads();
  const double y = PI / P.L, lnk0r0 = log(s_kcrc * exp(-P.L));
  for (int m = threadIdx.x; m <= N / 2; m += blockDim.x) {
    double ur, ui;
    general_coef(mu, P.q + P.plaw, P.sph, P.deriv, m * y, lnk0r0, 1, ur, ui);
    if ((N % 2) == 0 && m == N / 2) ui = 0.0;
    P.u[(size_t)imu * (N / 2 + 1) + m] = make_double2(ur, ui);
  }
  if (imu == 0) {
    for (int j = threadIdx.x; j < N; j += blockDim.x) {
      double sn, cs;
      sincospi(2.0 * j / N, &sn, &cs);
      P.w[j] = make_double2(cs, sn);
    }
    for (int i = threadIdx.x; i < P.npk * N; i += blockDim.x)
      P.a[i] = pow(P.k[i % N], -1.0 * P.sph - P.q) * P.f[i];
  }
}

// B[t][m] = sum_i a[t][i] exp(-2 pi i m i / N): one warp per (t, m)
__global__ void fg_forward_kernel(FGProblem P) {
  const int N = P.N, nh = N / 2 + 1;
  const int lane = threadIdx.x & 31;
  const long long gw = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (gw >= (long long)P.npk * nh) return;
  const int t = (int)(gw / nh), m = (int)(gw % nh);
  const double* a = P.a + (size_t)t * N;
  double sr = 0.0, si = 0.0;
  for (int i = lane; i < N; i += 32) {
    const double2 tw = P.w[(int)(((long long)i * m) % N)];
    sr += a[i] * tw.x;
    si -= a[i] * tw.y;
  }
  sr = warp_sum(sr);
  si = warp_sum(si);
  if (lane == 0) P.B[(size_t)t * nh + m] = make_double2(sr, si);
}

// out[imu][1 + t][i] = r_i^(-sph - q) (sqrt(pi) / 4)^sph Re sum_m B[t][m] u[imu][m] / N exp(+2 pi i m n / N), n = N-1-i;
// out[imu][0][i] = r_i = kcrc / k[N-1-i]: one warp per (imu, t, n)
__global__ void fg_backward_kernel(FGProblem P) {
  const int N = P.N, nh = N / 2 + 1, half = (N - 1) / 2;
  const int lane = threadIdx.x & 31;
  const long long gw = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (gw >= (long long)P.nmu * P.npk * N) return;
  const int n = (int)(gw % N), t = (int)((gw / N) % P.npk), imu = (int)(gw / ((long long)N * P.npk));
  const double2* B = P.B + (size_t)t * nh;
  const double2* u = P.u + (size_t)imu * nh;
  double sr = 0.0;
  for (int m = 1 + lane; m <= half; m += 32) {
    const double2 tw = P.w[(int)(((long long)m * n) % N)];
    double br = B[m].x, bi = B[m].y;
    cmul(br, bi, u[m].x / (double)N, u[m].y / (double)N);
    sr += br * tw.x - bi * tw.y;
  }
  sr = warp_sum(sr);
  if (lane == 0) {
    double c = B[0].x * (u[0].x / (double)N) - B[0].y * (u[0].y / (double)N) + 2.0 * sr;
    if ((N % 2) == 0) {
      const double bn = B[N / 2].x * (u[N / 2].x / (double)N);  // both factors are real at the Nyquist frequency
      c += (n & 1) ? -bn : bn;
    }
    const int i = N - 1 - n;
    const double r = P.kcrc[imu] / P.k[N - 1 - i];
    double* o = P.out + (size_t)imu * (P.npk + 1) * N;
    if (t == 0) o[i] = r;
    o[(size_t)(1 + t) * N + i] = pow(r, -1.0 * P.sph - P.q) * c * (P.sph ? 0.44311346272637900682 : 1.0);
  }
}

inline size_t up256(size_t b) { return (b + 255) / 256 * 256; }

}  // namespace

size_t correlation_work_bytes(int method, int n_ell, int n_theta, int n_ell_corr, double ell_max_corr,
                              int n_iteration) {
  size_t b = up256((size_t)n_ell * sizeof(double4)) + 256;
  if (method == CORR_FFTLOG) {
    const size_t N = (size_t)n_ell_corr;
    b += up256(N * 8) * 3 + up256(N * 16) * 3 + up256(N * 32) + 256;  // a, th, wth | u, w, b | cr | scal
  } else if (method == CORR_BESSEL) {
    const size_t warps = (size_t)std::min((n_theta + BW - 1) / BW, 148 * 8) * BW;
    b += up256(warps * 4 * n_iteration * 8);
  } else {
    b += up256(((size_t)ell_max_corr + 1) * 8) * 4;  // cl_arr, ra, rb, clw
  }
  return b;
}

size_t fftlog_general_work_bytes(int nmu, int npk, int N) {
  const size_t nh = (size_t)N / 2 + 1;
  return up256((size_t)nmu * 8) + up256((size_t)N * 8) + up256((size_t)npk * N * 8) * 2 + up256((size_t)N * 16) +
         up256((size_t)nmu * nh * 16) + up256((size_t)npk * nh * 16) + up256((size_t)nmu * 8) +
         up256((size_t)nmu * (npk + 1) * N * 8);
}

// work: fftlog_general_work_bytes(); the inputs are copied from the host here, the result stays at *out_dev
cudaError_t launch_fftlog_general(int nmu, const double* mu, int npk, int N, const double* k, const double* f, double q,
                                  int sph, double deriv, double plaw, char* work, double** out_dev,
                                  cudaStream_t stream) {
  FGProblem P;
  char* p = work;
  double* d_mu = reinterpret_cast<double*>(p); p += up256((size_t)nmu * 8);
  double* d_k = reinterpret_cast<double*>(p); p += up256((size_t)N * 8);
  double* d_f = reinterpret_cast<double*>(p); p += up256((size_t)npk * N * 8);
  P.a = reinterpret_cast<double*>(p); p += up256((size_t)npk * N * 8);
  P.w = reinterpret_cast<double2*>(p); p += up256((size_t)N * 16);
  const size_t nh = (size_t)N / 2 + 1;
  P.u = reinterpret_cast<double2*>(p); p += up256((size_t)nmu * nh * 16);
  P.B = reinterpret_cast<double2*>(p); p += up256((size_t)npk * nh * 16);
  P.kcrc = reinterpret_cast<double*>(p); p += up256((size_t)nmu * 8);
  P.out = reinterpret_cast<double*>(p);
  cudaMemcpyAsync(d_mu, mu, (size_t)nmu * 8, cudaMemcpyHostToDevice, stream);
  cudaMemcpyAsync(d_k, k, (size_t)N * 8, cudaMemcpyHostToDevice, stream);
  cudaMemcpyAsync(d_f, f, (size_t)npk * N * 8, cudaMemcpyHostToDevice, stream);
  P.nmu = nmu; P.npk = npk; P.N = N; P.sph = sph;
  P.q = q - 1.0 * sph; P.deriv = deriv; P.plaw = plaw;
  P.L = log(k[N - 1] / k[0]) * N / (N - 1.);
  P.mu = d_mu; P.k = d_k; P.f = d_f;
  fg_setup_kernel<<<nmu, 128, 0, stream>>>(P);
  const long long wf = (long long)npk * (long long)nh, wb = (long long)nmu * npk * N;
  fg_forward_kernel<<<(unsigned)((wf * 32 + 255) / 256), 256, 0, stream>>>(P);
  fg_backward_kernel<<<(unsigned)((wb * 32 + 255) / 256), 256, 0, stream>>>(P);
  *out_dev = P.out;
  return cudaGetLastError();
}

cudaError_t launch_correlation(const CorrProblem& P, cudaStream_t stream) {
  if (P.n_theta == 0) return cudaSuccess;
  char* p = P.work;
  double4* cr = reinterpret_cast<double4*>(p); p += up256((size_t)P.n_ell * sizeof(double4));
  double* taper = reinterpret_cast<double*>(p); p += 256;
  if (P.do_taper) cudaMemcpyAsync(taper, P.taper, 4 * sizeof(double), cudaMemcpyHostToDevice, stream);
  corr_akima_kernel<<<(P.n_ell + 127) / 128, 128, 0, stream>>>(P.n_ell, P.ell, P.cls, cr);
  CorrF1D f;
  f.x = P.ell; f.cr = cr; f.n = P.n_ell; f.hi_loglog = 1;
  f.y0 = P.y0; f.yf = 0.0; f.x_ini = P.x_ini; f.x_end = P.x_end; f.y_end = P.y_end; f.der_hi = P.der_hi;
  if (P.method == CORR_FFTLOG) {
    const int N = P.n_ell_corr;
    double* a = reinterpret_cast<double*>(p); p += up256((size_t)N * 8);
    double* th = reinterpret_cast<double*>(p); p += up256((size_t)N * 8);
    double* wth = reinterpret_cast<double*>(p); p += up256((size_t)N * 8);
    double2* u = reinterpret_cast<double2*>(p); p += up256((size_t)N * 16);
    double2* w = reinterpret_cast<double2*>(p); p += up256((size_t)N * 16);
    double2* b = reinterpret_cast<double2*>(p); p += up256((size_t)N * 16);
    double4* ocr = reinterpret_cast<double4*>(p); p += up256((size_t)N * 32);
    double* scal = reinterpret_cast<double*>(p);
    const double mu = P.corr_type == CORR_GL ? 2.0 : (P.corr_type == CORR_LM ? 4.0 : 0.0);
    corr_eval_kernel<<<(N + 127) / 128, 128, 0, stream>>>(f, N, P.l_arr, a, P.do_taper, taper, 1);
    fftlog_setup_kernel<<<(N + 127) / 128, 128, 0, stream>>>(N, mu, P.L, u, w, scal);
    fftlog_forward_kernel<<<N / 2 + 1, FT, 0, stream>>>(N, a, w, u, b);
    fftlog_backward_kernel<<<N, FT, 0, stream>>>(N, b, w, scal, P.k0, P.L, th, wth);
    corr_akima_kernel<<<(N + 127) / 128, 128, 0, stream>>>(N, th, wth, ocr);
    corr_output_kernel<<<(P.n_theta + 127) / 128, 128, 0, stream>>>(N, th, wth, ocr, P.n_theta, P.theta, P.wtheta);
  } else if (P.method == CORR_BESSEL) {
    qagw_upload_tables();
    const int nblk = std::min((P.n_theta + BW - 1) / BW, 148 * 8);
    corr_bessel_kernel<<<nblk, BW * 32, 0, stream>>>(f, P, reinterpret_cast<double*>(p));
  } else {
    const int lmax = (int)P.ell_max_corr;
    const size_t tb = up256(((size_t)lmax + 1) * 8);
    double* cl_arr = reinterpret_cast<double*>(p);
    double* ra = reinterpret_cast<double*>(p + tb);
    double* rb = reinterpret_cast<double*>(p + 2 * tb);
    double* clw = reinterpret_cast<double*>(p + 3 * tb);
    corr_eval_kernel<<<(lmax + 1 + 127) / 128, 128, 0, stream>>>(f, lmax + 1, nullptr, cl_arr, P.do_taper, taper, 0);
    corr_legendre_tables<<<(lmax + 1 + 127) / 128, 128, 0, stream>>>(lmax, P.corr_type, cl_arr, ra, rb, clw);
    // neighbouring thetas share a warp: their table reads coalesce into broadcasts
    corr_legendre_kernel<<<(P.n_theta + 31) / 32, 32, 0, stream>>>(lmax, ra, rb, clw, P);
  }
  return cudaGetLastError();
}

}  // namespace cclb
