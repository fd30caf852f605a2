// Device table descriptors for the B200 Limber C_ell path.
//
// Every reference object the hot path reads is flattened into
// plain-old-data descriptors holding device pointers into one HBM arena:
//   ccl_f1d_t / gsl akima spline      -> Spline1D   (include/ccl_f1d.h:25-34)
//   gsl cspline (factorised f2d)      -> Spline1D   (src/ccl_f2d.c:152,161)
//   ccl_f2d_t                          -> F2D        (include/ccl_f2d.h:29-44)
//   ccl_cl_tracer_t                    -> Tracer     (include/ccl_tracers.h:15-22)
//   ccl_cosmology (a(chi), growth)     -> Cosmo      (include/ccl_core.h:260)
// Spline polynomial coefficients are precomputed per interval as one 48-byte record
// {x_i, x_{i+1}, y_i, b_i, c_i, d_i} so that an evaluation is one index guess plus three
// fp64x2 loads; the bicubic P(k,a) keeps {z, zx, zy, zxy} interleaved per node (32 bytes).
#pragma once
#include <cuda_runtime.h>

namespace cclb {

// GSL / CCL status codes reproduced on the device (include/ccl_error.h:10-48, gsl_errno.h)
enum : int {
  ST_OK = 0,
  ST_GSL_EDOM = 1,
  ST_GSL_EINVAL = 4,
  ST_GSL_EFAILED = 5,
  ST_GSL_EMAXITER = 11,
  ST_GSL_EBADTOL = 13,
  ST_GSL_EROUND = 18,
  ST_GSL_ESING = 21,
  ST_GSL_EDIVERGE = 22,
  ST_ERR_MEMORY = 1025,
  ST_ERR_INCONSISTENT = 1027,
  ST_ERR_SPLINE = 1028,
  ST_ERR_SPLINE_EV = 1029,
  ST_ERR_INTEG = 1030,
  ST_ERR_COMPUTECHI = 1033,
  ST_ERR_NOT_IMPLEMENTED = 1040,
  ST_ERR_GROWTH_INIT = 1058,
  ST_ERR_DISTANCES_INIT = 1059,
};

enum : int { GROWTH_CCL = 401, GROWTH_CONST = 403, GROWTH_NONE = 404 };  // include/ccl_f2d.h:13-18
enum : int { INTEG_QAG = 500, INTEG_SPLINE = 501 };                      // include/ccl_utils.h:12-15

// One interval of a piecewise-cubic 1-D interpolant: p(x) = y + dx (b + dx (c + dx d)),
// dx = x - x0, valid for x0 <= x < x1 (GSL's interval convention; last interval closed).
// Stored as two parallel arrays: the interval's end points {x0, x1} (16 B, one LDG.128) and its
// coefficients {y, b, c, d} (32 B on a 32-byte boundary, one LDG.256).  An index guess is verified
// against the interval's own end points without touching a node array, and splines on identical
// grids (the lensing kernels of all source bins, the bias transfers of all lens bins) can share one
// interval search and then read only their 32-byte coefficient records (limber_fast.cu).
struct Rec1 {
  double x0, x1, y, b, c, d;
};

// One cell edge of a bicubic axis: node, next node, 1 / spacing.
struct __align__(32) AxisRec {
  double x0, x1, inv_dx, pad;
};

// Index guess for a strictly increasing grid of n nodes:
//   uniform != 0 : i = (x - x0) * inv_step        (grid is uniform to within +-1 interval)
//   uniform == 0 : i = lut[(x - x0) * inv_bin]    (nlut uniform bins -> interval of left edge)
// followed by a +-1 walk on the records until x0 <= x < x1, which reproduces
// gsl_interp_bsearch exactly for any grid.
struct Spline1D {
  const double2* xr;  // [n-1] {x_i, x_{i+1}}
  const double4* cr;  // [n-1] {y_i, b_i, c_i, d_i}
  const int* lut;   // [nlut] (unused when uniform)
  int n, nlut, uniform, pad;
  double x0, xn;     // x[0], x[n-1]
  double inv_step;   // (n-1)/(xn-x0)
  double inv_bin;    // nlut/(xn-x0)
};

struct Axis {
  const AxisRec* rec;  // [n-1]
  const int* lut;
  int n, nlut, uniform, pad;
  double x0, xn, inv_step, inv_bin;
};

struct F2D {
  int is_factorizable, is_k_constant, is_a_constant, is_log;
  int extrap_order_lok, extrap_order_hik, extrap_linear_growth, growth_exponent;
  int has_fk, has_fa, has_fka, pad0;
  double growth_factor_0;
  double lkmin, lkmax, amin, amax;
  Spline1D fk;  // natural cspline in ln k (factorised form)
  Spline1D fa;  // natural cspline in a   (factorised form)
  // bicubic form: tab[ia*nk + ik] = {z, dz/dlnk, dz/da, d2z/dlnk da} (GSL zx, zy, zxy)
  const double4* tab;
  Axis gk;  // ln k axis
  Axis ga;  // a axis
};

struct Tracer {
  int der_bessel, der_angles, has_kernel, has_transfer;
  double chi_min, chi_max;
  // an a-only transfer whose samples are all equal (a constant galaxy bias) evaluates to exactly that
  // constant through its natural cubic spline: the fast kernel folds it into the ell prefactor
  int const_transfer, pad0;
  double const_transfer_value;
  Spline1D kernel;  // Akima W(chi); value 0 at and outside the end nodes (src/ccl_f1d.c:137-161)
  F2D transfer;
};

struct Cosmo {
  Spline1D achi;    // Akima a(chi)
  Spline1D growth;  // Akima D(a), normalised to D(1)=1
  int has_achi, has_growth;
  const F2D* psp;          // P(k,a) used for this cosmology
  const Tracer* tracers;   // all sub-tracers of this cosmology
  int ntracers, pad0;
};

// Parameters the path reads from ccl_spline_params / ccl_gsl_params
struct LimberParams {
  double K_MAX, K_MIN, DLOGK_INTEGRATION, LIMBER_EPSREL;
  int N_ITERATION;
  int nk_cap;  // capacity (nodes) of the per-warp integrand buffer
  int force_cquad;  // tests: send every 'qag_quad' C_ell through the CQUAD fallback (CCL_B200_FORCE_CQUAD)
};

// One launch: cl[(ic*npair + ip)*nl + il]
struct LimberProblem {
  const Cosmo* cosmos;
  const int2* colls;  // [ncoll] {first sub-tracer, count}
  const int2* pairs;  // [npair] {collection 1, collection 2}
  const double* ells;
  int ncosmo, npair, nl, ncoll;
  LimberParams par;
  double* cl;
  int* status;         // [ncosmo*npair] 0 or CCL_ERROR_INTEG
  int* status_detail;  // [ncosmo*npair] first underlying code
  unsigned long long* counters;  // [0] integrand evaluations (qag only)
};

// ---- Limber covariance (covariance.cu): ccl_f3d_t and one ccl_angular_cl_covariance call ----
// ccl_f3d_t (include/ccl_f3d.h): product form = two (ln k, a) tables, planes[0] and planes[1]; otherwise na bicubic
// (ln k1, ln k2) planes, linearly interpolated in a (src/ccl_f3d.c:273-392)
struct F3D {
  int na, is_product, extrap_order_lok, extrap_order_hik;
  int extrap_linear_growth, is_log, growth_exponent, pad;
  double growth_factor_0, lkmin, lkmax;
  const double* a_arr;  // [na]
  const F2D* planes;
};
struct CovProblem {
  const Cosmo* cosmo;
  const Tracer* tracers;  // the four collections, concatenated
  int2 coll[4];           // {first, count} of trc1..trc4
  const F3D* tsp;
  const double* l1;
  const double* l2;
  int nl1, nl2, chipow, has_ker;
  Spline1D ker;           // kernel_extra: Akima f(a), constant extrapolation (pyccl/ccl_covs.i:95-118)
  double ker_y0, ker_yf;
  double chimin, chimax, prefactor;
  double* cov;            // [nl1 * nl2], cov[l1 + nl1 * l2]
  int* status;            // [2]: 0 / CCL_ERROR_INTEG, first underlying code
};
size_t covariance_scratch_bytes(int nchi, int n_iteration);
cudaError_t launch_covariance(const CovProblem& prob, int method, int nchi, int n_iteration, double epsrel,
                              double* scratch, cudaStream_t stream);

// ---- C_ell -> xi(theta): ccl_correlation (src/ccl_correlation.c:14-440), SURVEY 8f-4 ----
enum CorrKind : int {
  CORR_LGNDRE = 1001, CORR_FFTLOG = 1002, CORR_BESSEL = 1003,      // include/ccl_correlation.h:6-8
  CORR_GG = 2001, CORR_GL = 2002, CORR_LP = 2003, CORR_LM = 2004,  // :9-12
};
struct CorrProblem {
  // inputs on the device
  const double* ell;      // [n_ell]
  const double* cls;      // [n_ell]
  const double* theta;    // [n_theta], degrees
  const double* l_arr;    // [n_ell_corr] ccl_log_spacing(ELL_MIN_CORR, ELL_MAX_CORR, N_ELL_CORR) (FFTLog only)
  int n_ell, n_theta, n_ell_corr, corr_type, method, do_taper, n_iteration, pad;
  double taper[4];
  double ell_max_corr, epsrel;
  double L;               // log(l_arr[N-1] / l_arr[0]) * N / (N - 1)  (src/ccl_fftlog.c:206)
  double der_hi;          // logx_logy slope of the C_ell spline above its last node (src/ccl_f1d.c:97-102)
  double x_ini, x_end, y0, y_end;  // ell[0], ell[n-1], cls[0], cls[n-1] (host copies of the spline's end values)
  double k0;              // l_arr[0]
  // outputs / work space on the device
  double* wtheta;         // [n_theta]
  int* status;            // [1]: OR of the GSL codes of the 'bessel' integrals
  char* work;             // correlation_work_bytes()
};
size_t correlation_work_bytes(int method, int n_ell, int n_theta, int n_ell_corr, double ell_max_corr, int n_iteration);
cudaError_t launch_correlation(const CorrProblem& prob, cudaStream_t stream);
// generalised FFTLog (ccl_fftlog_ComputeXi_general, src/ccl_fftlog.c:330-503, 525-535), batched over mu
size_t fftlog_general_work_bytes(int nmu, int npk, int N);
cudaError_t launch_fftlog_general(int nmu, const double* mu, int npk, int N, const double* k, const double* f, double q,
                                  int sph, double deriv, double plaw, char* work, double** out_dev, cudaStream_t stream);

// ---- fast 'spline' path: ln k node grids shared by the pairs of one cosmology ----
// Two pairs of one cosmology share a node grid for every ell when their (chi_min, chi_max) from
// get_k_interval (src/ccl_cls.c:64-100) coincide; a(chi), k P(k,a) and each collection's
// Delta(node) are then evaluated once per grid instead of once per pair.
struct GridDesc {
  double chi_min, chi_max;
  int coll_begin, ncoll;  // into gcolls (collection indices used by this grid's pairs)
  int pair_begin, npair;  // into gpairs
  int sub_begin, nsub;    // into gsubs: the task's flattened sub-tracer list, ordered by spline grid
  int has_rsd, pad;       // some sub-tracer of the task has der_bessel 1 or 2
};
struct GridPair {
  int out;     // pair index in the caller's list
  short a, b;  // slots in the grid's collection list
};
struct FastPlan {
  GridDesc* grids;  // [ncosmo][gstride]
  GridPair* gpairs; // [ncosmo][gstride]
  int* gcolls;      // [ncosmo][cstride]
  int* ngrids;      // [ncosmo]
  unsigned* gsubs;  // [ncosmo][sstride] sub-tracer entries of every grid task (limber_fast.cu)
  int gstride, cstride, sstride;
  int ntr_max;  // largest sub-tracer count of any cosmology of the launch (sizes shared memory)
  int has_rsd;  // some sub-tracer has der_bessel 1 or 2: launch the kernel variant that handles them
  int regroup;  // 0: the grid tasks in this plan are still valid (same tables, collections and pairs): skip the grouping pass
  unsigned long long* redo;  // [0] = count, [1..redo_cap] = task ids left to the general kernel
  unsigned long long redo_cap;
};

// ---- table-preparation jobs (device kernels fill the derived regions of the arena) ----
struct AkimaJob { const double* x; const double* y; double2* xr; double4* cr; int n; };
struct CsplineJob { const double* x; const double* y; double2* xr; double4* cr; double* scratch; int n; };
struct LutJob { const double* x; int* lut; int n; int nlut; };
struct AxisJob { const double* x; AxisRec* rec; int n; };
struct BicubicJob {
  const double* xk;  // [nk]
  const double* ya;  // [na]
  const double* z;   // [na*nk]
  double4* tab;      // [na*nk]
  double* scratch;   // [4*(nk+na)] alpha, gamma, 1/h, 1/alpha of both axes (shared by jobs on the same grids)
  int nk, na;
  int factor;        // this job computes the axis factors (the first job of each distinct grid pair)
};

// launchers (kernels.cu)
void launch_prep(const AkimaJob* d_ak, int nak, const CsplineJob* d_cs, int ncs, const LutJob* d_lut,
                 int nlut, const AxisJob* d_ax, int nax, const BicubicJob* d_bc, int nbc, int max_nk,
                 int max_na, int max_ncs, cudaStream_t stream);
cudaError_t launch_limber(const LimberProblem& prob, int method, double* qag_ws, size_t qag_ws_bytes,
                          cudaStream_t stream);
size_t limber_qag_workspace_bytes(int n_iteration);
// grid-sharing 'qag_quad' (limber_qag_fast.cu): its workspace follows the general kernel's in the same buffer
size_t limber_qag_fast_workspace_bytes(int n_iteration);
void limber_qag_cq_list(double* qag_ws, int n_iteration, unsigned long long** list, unsigned long long* cap,
                        cudaStream_t stream);
void launch_limber_qag_list(const LimberProblem& prob, double* qag_ws, size_t qag_ws_bytes,
                            const unsigned long long* list, unsigned long long cap, cudaStream_t stream);
// fast path (limber_fast.cu): eligibility limits, plan sizing, launch (group + fast + redo kernels)
constexpr int FAST_MAX_TRACERS = 48;  // sub-tracers per cosmology
constexpr int FAST_MAX_COLLS = 48;    // collections per call
constexpr int FAST_MAX_PAIRS = 2048;  // pairs per call
size_t limber_fast_plan_bytes(int ncosmo, int npair, unsigned long long redo_cap);
FastPlan limber_fast_plan_carve(void* base, int ncosmo, int npair, unsigned long long redo_cap);
cudaError_t launch_limber_fast(const LimberProblem& prob, const FastPlan& plan, cudaStream_t stream);
void launch_limber_group(const LimberProblem& prob, const FastPlan& plan, cudaStream_t stream);
cudaError_t launch_limber_qag_fast(const LimberProblem& prob, const FastPlan& plan, double* ws, size_t ws_bytes,
                                   cudaStream_t stream);
// the general one-warp-per-C_ell kernel over the plan's redo list (kernels.cu)
void launch_limber_redo(const LimberProblem& prob, const FastPlan& plan, cudaStream_t stream);
int limber_nk_cap(const LimberParams& par);
double measure_fp64_peak_tflops(cudaStream_t stream);
// vectorised accessors (ccl_scale_factor_of_chis, pk2d_eval_multi, cl_tracer_get_kernel/_transfer)
void launch_eval_achi(const Cosmo* d_cs, int n, const double* d_chi, double* d_out, int* d_status,
                      cudaStream_t stream);
void launch_eval_f2d(const F2D* d_f, const Cosmo* d_cs, int n, const double* d_lk, const double* d_a,
                     double* d_out, int* d_status, cudaStream_t stream, int deriv = 0);  // deriv: d ln f / d ln k
void launch_eval_tracer_kernel(const Tracer* d_tr, int n, const double* d_chi, double* d_out,
                               cudaStream_t stream);
void launch_eval_tracer_transfer(const Tracer* d_tr, int n, const double* d_lk, const double* d_a,
                                 double* d_out, int* d_status, cudaStream_t stream);

// device table builders (builders.cu)
void launch_build_background(int ncosmo, int na, const double* d_a, const double* d_params, double dchi,
                             double epsrel, int nmax, double a_init, int stride, double* d_E, double* d_chi,
                             int* d_n, double* d_achi_chi, double* d_achi_a, double* d_growth, double* d_fgrowth,
                             double* d_growth0, cudaStream_t stream);

void launch_build_linpower(int ncosmo, int nk, const double* d_k, int na_pk, const double* d_apk, int na_bg,
                           const double* d_abg, const double* d_growth, const double* d_params, int model,
                           double lg_kmin, double lg_kmax, double* d_logp, double* d_sigma8, cudaStream_t stream);

void launch_build_bg_transfer(int ncosmo, int na, const double* d_abg, const double* d_D, const double* d_f,
                              const double* d_params, int kind, int n, const double* d_a_ka, const double* d_amp,
                              double* d_out, cudaStream_t stream);
void launch_build_kappa_kernel(int ncosmo, int na, const double* d_abg, const double* d_chibg, const int* d_nachi,
                               int stride, const double* d_achi_chi, const double* d_achi_a, const double* d_params,
                               double z_source, int n, double* d_chi, double* d_w, cudaStream_t stream);
void launch_build_tracer_kernels(int ncosmo, int na, const double* d_abg, const double* d_chibg, const int* d_nachi,
                                 int stride, const double* d_achi_chi, const double* d_achi_a, const double* d_params,
                                 int nz, const double* d_z, int ntr, const double* d_nz, const double* d_sz,
                                 const double* d_inv_norm, const double* d_lens_scale, int nchi, double* d_chi_z, double* d_w_nc, double* d_chi_l,
                                 double* d_w_l, cudaStream_t stream);

void launch_build_halofit(int ncosmo, int nk, const double* d_lk, int na, const double* d_a, const double* d_logplin,
                          const double* d_params, double* d_logpnl, double* d_rsigma, int* d_fail,
                          cudaStream_t stream);

void launch_tracer_ranges(int ncosmo, int ntr, int n, const double* d_x, int x_per_tracer, const double* d_w,
                          double* d_out, cudaStream_t stream);
void launch_chi_at(int ncosmo, int na, const double* d_abg, const double* d_chi, double a_eval, double* d_out,
                   cudaStream_t stream);
void launch_log(int n, const double* d_x, double* d_out, cudaStream_t stream);

void launch_scale_rows(int ncosmo, int ntr, int n, double* d_w, const double* d_scale, cudaStream_t stream);

}  // namespace cclb
