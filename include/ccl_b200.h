/*
 * ccl_b200.h -- C ABI of the B200-native Limber angular power spectrum path.
 *
 * Drop-in boundary for ccl_angular_cls_limber (reference: /root/reference/src/ccl_cls.c:265-363,
 * declared include/ccl_cls.h:20-26, bound by pyccl/ccl_cls.i:98-112) and for the object
 * constructors that sit on the same boundary.  Plain pointers and sizes only; every entry
 * takes `int *status` with the reference's convention (pyccl/ccl.i:34): in/out, 0 = OK,
 * CCL_ERROR_* / GSL codes otherwise, and a call is a no-op when *status != 0 on entry.
 *
 * There is NO CPU fallback: every compute entry runs hand-written sm_100a kernels and fails
 * with CCL_B200_ERROR_CUDA when no usable GPU is present.
 *
 * Ownership: input arrays are borrowed for the duration of the call and copied into the
 * object's HBM arena; objects are freed by their *_free function.  Thread-safety matches the
 * reference: objects are immutable after construction; concurrent calls on distinct objects
 * are safe; one object must not be freed while a call uses it.
 */
#ifndef CCL_B200_H
#define CCL_B200_H

#ifdef __cplusplus
extern "C" {
#endif

/* status codes shared with the reference (include/ccl_error.h:10-48) */
#define CCL_B200_ERROR_MEMORY 1025
#define CCL_B200_ERROR_LINSPACE 1026
#define CCL_B200_ERROR_INCONSISTENT 1027
#define CCL_B200_ERROR_SPLINE 1028
#define CCL_B200_ERROR_SPLINE_EV 1029
#define CCL_B200_ERROR_INTEG 1030
#define CCL_B200_ERROR_NOT_IMPLEMENTED 1040
#define CCL_B200_ERROR_GROWTH_INIT 1058
#define CCL_B200_ERROR_DISTANCES_INIT 1059
/* library-specific: CUDA runtime failure / no device / extension unusable */
#define CCL_B200_ERROR_CUDA 2001

/* ccl_integration_t (include/ccl_utils.h:12-15) */
#define CCL_B200_INTEGRATION_QAG_QUAD 500
#define CCL_B200_INTEGRATION_SPLINE 501
/* ccl_f2d_extrap_growth_t (include/ccl_f2d.h:13-18) */
#define CCL_B200_F2D_CCLGROWTH 401
#define CCL_B200_F2D_CONSTANTGROWTH 403
#define CCL_B200_F2D_NO_EXTRAPOL 404

typedef struct ccl_b200_cosmology ccl_b200_cosmology;    /* ccl_cosmology, as seen by the path */
typedef struct ccl_b200_f2d ccl_b200_f2d;                /* ccl_f2d_t */
typedef struct ccl_b200_f3d ccl_b200_f3d;                /* ccl_f3d_t (Tk3D) */
typedef struct ccl_b200_tracer ccl_b200_tracer;          /* ccl_cl_tracer_t */
typedef struct ccl_b200_collection ccl_b200_collection;  /* ccl_cl_tracer_collection_t */
typedef struct ccl_b200_batch ccl_b200_batch;            /* N cosmologies' table packs in one arena */

/* The spline/gsl parameters the path reads (include/ccl_core.h:94-178; defaults
 * src/ccl_core.c:67-141).  Snapshotted by value at cosmology creation like the reference
 * (src/ccl_core.c:247-248). */
typedef struct {
  double K_MAX;                     /* 1e3   */
  double K_MIN;                     /* 5e-5  */
  double DLOGK_INTEGRATION;         /* 0.025 */
  double INTEGRATION_LIMBER_EPSREL; /* 1e-4  */
  int N_ITERATION;                  /* 1000  */
  /* GSL key of the Gauss-Kronrod rule of 'qag_quad' (src/ccl_cls.c:240; gsl_integration.h: 1..6 = GK15, 21,
   * 31, 41, 51, 61).  Default GSL_INTEG_GAUSS41 = 4 (src/ccl_core.c:40,71), the only rule built on the
   * device: any other value makes a 'qag_quad' call fail with CCL_B200_ERROR_NOT_IMPLEMENTED instead of
   * silently integrating with GK41. */
  int INTEGRATION_LIMBER_GAUSS_KRONROD_POINTS;
  double A_SPLINE_MIN;              /* 0.1: chi_max of kernel-less tracers = chi(A_SPLINE_MIN) */
  double DCHI_INTEGRATION;          /* 5.0: chi sampling of the covariance 'spline' method (src/ccl_cls.c:418) */
  /* read by ccl_b200_correlation (src/ccl_correlation.c:54-57,243-246; defaults src/ccl_core.c:69-70,129-131) */
  double ELL_MIN_CORR;              /* 0.01  */
  double ELL_MAX_CORR;              /* 60000 */
  int N_ELL_CORR;                   /* 5000  */
  int INTEGRATION_GAUSS_KRONROD_POINTS; /* GSL key of the 'bessel' method's QAG rule; 4 = GK41, the only one built */
  double INTEGRATION_EPSREL;        /* 1e-4  */
} ccl_b200_params;

/* ---- library ---- */
/* Select the CUDA device used by subsequently created objects (per calling thread). */
void ccl_b200_set_device(int device, int *status);
int ccl_b200_device_count(void);
/* Human-readable message of the last failure on this thread (ccl_cosmology.status_message). */
const char *ccl_b200_last_error(void);
const char *ccl_b200_version(void);
void ccl_b200_default_params(ccl_b200_params *p);
/* Measured fp64 FMA throughput of the current device in TFLOP/s (DFMA microbenchmark, best of
 * 5): the roofline denominator of the Limber kernels (MEASURED_PEAKS.json has no fp64 figure). */
double ccl_b200_measure_fp64_peak(int *status);

/* ---- cosmology: background tables (src/ccl_background.c) ---- */
ccl_b200_cosmology *ccl_b200_cosmology_new(const ccl_b200_params *params, int *status);
void ccl_b200_cosmology_free(ccl_b200_cosmology *cosmo);
/* Install the a(chi) Akima table that ccl_cosmology_compute_distances leaves in
 * cosmo->data.achi (src/ccl_background.c:521-571): chi increasing, chi[0] = 0. */
void ccl_b200_cosmology_set_achi(ccl_b200_cosmology *cosmo, int n, const double *chi, const double *a,
                                 int *status);
/* ccl_cosmology_distances_from_input (src/ccl_background.c:595-680): a increasing, chi(a)
 * decreasing; the a(chi) spline is built on the reversed arrays.  Also sets A_SPLINE_MIN = a[0]. */
void ccl_b200_cosmology_distances_from_input(ccl_b200_cosmology *cosmo, int na, const double *a,
                                             const double *chi_a, int *status);
/* Growth factor table D(a) normalised to D(1) = 1 (cosmo->data.growth,
 * src/ccl_background.c:763-1002 / ccl_cosmology_growth_from_input :686-756).  Only read when a
 * P(k) with ccl_f2d_cclgrowth is evaluated below its amin (src/ccl_f2d.c:351-370). */
void ccl_b200_cosmology_set_growth(ccl_b200_cosmology *cosmo, int n, const double *a, const double *growth,
                                   int *status);
/* chi(A_SPLINE_MIN) used as chi_max of kernel-less tracers (src/ccl_tracers.c:632). */
void ccl_b200_cosmology_set_chi_max_nokernel(ccl_b200_cosmology *cosmo, double chi);
/* ccl_scale_factor_of_chis (src/ccl_background.c:1279-1288), evaluated on the device. */
void ccl_b200_scale_factor_of_chis(ccl_b200_cosmology *cosmo, int n, const double *chi, double *a_out,
                                   int *status);

/* ---- ccl_f2d_t (include/ccl_f2d.h:65-89, src/ccl_f2d.c:92-201) ---- */
/* Same argument order and NULL conventions as ccl_f2d_t_new; interp_type is always bicubic. */
ccl_b200_f2d *ccl_b200_f2d_new(int na, const double *a_arr, int nk, const double *lk_arr,
                               const double *fka_arr, const double *fk_arr, const double *fa_arr,
                               int is_factorizable, int extrap_order_lok, int extrap_order_hik,
                               int extrap_linear_growth, int is_fka_log, double growth_factor_0,
                               int growth_exponent, int *status);
/* set_pk2d_new_from_arrays (pyccl/ccl_pk2d.i:17-28) */
ccl_b200_f2d *ccl_b200_pk2d_new_from_arrays(const double *lk_arr, int nk, const double *a_arr, int na,
                                            const double *pk_arr, int order_lok, int order_hik,
                                            int is_logp, int *status);
void ccl_b200_f2d_free(ccl_b200_f2d *f2d);
/* pk2d_eval_multi (pyccl/ccl_pk2d.i:55-62): out[i] = ccl_f2d_t_eval(f2d, lk[i], a, cosmo).
 * cosmo may be NULL when no growth extrapolation is needed. */
void ccl_b200_f2d_eval_multi(ccl_b200_f2d *f2d, const double *lk, int nk, double a, ccl_b200_cosmology *cosmo,
                             double *out, int *status);
/* pk2d_der_eval_multi (pyccl/ccl_pk2d.i:58-64) -> ccl_f2d_t_dlogf_dlk_eval (src/ccl_f2d.c:375-531): d ln f / d ln k */
void ccl_b200_f2d_dlogf_dlk_eval_multi(ccl_b200_f2d *f2d, const double *lk, int nk, double a,
                                       ccl_b200_cosmology *cosmo, double *out, int *status);

/* ---- ccl_cl_tracer_t (include/ccl_tracers.h:46-59, src/ccl_tracers.c:569-691) ---- */
ccl_b200_tracer *ccl_b200_cl_tracer_t_new(ccl_b200_cosmology *cosmo, int der_bessel, int der_angles,
                                          int n_w, const double *chi_w, const double *w_w, int na_ka,
                                          const double *a_ka, int nk_ka, const double *lk_ka,
                                          const double *fka_arr, const double *fk_arr,
                                          const double *fa_arr, int is_fka_log, int is_factorizable,
                                          int extrap_order_lok, int extrap_order_hik, int *status);
void ccl_b200_cl_tracer_t_free(ccl_b200_tracer *tr);
double ccl_b200_cl_tracer_t_chi_min(const ccl_b200_tracer *tr);
double ccl_b200_cl_tracer_t_chi_max(const ccl_b200_tracer *tr);
/* ccl_cl_tracer_t_get_f_ell / _get_kernel / _get_transfer (src/ccl_tracers.c:703-749), vectorised
 * like pyccl/ccl_tracers.i cl_tracer_get_*; kernel and transfer run on the device. */
void ccl_b200_cl_tracer_get_f_ell(const ccl_b200_tracer *tr, int n, const double *ell, double *out,
                                  int *status);
void ccl_b200_cl_tracer_get_kernel(ccl_b200_tracer *tr, int n, const double *chi, double *out, int *status);
void ccl_b200_cl_tracer_get_transfer(ccl_b200_tracer *tr, int nk, const double *lk, int na, const double *a,
                                     double *out /* [nk*na] */, int *status);

/* ---- ccl_cl_tracer_collection_t (src/ccl_tracers.c:13-50) ---- */
ccl_b200_collection *ccl_b200_cl_tracer_collection_t_new(int *status);
void ccl_b200_cl_tracer_collection_t_free(ccl_b200_collection *trc);
void ccl_b200_add_cl_tracer_to_collection(ccl_b200_collection *trc, ccl_b200_tracer *tr, int *status);

/* ---- the hot path ---- */
/* ccl_angular_cls_limber (include/ccl_cls.h:20-26): same argument order and semantics.  On an
 * integration failure cl_out[l] = NaN and *status = CCL_ERROR_INTEG (src/ccl_cls.c:343-345);
 * without distances *status = CCL_ERROR_DISTANCES_INIT (:274-280). */
void ccl_b200_angular_cls_limber(ccl_b200_cosmology *cosmo, ccl_b200_collection *trc1,
                                 ccl_b200_collection *trc2, ccl_b200_f2d *psp, int nl_out,
                                 const double *l_out, double *cl_out, int integration_method,
                                 int *status);

/* ---- batched entry: (ncosmo, npair, nl) in one launch ---- */
/* ---- Limber covariance ----
 * ccl_f3d_t_new (include/ccl_f3d.h, src/ccl_f3d.c:161-271) minus interp_type (bicubic only): tkka_arr[na][nk][nk]
 * (index [ia][ik2][ik1]) or, in product form, fka1_arr / fka2_arr [na][nk].  pyccl binds it through
 * tk3d_new_from_arrays / tk3d_new_factorizable (pyccl/ccl_tk3d.i:17-44: constant growth, growth_factor_0 = 1,
 * exponent 4). */
ccl_b200_f3d *ccl_b200_f3d_new(int na, const double *a_arr, int nk, const double *lk_arr, const double *tkka_arr,
                               const double *fka1_arr, const double *fka2_arr, int is_product,
                               int extrap_order_lok, int extrap_order_hik, int extrap_linear_growth,
                               int is_tkka_log, double growth_factor_0, int growth_exponent, int *status);
void ccl_b200_f3d_free(ccl_b200_f3d *f3d);
/* ccl_angular_cl_covariance (include/ccl_cls.h, src/ccl_cls.c:490-612), argument for argument; kernel_extra (a
 * ccl_f1d_t in the reference) is passed as the (a, f) arrays angular_cov_ssc_vec builds it from
 * (pyccl/ccl_covs.i:95-118: Akima, constant extrapolation), n_ker = 0 for none (angular_cov_vec, :53-70).
 * cov_out[l1 + nl1 * l2]; NaN + CCL_ERROR_INTEG on failure. */
void ccl_b200_angular_cl_covariance(ccl_b200_cosmology *cosmo, ccl_b200_collection *trc1,
                                    ccl_b200_collection *trc2, ccl_b200_collection *trc3,
                                    ccl_b200_collection *trc4, ccl_b200_f3d *tsp, int nl1_out,
                                    const double *l1_out, int nl2_out, const double *l2_out, double *cov_out,
                                    int integration_method, int chi_exponent, int n_ker, const double *a_ker,
                                    const double *f_ker, double prefactor_extra, int *status);

/* ---- C_ell -> xi(theta) ----
 * ccl_correlation (include/ccl_correlation.h:14-44, src/ccl_correlation.c:410-428), argument for argument:
 * flag_method CCL_CORR_LGNDRE / _FFTLOG / _BESSEL = 1001 / 1002 / 1003, corr_type CCL_CORR_GG / _GL / _LP / _LM =
 * 2001..2004, theta in degrees.  pyccl binds it through correlation_vec (pyccl/ccl_correlation.i:51-64:
 * do_taper_cl = 0, taper_cl_limits = NULL).  cosmo supplies the spline / gsl parameters only. */
void ccl_b200_correlation(ccl_b200_cosmology *cosmo, int n_ell, const double *ell, const double *cls, int n_theta,
                          const double *theta, double *wtheta, int corr_type, int do_taper_cl,
                          const double *taper_cl_limits, int flag_method, int *status);

/* Generalised FFTLog: fftlog_transform_general (pyccl/ccl_fftlog.i:63-90) -> ccl_fftlog_ComputeXi_general
 * (include/ccl_fftlog.h, src/ccl_fftlog.c:330-503,532-535), with one addition: a leading batch over mu, because
 * its caller (the FKEM non-Limber integrator, pyccl/nonlimber/_nonlimber_FKEM.py:140-151) transforms the same
 * arrays for every multipole.  For each mu: output[imu][0][:] = k, output[imu][1 + j][:] = transform of
 * fk_in[j][:]; nmu = 1 is the reference's call. */
void ccl_b200_fftlog_transform_general(int nmu, const double *mu, int npk, const double *k_in, int n_in_k,
                                       const double *fk_in, double q, int spherical_bessel, double bessel_deriv,
                                       double plaw, double *output, int *status);

/* A batch owns one HBM arena holding the table packs of ncosmo cosmologies.  Slots are filled
 * with the same flat arrays as the single-object constructors, then committed (one H2D copy of
 * the raw region + device kernels that build every spline coefficient table in place). */
ccl_b200_batch *ccl_b200_batch_new(int ncosmo, const ccl_b200_params *params, int *status);
void ccl_b200_batch_free(ccl_b200_batch *b);
void ccl_b200_batch_set_achi(ccl_b200_batch *b, int icosmo, int n, const double *chi, const double *a,
                             int *status);
void ccl_b200_batch_set_growth(ccl_b200_batch *b, int icosmo, int n, const double *a, const double *growth,
                               int *status);
void ccl_b200_batch_set_chi_max_nokernel(ccl_b200_batch *b, int icosmo, double chi);
void ccl_b200_batch_set_pk2d(ccl_b200_batch *b, int icosmo, const double *lk_arr, int nk,
                             const double *a_arr, int na, const double *pk_arr, int order_lok,
                             int order_hik, int is_logp, int *status);
/* Returns the index of the new sub-tracer inside cosmology icosmo (-1 on failure). */
int ccl_b200_batch_add_tracer(ccl_b200_batch *b, int icosmo, int der_bessel, int der_angles, int n_w,
                              const double *chi_w, const double *w_w, int na_ka, const double *a_ka,
                              int nk_ka, const double *lk_ka, const double *fka_arr, const double *fk_arr,
                              const double *fa_arr, int is_fka_log, int is_factorizable,
                              int extrap_order_lok, int extrap_order_hik, int *status);
/* Stacked setters for likelihood batches: one call fills slots [first, first+count).  Value arrays
 * carry a leading cosmology axis ([count][n] row-major); abscissa arrays have a stride between
 * consecutive cosmologies, 0 meaning one grid shared by all.  Same per-slot semantics as the
 * single-slot setters above (and as ccl_cl_tracer_t_new / set_pk2d_new_from_arrays); the copies into
 * the pinned staging arena run on several host threads.  A stacked array that lives in page-locked
 * host memory (cudaHostAlloc / cudaHostRegister) is not staged: ccl_b200_batch_commit copies it with
 * one DMA straight from the caller's buffer, which must then stay valid and unchanged until commit
 * returns (pageable arrays are only borrowed for the call, like every other input). */
void ccl_b200_batch_set_achi_stacked(ccl_b200_batch *b, int first, int count, const int *n_per /* NULL: stride */,
                                     int stride, const double *chi, const double *a, int *status);
void ccl_b200_batch_set_growth_stacked(ccl_b200_batch *b, int first, int count, int n, int stride_a,
                                       const double *a, const double *growth, int *status);
void ccl_b200_batch_set_chi_max_nokernel_stacked(ccl_b200_batch *b, int first, int count, const double *chi);
void ccl_b200_batch_set_pk2d_stacked(ccl_b200_batch *b, int first, int count, const double *lk_arr, int nk,
                                     const double *a_arr, int na, const double *pk_arr /* [count][na*nk] */,
                                     int order_lok, int order_hik, int is_logp, int *status);
int ccl_b200_batch_add_tracer_stacked(ccl_b200_batch *b, int first, int count, int der_bessel, int der_angles,
                                      int n_w, int stride_chi, const double *chi_w, const double *w_w,
                                      int na_ka, int stride_a, const double *a_ka, int nk_ka,
                                      const double *lk_ka, const double *fka_arr, const double *fk_arr,
                                      const double *fa_arr, int is_fka_log, int is_factorizable,
                                      int extrap_order_lok, int extrap_order_hik, int *status);
void ccl_b200_batch_tracer_chi_range(const ccl_b200_batch *b, int icosmo, int itracer, double *chi_min,
                                     double *chi_max);
/* Forget all slots but keep the arena (re-use for the next set of cosmologies). */
void ccl_b200_batch_reset(ccl_b200_batch *b);
void ccl_b200_batch_commit(ccl_b200_batch *b, int *status);
/* Bytes uploaded by the last commit, and bytes of HBM the arena holds. */
unsigned long long ccl_b200_batch_h2d_bytes(const ccl_b200_batch *b);
unsigned long long ccl_b200_batch_arena_bytes(const ccl_b200_batch *b);

/* Collections are [coll_start[i], coll_start[i]+coll_count[i]) ranges of sub-tracer indices,
 * identical for every cosmology of the batch; pair p correlates collections pair1[p], pair2[p].
 * cl_out[(icosmo*npair + p)*nl + l]; status_out[icosmo*npair + p] is 0 or CCL_ERROR_INTEG.
 * *status is set to CCL_ERROR_INTEG if any entry failed. */
void ccl_b200_batch_angular_cls_limber(ccl_b200_batch *b, int ncoll, const int *coll_start,
                                       const int *coll_count, int npair, const int *pair1,
                                       const int *pair2, int nl, const double *ell, int integration_method,
                                       double *cl_out, int *status_out, int *status);
/* The same call in two halves, so that a host thread can queue the next chunk of a pipelined batch while this
 * one is still running: _start enqueues the Limber kernels and the device-to-host copy of every C_ell into
 * cl_out (which should be page-locked and must stay valid) on the batch's own stream and returns; _wait blocks
 * until they have finished and delivers the per-(cosmology, pair) status words like the synchronous entry. */
void ccl_b200_batch_angular_cls_limber_start(ccl_b200_batch *b, int ncoll, const int *coll_start,
                                             const int *coll_count, int npair, const int *pair1,
                                             const int *pair2, int nl, const double *ell, int integration_method,
                                             double *cl_out, int *status);
void ccl_b200_batch_angular_cls_limber_wait(ccl_b200_batch *b, int *status_out, int *status);
/* Same, asynchronous on `stream` (a cudaStream_t, may be NULL) with DEVICE output buffers;
 * d_status_out holds 2*ncosmo*npair ints (status, then first underlying code). */
void ccl_b200_batch_angular_cls_limber_dev(ccl_b200_batch *b, int ncoll, const int *coll_start,
                                           const int *coll_count, int npair, const int *pair1,
                                           const int *pair2, int nl, const double *ell,
                                           int integration_method, double *d_cl_out, int *d_status_out,
                                           void *stream, int *status);
/* Number of ln k integrand nodes the 'spline' method evaluates for this problem
 * (sum over cosmology, pair, ell of nk; src/ccl_cls.c:194-195) -- the unit of the roofline
 * accounting.  Computed on the host from the chi ranges; qag counts come from the device. */
unsigned long long ccl_b200_batch_spline_nodes(const ccl_b200_batch *b, int ncoll, const int *coll_start,
                                               const int *coll_count, int npair, const int *pair1,
                                               const int *pair2, int nl, const double *ell);
/* Integrand evaluations of the last qag_quad launch on this batch. */
unsigned long long ccl_b200_batch_last_qag_evals(const ccl_b200_batch *b);
/* Integrand abscissae the last 'qag_quad' launch actually evaluated: smaller than last_qag_evals when the
 * grid-sharing kernel served panels from its cache (pairs with equal chi ranges request the same panels). */
unsigned long long ccl_b200_batch_last_qag_nodes_computed(const ccl_b200_batch *b);
/* 1 if the last 'spline' launch ran the grid-sharing kernels (pairs of one cosmology whose
 * get_k_interval ranges coincide share a(chi), k P(k,a) and tracer evaluations; same node values
 * and integrand arithmetic as src/ccl_cls.c:166-225), 0 if it ran the general kernel.  Setting the
 * environment variable CCL_B200_NO_FAST forces the general kernel. */
int ccl_b200_batch_last_used_fast_path(const ccl_b200_batch *b);

/* ---- device builders of the tables the path reads (batched over cosmologies) ---- */
/* One cosmology = CCL_B200_NPAR doubles, derived like pyccl/cosmology.py:398-496 (m_nu = 0):
 * [0] Omega_c [1] Omega_b [2] Omega_l [3] Omega_k [4] Omega_g [5] Omega_nu_rel [6] w0 [7] wa [8] h
 * [9] Omega_m [10] n_s [11] sigma8 [12] T_CMB [13] k_sign [14] sqrtk [15] reserved */
#define CCL_B200_NPAR 16
/* ccl_cosmology_compute_distances + ccl_cosmology_compute_growth (src/ccl_background.c:417-589,
 * 763-1002) for ncosmo cosmologies on the device.  a_grid[na] is the shared linlog a-grid ending at
 * a = 1.  Host outputs, row-major with a leading cosmology axis: E[na] = H/H0, chi[na] (Mpc), the
 * a(chi) table on a uniform chi grid of step <= dchi (n_achi[i] nodes in rows of `stride`; a negative
 * count means stride was too small), growth[na] (D, D(1) = 1), fgrowth[na], growth0 (D unnormalised
 * at a = 1).  root_epsrel / root_niter are ccl_gsl_params.ROOT_EPSREL / ROOT_N_ITERATION (the Newton
 * stopping rule decides the a(chi) node values), a_growth_init is EPS_SCALEFAC_GROWTH. */
void ccl_b200_build_background(int ncosmo, const double *params, int na, const double *a_grid, double dchi,
                               double root_epsrel, int root_niter, double a_growth_init, int stride,
                               double *E, double *chi, int *n_achi, double *achi_chi, double *achi_a,
                               double *growth, double *fgrowth, double *growth0, int *status);

/* ccl_compute_linpower_analytic (src/ccl_power.c:28-146) on the device for ncosmo cosmologies:
 * log P(k, a) = log(k^ns T(k)^2) + 2 log D(a) + 2 log(sigma8 / sigma8_unnormalised) on the shared grids
 * k_grid[nk] (1/Mpc) x a_pk[na_pk].  model: 0 = BBKS (src/ccl_bbks.c), 1 = Eisenstein-Hu,
 * 2 = Eisenstein-Hu without wiggles (src/ccl_eh.c).  growth[ncosmo][na_bg] is D(a) on a_bg[na_bg]
 * (e.g. from ccl_b200_build_background), read through its Akima spline like ccl_growth_factor.
 * sigma8 of the unnormalised table is integrated over [k_min, k_max] like ccl_sigmaR
 * (src/ccl_power.c:390-454).  Host outputs: logp[ncosmo][na_pk][nk], sigma8_unnorm[ncosmo]. */
void ccl_b200_build_linear_power(int ncosmo, const double *params, int model, int nk, const double *k_grid,
                                 int na_pk, const double *a_pk, int na_bg, const double *a_bg,
                                 const double *growth, double k_min, double k_max, double *logp,
                                 double *sigma8_unnorm, int *status);

/* Radial kernels of n(z)-defined tracers for ncosmo cosmologies on the device.  Shared by all
 * cosmologies: the redshift grid z[nz], ntr distributions nz_arr[ntr][nz], optional magnification
 * bias sz_arr[ntr][nz] on the same grid (NULL: none) and inv_norm[ntr] = 1 / integral of n(z)
 * (get_nz_norm, src/ccl_tracers.c:60-119; cosmology-independent).  Per cosmology: params, the chi(a)
 * table chi_bg[ncosmo][na] on a_bg[na] and the a(chi) table (n_achi, stride, achi_chi, achi_a), e.g.
 * from ccl_b200_build_background.  Host outputs: chi_z[ncosmo][nz] (chi at the n(z) nodes = abscissa
 * of the number-counts kernels), w_nc[ncosmo][ntr][nz] = p(z) H(z) (ccl_get_number_counts_kernel,
 * src/ccl_tracers.c:121-160), chi_l[ncosmo][nchi] and w_l[ncosmo][ntr][nchi]: the lensing kernel
 * nodes (ccl_get_chis_lensing_kernel, :470-479) and values by Akima-spline integration with the
 * trapezoid edge term (integrate_lensing_kernel_spline, :347-454).  nchi <= 256. */
void ccl_b200_build_tracer_kernels(int ncosmo, const double *params, int na, const double *a_bg,
                                   const double *chi_bg, const int *n_achi, int stride, const double *achi_chi,
                                   const double *achi_a, int nz, const double *z, int ntr, const double *nz_arr,
                                   const double *sz_arr, const double *inv_norm, int nchi, double *chi_z,
                                   double *w_nc, double *chi_l, double *w_l, int *status);

/* ccl_apply_halofit (src/ccl_power.c:171-232; src/ccl_halofit.c) on the device: log P_NL on the grid of
 * the linear table logplin[ncosmo][na][nk] (natural log of P, ln k grid lk[nk], scale factors a[na]).
 * Host outputs: logpnl[ncosmo][na][nk] and the non-linear scales rsigma[ncosmo][na] (Mpc).  *status is
 * CCL_B200_ERROR_INTEG (and the row NaN) where sigma(R) = 1 has no root in [1e-64, 1e4] Mpc. */
void ccl_b200_build_halofit(int ncosmo, const double *params, int nk, const double *lk, int na, const double *a,
                            const double *logplin, double *logpnl, double *rsigma, int *status);

/* ---- device-resident tables: parameters -> tables -> batch without a host round trip ---- */
typedef struct ccl_b200_tables ccl_b200_tables;
/* Runs the four builders above for ncosmo cosmologies and keeps every table in HBM.  Arguments as in the
 * host-in/out entries; halofit != 0 also builds the non-linear P(k); nz_arr / sz_arr / inv_norm / lens_scale
 * describe ntr redshift distributions (lens_scale[ntr], may be NULL, multiplies the lensing kernel; 0 = not built; e.g.
 * -2 for a magnification term).  The object must outlive ccl_b200_batch_commit of every batch filled
 * from it (the preparation kernels read its tables in place).  The halofit pass is still running on its own stream
 * when the call returns -- the caller's batch planning overlaps it; ccl_b200_batch_set_from_tables orders the batch
 * behind it and the following ccl_b200_batch_commit reports CCL_B200_ERROR_INTEG if sigma(R) = 1 had no root. */
ccl_b200_tables *ccl_b200_tables_build(int ncosmo, const double *params, int na, const double *a_grid,
                                       double dchi, double root_epsrel, int root_niter, double a_growth_init,
                                       int achi_stride, double a_spline_min, int model, int halofit, int nk,
                                       const double *k_grid, const double *lk_grid /* ln k_grid */, int na_pk,
                                       const double *a_pk, double k_min, double k_max, int nz, const double *z, int ntr, const double *nz_arr,
                                       const double *sz_arr, const double *inv_norm, const double *lens_scale,
                                       int nchi, int *status);
void ccl_b200_tables_free(ccl_b200_tables *t);
/* Fill batch slots [first, first + ncosmo) with the a(chi) and growth tables, chi(A_SPLINE_MIN) and the
 * P(k,a) of the tables object (non-linear if built, unless use_linear != 0). */
void ccl_b200_batch_set_from_tables(ccl_b200_batch *b, int first, const ccl_b200_tables *t, int order_lok,
                                    int order_hik, int use_linear, int *status);
/* Add one sub-tracer per slot whose radial kernel is a device-built one: kind 0 = number-counts kernel of
 * distribution itr (abscissa chi(z)), kind 1 = its lensing kernel.  Optional a-only transfer shared by all
 * cosmologies (a_ka[na_ka], fa_arr[na_ka]; NULL = none), e.g. a galaxy bias b(a).  Returns the index of
 * the new sub-tracer (same in every slot) or -1. */
int ccl_b200_batch_add_tracer_from_tables(ccl_b200_batch *b, int first, const ccl_b200_tables *t, int kind,
                                          int itr, int der_bessel, int der_angles, int na_ka, const double *a_ka,
                                          const double *fa_arr, int *status);

/* Same with an a-only transfer that depends on the cosmology through its growth tables, built on the device
 * on the shared nodes a_ka[na_ka]: bg_kind 1 = redshift-space distortions, -f(a) (pyccl/tracers.py:895-899, use
 * with der_bessel = 2); bg_kind 2 = intrinsic alignments, -amp[j] Omega_m / D(a_ka[j]) with
 * amp[j] = A_IA(z_j) 5e-14 rho_crit (pyccl/tracers.py:967-987, use with der_bessel = -1, der_angles = 2). */
int ccl_b200_batch_add_tracer_from_tables_bg(ccl_b200_batch *b, int first, ccl_b200_tables *t, int kind, int itr,
                                             int der_bessel, int der_angles, int bg_kind, int na_ka,
                                             const double *a_ka, const double *amp, int *status);
/* One CMB-lensing sub-tracer per slot (pyccl CMBLensingTracer, pyccl/tracers.py:991-1014): the kappa kernel
 * of ccl_get_kappa_kernel (src/ccl_tracers.c:554-567) on linspace(0, chi(z_source), n_samples) is built on
 * the device from the object's chi(a) and a(chi) tables and kept by the object; der_bessel = -1,
 * der_angles = 1.  Returns the index of the new sub-tracer or -1. */
int ccl_b200_batch_add_kappa_tracer_from_tables(ccl_b200_batch *b, int first, ccl_b200_tables *t, double z_source,
                                                int n_samples, int *status);

#ifdef __cplusplus
}
#endif
#endif
