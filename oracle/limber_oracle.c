/*
 * oracle/limber_oracle.c -- TEST INFRASTRUCTURE ONLY. NOT PART OF THE PRODUCT.
 *
 * CPU restatement (kind: "port") of the reference's Limber angular-power-spectrum path,
 * used only by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs as the checker.  The product (ccl_b200/) never links, imports or
 * executes anything in this directory.
 *
 * PARITY STATUS: the reference's arithmetic for this path lives in GSL (>= 2.1, fallback
 * tarball 2.4: /root/reference/CMakeLists.txt:118, cmake/Modules/BuildGSL.cmake:3), which is
 * not vendored under /root/reference and is not installed in the build container, so the
 * reference itself is unbuildable here.  The GSL algorithms (interpolation/akima.c,
 * cspline.c, linalg/tridiag.c, interpolation/bsearch.h, integ_eval.h, interp2d/bicubic.c,
 * integration/qag.c, qk.c, qk41.c, qpsrt.c, util.c, cquad.c with its constant tables
 * regenerated from their definitions by tools/gen_cquad.py) are restated from their published
 * sources.  Pins: scipy Akima1DInterpolator / CubicSpline(natural) (tests/test_oracle.py),
 * the reference's 23 golden C_ell benchmark curves (benchmarks/data/run_*_log_cl_*.txt via
 * tests/golden/), the closed-form power-law C_ell of benchmarks/test_tracers.py:6-18.
 * No reference fixture pins this path at 1e-10 ("parity unpinned" at that level; see
 * DESIGN.md).
 *
 * Reference files followed (all under /root/reference):
 *   src/ccl_cls.c:64-363        get_k_interval, transfer_limber_*, cl_integrand,
 *                               integ_cls_limber_spline / _qag_quad, ccl_angular_cls_limber
 *   src/ccl_tracers.c:569-749   ccl_cl_tracer_t_new (chi range rule), get_f_ell/kernel/transfer
 *   src/ccl_f1d.c:15-191        ccl_f1d_t_new / eval (constant extrapolation branch)
 *   src/ccl_f2d.c:92-373        ccl_f2d_t_new / eval
 *   src/ccl_utils.c:17-37       ccl_linear_spacing
 *   src/ccl_utils.c:274-345     ccl_integ_spline
 *   src/ccl_background.c:1251-1325  ccl_scale_factor_of_chi, ccl_growth_factor
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <float.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "gk41_table.h"

#define O_SUCCESS 0
#define O_GSL_EDOM 1
#define O_GSL_EINVAL 4
#define O_GSL_EFAILED 5
#define O_GSL_EMAXITER 11
#define O_GSL_EBADTOL 13
#define O_GSL_EROUND 18
#define O_GSL_ESING 21
#define O_ERR_MEMORY 1025
#define O_ERR_INCONSISTENT 1027
#define O_ERR_SPLINE 1028
#define O_ERR_SPLINE_EV 1029
#define O_ERR_INTEG 1030
#define O_ERR_COMPUTECHI 1033
#define O_ERR_NOT_IMPLEMENTED 1040
#define O_ERR_GROWTH_INIT 1058
#define O_ERR_DISTANCES_INIT 1059
#define O_F2D_CCLGROWTH 401
#define O_F2D_CONSTGROWTH 403
#define O_F2D_NOEXTRAPOL 404
#define O_INTEG_QAG 500
#define O_INTEG_SPLINE 501

/* ------------------------------------------------------------------ */
/* GSL interpolation/bsearch.h: gsl_interp_bsearch                     */
/* ------------------------------------------------------------------ */
static size_t o_bsearch(const double *xa, double x, size_t ilo, size_t ihi) {
  while (ihi > ilo + 1) {
    size_t i = (ihi + ilo) / 2;
    if (xa[i] > x) ihi = i; else ilo = i;
  }
  return ilo;
}

/* ------------------------------------------------------------------ */
/* GSL interpolation/akima.c (non-periodic)                            */
/* ------------------------------------------------------------------ */
typedef struct {
  int n;
  double *x, *y, *b, *c, *d;
} o_akima;

void o_akima_free(o_akima *s) {
  if (!s) return;
  free(s->x); free(s->y); free(s->b); free(s->c); free(s->d); free(s);
}

/* gsl_spline_alloc(gsl_interp_akima, n) returns NULL for n < 5 (min_size) */
o_akima *o_akima_new(int n, const double *x, const double *y) {
  if (n < 5) return NULL;
  o_akima *s = calloc(1, sizeof(o_akima));
  s->n = n;
  s->x = malloc(sizeof(double) * n);
  s->y = malloc(sizeof(double) * n);
  s->b = malloc(sizeof(double) * n);
  s->c = malloc(sizeof(double) * n);
  s->d = malloc(sizeof(double) * n);
  memcpy(s->x, x, sizeof(double) * n);
  memcpy(s->y, y, sizeof(double) * n);
  double *mm = malloc(sizeof(double) * (n + 4));
  double *m = mm + 2;
  for (int i = 0; i <= n - 2; i++)
    m[i] = (y[i + 1] - y[i]) / (x[i + 1] - x[i]);
  m[-2] = 3.0 * m[0] - 2.0 * m[1];
  m[-1] = 2.0 * m[0] - m[1];
  m[n - 1] = 2.0 * m[n - 2] - m[n - 3];
  m[n] = 3.0 * m[n - 2] - 2.0 * m[n - 3];
  for (int i = 0; i < n - 1; i++) {
    const double NE = fabs(m[i + 1] - m[i]) + fabs(m[i - 1] - m[i - 2]);
    if (NE == 0.0) {
      s->b[i] = m[i]; s->c[i] = 0.0; s->d[i] = 0.0;
    } else {
      const double h_i = x[i + 1] - x[i];
      const double NE_next = fabs(m[i + 2] - m[i + 1]) + fabs(m[i] - m[i - 1]);
      const double alpha_i = fabs(m[i - 1] - m[i - 2]) / NE;
      double alpha_ip1, tL_ip1;
      if (NE_next == 0.0) {
        tL_ip1 = m[i];
      } else {
        alpha_ip1 = fabs(m[i] - m[i - 1]) / NE_next;
        tL_ip1 = (1.0 - alpha_ip1) * m[i] + alpha_ip1 * m[i + 1];
      }
      s->b[i] = (1.0 - alpha_i) * m[i - 1] + alpha_i * m[i];
      s->c[i] = (3.0 * m[i] - 2.0 * s->b[i] - tL_ip1) / h_i;
      s->d[i] = (s->b[i] + tL_ip1 - 2.0 * m[i]) / (h_i * h_i);
    }
  }
  free(mm);
  return s;
}

/* gsl_spline_eval_e: EDOM + NaN outside [xmin, xmax] */
int o_akima_eval_e(const o_akima *s, double x, double *y) {
  if (x < s->x[0] || x > s->x[s->n - 1]) { *y = NAN; return O_GSL_EDOM; }
  size_t i = o_bsearch(s->x, x, 0, s->n - 1);
  const double delx = x - s->x[i];
  *y = s->y[i] + delx * (s->b[i] + delx * (s->c[i] + s->d[i] * delx));
  return O_SUCCESS;
}

/* GSL interpolation/integ_eval.h */
static double o_integ_eval(double ai, double bi, double ci, double di, double xi,
                           double a, double b) {
  const double r1 = a - xi;
  const double r2 = b - xi;
  const double r12 = r1 + r2;
  const double bterm = 0.5 * bi * r12;
  const double cterm = (1.0 / 3.0) * ci * (r1 * r1 + r2 * r2 + r1 * r2);
  const double dterm = 0.25 * di * r12 * (r1 * r1 + r2 * r2);
  return (b - a) * (ai + bterm + cterm + dterm);
}

/* akima_eval_integ behind gsl_spline_eval_integ_e */
int o_akima_integ_e(const o_akima *s, double a, double b, double *result) {
  if (a > b || a < s->x[0] || b > s->x[s->n - 1]) { *result = NAN; return O_GSL_EDOM; }
  if (a == b) { *result = 0.0; return O_SUCCESS; }
  size_t index_a = o_bsearch(s->x, a, 0, s->n - 1);
  size_t index_b = o_bsearch(s->x, b, 0, s->n - 1);
  double res = 0.0;
  for (size_t i = index_a; i <= index_b; i++) {
    const double x_hi = s->x[i + 1];
    const double x_lo = s->x[i];
    const double y_lo = s->y[i];
    const double dx = x_hi - x_lo;
    if (dx != 0.0) {
      if (i == index_a || i == index_b) {
        double x1 = (i == index_a) ? a : x_lo;
        double x2 = (i == index_b) ? b : x_hi;
        res += o_integ_eval(y_lo, s->b[i], s->c[i], s->d[i], x_lo, x1, x2);
      } else {
        res += dx * (y_lo + dx * (0.5 * s->b[i] + dx * (s->c[i] / 3.0 + 0.25 * s->d[i] * dx)));
      }
    } else {
      *result = 0.0;
      return O_GSL_EINVAL;
    }
  }
  *result = res;
  return O_SUCCESS;
}

/* ------------------------------------------------------------------ */
/* GSL interpolation/cspline.c (natural) + linalg/tridiag.c            */
/* ------------------------------------------------------------------ */
typedef struct {
  int n;
  double *x, *y, *c;
} o_cspline;

void o_cspline_free(o_cspline *s) {
  if (!s) return;
  free(s->x); free(s->y); free(s->c); free(s);
}

static void o_solve_symm_tridiag(const double *diag, const double *offdiag,
                                 const double *b, double *x, size_t N) {
  double *gamma = malloc(N * sizeof(double));
  double *alpha = malloc(N * sizeof(double));
  double *c = malloc(N * sizeof(double));
  double *z = malloc(N * sizeof(double));
  size_t i, j;
  alpha[0] = diag[0];
  gamma[0] = offdiag[0] / alpha[0];
  for (i = 1; i < N - 1; i++) {
    alpha[i] = diag[i] - offdiag[i - 1] * gamma[i - 1];
    gamma[i] = offdiag[i] / alpha[i];
  }
  if (N > 1) alpha[N - 1] = diag[N - 1] - offdiag[N - 2] * gamma[N - 2];
  z[0] = b[0];
  for (i = 1; i < N; i++) z[i] = b[i] - gamma[i - 1] * z[i - 1];
  for (i = 0; i < N; i++) c[i] = z[i] / alpha[i];
  x[N - 1] = c[N - 1];
  if (N >= 2) {
    for (i = N - 2, j = 0; j <= N - 2; j++, i--) x[i] = c[i] - gamma[i] * x[i + 1];
  }
  free(gamma); free(alpha); free(c); free(z);
}

/* gsl_interp_cspline min_size is 3 */
o_cspline *o_cspline_new(int n, const double *xa, const double *ya) {
  if (n < 3) return NULL;
  o_cspline *s = calloc(1, sizeof(o_cspline));
  s->n = n;
  s->x = malloc(sizeof(double) * n);
  s->y = malloc(sizeof(double) * n);
  s->c = malloc(sizeof(double) * n);
  memcpy(s->x, xa, sizeof(double) * n);
  memcpy(s->y, ya, sizeof(double) * n);
  size_t max_index = n - 1;
  size_t sys_size = max_index - 1;
  double *g = malloc(sizeof(double) * n);
  double *diag = malloc(sizeof(double) * n);
  double *offdiag = malloc(sizeof(double) * n);
  s->c[0] = 0.0;
  s->c[max_index] = 0.0;
  for (size_t i = 0; i < sys_size; i++) {
    const double h_i = xa[i + 1] - xa[i];
    const double h_ip1 = xa[i + 2] - xa[i + 1];
    const double ydiff_i = ya[i + 1] - ya[i];
    const double ydiff_ip1 = ya[i + 2] - ya[i + 1];
    const double g_i = (h_i != 0.0) ? 1.0 / h_i : 0.0;
    const double g_ip1 = (h_ip1 != 0.0) ? 1.0 / h_ip1 : 0.0;
    offdiag[i] = h_ip1;
    diag[i] = 2.0 * (h_ip1 + h_i);
    g[i] = 3.0 * (ydiff_ip1 * g_ip1 - ydiff_i * g_i);
  }
  if (sys_size == 1) {
    s->c[1] = g[0] / diag[0];
  } else {
    o_solve_symm_tridiag(diag, offdiag, g, s->c + 1, sys_size);
  }
  free(g); free(diag); free(offdiag);
  return s;
}

static void o_cs_coeff(const double *c_array, double dy, double dx, size_t index,
                       double *b, double *c, double *d) {
  const double c_i = c_array[index];
  const double c_ip1 = c_array[index + 1];
  *b = (dy / dx) - dx * (c_ip1 + 2.0 * c_i) / 3.0;
  *c = c_i;
  *d = (c_ip1 - c_i) / (3.0 * dx);
}

/* which: 0 value, 1 first derivative, 2 second derivative */
int o_cspline_eval_e(const o_cspline *s, double x, int which, double *out) {
  if (x < s->x[0] || x > s->x[s->n - 1]) { *out = NAN; return O_GSL_EDOM; }
  size_t i = o_bsearch(s->x, x, 0, s->n - 1);
  const double x_hi = s->x[i + 1], x_lo = s->x[i];
  const double dx = x_hi - x_lo;
  if (dx > 0.0) {
    const double y_lo = s->y[i], y_hi = s->y[i + 1];
    const double dy = y_hi - y_lo;
    double delx = x - x_lo;
    double b_i, c_i, d_i;
    o_cs_coeff(s->c, dy, dx, i, &b_i, &c_i, &d_i);
    if (which == 0) *out = y_lo + delx * (b_i + delx * (c_i + delx * d_i));
    else if (which == 1) *out = b_i + delx * (2.0 * c_i + 3.0 * d_i * delx);
    else *out = 2.0 * c_i + 6.0 * d_i * delx;
    return O_SUCCESS;
  }
  *out = 0.0;
  return O_GSL_EINVAL;
}

/* ------------------------------------------------------------------ */
/* GSL interp2d/bicubic.c                                              */
/* ------------------------------------------------------------------ */
typedef struct {
  int nx, ny;
  double *x, *y, *z, *zx, *zy, *zxy;
} o_bicubic;

void o_bicubic_free(o_bicubic *s) {
  if (!s) return;
  free(s->x); free(s->y); free(s->z); free(s->zx); free(s->zy); free(s->zxy); free(s);
}

#define IDX2D(i, j, nx) ((j) * (nx) + (i))

/* z[j*nx + i] = f(x_i, y_j) (gsl_spline2d layout; CCL passes x=ln k, y=a, src/ccl_f2d.c:194) */
o_bicubic *o_bicubic_new(int nx, int ny, const double *xa, const double *ya, const double *za) {
  if (nx < 4 || ny < 4) return NULL; /* gsl_interp2d_bicubic min_size 4 */
  o_bicubic *s = calloc(1, sizeof(o_bicubic));
  s->nx = nx; s->ny = ny;
  s->x = malloc(sizeof(double) * nx);
  s->y = malloc(sizeof(double) * ny);
  size_t nn = (size_t)nx * ny;
  s->z = malloc(sizeof(double) * nn);
  s->zx = malloc(sizeof(double) * nn);
  s->zy = malloc(sizeof(double) * nn);
  s->zxy = malloc(sizeof(double) * nn);
  memcpy(s->x, xa, sizeof(double) * nx);
  memcpy(s->y, ya, sizeof(double) * ny);
  memcpy(s->z, za, sizeof(double) * nn);
  double *tx = malloc(sizeof(double) * (nx > ny ? nx : ny));
  /* zx: natural cspline along x for every y row, first derivative at the nodes */
  for (int j = 0; j < ny; j++) {
    for (int i = 0; i < nx; i++) tx[i] = za[IDX2D(i, j, nx)];
    o_cspline *sp = o_cspline_new(nx, xa, tx);
    for (int i = 0; i < nx; i++) o_cspline_eval_e(sp, xa[i], 1, &s->zx[IDX2D(i, j, nx)]);
    o_cspline_free(sp);
  }
  /* zy: along y for every x column */
  for (int i = 0; i < nx; i++) {
    for (int j = 0; j < ny; j++) tx[j] = za[IDX2D(i, j, nx)];
    o_cspline *sp = o_cspline_new(ny, ya, tx);
    for (int j = 0; j < ny; j++) o_cspline_eval_e(sp, ya[j], 1, &s->zy[IDX2D(i, j, nx)]);
    o_cspline_free(sp);
  }
  /* zxy: x-derivative of the zy plane (spline along x of zy for every y row) */
  for (int j = 0; j < ny; j++) {
    for (int i = 0; i < nx; i++) tx[i] = s->zy[IDX2D(i, j, nx)];
    o_cspline *sp = o_cspline_new(nx, xa, tx);
    for (int i = 0; i < nx; i++) o_cspline_eval_e(sp, xa[i], 1, &s->zxy[IDX2D(i, j, nx)]);
    o_cspline_free(sp);
  }
  free(tx);
  return s;
}

/* The bicubic patch of GSL's interp2d/bicubic.c on one cell, from its 16 corner values (already scaled by the
 * cell size).  which: 0 value, 1 d/dx, 2 d2/dx2 (bicubic_eval / _deriv_x / _deriv_xx). */
static void o_bicubic_cell(double zminmin, double zminmax, double zmaxmin, double zmaxmax, double zxminmin,
                           double zxminmax, double zxmaxmin, double zxmaxmax, double zyminmin, double zyminmax,
                           double zymaxmin, double zymaxmax, double zxyminmin, double zxyminmax,
                           double zxymaxmin, double zxymaxmax, double t, double u, double dt, int which,
                           double *out) {
  const double t0 = 1, t1 = t, t2 = t * t, t3 = t * t2;
  const double u0 = 1, u1 = u, u2 = u * u, u3 = u * u2;
  double v, z = 0;
  if (which == 0) {
    v = zminmin; z += v * t0 * u0;
    v = zyminmin; z += v * t0 * u1;
    v = -3 * zminmin + 3 * zminmax - 2 * zyminmin - zyminmax; z += v * t0 * u2;
    v = 2 * zminmin - 2 * zminmax + zyminmin + zyminmax; z += v * t0 * u3;
    v = zxminmin; z += v * t1 * u0;
    v = zxyminmin; z += v * t1 * u1;
    v = -3 * zxminmin + 3 * zxminmax - 2 * zxyminmin - zxyminmax; z += v * t1 * u2;
    v = 2 * zxminmin - 2 * zxminmax + zxyminmin + zxyminmax; z += v * t1 * u3;
    v = -3 * zminmin + 3 * zmaxmin - 2 * zxminmin - zxmaxmin; z += v * t2 * u0;
    v = -3 * zyminmin + 3 * zymaxmin - 2 * zxyminmin - zxymaxmin; z += v * t2 * u1;
    v = 9 * zminmin - 9 * zmaxmin + 9 * zmaxmax - 9 * zminmax + 6 * zxminmin + 3 * zxmaxmin
        - 3 * zxmaxmax - 6 * zxminmax + 6 * zyminmin - 6 * zymaxmin - 3 * zymaxmax
        + 3 * zyminmax + 4 * zxyminmin + 2 * zxymaxmin + zxymaxmax + 2 * zxyminmax;
    z += v * t2 * u2;
    v = -6 * zminmin + 6 * zmaxmin - 6 * zmaxmax + 6 * zminmax - 4 * zxminmin - 2 * zxmaxmin
        + 2 * zxmaxmax + 4 * zxminmax - 3 * zyminmin + 3 * zymaxmin + 3 * zymaxmax
        - 3 * zyminmax - 2 * zxyminmin - zxymaxmin - zxymaxmax - 2 * zxyminmax;
    z += v * t2 * u3;
    v = 2 * zminmin - 2 * zmaxmin + zxminmin + zxmaxmin; z += v * t3 * u0;
    v = 2 * zyminmin - 2 * zymaxmin + zxyminmin + zxymaxmin; z += v * t3 * u1;
    v = -6 * zminmin + 6 * zmaxmin - 6 * zmaxmax + 6 * zminmax - 3 * zxminmin - 3 * zxmaxmin
        + 3 * zxmaxmax + 3 * zxminmax - 4 * zyminmin + 4 * zymaxmin + 2 * zymaxmax
        - 2 * zyminmax - 2 * zxyminmin - 2 * zxymaxmin - zxymaxmax - zxyminmax;
    z += v * t3 * u2;
    v = 4 * zminmin - 4 * zmaxmin + 4 * zmaxmax - 4 * zminmax + 2 * zxminmin + 2 * zxmaxmin
        - 2 * zxmaxmax - 2 * zxminmax + 2 * zyminmin - 2 * zymaxmin - 2 * zymaxmax
        + 2 * zyminmax + zxyminmin + zxymaxmin + zxymaxmax + zxyminmax;
    z += v * t3 * u3;
    *out = z;
  } else if (which == 1) {
    /* bicubic_deriv_x: d/dt of the same polynomial, times dt */
    v = zxminmin; z += v * t0 * u0;
    v = zxyminmin; z += v * t0 * u1;
    v = -3 * zxminmin + 3 * zxminmax - 2 * zxyminmin - zxyminmax; z += v * t0 * u2;
    v = 2 * zxminmin - 2 * zxminmax + zxyminmin + zxyminmax; z += v * t0 * u3;
    v = -3 * zminmin + 3 * zmaxmin - 2 * zxminmin - zxmaxmin; z += 2 * v * t1 * u0;
    v = -3 * zyminmin + 3 * zymaxmin - 2 * zxyminmin - zxymaxmin; z += 2 * v * t1 * u1;
    v = 9 * zminmin - 9 * zmaxmin + 9 * zmaxmax - 9 * zminmax + 6 * zxminmin + 3 * zxmaxmin
        - 3 * zxmaxmax - 6 * zxminmax + 6 * zyminmin - 6 * zymaxmin - 3 * zymaxmax
        + 3 * zyminmax + 4 * zxyminmin + 2 * zxymaxmin + zxymaxmax + 2 * zxyminmax;
    z += 2 * v * t1 * u2;
    v = -6 * zminmin + 6 * zmaxmin - 6 * zmaxmax + 6 * zminmax - 4 * zxminmin - 2 * zxmaxmin
        + 2 * zxmaxmax + 4 * zxminmax - 3 * zyminmin + 3 * zymaxmin + 3 * zymaxmax
        - 3 * zyminmax - 2 * zxyminmin - zxymaxmin - zxymaxmax - 2 * zxyminmax;
    z += 2 * v * t1 * u3;
    v = 2 * zminmin - 2 * zmaxmin + zxminmin + zxmaxmin; z += 3 * v * t2 * u0;
    v = 2 * zyminmin - 2 * zymaxmin + zxyminmin + zxymaxmin; z += 3 * v * t2 * u1;
    v = -6 * zminmin + 6 * zmaxmin - 6 * zmaxmax + 6 * zminmax - 3 * zxminmin - 3 * zxmaxmin
        + 3 * zxmaxmax + 3 * zxminmax - 4 * zyminmin + 4 * zymaxmin + 2 * zymaxmax
        - 2 * zyminmax - 2 * zxyminmin - 2 * zxymaxmin - zxymaxmax - zxyminmax;
    z += 3 * v * t2 * u2;
    v = 4 * zminmin - 4 * zmaxmin + 4 * zmaxmax - 4 * zminmax + 2 * zxminmin + 2 * zxmaxmin
        - 2 * zxmaxmax - 2 * zxminmax + 2 * zyminmin - 2 * zymaxmin - 2 * zymaxmax
        + 2 * zyminmax + zxyminmin + zxymaxmin + zxymaxmax + zxyminmax;
    z += 3 * v * t2 * u3;
    *out = z * dt;
  } else {
    /* bicubic_deriv_xx */
    v = -3 * zminmin + 3 * zmaxmin - 2 * zxminmin - zxmaxmin; z += 2 * v * t0 * u0;
    v = -3 * zyminmin + 3 * zymaxmin - 2 * zxyminmin - zxymaxmin; z += 2 * v * t0 * u1;
    v = 9 * zminmin - 9 * zmaxmin + 9 * zmaxmax - 9 * zminmax + 6 * zxminmin + 3 * zxmaxmin
        - 3 * zxmaxmax - 6 * zxminmax + 6 * zyminmin - 6 * zymaxmin - 3 * zymaxmax
        + 3 * zyminmax + 4 * zxyminmin + 2 * zxymaxmin + zxymaxmax + 2 * zxyminmax;
    z += 2 * v * t0 * u2;
    v = -6 * zminmin + 6 * zmaxmin - 6 * zmaxmax + 6 * zminmax - 4 * zxminmin - 2 * zxmaxmin
        + 2 * zxmaxmax + 4 * zxminmax - 3 * zyminmin + 3 * zymaxmin + 3 * zymaxmax
        - 3 * zyminmax - 2 * zxyminmin - zxymaxmin - zxymaxmax - 2 * zxyminmax;
    z += 2 * v * t0 * u3;
    v = 2 * zminmin - 2 * zmaxmin + zxminmin + zxmaxmin; z += 6 * v * t1 * u0;
    v = 2 * zyminmin - 2 * zymaxmin + zxyminmin + zxymaxmin; z += 6 * v * t1 * u1;
    v = -6 * zminmin + 6 * zmaxmin - 6 * zmaxmax + 6 * zminmax - 3 * zxminmin - 3 * zxmaxmin
        + 3 * zxmaxmax + 3 * zxminmax - 4 * zyminmin + 4 * zymaxmin + 2 * zymaxmax
        - 2 * zyminmax - 2 * zxyminmin - 2 * zxymaxmin - zxymaxmax - zxyminmax;
    z += 6 * v * t1 * u2;
    v = 4 * zminmin - 4 * zmaxmin + 4 * zmaxmax - 4 * zminmax + 2 * zxminmin + 2 * zxmaxmin
        - 2 * zxmaxmax - 2 * zxminmax + 2 * zyminmin - 2 * zymaxmin - 2 * zymaxmax
        + 2 * zyminmax + zxyminmin + zxymaxmin + zxymaxmax + zxyminmax;
    z += 6 * v * t1 * u3;
    *out = z * dt * dt;
  }
}


/* which: 0 value, 1 d/dx, 2 d2/dx2, 3 d/dy (bicubic_deriv_y: the x-derivative of the patch with the roles of
 * the two axes exchanged) */
int o_bicubic_eval_e(const o_bicubic *s, double x, double y, int which, double *out) {
  if (x < s->x[0] || x > s->x[s->nx - 1] || y < s->y[0] || y > s->y[s->ny - 1]) {
    *out = NAN;
    return O_GSL_EDOM;
  }
  const int nx = s->nx;
  size_t xi = o_bsearch(s->x, x, 0, s->nx - 1);
  size_t yi = o_bsearch(s->y, y, 0, s->ny - 1);
  const double xmin = s->x[xi], xmax = s->x[xi + 1];
  const double ymin = s->y[yi], ymax = s->y[yi + 1];
  const double zminmin = s->z[IDX2D(xi, yi, nx)];
  const double zminmax = s->z[IDX2D(xi, yi + 1, nx)];
  const double zmaxmin = s->z[IDX2D(xi + 1, yi, nx)];
  const double zmaxmax = s->z[IDX2D(xi + 1, yi + 1, nx)];
  const double dx = xmax - xmin, dy = ymax - ymin;
  const double t = (x - xmin) / dx, u = (y - ymin) / dy;
  const double dt = 1. / dx, du = 1. / dy;
  const double zxminmin = s->zx[IDX2D(xi, yi, nx)] / dt;
  const double zxminmax = s->zx[IDX2D(xi, yi + 1, nx)] / dt;
  const double zxmaxmin = s->zx[IDX2D(xi + 1, yi, nx)] / dt;
  const double zxmaxmax = s->zx[IDX2D(xi + 1, yi + 1, nx)] / dt;
  const double zyminmin = s->zy[IDX2D(xi, yi, nx)] / du;
  const double zyminmax = s->zy[IDX2D(xi, yi + 1, nx)] / du;
  const double zymaxmin = s->zy[IDX2D(xi + 1, yi, nx)] / du;
  const double zymaxmax = s->zy[IDX2D(xi + 1, yi + 1, nx)] / du;
  const double zxyminmin = s->zxy[IDX2D(xi, yi, nx)] / (dt * du);
  const double zxyminmax = s->zxy[IDX2D(xi, yi + 1, nx)] / (dt * du);
  const double zxymaxmin = s->zxy[IDX2D(xi + 1, yi, nx)] / (dt * du);
  const double zxymaxmax = s->zxy[IDX2D(xi + 1, yi + 1, nx)] / (dt * du);
  if (which == 3)
    o_bicubic_cell(zminmin, zmaxmin, zminmax, zmaxmax, zyminmin, zymaxmin, zyminmax, zymaxmax, zxminmin, zxmaxmin,
                   zxminmax, zxmaxmax, zxyminmin, zxymaxmin, zxyminmax, zxymaxmax, u, t, du, 1, out);
  else
    o_bicubic_cell(zminmin, zminmax, zmaxmin, zmaxmax, zxminmin, zxminmax, zxmaxmin, zxmaxmax, zyminmin, zyminmax,
                   zymaxmin, zymaxmax, zxyminmin, zxyminmax, zxymaxmin, zxymaxmax, t, u, dt, which, out);
  return O_SUCCESS;
}

/* ------------------------------------------------------------------ */
/* ccl_f1d_t (src/ccl_f1d.c), constant extrapolation only              */
/* ------------------------------------------------------------------ */
typedef struct {
  o_akima *spl;
  double y0, yf, x_ini, x_end;
} o_f1d;

void o_f1d_free(o_f1d *f) {
  if (!f) return;
  o_akima_free(f->spl);
  free(f);
}

o_f1d *o_f1d_new(int n, const double *x, const double *y, double y0, double yf) {
  o_akima *s = o_akima_new(n, x, y);
  if (!s) return NULL;
  o_f1d *f = calloc(1, sizeof(o_f1d));
  f->spl = s; f->y0 = y0; f->yf = yf; f->x_ini = x[0]; f->x_end = x[n - 1];
  return f;
}

/* src/ccl_f1d.c:118-191: the end nodes themselves take the extrapolation value */
double o_f1d_eval(const o_f1d *f, double x) {
  if (x <= f->x_ini) return f->y0;
  else if (x >= f->x_end) return f->yf;
  double y;
  if (o_akima_eval_e(f->spl, x, &y) != O_SUCCESS) return NAN;
  return y;
}

/* ------------------------------------------------------------------ */
/* cosmology subset: a(chi) and growth splines                         */
/* ------------------------------------------------------------------ */
typedef struct {
  o_akima *achi;    /* x = chi increasing, y = a */
  o_akima *growth;  /* x = a, y = D(a) normalised to D(1)=1; may be NULL */
  double K_MAX, K_MIN, DLOGK_INTEGRATION;
  double LIMBER_EPSREL;
  int N_ITERATION;
} o_cosmo;

o_cosmo *o_cosmo_new(int n_achi, const double *chi, const double *a,
                     int n_growth, const double *a_g, const double *d_g,
                     double k_max, double k_min, double dlogk, double epsrel, int n_iter) {
  o_cosmo *c = calloc(1, sizeof(o_cosmo));
  c->achi = (n_achi > 0) ? o_akima_new(n_achi, chi, a) : NULL;
  c->growth = (n_growth > 0) ? o_akima_new(n_growth, a_g, d_g) : NULL;
  c->K_MAX = k_max; c->K_MIN = k_min; c->DLOGK_INTEGRATION = dlogk;
  c->LIMBER_EPSREL = epsrel; c->N_ITERATION = n_iter;
  return c;
}

void o_cosmo_free(o_cosmo *c) {
  if (!c) return;
  o_akima_free(c->achi); o_akima_free(c->growth); free(c);
}

/* src/ccl_background.c:1251-1277 */
double o_scale_factor_of_chi(const o_cosmo *c, double chi, int *status) {
  if ((chi < 1.e-8) && (chi >= 0.)) return 1.;
  else if (chi < 0.) { *status = O_ERR_COMPUTECHI; return NAN; }
  if (!c->achi) { *status = O_ERR_DISTANCES_INIT; return NAN; }
  double a;
  int st = o_akima_eval_e(c->achi, chi, &a);
  if (st != O_SUCCESS) *status |= st;
  return a;
}

/* src/ccl_background.c:1290-1325 */
double o_growth_factor(const o_cosmo *c, double a, int *status) {
  if (a == 1.) return 1.;
  else if (a > 1.) { *status = O_ERR_COMPUTECHI; return NAN; }
  if (!c->growth) { *status = O_ERR_GROWTH_INIT; return NAN; }
  double D;
  int st = o_akima_eval_e(c->growth, a, &D);
  if (st != O_SUCCESS) { *status |= st; return NAN; }
  return D;
}

/* ------------------------------------------------------------------ */
/* ccl_f2d_t (src/ccl_f2d.c:92-373)                                    */
/* ------------------------------------------------------------------ */
typedef struct {
  double lkmin, lkmax, amin, amax;
  int is_factorizable, is_k_constant, is_a_constant;
  int extrap_order_lok, extrap_order_hik, extrap_linear_growth, is_log;
  double growth_factor_0;
  int growth_exponent;
  o_cspline *fk, *fa;
  o_bicubic *fka;
} o_f2d;

void o_f2d_free(o_f2d *f) {
  if (!f) return;
  o_cspline_free(f->fk); o_cspline_free(f->fa); o_bicubic_free(f->fka); free(f);
}

/* NULL array pointers mean "absent" exactly as in ccl_f2d_t_new */
o_f2d *o_f2d_new(int na, const double *a_arr, int nk, const double *lk_arr,
                 const double *fka_arr, const double *fk_arr, const double *fa_arr,
                 int is_factorizable, int extrap_order_lok, int extrap_order_hik,
                 int extrap_linear_growth, int is_fka_log, double growth_factor_0,
                 int growth_exponent, int *status) {
  o_f2d *f = calloc(1, sizeof(o_f2d));
  is_factorizable = is_factorizable || (a_arr == NULL) || (lk_arr == NULL) || (fka_arr == NULL);
  f->is_factorizable = is_factorizable;
  f->is_k_constant = ((lk_arr == NULL) || ((fka_arr == NULL) && (fk_arr == NULL)));
  f->is_a_constant = ((a_arr == NULL) || ((fka_arr == NULL) && (fa_arr == NULL)));
  f->extrap_order_lok = extrap_order_lok;
  f->extrap_order_hik = extrap_order_hik;
  f->extrap_linear_growth = extrap_linear_growth;
  f->is_log = is_fka_log;
  f->growth_factor_0 = growth_factor_0;
  f->growth_exponent = growth_exponent;
  if (!f->is_k_constant) { f->lkmin = lk_arr[0]; f->lkmax = lk_arr[nk - 1]; }
  if (!f->is_a_constant) { f->amin = a_arr[0]; f->amax = a_arr[na - 1]; }
  if ((extrap_order_lok > 2) || (extrap_order_lok < 0) || (extrap_order_hik > 2) ||
      (extrap_order_hik < 0))
    *status = O_ERR_INCONSISTENT;
  if ((extrap_linear_growth != O_F2D_CCLGROWTH) && (extrap_linear_growth != O_F2D_CONSTGROWTH) &&
      (extrap_linear_growth != O_F2D_NOEXTRAPOL))
    *status = O_ERR_INCONSISTENT;
  if (*status == 0) {
    if (f->is_factorizable) {
      if (!f->is_k_constant) {
        f->fk = o_cspline_new(nk, lk_arr, fk_arr);
        if (!f->fk) *status = O_ERR_MEMORY;
      }
      if (!f->is_a_constant) {
        f->fa = o_cspline_new(na, a_arr, fa_arr);
        if (!f->fa) *status = O_ERR_MEMORY;
      }
    } else if (!(f->is_k_constant || f->is_a_constant)) {
      f->fka = o_bicubic_new(nk, na, lk_arr, a_arr, fka_arr);
      if (!f->fka) *status = O_ERR_MEMORY;
    }
  }
  return f;
}

double o_f2d_eval(const o_f2d *f2d, double lk, double a, const o_cosmo *cosmo, int *status) {
  int is_hiz, is_loz;
  double a_ev = a;
  if (f2d->is_a_constant) { is_hiz = 0; is_loz = 0; }
  else {
    is_hiz = a < f2d->amin;
    is_loz = a > f2d->amax;
    if (is_loz) {
      if (f2d->extrap_linear_growth == O_F2D_NOEXTRAPOL) { *status = O_ERR_SPLINE_EV; return NAN; }
      a_ev = f2d->amax;
    } else if (is_hiz) {
      if (f2d->extrap_linear_growth == O_F2D_NOEXTRAPOL) { *status = O_ERR_SPLINE_EV; return NAN; }
      a_ev = f2d->amin;
    }
  }
  int is_hik, is_lok;
  double fka_pre, fka_post;
  double lk_ev = lk;
  if (f2d->is_k_constant) { is_hik = 0; is_lok = 0; }
  else {
    is_hik = lk > f2d->lkmax;
    is_lok = lk < f2d->lkmin;
    if (is_hik) lk_ev = f2d->lkmax;
    else if (is_lok) lk_ev = f2d->lkmin;
  }
  int spstatus = 0;
  if (f2d->is_factorizable) {
    double fk, fa;
    if (f2d->fk == NULL) fk = f2d->is_log ? 0 : 1;
    else spstatus |= o_cspline_eval_e(f2d->fk, lk_ev, 0, &fk);
    if (f2d->fa == NULL) fa = f2d->is_log ? 0 : 1;
    else spstatus |= o_cspline_eval_e(f2d->fa, a_ev, 0, &fa);
    fka_pre = f2d->is_log ? fk + fa : fk * fa;
  } else {
    if (f2d->fka == NULL) fka_pre = f2d->is_log ? 0 : 1;
    else spstatus = o_bicubic_eval_e(f2d->fka, lk_ev, a_ev, 0, &fka_pre);
  }
  if (spstatus) { *status = O_ERR_SPLINE_EV; return NAN; }

  if (is_hik || is_lok) {
    int order = is_hik ? f2d->extrap_order_hik : f2d->extrap_order_lok;
    fka_post = fka_pre;
    if (order > 0) {
      double pd;
      double dlk = lk - lk_ev;
      if (f2d->is_factorizable) spstatus = o_cspline_eval_e(f2d->fk, lk_ev, 1, &pd);
      else spstatus = o_bicubic_eval_e(f2d->fka, lk_ev, a_ev, 1, &pd);
      if (spstatus) { *status = O_ERR_SPLINE_EV; return NAN; }
      fka_post += pd * dlk;
      if (order > 1) {
        if (f2d->is_factorizable) spstatus = o_cspline_eval_e(f2d->fk, lk_ev, 2, &pd);
        else spstatus = o_bicubic_eval_e(f2d->fka, lk_ev, a_ev, 2, &pd);
        if (spstatus) { *status = O_ERR_SPLINE_EV; return NAN; }
        fka_post += pd * dlk * dlk * 0.5;
      }
    }
  } else
    fka_post = fka_pre;

  if (f2d->is_log) fka_post = exp(fka_post);

  if (is_hiz) {
    double gz;
    if (f2d->extrap_linear_growth == O_F2D_CCLGROWTH) {
      if (!cosmo || !cosmo->growth) { *status = O_ERR_GROWTH_INIT; return NAN; }
      gz = o_growth_factor(cosmo, a, status) / o_growth_factor(cosmo, a_ev, status);
    } else
      gz = f2d->growth_factor_0;
    fka_post *= pow(gz, f2d->growth_exponent);
  }
  return fka_post;
}

/* ccl_f2d_t_dlogf_dlk_eval (src/ccl_f2d.c:375-531): d ln f / d ln k; in k-extrapolation the slope of the Taylor
 * branch (orders > 1 add the second derivative times the distance), for non-log tables divided by f itself */
double o_f2d_dlogf_dlk_eval(const o_f2d *f2d, double lk, double a, const o_cosmo *cosmo, int *status) {
  if (f2d->is_k_constant) return 0;
  double inv_pk0 = 1.;
  if (!f2d->is_log) {
    inv_pk0 = o_f2d_eval(f2d, lk, a, cosmo, status);
    if (inv_pk0 == 0) return 0;
    inv_pk0 = 1. / inv_pk0;
  }
  int is_hiz, is_loz;
  double a_ev = a;
  if (f2d->is_a_constant) { is_hiz = 0; is_loz = 0; }
  else {
    is_hiz = a < f2d->amin;
    is_loz = a > f2d->amax;
    if (is_loz || is_hiz) {
      if (f2d->extrap_linear_growth == O_F2D_NOEXTRAPOL) { *status = O_ERR_SPLINE_EV; return NAN; }
      a_ev = is_loz ? f2d->amax : f2d->amin;
    }
  }
  double fka_pre, fka_post, lk_ev = lk;
  const int is_hik = lk > f2d->lkmax, is_lok = lk < f2d->lkmin;
  if (is_hik) lk_ev = f2d->lkmax;
  else if (is_lok) lk_ev = f2d->lkmin;
  int spstatus = 0;
  if (f2d->is_factorizable) {
    double fk, fa;
    if (f2d->fk == NULL) fk = f2d->is_log ? 0 : 1;
    else spstatus |= o_cspline_eval_e(f2d->fk, lk_ev, 1, &fk);
    if (f2d->fa == NULL) fa = f2d->is_log ? 0 : 1;
    else spstatus |= o_cspline_eval_e(f2d->fa, a_ev, 0, &fa);
    fka_pre = f2d->is_log ? fk + fa : fk * fa;
  } else {
    if (f2d->fka == NULL) fka_pre = f2d->is_log ? 0 : 1;
    else spstatus = o_bicubic_eval_e(f2d->fka, lk_ev, a_ev, 1, &fka_pre);
  }
  if (spstatus) { *status = O_ERR_SPLINE_EV; return NAN; }
  fka_post = fka_pre;
  if (is_hik || is_lok) {
    const int order = is_hik ? f2d->extrap_order_hik : f2d->extrap_order_lok;
    if (order > 1) {
      double pd;
      if (f2d->is_factorizable) spstatus = o_cspline_eval_e(f2d->fk, lk_ev, 2, &pd);
      else spstatus = o_bicubic_eval_e(f2d->fka, lk_ev, a_ev, 2, &pd);
      if (spstatus) { *status = O_ERR_SPLINE_EV; return NAN; }
      fka_post += pd * (lk - lk_ev);
    }
  }
  if (!f2d->is_log) {
    fka_post *= inv_pk0;
    if (is_hiz) {
      double gz;
      if (f2d->extrap_linear_growth == O_F2D_CCLGROWTH) {
        if (!cosmo || !cosmo->growth) { *status = O_ERR_GROWTH_INIT; return NAN; }
        gz = o_growth_factor(cosmo, a, status) / o_growth_factor(cosmo, a_ev, status);
      } else
        gz = f2d->growth_factor_0;
      fka_post *= pow(gz, f2d->growth_exponent);
    }
  }
  return fka_post;
}

/* ------------------------------------------------------------------ */
/* ccl_cl_tracer_t (src/ccl_tracers.c:569-749)                         */
/* ------------------------------------------------------------------ */
#define O_FRAC_RELEVANT 5E-4 /* include/ccl_tracers.h:13 */

typedef struct {
  int der_bessel, der_angles;
  o_f1d *kernel;
  o_f2d *transfer;
  double chi_min, chi_max;
} o_tracer;

void o_tracer_free(o_tracer *t) {
  if (!t) return;
  o_f1d_free(t->kernel); o_f2d_free(t->transfer); free(t);
}

/* chi_max_nokernel = comoving distance to A_SPLINE_MIN, computed by the caller
 * (src/ccl_tracers.c:632). */
o_tracer *o_tracer_new(int der_bessel, int der_angles, int n_w, const double *chi_w,
                       const double *w_w, int na_ka, const double *a_ka, int nk_ka,
                       const double *lk_ka, const double *fka_arr, const double *fk_arr,
                       const double *fa_arr, int is_fka_log, int is_factorizable,
                       int extrap_order_lok, int extrap_order_hik, double chi_max_nokernel,
                       int *status) {
  if ((der_angles < 0) || (der_angles > 2)) *status = O_ERR_INCONSISTENT;
  if ((der_bessel < -1) || (der_bessel > 2)) *status = O_ERR_INCONSISTENT;
  if (*status) return NULL;
  o_tracer *tr = calloc(1, sizeof(o_tracer));
  tr->der_angles = der_angles;
  tr->der_bessel = der_bessel;
  tr->chi_min = 0;
  tr->chi_max = 1E15;
  if ((n_w > 0) && (chi_w != NULL) && (w_w != NULL)) {
    tr->kernel = o_f1d_new(n_w, chi_w, w_w, 0, 0);
    if (tr->kernel == NULL) *status = O_ERR_MEMORY;
  }
  if (*status == 0) {
    if (tr->kernel == NULL) {
      tr->chi_min = 0;
      tr->chi_max = chi_max_nokernel;
    } else {
      int ichi;
      double w_max = fabs(w_w[0]);
      for (ichi = 0; ichi < n_w; ichi++)
        if (fabs(w_w[ichi]) >= w_max) w_max = fabs(w_w[ichi]);
      w_max *= O_FRAC_RELEVANT;
      tr->chi_min = chi_w[0];
      tr->chi_max = chi_w[n_w - 1];
      for (ichi = 0; ichi < n_w - 1; ichi++)
        if (fabs(w_w[ichi + 1]) >= w_max) { tr->chi_min = chi_w[ichi]; break; }
      for (ichi = n_w - 1; ichi >= 1; ichi--)
        if (fabs(w_w[ichi - 1]) >= w_max) { tr->chi_max = chi_w[ichi]; break; }
    }
  }
  if (*status == 0) {
    if ((fka_arr != NULL) || (fk_arr != NULL) || (fa_arr != NULL)) {
      tr->transfer = o_f2d_new(na_ka, a_ka, nk_ka, lk_ka, fka_arr, fk_arr, fa_arr,
                               is_factorizable, extrap_order_lok, extrap_order_hik,
                               O_F2D_CONSTGROWTH, is_fka_log, 1, 0, status);
    }
  }
  return tr;
}

double o_tracer_chi_min(const o_tracer *t) { return t->chi_min; }
double o_tracer_chi_max(const o_tracer *t) { return t->chi_max; }

double o_tracer_f_ell(const o_tracer *tr, double ell) {
  if (tr->der_angles == 1) return ell * (ell + 1.);
  else if (tr->der_angles == 2) {
    if (ell <= 1) return 0;
    else if (ell <= 10) return sqrt((ell + 2) * (ell + 1) * ell * (ell - 1));
    else {
      double lp1h = ell + 0.5;
      double lp1h2 = lp1h * lp1h;
      if (ell <= 1000) return lp1h2 * (1 - 1.25 / lp1h2);
      else return lp1h2;
    }
  }
  return 1;
}

double o_tracer_kernel(const o_tracer *tr, double chi) {
  return tr->kernel ? o_f1d_eval(tr->kernel, chi) : 1;
}

double o_tracer_transfer(const o_tracer *tr, double lk, double a, int *status) {
  return tr->transfer ? o_f2d_eval(tr->transfer, lk, a, NULL, status) : 1;
}

/* ------------------------------------------------------------------ */
/* src/ccl_cls.c:64-263                                                */
/* ------------------------------------------------------------------ */
typedef struct {
  double l;
  const o_cosmo *cosmo;
  int n1, n2;
  o_tracer *const *trc1;
  o_tracer *const *trc2;
  const o_f2d *psp;
  int *status;
  long neval;
} o_integ_par;

/* Test knob (not in the reference): shift lkmin by n units in the last place.  With the default
 * tracers chi_max IS the last node of a kernel spline, so the integrand's first node sits exactly
 * on a support edge, chi_0 = (l+1/2)/exp(log((l+1/2)/chi_max)), where ccl_f1d_t_eval jumps from
 * W(chi_max) to 0 (src/ccl_f1d.c:159-161).  Which side chi_0 falls on is decided by the last bit
 * of libm's log/exp, so the reference's own value is only defined up to that choice; the parity
 * tests use this knob to enumerate the outcomes a <= 1 ulp libm can produce. */
static int o_lkmin_ulps = 0;
void o_set_lkmin_ulps(int n) { o_lkmin_ulps = n; }

static void o_get_k_interval(const o_cosmo *cosmo, int n1, o_tracer *const *t1, int n2,
                             o_tracer *const *t2, double l, double *lkmin, double *lkmax) {
  double chi_min1 = 1E15, chi_max1 = -1E15;
  for (int i = 0; i < n1; i++) {
    if (t1[i]->chi_min < chi_min1) chi_min1 = t1[i]->chi_min;
    if (t1[i]->chi_max > chi_max1) chi_max1 = t1[i]->chi_max;
  }
  double chi_min2 = 1E15, chi_max2 = -1E15;
  for (int i = 0; i < n2; i++) {
    if (t2[i]->chi_min < chi_min2) chi_min2 = t2[i]->chi_min;
    if (t2[i]->chi_max > chi_max2) chi_max2 = t2[i]->chi_max;
  }
  double chi_min = fmax(chi_min1, chi_min2);
  double chi_max = fmin(chi_max1, chi_max2);
  if (chi_min <= 0) chi_min = 0.5 * (l + 0.5) / cosmo->K_MAX;
  *lkmax = log(fmin(cosmo->K_MAX, 2 * (l + 0.5) / chi_min));
  *lkmin = log(fmax(cosmo->K_MIN, (l + 0.5) / chi_max));
  for (int i = 0; i < abs(o_lkmin_ulps); i++)
    *lkmin = nextafter(*lkmin, o_lkmin_ulps > 0 ? INFINITY : -INFINITY);
}

static double o_transfer_limber_single(const o_tracer *tr, double l, double lk, double k,
                                       double chi_l, double a_l, const o_cosmo *cosmo,
                                       const o_f2d *psp, int *status) {
  double dd = 0;
  double w = o_tracer_kernel(tr, chi_l);
  double t = o_tracer_transfer(tr, lk, a_l, status);
  double fl = o_tracer_f_ell(tr, l);
  if (tr->der_bessel < 1) {
    dd = w * t;
    if (tr->der_bessel == -1) {
      double lp1h = l + 0.5;
      dd /= (lp1h * lp1h);
    }
  } else {
    double lp1h = l + 0.5;
    double lp3h = l + 1.5;
    double chi_lp = lp3h / k;
    double a_lp = o_scale_factor_of_chi(cosmo, chi_lp, status);
    double pk_ratio = fabs(o_f2d_eval(psp, lk, a_lp, cosmo, status) /
                           o_f2d_eval(psp, lk, a_l, cosmo, status));
    double w_p = o_tracer_kernel(tr, chi_lp);
    double t_p = o_tracer_transfer(tr, lk, a_lp, status);
    double sqell = sqrt(lp1h * pk_ratio / lp3h);
    if (tr->der_bessel == 1) dd = l * w * t / lp1h - sqell * w_p * t_p;
    else dd = sqell * 2 * w_p * t_p / lp3h - (0.25 + 2 * l) * w * t / (lp1h * lp1h);
  }
  return dd * fl;
}

static double o_transfer_limber_wrap(double l, double lk, double k, double chi, double a, int n,
                                     o_tracer *const *trc, const o_cosmo *cosmo,
                                     const o_f2d *psp, int *status) {
  double transfer = 0;
  for (int itr = 0; itr < n; itr++) {
    transfer += o_transfer_limber_single(trc[itr], l, lk, k, chi, a, cosmo, psp, status);
    if (*status != 0) return -1;
  }
  return transfer;
}

static double o_cl_integrand(double lk, void *params) {
  o_integ_par *p = (o_integ_par *)params;
  p->neval++;
  double k = exp(lk);
  double chi = (p->l + 0.5) / k;
  double a = o_scale_factor_of_chi(p->cosmo, chi, p->status);
  double d1 = o_transfer_limber_wrap(p->l, lk, k, chi, a, p->n1, p->trc1, p->cosmo, p->psp, p->status);
  if (d1 == 0) return 0;
  double d2 = o_transfer_limber_wrap(p->l, lk, k, chi, a, p->n2, p->trc2, p->cosmo, p->psp, p->status);
  if (d2 == 0) return 0;
  double pk = o_f2d_eval(p->psp, lk, a, p->cosmo, p->status);
  return k * pk * d1 * d2;
}

/* src/ccl_utils.c:17-37 */
static double *o_linear_spacing(double xmin, double xmax, int N) {
  double dx = (xmax - xmin) / (N - 1.);
  double *x = malloc(sizeof(double) * N);
  for (int i = 0; i < N; i++) x[i] = xmin + dx * i;
  x[0] = xmin;
  x[N - 1] = xmax;
  return x;
}

/* src/ccl_utils.c:274-345 with ny = 1, T = akima */
void o_integ_spline(int nx, const double *x, const double *y, double a, double b,
                    double *result, int *status) {
  if (b == a) { *result = 0; return; }
  if (b < a) { b = x[nx - 1]; a = x[0]; }
  if ((b > x[nx - 1]) || (a < x[0])) { *status = O_ERR_SPLINE; return; }
  if (*status == 0) {
    o_akima *s = o_akima_new(nx, x, y);
    if (s == NULL) { *status = O_ERR_MEMORY; return; }
    int sstat = o_akima_integ_e(s, a, b, result);
    if (sstat) { *status = O_ERR_SPLINE_EV; *result = NAN; }
    o_akima_free(s);
  }
}

int o_limber_nk(double lkmin, double lkmax, double dlogk) {
  return (int)(fmax((lkmax - lkmin) / dlogk + 0.5, 1)) + 1;
}

static void o_integ_cls_limber_spline(const o_cosmo *cosmo, o_integ_par *ipar, double lkmin,
                                      double lkmax, double *result, int *status) {
  int nk = o_limber_nk(lkmin, lkmax, cosmo->DLOGK_INTEGRATION);
  double *lk_arr = o_linear_spacing(lkmin, lkmax, nk);
  double *fk_arr = malloc(nk * sizeof(double));
  if (*status == 0) {
    for (int ik = 0; ik < nk; ik++) {
      fk_arr[ik] = o_cl_integrand(lk_arr[ik], ipar);
      if (*(ipar->status)) { *status = *(ipar->status); break; }
    }
  }
  if (*status == 0) o_integ_spline(nk, lk_arr, fk_arr, 1, -1, result, status);
  free(fk_arr);
  free(lk_arr);
}

/* ------------------------------------------------------------------ */
/* GSL integration/qk.c + qk41.c + qag.c + workspace (qpsrt.c, util.c) */
/* ------------------------------------------------------------------ */
typedef double (*o_func)(double, void *);

static double o_rescale_error(double err, const double result_abs, const double result_asc) {
  err = fabs(err);
  if (result_asc != 0 && err != 0) {
    double scale = pow((200 * err / result_asc), 1.5);
    if (scale < 1) err = result_asc * scale;
    else err = result_asc;
  }
  if (result_abs > DBL_MIN / (50 * DBL_EPSILON)) {
    double min_err = 50 * DBL_EPSILON * result_abs;
    if (min_err > err) err = min_err;
  }
  return err;
}

static void o_qk41(o_func f, void *par, double a, double b, double *result, double *abserr,
                   double *resabs, double *resasc) {
  const int n = 21;
  const double *xgk = GK41_XGK, *wg = GK41_WG, *wgk = GK41_WGK;
  double fv1[21], fv2[21];
  const double center = 0.5 * (a + b);
  const double half_length = 0.5 * (b - a);
  const double abs_half_length = fabs(half_length);
  const double f_center = f(center, par);
  double result_gauss = 0;
  double result_kronrod = f_center * wgk[n - 1];
  double result_abs = fabs(result_kronrod);
  double result_asc = 0;
  double mean = 0, err = 0;
  int j;
  if (n % 2 == 0) result_gauss = f_center * wg[n / 2 - 1];
  for (j = 0; j < (n - 1) / 2; j++) {
    const int jtw = j * 2 + 1;
    const double abscissa = half_length * xgk[jtw];
    const double fval1 = f(center - abscissa, par);
    const double fval2 = f(center + abscissa, par);
    const double fsum = fval1 + fval2;
    fv1[jtw] = fval1;
    fv2[jtw] = fval2;
    result_gauss += wg[j] * fsum;
    result_kronrod += wgk[jtw] * fsum;
    result_abs += wgk[jtw] * (fabs(fval1) + fabs(fval2));
  }
  for (j = 0; j < n / 2; j++) {
    int jtwm1 = j * 2;
    const double abscissa = half_length * xgk[jtwm1];
    const double fval1 = f(center - abscissa, par);
    const double fval2 = f(center + abscissa, par);
    fv1[jtwm1] = fval1;
    fv2[jtwm1] = fval2;
    result_kronrod += wgk[jtwm1] * (fval1 + fval2);
    result_abs += wgk[jtwm1] * (fabs(fval1) + fabs(fval2));
  }
  mean = result_kronrod * 0.5;
  result_asc = wgk[n - 1] * fabs(f_center - mean);
  for (j = 0; j < n - 1; j++)
    result_asc += wgk[j] * (fabs(fv1[j] - mean) + fabs(fv2[j] - mean));
  err = (result_kronrod - result_gauss) * half_length;
  result_kronrod *= half_length;
  result_abs *= abs_half_length;
  result_asc *= abs_half_length;
  *result = result_kronrod;
  *resabs = result_abs;
  *resasc = result_asc;
  *abserr = o_rescale_error(err, result_abs, result_asc);
}

typedef struct {
  size_t limit, size, nrmax, i, maximum_level;
  double *alist, *blist, *rlist, *elist;
  size_t *order, *level;
} o_ws;

static o_ws *o_ws_alloc(size_t n) {
  o_ws *w = calloc(1, sizeof(o_ws));
  w->limit = n;
  w->alist = malloc(n * sizeof(double));
  w->blist = malloc(n * sizeof(double));
  w->rlist = malloc(n * sizeof(double));
  w->elist = malloc(n * sizeof(double));
  w->order = malloc(n * sizeof(size_t));
  w->level = malloc(n * sizeof(size_t));
  return w;
}

static void o_ws_free(o_ws *w) {
  if (!w) return;
  free(w->alist); free(w->blist); free(w->rlist); free(w->elist); free(w->order); free(w->level);
  free(w);
}

/* integration/qpsrt.c */
static void o_qpsrt(o_ws *workspace) {
  const size_t last = workspace->size - 1;
  const size_t limit = workspace->limit;
  double *elist = workspace->elist;
  size_t *order = workspace->order;
  double errmax, errmin;
  int i, k, top;
  size_t i_nrmax = workspace->nrmax;
  size_t i_maxerr = order[i_nrmax];
  if (last < 2) {
    order[0] = 0;
    order[1] = 1;
    workspace->i = i_maxerr;
    return;
  }
  errmax = elist[i_maxerr];
  while (i_nrmax > 0 && errmax > elist[order[i_nrmax - 1]]) {
    order[i_nrmax] = order[i_nrmax - 1];
    i_nrmax--;
  }
  if (last < (limit / 2 + 2)) top = last;
  else top = limit - last + 1;
  i = i_nrmax + 1;
  while (i < top && errmax < elist[order[i]]) {
    order[i - 1] = order[i];
    i++;
  }
  order[i - 1] = i_maxerr;
  errmin = elist[last];
  k = top - 1;
  while (k > i - 2 && errmin >= elist[order[k]]) {
    order[k + 1] = order[k];
    k--;
  }
  order[k + 1] = last;
  i_maxerr = order[i_nrmax];
  workspace->i = i_maxerr;
  workspace->nrmax = i_nrmax;
}

static void o_ws_update(o_ws *workspace, double a1, double b1, double area1, double error1,
                        double a2, double b2, double area2, double error2) {
  double *alist = workspace->alist, *blist = workspace->blist;
  double *rlist = workspace->rlist, *elist = workspace->elist;
  size_t *level = workspace->level;
  const size_t i_max = workspace->i;
  const size_t i_new = workspace->size;
  const size_t new_level = workspace->level[i_max] + 1;
  if (error2 > error1) {
    alist[i_max] = a2;
    rlist[i_max] = area2;
    elist[i_max] = error2;
    level[i_max] = new_level;
    alist[i_new] = a1;
    blist[i_new] = b1;
    rlist[i_new] = area1;
    elist[i_new] = error1;
    level[i_new] = new_level;
  } else {
    blist[i_max] = b1;
    rlist[i_max] = area1;
    elist[i_max] = error1;
    level[i_max] = new_level;
    alist[i_new] = a2;
    blist[i_new] = b2;
    rlist[i_new] = area2;
    elist[i_new] = error2;
    level[i_new] = new_level;
  }
  workspace->size++;
  if (new_level > workspace->maximum_level) workspace->maximum_level = new_level;
  o_qpsrt(workspace);
}

static int o_subinterval_too_small(double a1, double a2, double b2) {
  const double e = DBL_EPSILON;
  const double u = DBL_MIN;
  double tmp = (1 + 100 * e) * (fabs(a2) + 1000 * u);
  return fabs(a1) <= tmp && fabs(b2) <= tmp;
}

/* gsl_integration_qag with key = GSL_INTEG_GAUSS41 */
int o_qag41(o_func f, void *par, double a, double b, double epsabs, double epsrel, size_t limit,
            double *result, double *abserr, int *nintervals) {
  double area, errsum;
  double result0, abserr0, resabs0, resasc0;
  double tolerance;
  size_t iteration = 0;
  int roundoff_type1 = 0, roundoff_type2 = 0, error_type = 0;
  double round_off;
  o_ws *workspace = o_ws_alloc(limit);
  workspace->size = 0; workspace->nrmax = 0; workspace->i = 0;
  workspace->alist[0] = a; workspace->blist[0] = b;
  workspace->rlist[0] = 0.0; workspace->elist[0] = 0.0;
  workspace->order[0] = 0; workspace->level[0] = 0; workspace->maximum_level = 0;
  *result = 0; *abserr = 0;
  if (nintervals) *nintervals = 1;
  if (epsabs <= 0 && (epsrel < 50 * DBL_EPSILON || epsrel < 0.5e-28)) {
    o_ws_free(workspace);
    return O_GSL_EBADTOL;
  }
  o_qk41(f, par, a, b, &result0, &abserr0, &resabs0, &resasc0);
  workspace->size = 1;
  workspace->rlist[0] = result0;
  workspace->elist[0] = abserr0;
  tolerance = fmax(epsabs, epsrel * fabs(result0));
  { volatile double ro = 50 * DBL_EPSILON * resabs0; round_off = ro; }
  if (abserr0 <= round_off && abserr0 > tolerance) {
    *result = result0; *abserr = abserr0; o_ws_free(workspace);
    return O_GSL_EROUND;
  } else if ((abserr0 <= tolerance && abserr0 != resasc0) || abserr0 == 0.0) {
    *result = result0; *abserr = abserr0; o_ws_free(workspace);
    return O_SUCCESS;
  } else if (limit == 1) {
    *result = result0; *abserr = abserr0; o_ws_free(workspace);
    return O_GSL_EMAXITER;
  }
  area = result0;
  errsum = abserr0;
  iteration = 1;
  do {
    double a1, b1, a2, b2;
    double a_i, b_i, r_i, e_i;
    double area1 = 0, area2 = 0, area12 = 0;
    double error1 = 0, error2 = 0, error12 = 0;
    double resasc1, resasc2, resabs1, resabs2;
    a_i = workspace->alist[workspace->i];
    b_i = workspace->blist[workspace->i];
    r_i = workspace->rlist[workspace->i];
    e_i = workspace->elist[workspace->i];
    a1 = a_i;
    b1 = 0.5 * (a_i + b_i);
    a2 = b1;
    b2 = b_i;
    o_qk41(f, par, a1, b1, &area1, &error1, &resabs1, &resasc1);
    o_qk41(f, par, a2, b2, &area2, &error2, &resabs2, &resasc2);
    area12 = area1 + area2;
    error12 = error1 + error2;
    errsum += (error12 - e_i);
    area += area12 - r_i;
    if (resasc1 != error1 && resasc2 != error2) {
      double delta = r_i - area12;
      if (fabs(delta) <= 1.0e-5 * fabs(area12) && error12 >= 0.99 * e_i) roundoff_type1++;
      if (iteration >= 10 && error12 > e_i) roundoff_type2++;
    }
    tolerance = fmax(epsabs, epsrel * fabs(area));
    if (errsum > tolerance) {
      if (roundoff_type1 >= 6 || roundoff_type2 >= 20) error_type = 2;
      if (o_subinterval_too_small(a1, a2, b2)) error_type = 3;
    }
    o_ws_update(workspace, a1, b1, area1, error1, a2, b2, area2, error2);
    iteration++;
  } while (iteration < limit && !error_type && errsum > tolerance);
  {
    double result_sum = 0;
    for (size_t k = 0; k < workspace->size; k++) result_sum += workspace->rlist[k];
    *result = result_sum;
  }
  *abserr = errsum;
  if (nintervals) *nintervals = (int)workspace->size;
  o_ws_free(workspace);
  if (errsum <= tolerance) return O_SUCCESS;
  else if (error_type == 2) return O_GSL_EROUND;
  else if (error_type == 3) return O_GSL_ESING;
  else if (iteration == limit) return O_GSL_EMAXITER;
  return O_GSL_EFAILED;
}

/* --------------------------------------------------------------------------------------------
 * gsl_integration_cquad (GSL 2.x integration/cquad.c; P. Gonnet, "Increasing the reliability of
 * adaptive quadrature using explicit interpolants", ACM TOMS 37 (2010)): doubly-adaptive
 * Clenshaw-Curtis quadrature on nested 5/9/17/33 Chebyshev nodes with the interpolants kept in an
 * orthonormal Legendre basis.  Restated from the published algorithm and GSL's routine; the constant
 * tables are regenerated from their definitions (tools/gen_cquad.py).  The reference calls it only
 * after gsl_integration_qag returned GSL_EROUND (src/ccl_cls.c:245-259).  GSL's removal of
 * non-finite function values (`downdate`) is not restated: a non-finite integrand is an error here.
 * parity unpinned: no reference vector exercises this branch.
 * ------------------------------------------------------------------------------------------ */
#define CQUAD_TABLE static const
#include "cquad_tables.h"
#define O_GSL_EDIVERGE 22

typedef struct {
  double a, b;
  double c[64];
  double fx[33];
  double igral, err;
  int depth, rdepth, ndiv;
} o_cquad_ival;

static void o_cq_vinvfx(const double *fx, double *c, int d) {
  static const int nn[4] = {4, 8, 16, 32}, skip[4] = {8, 4, 2, 1};
  const double *V = d == 0 ? CQ_V1INV : d == 1 ? CQ_V2INV : d == 2 ? CQ_V3INV : CQ_V4INV;
  const int m = nn[d] + 1;
  for (int i = 0; i < m; i++) {
    c[i] = 0.0;
    for (int j = 0; j < m; j++) c[i] += V[i * m + j] * fx[j * skip[d]];
  }
}

static void o_cq_sift_down(const o_cquad_ival *ivals, size_t *heap, int i, int n) {
  while (2 * i + 1 < n) {
    int j = 2 * i + 1;
    if (j + 1 < n && ivals[heap[j + 1]].err >= ivals[heap[j]].err) j++;
    if (ivals[heap[j]].err <= ivals[heap[i]].err) break;
    size_t t = heap[j]; heap[j] = heap[i]; heap[i] = t;
    i = j;
  }
}

static int o_cquad(o_func f, void *par, double a, double b, double epsabs, double epsrel, size_t wsize,
                   double *result, double *abserr, size_t *nevals) {
  static const int n[4] = {4, 8, 16, 32};
  static const int skip[4] = {8, 4, 2, 1};
  static const int idx[4] = {0, 5, 14, 31};
  const double w = M_SQRT2 / 2;
  const int ndiv_max = 20;
  const double *xi = CQ_XI;
  if (epsabs < 0.0 || epsrel < 0.0) return O_GSL_EBADTOL;
  if (epsabs <= 0 && epsrel < DBL_EPSILON) return O_GSL_EBADTOL;
  if (wsize < 3) return O_GSL_EDOM;
  o_cquad_ival *ivals = (o_cquad_ival *)malloc(sizeof(o_cquad_ival) * wsize);
  size_t *heap = (size_t *)malloc(sizeof(size_t) * wsize);
  if (!ivals || !heap) { free(ivals); free(heap); return O_ERR_MEMORY; }
  int neval = 0, bad = 0;
  double nc, ncdiff, temp;

  /* first interval: the 33-point rule straight away */
  o_cquad_ival *iv = &ivals[0], *ivl, *ivr;
  double m = (a + b) / 2, h = (b - a) / 2;
  for (int i = 0; i <= n[3]; i++) {
    iv->fx[i] = f(m + xi[i] * h, par);
    neval++;
    if (!isfinite(iv->fx[i])) bad = 1;
  }
  o_cq_vinvfx(iv->fx, &iv->c[idx[0]], 0);
  o_cq_vinvfx(iv->fx, &iv->c[idx[3]], 3);
  o_cq_vinvfx(iv->fx, &iv->c[idx[2]], 2);
  iv->a = a; iv->b = b; iv->depth = 3; iv->rdepth = 1; iv->ndiv = 0;
  iv->igral = 2 * h * iv->c[idx[3]] * w;
  nc = 0.0;
  for (int i = n[2] + 1; i <= n[3]; i++) { temp = iv->c[idx[3] + i]; nc += temp * temp; }
  ncdiff = nc;
  for (int i = 0; i <= n[2]; i++) {
    temp = iv->c[idx[2] + i] - iv->c[idx[3] + i];
    ncdiff += temp * temp;
    nc += iv->c[idx[3] + i] * iv->c[idx[3] + i];
  }
  ncdiff = sqrt(ncdiff);
  nc = sqrt(nc);
  iv->err = ncdiff * 2 * h;
  if (ncdiff / nc > 0.1 && iv->err < 2 * h * nc) iv->err = 2 * h * nc;
  for (size_t i = 0; i < wsize; i++) heap[i] = i;
  double igral = iv->igral, err = iv->err, igral_final = 0.0, err_final = 0.0;
  int nivals = 1, status = O_SUCCESS;

  while (!bad && nivals > 0 && err > 0.0 && !(err <= fabs(igral) * epsrel || err <= epsabs) &&
         !(err_final > fabs(igral) * epsrel || err - err_final < fabs(igral) * epsrel)) {
    int split, d;
    iv = &ivals[heap[0]];  /* the interval with the largest error */
    m = (iv->a + iv->b) / 2;
    h = (iv->b - iv->a) / 2;
    if (iv->depth < 3) {  /* raise the degree first */
      d = ++iv->depth;
      for (int i = skip[d]; i <= 32; i += 2 * skip[d]) {
        iv->fx[i] = f(m + xi[i] * h, par);
        neval++;
        if (!isfinite(iv->fx[i])) bad = 1;
      }
      o_cq_vinvfx(iv->fx, &iv->c[idx[d]], d);
      nc = 0.0;
      for (int i = n[d - 1] + 1; i <= n[d]; i++) { temp = iv->c[idx[d] + i]; nc += temp * temp; }
      ncdiff = nc;
      for (int i = 0; i <= n[d - 1]; i++) {
        temp = iv->c[idx[d - 1] + i] - iv->c[idx[d] + i];
        ncdiff += temp * temp;
        nc += iv->c[idx[d] + i] * iv->c[idx[d] + i];
      }
      ncdiff = 2 * h * sqrt(ncdiff);
      nc = sqrt(nc);
      iv->err = ncdiff;
      iv->igral = 2 * h * w * iv->c[idx[d]];
      split = (nc > 0 && ncdiff / nc > 0.1);
    } else
      split = 1;  /* maximum degree reached */

    if ((m + h * xi[0]) >= (m + h * xi[1]) || (m + h * xi[31]) >= (m + h * xi[32]) ||
        iv->err < fabs(iv->igral) * DBL_EPSILON * 10) {
      /* drop the interval: keep its contribution */
      err_final += iv->err;
      igral_final += iv->igral;
      size_t t = heap[nivals - 1]; heap[nivals - 1] = heap[0]; heap[0] = t;
      nivals--;
      o_cq_sift_down(ivals, heap, 0, nivals);
    } else if (split) {
      d = iv->depth;
      for (int side = 0; side < 2; side++) {
        o_cquad_ival *ic = &ivals[heap[nivals++]];
        const double *T = side == 0 ? CQ_TLEFT : CQ_TRIGHT;
        if (side == 0) { ivl = ic; ic->a = iv->a; ic->b = m; ic->fx[0] = iv->fx[0]; ic->fx[32] = iv->fx[16]; }
        else { ivr = ic; ic->a = m; ic->b = iv->b; ic->fx[0] = iv->fx[16]; ic->fx[32] = iv->fx[32]; }
        ic->depth = 0;
        ic->rdepth = iv->rdepth + 1;
        for (int i = skip[0]; i < 32; i += skip[0]) {
          ic->fx[i] = f((ic->a + ic->b) / 2 + xi[i] * h / 2, par);
          neval++;
          if (!isfinite(ic->fx[i])) bad = 1;
        }
        o_cq_vinvfx(ic->fx, ic->c, 0);
        for (int i = 0; i <= n[d]; i++) {
          ic->c[idx[d] + i] = 0.0;
          for (int j = i; j <= n[d]; j++) ic->c[idx[d] + i] += T[i * 33 + j] * iv->c[idx[d] + j];
        }
        ncdiff = 0.0;
        for (int i = 0; i <= n[0]; i++) { temp = ic->c[i] - ic->c[idx[d] + i]; ncdiff += temp * temp; }
        for (int i = n[0] + 1; i <= n[d]; i++) { temp = ic->c[idx[d] + i]; ncdiff += temp * temp; }
        ncdiff = sqrt(ncdiff);
        ic->err = ncdiff * h;
        ic->ndiv = iv->ndiv + (fabs(iv->c[0]) > 0 && ic->c[0] / iv->c[0] > 2);
        if (ic->ndiv > ndiv_max && 2 * ic->ndiv > ic->rdepth) {
          *result = (igral >= 0) ? INFINITY : -INFINITY;
          status = O_GSL_EDIVERGE;
          goto done;
        }
        ic->igral = h * w * ic->c[0];
      }
      (void)ivl; (void)ivr;
      /* the parent leaves the heap; the right child takes its place, the left one is the last entry */
      size_t t = heap[nivals - 1]; heap[nivals - 1] = heap[0]; heap[0] = t;
      nivals--;
      o_cq_sift_down(ivals, heap, 0, nivals - 1);
      int i = nivals - 1;
      while (i > 0) {
        int j = (i - 1) / 2;
        if (ivals[heap[j]].err < ivals[heap[i]].err) {
          t = heap[j]; heap[j] = heap[i]; heap[i] = t;
          i = j;
        } else
          break;
      }
    } else
      o_cq_sift_down(ivals, heap, 0, nivals);

    /* heap about to overflow: retire the last intervals */
    while (nivals > (int)wsize - 2) {
      iv = &ivals[heap[nivals - 1]];
      err_final += iv->err;
      igral_final += iv->igral;
      nivals--;
    }
    igral = igral_final;
    err = err_final;
    for (int i = 0; i < nivals; i++) { igral += ivals[heap[i]].igral; err += ivals[heap[i]].err; }
  }
  *result = igral;
done:
  if (abserr) *abserr = err;
  if (nevals) *nevals = (size_t)neval;
  free(ivals);
  free(heap);
  if (bad) { *result = NAN; return O_GSL_EFAILED; }
  return status;
}

static int o_force_cquad = 0;
/* oracle-only: send every 'qag_quad' C_ell through CQUAD (tests of the fallback branch) */
void o_set_force_cquad(int on) { o_force_cquad = on; }

static void o_integ_cls_limber_qag_quad(const o_cosmo *cosmo, o_integ_par *ipar, double lkmin,
                                        double lkmax, double *result, double *eresult,
                                        int *status) {
  int gslstatus = o_force_cquad ? O_GSL_EROUND
                                : o_qag41(o_cl_integrand, ipar, lkmin, lkmax, 0, cosmo->LIMBER_EPSREL,
                                          cosmo->N_ITERATION, result, eresult, NULL);
  if (gslstatus == O_GSL_EROUND) {  /* src/ccl_cls.c:245-259 */
    size_t nevals = 0;
    gslstatus = o_cquad(o_cl_integrand, ipar, lkmin, lkmax, 0, cosmo->LIMBER_EPSREL,
                        (size_t)cosmo->N_ITERATION, result, eresult, &nevals);
  }
  if (*status == 0) *status = gslstatus;
}

/* ------------------------------------------------------------------ */
/* ccl_f3d_t (src/ccl_f3d.c) and ccl_angular_cl_covariance (src/ccl_cls.c:377-612) */
/* ------------------------------------------------------------------ */
typedef struct {
  double lkmin, lkmax;
  int na;
  double *a_arr;
  int is_product, extrap_order_lok, extrap_order_hik, extrap_linear_growth, is_log;
  double growth_factor_0;
  int growth_exponent;
  o_f2d *fka_1, *fka_2;
  o_bicubic **tkka; /* [na] bicubic (lk1, lk2) planes */
} o_f3d;

void o_f3d_free(o_f3d *f) {
  if (!f) return;
  o_f2d_free(f->fka_1);
  o_f2d_free(f->fka_2);
  if (f->tkka) {
    for (int ia = 0; ia < f->na; ia++) o_bicubic_free(f->tkka[ia]);
    free(f->tkka);
  }
  free(f->a_arr);
  free(f);
}

/* ccl_f3d_t_new (src/ccl_f3d.c:161-271), interp_type = ccl_f2d_3 */
o_f3d *o_f3d_new(int na, const double *a_arr, int nk, const double *lk_arr, const double *tkka_arr,
                 const double *fka1_arr, const double *fka2_arr, int is_product, int extrap_order_lok,
                 int extrap_order_hik, int extrap_linear_growth, int is_tkka_log, double growth_factor_0,
                 int growth_exponent, int *status) {
  o_f3d *f = calloc(1, sizeof(o_f3d));
  is_product = is_product || (tkka_arr == NULL);
  f->is_product = is_product;
  f->extrap_order_lok = extrap_order_lok;
  f->extrap_order_hik = extrap_order_hik;
  f->extrap_linear_growth = extrap_linear_growth;
  f->is_log = is_tkka_log;
  f->growth_factor_0 = growth_factor_0;
  f->growth_exponent = growth_exponent;
  f->lkmin = lk_arr[0];
  f->lkmax = lk_arr[nk - 1];
  f->na = na;
  f->a_arr = malloc(na * sizeof(double));
  memcpy(f->a_arr, a_arr, na * sizeof(double));
  if ((extrap_order_lok > 1) || (extrap_order_lok < 0) || (extrap_order_hik > 1) || (extrap_order_hik < 0))
    *status = O_ERR_INCONSISTENT;
  if ((extrap_linear_growth != O_F2D_CCLGROWTH) && (extrap_linear_growth != O_F2D_CONSTGROWTH) &&
      (extrap_linear_growth != O_F2D_NOEXTRAPOL))
    *status = O_ERR_INCONSISTENT;
  if (f->is_product) {
    if ((fka1_arr == NULL) || (fka2_arr == NULL)) *status = O_ERR_INCONSISTENT;
  } else if (tkka_arr == NULL)
    *status = O_ERR_INCONSISTENT;
  if (*status == 0) {
    if (f->is_product) {
      f->fka_1 = o_f2d_new(na, a_arr, nk, lk_arr, fka1_arr, NULL, NULL, 0, extrap_order_lok, extrap_order_hik,
                           extrap_linear_growth, is_tkka_log, growth_factor_0, growth_exponent / 2, status);
      f->fka_2 = o_f2d_new(na, a_arr, nk, lk_arr, fka2_arr, NULL, NULL, 0, extrap_order_lok, extrap_order_hik,
                           extrap_linear_growth, is_tkka_log, growth_factor_0, growth_exponent / 2, status);
    } else {
      f->tkka = calloc(na, sizeof(o_bicubic *));
      for (int ia = 0; ia < na; ia++) {
        f->tkka[ia] = o_bicubic_new(nk, nk, lk_arr, lk_arr, tkka_arr + (size_t)ia * nk * nk);
        if (!f->tkka[ia]) { *status = O_ERR_SPLINE; break; }
      }
    }
  }
  return f;
}

/* ccl_find_a_index (src/ccl_f3d.c:46-88): the interval [a_ia, a_ia+1) holding a; na-1 at and above the last node */
static int o_find_a_index(const o_f3d *f, double a) {
  if (a >= f->a_arr[f->na - 1]) return f->na - 1;
  if (a <= f->a_arr[0]) return 0;
  int ia = 0;
  while (ia < f->na - 2 && a >= f->a_arr[ia + 1]) ia++;
  return ia;
}

/* ccl_f3d_t_eval (src/ccl_f3d.c:273-392) */
double o_f3d_eval(const o_f3d *f3d, double lk1, double lk2, double a, const o_cosmo *cosmo, int *status) {
  double tkka_post = 0;
  double a_ev = a;
  int is_hiz = a < f3d->a_arr[0];
  int is_loz = a > f3d->a_arr[f3d->na - 1];
  if (is_loz) {
    if (f3d->extrap_linear_growth == O_F2D_NOEXTRAPOL) { *status = O_ERR_SPLINE_EV; return NAN; }
    a_ev = f3d->a_arr[f3d->na - 1];
  } else if (is_hiz) {
    if (f3d->extrap_linear_growth == O_F2D_NOEXTRAPOL) { *status = O_ERR_SPLINE_EV; return NAN; }
    a_ev = f3d->a_arr[0];
  }
  if (f3d->is_product) {
    double fka1 = o_f2d_eval(f3d->fka_1, lk1, a_ev, cosmo, status);
    double fka2 = o_f2d_eval(f3d->fka_2, lk2, a_ev, cosmo, status);
    if (*status != 0) return NAN;
    tkka_post = fka1 * fka2;
  } else {
    double lk1_ev = lk1;
    int is_hik1 = lk1 > f3d->lkmax, is_lok1 = lk1 < f3d->lkmin;
    int extrap_k1 = (is_hik1 & (f3d->extrap_order_hik > 0)) || (is_lok1 & (f3d->extrap_order_lok > 0));
    if (is_hik1) lk1_ev = f3d->lkmax;
    else if (is_lok1) lk1_ev = f3d->lkmin;
    double lk2_ev = lk2;
    int is_hik2 = lk2 > f3d->lkmax, is_lok2 = lk2 < f3d->lkmin;
    int extrap_k2 = (is_hik2 & (f3d->extrap_order_hik > 0)) || (is_lok2 & (f3d->extrap_order_lok > 0));
    if (is_hik2) lk2_ev = f3d->lkmax;
    else if (is_lok2) lk2_ev = f3d->lkmin;
    int ia = o_find_a_index(f3d, a_ev);
    if (*status == 0) {
      int spstatus = 0;
      double tkka = 0, dtkka1 = 0, dtkka2 = 0;
      spstatus |= o_bicubic_eval_e(f3d->tkka[ia], lk1_ev, lk2_ev, 0, &tkka);
      if (extrap_k1) spstatus |= o_bicubic_eval_e(f3d->tkka[ia], lk1_ev, lk2_ev, 1, &dtkka1);
      if (extrap_k2) spstatus |= o_bicubic_eval_e(f3d->tkka[ia], lk1_ev, lk2_ev, 3, &dtkka2);
      if (ia < f3d->na - 1) {
        double tkka_p1, dtkka1_p1, dtkka2_p1;
        double h = (a_ev - f3d->a_arr[ia]) / (f3d->a_arr[ia + 1] - f3d->a_arr[ia]);
        spstatus |= o_bicubic_eval_e(f3d->tkka[ia + 1], lk1_ev, lk2_ev, 0, &tkka_p1);
        if (!spstatus) tkka = tkka * (1 - h) + tkka_p1 * h;
        if (extrap_k1) {
          spstatus |= o_bicubic_eval_e(f3d->tkka[ia + 1], lk1_ev, lk2_ev, 1, &dtkka1_p1);
          if (!spstatus) dtkka1 = dtkka1 * (1 - h) + dtkka1_p1 * h;
        }
        if (extrap_k2) {
          spstatus |= o_bicubic_eval_e(f3d->tkka[ia + 1], lk1_ev, lk2_ev, 3, &dtkka2_p1);
          if (!spstatus) dtkka2 = dtkka2 * (1 - h) + dtkka2_p1 * h;
        }
      }
      if (spstatus) { *status = O_ERR_SPLINE_EV; return NAN; }
      if (extrap_k1) tkka += dtkka1 * (lk1 - lk1_ev);
      if (extrap_k2) tkka += dtkka2 * (lk2 - lk2_ev);
      tkka_post = tkka;
    }
    if (f3d->is_log) tkka_post = exp(tkka_post);
  }
  if (is_hiz) {
    double gz;
    if (f3d->extrap_linear_growth == O_F2D_CCLGROWTH) {
      if (!cosmo || !cosmo->growth) { *status = O_ERR_GROWTH_INIT; return NAN; }
      gz = o_growth_factor(cosmo, a, status) / o_growth_factor(cosmo, a_ev, status);
    } else
      gz = f3d->growth_factor_0;
    tkka_post *= pow(gz, f3d->growth_exponent);
  }
  return tkka_post;
}

typedef struct {
  int chipow;
  double l1, l2;
  const o_cosmo *cosmo;
  int n1, n2, n3, n4;
  o_tracer *const *trc1, *const *trc2, *const *trc3, *const *trc4;
  const o_f3d *tsp;
  const o_f1d *ker_extra;
  int *status;
} o_cov_par;

/* transfer_limber_wrap(..., psp = NULL, ignore_jbes_deriv = 1) (src/ccl_cls.c:102-164): der_bessel 1/2 terms drop */
static double o_transfer_limber_noderiv(double l, double lk, double chi, double a, int n, o_tracer *const *trc,
                                        int *status) {
  double transfer = 0;
  for (int itr = 0; itr < n; itr++) {
    const o_tracer *tr = trc[itr];
    double dd = 0;
    double w = o_tracer_kernel(tr, chi);
    double t = o_tracer_transfer(tr, lk, a, status);
    double fl = o_tracer_f_ell(tr, l);
    if (tr->der_bessel < 1) {
      dd = w * t;
      if (tr->der_bessel == -1) {
        double lp1h = l + 0.5;
        dd /= (lp1h * lp1h);
      }
    }
    transfer += dd * fl;
    if (*status != 0) return -1;
  }
  return transfer;
}

/* cov_integrand (src/ccl_cls.c:377-411) */
static double o_cov_integrand(double chi, void *params) {
  o_cov_par *p = (o_cov_par *)params;
  double ker = 1;
  double k1 = (p->l1 + 0.5) / chi, k2 = (p->l2 + 0.5) / chi;
  double lk1 = log(k1), lk2 = log(k2);
  double a = o_scale_factor_of_chi(p->cosmo, chi, p->status);
  double d1 = o_transfer_limber_noderiv(p->l1, lk1, chi, a, p->n1, p->trc1, p->status);
  if (d1 == 0) return 0;
  double d2 = o_transfer_limber_noderiv(p->l1, lk1, chi, a, p->n2, p->trc2, p->status);
  if (d2 == 0) return 0;
  double d3 = o_transfer_limber_noderiv(p->l2, lk2, chi, a, p->n3, p->trc3, p->status);
  if (d3 == 0) return 0;
  double d4 = o_transfer_limber_noderiv(p->l2, lk2, chi, a, p->n4, p->trc4, p->status);
  if (d4 == 0) return 0;
  double tkk = o_f3d_eval(p->tsp, lk1, lk2, a, p->cosmo, p->status);
  if (p->ker_extra != NULL) ker = o_f1d_eval(p->ker_extra, a);
  return d1 * d2 * d3 * d4 * tkk * ker / pow(chi, p->chipow);
}

/* update_chi_limits (src/ccl_cls.c:34-62) */
static void o_update_chi_limits(int n, o_tracer *const *trc, double *chimin, double *chimax, int is_union) {
  double lo = 1E15, hi = -1E15;
  for (int i = 0; i < n; i++) {
    if (trc[i]->chi_min < lo) lo = trc[i]->chi_min;
    if (trc[i]->chi_max > hi) hi = trc[i]->chi_max;
  }
  if (is_union) {
    if (lo < *chimin) *chimin = lo;
    if (hi > *chimax) *chimax = hi;
  } else {
    if (lo > *chimin) *chimin = lo;
    if (hi < *chimax) *chimax = hi;
  }
}

/* ccl_angular_cl_covariance (src/ccl_cls.c:490-612); kernel_extra as the (a, f) arrays angular_cov_ssc_vec turns into a
 * ccl_f1d_t with constant extrapolation (pyccl/ccl_covs.i:95-118), n_ker = 0: none.  cov_out[lind1 + nl1 * lind2]. */
void o_angular_cl_covariance(const o_cosmo *cosmo, int n1, o_tracer *const *trc1, int n2, o_tracer *const *trc2,
                             int n3, o_tracer *const *trc3, int n4, o_tracer *const *trc4, const o_f3d *tsp,
                             int nl1, const double *l1_out, int nl2, const double *l2_out, double *cov_out,
                             int integration_method, int chi_exponent, int n_ker, const double *a_ker,
                             const double *f_ker, double prefactor_extra, double dchi, int *status) {
  if (!cosmo->achi) { *status = O_ERR_DISTANCES_INIT; return; }
  o_f1d *ker = n_ker > 0 ? o_f1d_new(n_ker, a_ker, f_ker, f_ker[0], f_ker[n_ker - 1]) : NULL;
  double chimin = 1E15, chimax = -1E15;
  o_update_chi_limits(n1, trc1, &chimin, &chimax, 1);
  o_update_chi_limits(n2, trc2, &chimin, &chimax, 0);
  o_update_chi_limits(n3, trc3, &chimin, &chimax, 0);
  o_update_chi_limits(n4, trc4, &chimin, &chimax, 0);
#pragma omp parallel
  {
    int clastatus = 0;
    int local_status = *status;
    o_cov_par ipar;
    ipar.cosmo = cosmo; ipar.n1 = n1; ipar.n2 = n2; ipar.n3 = n3; ipar.n4 = n4;
    ipar.trc1 = trc1; ipar.trc2 = trc2; ipar.trc3 = trc3; ipar.trc4 = trc4;
    ipar.tsp = tsp; ipar.ker_extra = ker; ipar.status = &clastatus; ipar.chipow = chi_exponent;
#pragma omp for schedule(dynamic)
    for (int lind1 = 0; lind1 < nl1; ++lind1) {
      ipar.l1 = l1_out[lind1];
      for (int lind2 = 0; lind2 < nl2; ++lind2) {
        if (local_status == 0) {
          double result = 0, eresult = 0;
          clastatus = 0;
          ipar.l2 = l2_out[lind2];
          if (integration_method == O_INTEG_QAG) {
            int gslstatus = o_qag41(o_cov_integrand, &ipar, chimin, chimax, 0, cosmo->LIMBER_EPSREL,
                                    cosmo->N_ITERATION, &result, &eresult, NULL);
            if (gslstatus == O_GSL_EROUND) {
              size_t nevals = 0;
              gslstatus = o_cquad(o_cov_integrand, &ipar, chimin, chimax, 0, cosmo->LIMBER_EPSREL,
                                  (size_t)cosmo->N_ITERATION, &result, &eresult, &nevals);
            }
            if (local_status == 0) local_status = gslstatus;
          } else if (integration_method == O_INTEG_SPLINE) {
            /* integ_cov_limber_spline (src/ccl_cls.c:413-452) */
            int nchi = (int)(fmax((chimax - chimin) / dchi + 0.5, 1)) + 1;
            double *chi_arr = o_linear_spacing(chimin, chimax, nchi);
            double *fchi = malloc(nchi * sizeof(double));
            for (int i = 0; i < nchi; i++) {
              fchi[i] = o_cov_integrand(chi_arr[i], &ipar);
              if (clastatus) { local_status = clastatus; break; }
            }
            if (local_status == 0) o_integ_spline(nchi, chi_arr, fchi, 1, -1, &result, &local_status);
            free(fchi);
            free(chi_arr);
          } else
            local_status = O_ERR_NOT_IMPLEMENTED;
          if ((clastatus == 0) && (local_status == 0)) {
            cov_out[lind1 + nl1 * lind2] = result * prefactor_extra;
          } else {
            cov_out[lind1 + nl1 * lind2] = NAN;
            local_status = O_ERR_INTEG;
          }
        }
      }
    }
    if (local_status) {
#pragma omp atomic write
      *status = local_status;
    }
  }
  o_f1d_free(ker);
}

/* src/ccl_cls.c:265-363.  neval_out (may be NULL) receives the integrand evaluations per ell. */
void o_angular_cls_limber(const o_cosmo *cosmo, int n1, o_tracer *const *trc1, int n2,
                          o_tracer *const *trc2, const o_f2d *psp, int nl_out,
                          const double *l_out, double *cl_out, int integration_method,
                          long *neval_out, int *status) {
  if (!cosmo->achi) { *status = O_ERR_DISTANCES_INIT; return; }
#pragma omp parallel
  {
    int clastatus, lind;
    o_integ_par ipar;
    int local_status = *status;
    double lkmin, lkmax, l, result, eresult;
    ipar.cosmo = cosmo; ipar.n1 = n1; ipar.n2 = n2; ipar.trc1 = trc1; ipar.trc2 = trc2;
    ipar.psp = psp; ipar.status = &clastatus;
#pragma omp for schedule(dynamic)
    for (lind = 0; lind < nl_out; ++lind) {
      if (local_status == 0) {
        l = l_out[lind];
        clastatus = 0;
        ipar.l = l;
        ipar.neval = 0;
        o_get_k_interval(cosmo, n1, trc1, n2, trc2, l, &lkmin, &lkmax);
        if (integration_method == O_INTEG_QAG)
          o_integ_cls_limber_qag_quad(cosmo, &ipar, lkmin, lkmax, &result, &eresult, &local_status);
        else if (integration_method == O_INTEG_SPLINE)
          o_integ_cls_limber_spline(cosmo, &ipar, lkmin, lkmax, &result, &local_status);
        else
          local_status = O_ERR_NOT_IMPLEMENTED;
        if (neval_out) neval_out[lind] = ipar.neval;
        if ((*ipar.status == 0) && (local_status == 0)) {
          cl_out[lind] = result / (l + 0.5);
        } else {
          cl_out[lind] = NAN;
          local_status = O_ERR_INTEG;
        }
      }
    }
    if (local_status) {
#pragma omp atomic write
      *status = local_status;
    }
  }
}

/* Unlike the reference (a failing ell poisons its OpenMP thread, src/ccl_cls.c:319), this
 * variant evaluates every ell independently and reports a per-ell status; used to check the
 * GPU's per-(pair, ell) status words. */
void o_angular_cls_limber_each(const o_cosmo *cosmo, int n1, o_tracer *const *trc1, int n2,
                               o_tracer *const *trc2, const o_f2d *psp, int nl_out,
                               const double *l_out, double *cl_out, int integration_method,
                               int *status_each) {
  for (int i = 0; i < nl_out; i++) {
    int st = 0;
    o_angular_cls_limber(cosmo, n1, trc1, n2, trc2, psp, 1, l_out + i, cl_out + i,
                         integration_method, NULL, &st);
    status_each[i] = st;
  }
}

void o_get_k_interval_pub(const o_cosmo *cosmo, int n1, o_tracer *const *t1, int n2,
                          o_tracer *const *t2, double l, double *lkmin, double *lkmax) {
  o_get_k_interval(cosmo, n1, t1, n2, t2, l, lkmin, lkmax);
}

/* ================================================================== */
/* C_ell -> xi(theta): ccl_correlation (src/ccl_correlation.c:14-440),  */
/* FFTLog (src/ccl_fftlog.c:53-106,199-328,525-530), SURVEY 8f-4       */
/* ================================================================== */
#define O_CORR_LGNDRE 1001
#define O_CORR_FFTLOG 1002
#define O_CORR_BESSEL 1003
#define O_CORR_GG 2001
#define O_CORR_GL 2002
#define O_CORR_LP 2003
#define O_CORR_LM 2004
#define O_ERR_LINSPACE 1026

/* ccl_f1d_t with the two extrapolation pairs ccl_correlation.c uses (src/ccl_f1d.c:15-191):
 * below x_ini: constant y0; above x_end: constant yf, or y_end (x/x_end)^der_hi (logx_logy) */
typedef struct {
  o_akima *spl;
  double y0, yf, x_ini, x_end, y_end, der_hi;
  int hi_loglog;
} o_f1dx;

static void o_f1dx_free(o_f1dx *f) {
  if (!f) return;
  o_akima_free(f->spl);
  free(f);
}

static o_f1dx *o_f1dx_new(int n, const double *x, const double *y, double y0, double yf, int hi_loglog,
                          int *status) {
  o_akima *s = o_akima_new(n, x, y); /* gsl_spline_alloc(akima, n < 5) fails: CCL_ERROR_MEMORY */
  if (!s) { *status = O_ERR_MEMORY; return NULL; }
  o_f1dx *f = calloc(1, sizeof(o_f1dx));
  f->spl = s; f->y0 = y0; f->yf = yf; f->x_ini = x[0]; f->x_end = x[n - 1]; f->y_end = y[n - 1];
  f->hi_loglog = hi_loglog;
  if (hi_loglog) { /* src/ccl_f1d.c:97-102 */
    if ((y[n - 1] * y[n - 2] <= 0) || (x[n - 1] * x[n - 2] <= 0)) *status = O_ERR_SPLINE;
    else f->der_hi = log(y[n - 1] / y[n - 2]) / log(x[n - 1] / x[n - 2]);
  }
  if (*status) { o_f1dx_free(f); return NULL; }
  return f;
}

static double o_f1dx_eval(const o_f1dx *f, double x) {
  if (x <= f->x_ini) return f->y0;
  else if (x >= f->x_end) return f->hi_loglog ? f->y_end * pow(x / f->x_end, f->der_hi) : f->yf;
  double y;
  if (o_akima_eval_e(f->spl, x, &y) != O_SUCCESS) return NAN;
  return y;
}

/* taper_cl (src/ccl_correlation.c:20-39) */
static void o_taper_cl(int n_ell, const double *ell, double *cl, const double *lim) {
  for (int i = 0; i < n_ell; i++) {
    if (ell[i] < lim[0] || ell[i] > lim[3]) { cl[i] = 0; continue; }
    if (ell[i] >= lim[1] && ell[i] <= lim[2]) continue;
    if (ell[i] < lim[1]) cl[i] *= cos((ell[i] - lim[1]) / (lim[1] - lim[0]) * M_PI / 2.);
    if (ell[i] > lim[2]) cl[i] *= cos((ell[i] - lim[2]) / (lim[3] - lim[2]) * M_PI / 2.);
  }
}

/* ccl_log_spacing (src/ccl_utils.c:109-148): running product, end points overwritten */
static double *o_log_spacing(double xmin, double xmax, int N) {
  if (N < 2 || !(xmin > 0 && xmax > 0)) return NULL;
  const double dlog_x = (log(xmax) - log(xmin)) / (N - 1.);
  double *x = malloc(sizeof(double) * N);
  const double xratio = exp(dlog_x);
  x[0] = xmin;
  for (int i = 1; i < N - 1; i++) x[i] = x[i - 1] * xratio;
  x[N - 1] = xmax;
  return x;
}

/* GSL specfunc (not in /root/reference: libgsl >= 2.1, CMakeLists.txt:118), restated from the published
 * algorithm: gsl_sf_lngamma_complex_e = Lanczos g = 7, 9 terms (gamma.c lngamma_lanczos_complex) for
 * Re z > 1/2, the reflection formula with gsl_sf_complex_logsin_e (trig.c) otherwise; the phase is
 * reduced to (-pi, pi]. */
static const double o_lanczos_7_c[9] = {
  0.99999999999980993227684700473478, 676.520368121885098567009190444019, -1259.13921672240287047156078755283,
  771.3234287776530788486528258894, -176.61502916214059906584551354, 12.507343278686904814458936853,
  -0.13857109526572011689554707, 9.984369578019570859563e-6, 1.50563273514931155834e-7};

static double o_angle_restrict_symm(double theta) {
  /* gsl_sf_angle_restrict_symm_err_e: 2 pi split in three parts */
  const double P1 = 4 * 7.8539812564849853515625e-01, P2 = 4 * 3.7748947079307981766760e-08,
               P3 = 4 * 2.6951514290790594840552e-15, TwoPi = 2 * (P1 + P2 + P3);
  const double y = (theta >= 0 ? 1 : -1) * 2 * floor(fabs(theta) / TwoPi);
  double r = ((theta - y * P1) - y * P2) - y * P3;
  if (r > M_PI) r = (((r - 2 * P1) - 2 * P2) - 2 * P3);
  else if (r < -M_PI) r = (((r + 2 * P1) + 2 * P2) + 2 * P3);
  return r;
}

static void o_complex_log(double zr, double zi, double *lnr, double *theta) {
  const double ax = fabs(zr), ay = fabs(zi);
  const double mn = ax < ay ? ax : ay, mx = ax < ay ? ay : ax;
  *lnr = log(mx) + 0.5 * log(1.0 + (mn / mx) * (mn / mx));
  *theta = atan2(zi, zr);
}

static void o_lngamma_lanczos_complex(double zr, double zi, double *yr, double *yi) {
  zr -= 1.0; /* Lanczos writes z! instead of Gamma(z) */
  double Ag_r = o_lanczos_7_c[0], Ag_i = 0.0;
  for (int k = 1; k <= 8; k++) {
    const double R = zr + k, I = zi;
    const double a = o_lanczos_7_c[k] / (R * R + I * I);
    Ag_r += a * R;
    Ag_i -= a * I;
  }
  double log1_r, log1_i, logAg_r, logAg_i;
  o_complex_log(zr + 7.5, zi, &log1_r, &log1_i);
  o_complex_log(Ag_r, Ag_i, &logAg_r, &logAg_i);
  /* (z+0.5)*log(z+7.5) - (z+7.5) + LogRootTwoPi_ + log(Ag(z)) */
  *yr = (zr + 0.5) * log1_r - zi * log1_i - (zr + 7.5) + 0.9189385332046727418 + logAg_r;
  *yi = o_angle_restrict_symm(zi * log1_r + (zr + 0.5) * log1_i - zi + logAg_i);
}

static void o_complex_logsin(double zr, double zi, double *lszr, double *lszi) {
  if (zi > 60.0) { *lszr = -M_LN2 + zi; *lszi = 0.5 * M_PI - zr; }
  else if (zi < -60.0) { *lszr = -M_LN2 - zi; *lszi = -0.5 * M_PI + zr; }
  else o_complex_log(sin(zr) * cosh(zi), cos(zr) * sinh(zi), lszr, lszi);
  *lszi = o_angle_restrict_symm(*lszi);
}

void o_lngamma_complex(double zr, double zi, double *lnr, double *arg) {
  if (zr <= 0.5) { /* reflection to the right half plane, stopping at 1/2 */
    double a, b, lnsin_r, lnsin_i;
    o_lngamma_lanczos_complex(1.0 - zr, -zi, &a, &b);
    o_complex_logsin(M_PI * zr, M_PI * zi, &lnsin_r, &lnsin_i);
    *lnr = 1.14472988584940017414 - lnsin_r - a; /* M_LNPI */
    *arg = o_angle_restrict_symm(-lnsin_i - b);
  } else
    o_lngamma_lanczos_complex(zr, zi, lnr, arg);
}

/* goodkr (src/ccl_fftlog.c:72-86) */
static double o_goodkr(int N, double mu, double q, double L, double kr) {
  const double xp = (mu + 1 + q) / 2, xm = (mu + 1 - q) / 2, y = M_PI * N / (2 * L);
  double lnr, argm, argp;
  o_lngamma_complex(xp, y, &lnr, &argp);
  o_lngamma_complex(xm, y, &lnr, &argm);
  const double arg = log(2 / kr) * N / L + (argp + argm) / M_PI;
  const double iarg = round(arg);
  if (arg != iarg) kr *= exp((arg - iarg) * L / N);
  return kr;
}

/* compute_u_coefficients (src/ccl_fftlog.c:95-125), q == 0 branch and the general one; u as (re, im) pairs */
static void o_compute_u(int N, double mu, double q, double L, double kcrc, double *u) {
  const double y = M_PI / L, k0r0 = kcrc * exp(-L), t = -2 * y * log(k0r0 / 2);
  if (q == 0) {
    const double x = (mu + 1) / 2;
    for (int m = 0; m <= N / 2; m++) {
      double lnr, phi;
      o_lngamma_complex(x, m * y, &lnr, &phi);
      u[2 * m] = 1.0 * cos(m * t + 2 * phi);
      u[2 * m + 1] = 1.0 * sin(m * t + 2 * phi);
    }
  } else {
    const double xp = (mu + 1 + q) / 2, xm = (mu + 1 - q) / 2;
    for (int m = 0; m <= N / 2; m++) {
      double lnrp, phip, lnrm, phim;
      o_lngamma_complex(xp, m * y, &lnrp, &phip);
      o_lngamma_complex(xm, -m * y, &lnrm, &phim);
      const double r = exp(q * M_LN2 + lnrp - lnrm), ph = m * t + phip - phim;
      u[2 * m] = r * cos(ph);
      u[2 * m + 1] = r * sin(ph);
    }
  }
  for (int m = N / 2 + 1; m < N; m++) { u[2 * m] = u[2 * (N - m)]; u[2 * m + 1] = -u[2 * (N - m) + 1]; }
  if ((N % 2) == 0) u[2 * (N / 2) + 1] = 0.0;
}

/* fht (src/ccl_fftlog.c:199-328) for one input array.  FFTW (absent here) is replaced by the O(N^2)
 * definition of the two DFTs it computes (forward sign -1, backward sign +1, unnormalised), with the
 * twiddle factors tabulated once. */
void o_fftlog_fht(int N, const double *k, const double *pk, double *r, double *xi, double dim, double mu,
                  double q, double kcrc, int noring) {
  const double L = log(k[N - 1] / k[0]) * N / (N - 1.);
  if (noring) kcrc = o_goodkr(N, mu, q, L, kcrc);
  double *u = malloc(sizeof(double) * 2 * N), *w = malloc(sizeof(double) * 2 * N);
  double *a = malloc(sizeof(double) * N), *b = malloc(sizeof(double) * 2 * N), *c = malloc(sizeof(double) * N);
  o_compute_u(N, mu, q, L, kcrc, u);
  for (int j = 0; j < N; j++) { w[2 * j] = cos(2 * M_PI * j / N); w[2 * j + 1] = sin(2 * M_PI * j / N); }
  const double k0r0 = kcrc * exp(-L);
  r[0] = k0r0 / k[0];
  for (int n = 1; n < N; n++) r[n] = r[0] * exp(n * L / N);
  const double one_over_2pi_dhalf = pow(2 * M_PI, -dim / 2);
  for (int i = 0; i < N; i++) a[i] = pow(k[i], dim / 2 - q) * pk[i];
#pragma omp parallel for schedule(static)
  for (int m = 0; m < N; m++) { /* b[m] = sum_i a[i] exp(-2 pi i m i / N), times u[m] / N */
    double sr = 0, si = 0;
    long idx = 0;
    for (int i = 0; i < N; i++) {
      sr += a[i] * w[2 * idx];
      si -= a[i] * w[2 * idx + 1];
      idx += m; if (idx >= N) idx -= N;
    }
    const double ur = u[2 * m] / (double)N, ui = u[2 * m + 1] / (double)N;
    b[2 * m] = sr * ur - si * ui;
    b[2 * m + 1] = sr * ui + si * ur;
  }
#pragma omp parallel for schedule(static)
  for (int n = 0; n < N; n++) { /* Re sum_m b[m] exp(+2 pi i m n / N) */
    double sr = 0;
    long idx = 0;
    for (int m = 0; m < N; m++) {
      sr += b[2 * m] * w[2 * idx] - b[2 * m + 1] * w[2 * idx + 1];
      idx += n; if (idx >= N) idx -= N;
    }
    c[n] = sr;
  }
  for (int i = 0; i < N; i++) /* reversed array, prefactor */
    xi[i] = one_over_2pi_dhalf * pow(r[i], -dim / 2 - q) * c[N - 1 - i];
  free(u); free(w); free(a); free(b); free(c);
}

/* goodkr_new_deriv (src/ccl_fftlog.c:134-164): low-ringing kr of the generalised transform.  cpow(2, z) and the
 * complex products are written out in (modulus, phase) / (re, im) form. */
static void o_cmul(double *ar, double *ai, double br, double bi) {
  const double r = *ar * br - *ai * bi, i = *ar * bi + *ai * br;
  *ar = r; *ai = i;
}

static void o_general_coef(double mu, double q, int sph, double deriv, double y, double lnk0r0, int with_k0r0,
                           double *ur, double *ui) {
  /* u = 2^(q + sph + 2 i y - deriv) [k0r0^(-2 i y)] Gamma(xp + i y) / Gamma(xm - i y)
   *     prod_{j=1..deriv} (q + 2 i y - (j - sph)) (-1)^deriv          (src/ccl_fftlog.c:146-156,181-192) */
  const double xp = (mu + 1 + q - deriv - 0.5 * sph) / 2, xm = (mu + 1 - q + deriv + 0.5 * sph) / 2;
  double lnrp, phip, lnrm, phim;
  o_lngamma_complex(xp, y, &lnrp, &phip);
  o_lngamma_complex(xm, -y, &lnrm, &phim);
  const double mod2 = pow(2.0, q + 1.0 * sph - deriv);
  double ph = 2 * y * M_LN2;
  if (with_k0r0) ph += -2 * y * lnk0r0;
  double r = mod2 * cos(ph), i = mod2 * sin(ph);
  const double g = exp(lnrp - lnrm);
  o_cmul(&r, &i, g * cos(phip - phim), g * sin(phip - phim));
  for (int j = 1; j <= (int)deriv; j++) o_cmul(&r, &i, q - (j - sph), 2 * y);
  const double sgn = pow(-1.0, deriv);
  *ur = r * sgn; *ui = i * sgn;
}

static double o_goodkr_new_deriv(int N, double mu, double q, double L, int sph, double deriv, double plaw, double kr) {
  q += plaw;
  double pr, pi_;
  o_general_coef(mu, q, sph, deriv, M_PI * N / (2 * L), 0.0, 0, &pr, &pi_);
  const double arg = atan2(pi_, pr) / M_PI; /* cimag(clog(prefac)) / pi */
  const double iarg = round(arg);
  if (arg != iarg) kr *= exp((arg - iarg) * L / N);
  return kr;
}

/* general_fht (src/ccl_fftlog.c:345-503) = ccl_fftlog_ComputeXi_general: for each of the npk input arrays
 *   xi(r_n) = int dk/k?  -- in the reference's words: a~(k) = int dr / (x k)^plaw a(r) (J / j)^(deriv)_mu(x k)
 * on the low-ringing output grid r_n = kcrc / k_{N-1-n}.  DFTs from their definition as in o_fftlog_fht. */
void o_fftlog_general(int npk, int N, const double *k, const double *pk, double *r, double *xi, double mu, double q,
                      int sph, double deriv, double plaw) {
  q = q - 1.0 * sph;
  double kcrc = mu + 1.0;
  mu = mu + 0.5 * sph;
  const double L = log(k[N - 1] / k[0]) * N / (N - 1.);
  kcrc = o_goodkr_new_deriv(N, mu, q, L, sph, deriv, plaw, kcrc);
  double *u = malloc(sizeof(double) * 2 * N), *w = malloc(sizeof(double) * 2 * N);
  double *a = malloc(sizeof(double) * N), *b = malloc(sizeof(double) * 2 * N), *c = malloc(sizeof(double) * N);
  { /* compute_u_coefficients_new_deriv (src/ccl_fftlog.c:171-197) */
    const double y = M_PI / L, lnk0r0 = log(kcrc * exp(-L));
    for (int m = 0; m <= N / 2; m++) o_general_coef(mu, q + plaw, sph, deriv, m * y, lnk0r0, 1, &u[2 * m], &u[2 * m + 1]);
    for (int m = N / 2 + 1; m < N; m++) { u[2 * m] = u[2 * (N - m)]; u[2 * m + 1] = -u[2 * (N - m) + 1]; }
    if ((N % 2) == 0) u[2 * (N / 2) + 1] = 0.0;
  }
  for (int j = 0; j < N; j++) { w[2 * j] = cos(2 * M_PI * j / N); w[2 * j + 1] = sin(2 * M_PI * j / N); }
  for (int n = 0; n < N; n++) r[n] = kcrc / k[N - 1 - n];
  for (int t = 0; t < npk; t++) {
    for (int i = 0; i < N; i++) a[i] = pow(k[i], -1 * sph - q) * pk[(size_t)t * N + i];
#pragma omp parallel for schedule(static)
    for (int m = 0; m < N; m++) {
      double sr = 0, si = 0;
      long idx = 0;
      for (int i = 0; i < N; i++) {
        sr += a[i] * w[2 * idx];
        si -= a[i] * w[2 * idx + 1];
        idx += m; if (idx >= N) idx -= N;
      }
      const double ur = u[2 * m] / (double)N, ui = u[2 * m + 1] / (double)N;
      b[2 * m] = sr * ur - si * ui;
      b[2 * m + 1] = sr * ui + si * ur;
    }
#pragma omp parallel for schedule(static)
    for (int n = 0; n < N; n++) {
      double sr = 0;
      long idx = 0;
      for (int m = 0; m < N; m++) {
        sr += b[2 * m] * w[2 * idx] - b[2 * m + 1] * w[2 * idx + 1];
        idx += n; if (idx >= N) idx -= N;
      }
      c[n] = sr;
    }
    for (int i = 0; i < N; i++)
      xi[(size_t)t * N + i] = pow(r[i], -1 * sph - q) * c[N - 1 - i] * pow(pow(M_PI, 0.5) / 4, sph);
  }
  free(u); free(w); free(a); free(b); free(c);
}

static int o_corr_bessel_order(int corr_type) {
  if (corr_type == O_CORR_GL) return 2;
  if (corr_type == O_CORR_LM) return 4;
  return 0;
}

/* ccl_tracer_corr_fftlog (src/ccl_correlation.c:47-162) */
static void o_corr_fftlog(int n_ell, const double *ell, const double *cls, int n_theta, const double *theta,
                          double *wtheta, int corr_type, int do_taper_cl, const double *taper_cl_limits,
                          double ell_min_corr, double ell_max_corr, int n_ell_corr, int *status) {
  double *l_arr = o_log_spacing(ell_min_corr, ell_max_corr, n_ell_corr);
  if (!l_arr) { *status = O_ERR_LINSPACE; return; }
  o_f1dx *cl_spl = o_f1dx_new(n_ell, ell, cls, cls[0], 0, 1, status);
  if (*status || !cl_spl) { free(l_arr); if (!*status) *status = O_ERR_MEMORY; return; }
  double *cl_arr = malloc(sizeof(double) * n_ell_corr), *th_arr = malloc(sizeof(double) * n_ell_corr),
         *wth_arr = malloc(sizeof(double) * n_ell_corr);
  for (int i = 0; i < n_ell_corr; i++) cl_arr[i] = o_f1dx_eval(cl_spl, l_arr[i]);
  o_f1dx_free(cl_spl);
  if (do_taper_cl) o_taper_cl(n_ell_corr, l_arr, cl_arr, taper_cl_limits);
  /* ccl_fftlog_ComputeXi2D(i_bessel, 0, ...) = fht(dim 2, mu, q 0, kcrc 1, noring 1) */
  o_fftlog_fht(n_ell_corr, l_arr, cl_arr, th_arr, wth_arr, 2., (double)o_corr_bessel_order(corr_type), 0., 1., 1);
  o_f1dx *wth_spl = o_f1dx_new(n_ell_corr, th_arr, wth_arr, wth_arr[0], 0, 0, status);
  if (wth_spl) {
    for (int i = 0; i < n_theta; i++) wtheta[i] = o_f1dx_eval(wth_spl, theta[i] * M_PI / 180.);
    o_f1dx_free(wth_spl);
  } else
    *status = O_ERR_MEMORY;
  free(l_arr); free(cl_arr); free(th_arr); free(wth_arr);
}

typedef struct { const o_f1dx *cl_spl; int i_bessel; double th; } o_corr_int_par;

/* corr_bessel_integrand (src/ccl_correlation.c:170-181); gsl_sf_bessel_Jn -> libm jn */
static double o_corr_bessel_integrand(double l, void *params) {
  const o_corr_int_par *p = (const o_corr_int_par *)params;
  return l * jn(p->i_bessel, l * p->th) * o_f1dx_eval(p->cl_spl, l);
}

/* ccl_tracer_corr_bessel (src/ccl_correlation.c:183-262) */
static void o_corr_bessel(int n_ell, const double *ell, const double *cls, int n_theta, const double *theta,
                          double *wtheta, int corr_type, double ell_max_corr, double epsrel, int n_iteration,
                          int *status) {
  o_f1dx *cl_spl = o_f1dx_new(n_ell, ell, cls, cls[0], 0, 1, status);
  if (!cl_spl) { *status = O_ERR_MEMORY; return; }
  int st_all = *status;
#pragma omp parallel for schedule(dynamic)
  for (int ith = 0; ith < n_theta; ith++) {
    o_corr_int_par cp = {cl_spl, o_corr_bessel_order(corr_type), theta[ith] * M_PI / 180};
    double result, eresult;
    const int gslstatus = o_qag41(o_corr_bessel_integrand, &cp, 0, ell_max_corr, 0, epsrel, (size_t)n_iteration,
                                  &result, &eresult, NULL);
    if (gslstatus != O_SUCCESS) {
#pragma omp atomic
      st_all |= gslstatus;
    }
    wtheta[ith] = result / (2 * M_PI);
  }
  *status = st_all;
  o_f1dx_free(cl_spl);
}

/* ccl_compute_legendre_polynomial (src/ccl_correlation.c:269-291): gsl_sf_legendre_Pl_array and
 * gsl_sf_legendre_Plm(j, 2, x) are the upward recurrences in l (legendre_poly.c); Plm restarts at P_2^2 =
 * 3 (1-x)(1+x) for every j, which produces the same values as one pass */
static void o_legendre_polynomial(int corr_type, double theta, int ell_max, double *Pl_theta) {
  const double cth = cos(theta * M_PI / 180);
  for (int j = 0; j <= ell_max; j++) Pl_theta[j] = 0.;
  if (corr_type == O_CORR_GG) {
    double p_ellm2 = 1.0, p_ellm1 = cth;
    Pl_theta[0] = 1.0;
    if (ell_max >= 1) Pl_theta[1] = cth;
    for (int ell = 2; ell <= ell_max; ell++) {
      const double p_ell = (cth * (2 * ell - 1) * p_ellm1 - (ell - 1) * p_ellm2) / ell;
      p_ellm2 = p_ellm1; p_ellm1 = p_ell;
      Pl_theta[ell] = p_ell;
    }
    for (int j = 0; j <= ell_max; j++) Pl_theta[j] *= (2 * j + 1);
  } else if (corr_type == O_CORR_GL) {
    const double root_factor = sqrt(1.0 - cth) * sqrt(1.0 + cth);
    double p_mm = 1.0, fact_coeff = 1.0;
    for (int i = 1; i <= 2; i++) { p_mm *= -fact_coeff * root_factor; fact_coeff += 2.0; }
    double p_ellm2 = p_mm, p_ellm1 = cth * 5 * p_mm;
    for (int j = 2; j <= ell_max; j++) {
      double p;
      if (j == 2) p = p_mm;
      else if (j == 3) p = p_ellm1;
      else {
        p = (cth * (2 * j - 1) * p_ellm1 - (j + 1) * p_ellm2) / (j - 2);
        p_ellm2 = p_ellm1; p_ellm1 = p;
      }
      Pl_theta[j] = p * ((2 * j + 1.) / ((j + 0.) * (j + 1.)));
    }
  }
}

/* ccl_tracer_corr_legendre (src/ccl_correlation.c:298-400) */
static void o_corr_legendre(int n_ell, const double *ell, const double *cls, int n_theta, const double *theta,
                            double *wtheta, int corr_type, int do_taper_cl, const double *taper_cl_limits,
                            double ell_max_corr, int *status) {
  if (corr_type == O_CORR_LM || corr_type == O_CORR_LP) { *status = O_ERR_NOT_IMPLEMENTED; return; }
  const int lmax = (int)ell_max_corr;
  o_f1dx *cl_spl = o_f1dx_new(n_ell, ell, cls, cls[0], 0, 1, status);
  if (!cl_spl) { *status = O_ERR_MEMORY; return; }
  double *l_arr = malloc(sizeof(double) * (lmax + 1)), *cl_arr = malloc(sizeof(double) * (lmax + 1));
  for (int i = 0; i <= lmax; i++) { l_arr[i] = (double)i; cl_arr[i] = o_f1dx_eval(cl_spl, l_arr[i]); }
  o_f1dx_free(cl_spl);
  if (do_taper_cl) o_taper_cl(lmax + 1, l_arr, cl_arr, taper_cl_limits);
#pragma omp parallel
  {
    double *Pl_theta = malloc(sizeof(double) * (lmax + 1));
#pragma omp for schedule(dynamic)
    for (int i = 0; i < n_theta; i++) {
      wtheta[i] = 0;
      o_legendre_polynomial(corr_type, theta[i], lmax, Pl_theta);
      for (int i_L = 1; i_L < lmax; i_L += 1) wtheta[i] += cl_arr[i_L] * Pl_theta[i_L];
      wtheta[i] /= (M_PI * 4);
    }
    free(Pl_theta);
  }
  free(l_arr); free(cl_arr);
}

/* ccl_correlation (src/ccl_correlation.c:410-440); the spline/gsl parameters it reads from cosmo are arguments */
void o_correlation(int n_ell, const double *ell, const double *cls, int n_theta, const double *theta,
                   double *wtheta, int corr_type, int do_taper_cl, const double *taper_cl_limits,
                   int flag_method, double ell_min_corr, double ell_max_corr, int n_ell_corr, double epsrel,
                   int n_iteration, int *status) {
  switch (flag_method) {
    case O_CORR_FFTLOG:
      o_corr_fftlog(n_ell, ell, cls, n_theta, theta, wtheta, corr_type, do_taper_cl, taper_cl_limits,
                    ell_min_corr, ell_max_corr, n_ell_corr, status);
      break;
    case O_CORR_LGNDRE:
      o_corr_legendre(n_ell, ell, cls, n_theta, theta, wtheta, corr_type, do_taper_cl, taper_cl_limits,
                      ell_max_corr, status);
      break;
    case O_CORR_BESSEL:
      o_corr_bessel(n_ell, ell, cls, n_theta, theta, wtheta, corr_type, ell_max_corr, epsrel, n_iteration,
                    status);
      break;
    default:
      *status = O_ERR_INCONSISTENT;
  }
}

/* test helper: integrate exp(-x^2) * cos(w x) type functions through the qag restatement */
typedef struct { int kind; double p; } o_testf_par;
static double o_testf(double x, void *v) {
  o_testf_par *p = (o_testf_par *)v;
  switch (p->kind) {
    case 0: return exp(-x * x) * cos(p->p * x);
    case 1: return pow(fabs(x), p->p);
    case 2: return 1.0 / (1.0 + p->p * x * x);
    default: return x;
  }
}
int o_qag41_test(int kind, double p, double a, double b, double epsabs, double epsrel, int limit,
                 double *result, double *abserr, int *nint) {
  o_testf_par par = {kind, p};
  return o_qag41(o_testf, &par, a, b, epsabs, epsrel, limit, result, abserr, nint);
}

int o_cquad_test(int kind, double p, double a, double b, double epsabs, double epsrel, int wsize,
                 double *result, double *abserr, int *nevals) {
  o_testf_par par = {kind, p};
  size_t ne = 0;
  const int st = o_cquad(o_testf, &par, a, b, epsabs, epsrel, (size_t)wsize, result, abserr, &ne);
  *nevals = (int)ne;
  return st;
}

void o_set_num_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

int o_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
