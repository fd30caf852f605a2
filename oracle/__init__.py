"""TEST INFRASTRUCTURE ONLY -- ctypes front-end of the CPU oracle (oracle/limber_oracle.c).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package.  The product (ccl_b200/) never does.  See limber_oracle.c for the list of
reference files (file:line) each function restates and for the parity-pin status.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

INTEG_QAG = 500      # include/ccl_utils.h:12-15 (ccl_integration_qag_quad)
INTEG_SPLINE = 501   # ccl_integration_spline
F2D_CCLGROWTH = 401  # include/ccl_f2d.h:13-18
F2D_CONSTGROWTH = 403
F2D_NOEXTRAPOL = 404

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


def build(force=False):
    """Compile liblimber_oracle.so with the committed Makefile (gcc, OpenMP if available)."""
    so = os.path.join(_HERE, "liblimber_oracle.so")
    src = os.path.join(_HERE, "limber_oracle.c")
    mk = os.path.join(_HERE, "Makefile")
    if force or not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(mk)):
        r = subprocess.run(["make", "-C", _HERE, "-B"], capture_output=True, text=True)
        if r.returncode != 0:
            # no OpenMP runtime: scalar build
            r = subprocess.run(["gcc", "-O3", "-fomit-frame-pointer", "-fPIC", "-std=gnu99", "-ffp-contract=off", "-shared",
                                "-o", so, src, "-lm"], capture_output=True, text=True)
            if r.returncode != 0:
                raise RuntimeError("oracle build failed:\n" + r.stderr)
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        vp = C.c_void_p
        L.o_akima_new.restype = vp
        L.o_akima_new.argtypes = [C.c_int, _dp, _dp]
        L.o_akima_free.argtypes = [vp]
        L.o_akima_eval_e.argtypes = [vp, C.c_double, _dp]
        L.o_akima_integ_e.argtypes = [vp, C.c_double, C.c_double, _dp]
        L.o_cspline_new.restype = vp
        L.o_cspline_new.argtypes = [C.c_int, _dp, _dp]
        L.o_cspline_free.argtypes = [vp]
        L.o_cspline_eval_e.argtypes = [vp, C.c_double, C.c_int, _dp]
        L.o_bicubic_new.restype = vp
        L.o_bicubic_new.argtypes = [C.c_int, C.c_int, _dp, _dp, _dp]
        L.o_bicubic_free.argtypes = [vp]
        L.o_bicubic_eval_e.argtypes = [vp, C.c_double, C.c_double, C.c_int, _dp]
        L.o_cosmo_new.restype = vp
        L.o_cosmo_new.argtypes = [C.c_int, _dp, _dp, C.c_int, _dp, _dp, C.c_double, C.c_double,
                                  C.c_double, C.c_double, C.c_int]
        L.o_cosmo_free.argtypes = [vp]
        L.o_scale_factor_of_chi.restype = C.c_double
        L.o_scale_factor_of_chi.argtypes = [vp, C.c_double, _ip]
        L.o_f2d_new.restype = vp
        L.o_f2d_new.argtypes = [C.c_int, _dp, C.c_int, _dp, _dp, _dp, _dp, C.c_int, C.c_int, C.c_int,
                                C.c_int, C.c_int, C.c_double, C.c_int, _ip]
        L.o_f2d_free.argtypes = [vp]
        L.o_f2d_eval.restype = C.c_double
        L.o_f2d_eval.argtypes = [vp, C.c_double, C.c_double, vp, _ip]
        L.o_f2d_dlogf_dlk_eval.restype = C.c_double
        L.o_f2d_dlogf_dlk_eval.argtypes = [vp, C.c_double, C.c_double, vp, _ip]
        L.o_tracer_new.restype = vp
        L.o_tracer_new.argtypes = [C.c_int, C.c_int, C.c_int, _dp, _dp, C.c_int, _dp, C.c_int, _dp,
                                   _dp, _dp, _dp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, _ip]
        L.o_tracer_free.argtypes = [vp]
        L.o_tracer_chi_min.restype = C.c_double
        L.o_tracer_chi_min.argtypes = [vp]
        L.o_tracer_chi_max.restype = C.c_double
        L.o_tracer_chi_max.argtypes = [vp]
        L.o_tracer_f_ell.restype = C.c_double
        L.o_tracer_f_ell.argtypes = [vp, C.c_double]
        L.o_tracer_kernel.restype = C.c_double
        L.o_tracer_kernel.argtypes = [vp, C.c_double]
        L.o_tracer_transfer.restype = C.c_double
        L.o_tracer_transfer.argtypes = [vp, C.c_double, C.c_double, _ip]
        L.o_angular_cls_limber.argtypes = [vp, C.c_int, C.POINTER(vp), C.c_int, C.POINTER(vp), vp,
                                           C.c_int, _dp, _dp, C.c_int, C.POINTER(C.c_long), _ip]
        L.o_angular_cls_limber_each.argtypes = [vp, C.c_int, C.POINTER(vp), C.c_int, C.POINTER(vp), vp,
                                                C.c_int, _dp, _dp, C.c_int, _ip]
        L.o_f3d_new.restype = vp
        L.o_f3d_new.argtypes = [C.c_int, _dp, C.c_int, _dp, _dp, _dp, _dp, C.c_int, C.c_int, C.c_int, C.c_int,
                                C.c_int, C.c_double, C.c_int, _ip]
        L.o_f3d_free.argtypes = [vp]
        L.o_f3d_eval.restype = C.c_double
        L.o_f3d_eval.argtypes = [vp, C.c_double, C.c_double, C.c_double, vp, _ip]
        L.o_angular_cl_covariance.argtypes = [vp, C.c_int, C.POINTER(vp), C.c_int, C.POINTER(vp), C.c_int,
                                              C.POINTER(vp), C.c_int, C.POINTER(vp), vp, C.c_int, _dp, C.c_int, _dp,
                                              _dp, C.c_int, C.c_int, C.c_int, _dp, _dp, C.c_double, C.c_double, _ip]
        L.o_correlation.argtypes = [C.c_int, _dp, _dp, C.c_int, _dp, _dp, C.c_int, C.c_int, _dp, C.c_int, C.c_double,
                                    C.c_double, C.c_int, C.c_double, C.c_int, _ip]
        L.o_lngamma_complex.argtypes = [C.c_double, C.c_double, _dp, _dp]
        L.o_fftlog_fht.argtypes = [C.c_int, _dp, _dp, _dp, _dp, C.c_double, C.c_double, C.c_double, C.c_double, C.c_int]
        L.o_fftlog_general.argtypes = [C.c_int, C.c_int, _dp, _dp, _dp, _dp, C.c_double, C.c_double, C.c_int,
                                       C.c_double, C.c_double]
        L.o_get_k_interval_pub.argtypes = [vp, C.c_int, C.POINTER(vp), C.c_int, C.POINTER(vp),
                                           C.c_double, _dp, _dp]
        L.o_set_lkmin_ulps.argtypes = [C.c_int]
        L.o_limber_nk.restype = C.c_int
        L.o_limber_nk.argtypes = [C.c_double, C.c_double, C.c_double]
        L.o_integ_spline.argtypes = [C.c_int, _dp, _dp, C.c_double, C.c_double, _dp, _ip]
        L.o_qag41_test.argtypes = [C.c_int, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double,
                                   C.c_int, _dp, _dp, _ip]
        L.o_cquad_test.argtypes = [C.c_int, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double,
                                   C.c_int, _dp, _dp, _ip]
        L.o_set_num_threads.argtypes = [C.c_int]
        L.o_set_force_cquad.argtypes = [C.c_int]
        L.o_num_threads.restype = C.c_int
        _LIB = L
    return _LIB


def _arr(x):
    if x is None:
        return None, None
    a = np.ascontiguousarray(x, dtype=np.float64)
    return a, a.ctypes.data_as(_dp)


class Akima:
    """GSL akima spline (interpolation/akima.c)."""

    def __init__(self, x, y):
        self._x, px = _arr(x)
        self._y, py = _arr(y)
        self.h = lib().o_akima_new(len(self._x), px, py)
        if not self.h:
            raise ValueError("akima needs >= 5 points")

    def __call__(self, x):
        out = np.empty(np.size(x))
        v = C.c_double()
        for i, xi in enumerate(np.atleast_1d(x)):
            lib().o_akima_eval_e(self.h, float(xi), C.byref(v))
            out[i] = v.value
        return out

    def integ(self, a, b):
        v = C.c_double()
        st = lib().o_akima_integ_e(self.h, float(a), float(b), C.byref(v))
        return v.value, st

    def __del__(self):
        if getattr(self, "h", None):
            lib().o_akima_free(self.h)


class CSpline:
    """GSL natural cubic spline (interpolation/cspline.c)."""

    def __init__(self, x, y):
        self._x, px = _arr(x)
        self._y, py = _arr(y)
        self.h = lib().o_cspline_new(len(self._x), px, py)

    def __call__(self, x, which=0):
        out = np.empty(np.size(x))
        v = C.c_double()
        for i, xi in enumerate(np.atleast_1d(x)):
            lib().o_cspline_eval_e(self.h, float(xi), which, C.byref(v))
            out[i] = v.value
        return out

    def __del__(self):
        if getattr(self, "h", None):
            lib().o_cspline_free(self.h)


class Bicubic:
    """GSL bicubic (interp2d/bicubic.c); z[j, i] = f(x_i, y_j)."""

    def __init__(self, x, y, z):
        self._x, px = _arr(x)
        self._y, py = _arr(y)
        self._z, pz = _arr(np.asarray(z).ravel())
        self.h = lib().o_bicubic_new(len(self._x), len(self._y), px, py, pz)

    def __call__(self, x, y, which=0):
        v = C.c_double()
        st = lib().o_bicubic_eval_e(self.h, float(x), float(y), which, C.byref(v))
        return v.value, st

    def __del__(self):
        if getattr(self, "h", None):
            lib().o_bicubic_free(self.h)


class Cosmo:
    """a(chi) + growth splines and the spline/gsl params the path reads (SURVEY 8a a14,a15,a18)."""

    def __init__(self, chi, a, a_growth=None, growth=None, K_MAX=1e3, K_MIN=5e-5, DLOGK=0.025,
                 LIMBER_EPSREL=1e-4, N_ITERATION=1000):
        self._chi, pchi = _arr(chi)
        self._a, pa = _arr(a)
        self._ag, pag = _arr(a_growth)
        self._g, pg = _arr(growth)
        ng = 0 if a_growth is None else len(self._ag)
        self.h = lib().o_cosmo_new(len(self._chi), pchi, pa, ng, pag, pg, K_MAX, K_MIN, DLOGK,
                                   LIMBER_EPSREL, N_ITERATION)

    def scale_factor_of_chi(self, chi):
        st = C.c_int(0)
        out = np.array([lib().o_scale_factor_of_chi(self.h, float(c), C.byref(st))
                        for c in np.atleast_1d(chi)])
        return out, st.value

    def __del__(self):
        if getattr(self, "h", None):
            lib().o_cosmo_free(self.h)


class F2D:
    """ccl_f2d_t (src/ccl_f2d.c:92-201); fka is [na, nk] row-major like the reference."""

    def __init__(self, a=None, lk=None, fka=None, fk=None, fa=None, is_factorizable=False,
                 extrap_order_lok=1, extrap_order_hik=2, extrap_linear_growth=F2D_CCLGROWTH,
                 is_log=False, growth_factor_0=1.0, growth_exponent=0):
        self._a, pa = _arr(a)
        self._lk, plk = _arr(lk)
        self._fka, pfka = _arr(None if fka is None else np.asarray(fka).ravel())
        self._fk, pfk = _arr(fk)
        self._fa, pfa = _arr(fa)
        st = C.c_int(0)
        self.h = lib().o_f2d_new(0 if a is None else len(self._a), pa,
                                 0 if lk is None else len(self._lk), plk, pfka, pfk, pfa,
                                 int(is_factorizable), extrap_order_lok, extrap_order_hik,
                                 extrap_linear_growth, int(is_log), growth_factor_0,
                                 growth_exponent, C.byref(st))
        self.status = st.value

    def __call__(self, lk, a, cosmo=None):
        st = C.c_int(0)
        v = lib().o_f2d_eval(self.h, float(lk), float(a), cosmo.h if cosmo else None, C.byref(st))
        return v, st.value

    def dlogf_dlk(self, lk, a, cosmo=None):
        """ccl_f2d_t_dlogf_dlk_eval (src/ccl_f2d.c:375-531)."""
        st = C.c_int(0)
        v = lib().o_f2d_dlogf_dlk_eval(self.h, float(lk), float(a), cosmo.h if cosmo else None, C.byref(st))
        return v, st.value

    def __del__(self):
        if getattr(self, "h", None):
            lib().o_f2d_free(self.h)


class F3D:
    """ccl_f3d_t (src/ccl_f3d.c:161-271) with the flags of tk3d_new_from_arrays / tk3d_new_factorizable
    (pyccl/ccl_tk3d.i): constant growth, growth_factor_0 = 1, exponent 4.  tkka is [na, nk, nk]."""

    def __init__(self, a, lk, tkka=None, fka1=None, fka2=None, extrap_order_lok=1, extrap_order_hik=1, is_log=True,
                 extrap_linear_growth=F2D_CONSTGROWTH, growth_factor_0=1.0, growth_exponent=4):
        self._a, pa = _arr(a)
        self._lk, plk = _arr(lk)
        self._t, pt = _arr(None if tkka is None else np.asarray(tkka).ravel())
        self._f1, p1 = _arr(None if fka1 is None else np.asarray(fka1).ravel())
        self._f2, p2 = _arr(None if fka2 is None else np.asarray(fka2).ravel())
        st = C.c_int(0)
        self.h = lib().o_f3d_new(len(self._a), pa, len(self._lk), plk, pt, p1, p2, int(tkka is None), extrap_order_lok,
                                 extrap_order_hik, extrap_linear_growth, int(is_log), growth_factor_0,
                                 growth_exponent, C.byref(st))
        self.status = st.value

    def __call__(self, lk1, lk2, a, cosmo=None):
        st = C.c_int(0)
        v = lib().o_f3d_eval(self.h, float(lk1), float(lk2), float(a), cosmo.h if cosmo else None, C.byref(st))
        return v, st.value

    def __del__(self):
        if getattr(self, "h", None):
            lib().o_f3d_free(self.h)


def angular_cl_cov(cosmo, trc1, trc2, trc3, trc4, tsp, ell1, ell2, method=INTEG_QAG, chi_exponent=6,
                   kernel_extra=None, prefactor=1.0, dchi=5.0):
    """ccl_angular_cl_covariance (src/ccl_cls.c:490-612).  Returns (cov[nl2, nl1], status): the flat output of the
    reference is cov_out[l1 + nl1 * l2]."""
    ell1 = np.ascontiguousarray(np.atleast_1d(ell1), dtype=np.float64)
    ell2 = np.ascontiguousarray(np.atleast_1d(ell2), dtype=np.float64)
    out = np.full(ell1.size * ell2.size, np.nan)
    ka = kf = None
    nker = 0
    if kernel_extra is not None:
        (ka, pka), (kf, pkf) = _arr(kernel_extra[0]), _arr(kernel_extra[1])
        nker = len(ka)
    else:
        pka = pkf = None
    st = C.c_int(0)
    lib().o_angular_cl_covariance(cosmo.h, len(trc1), _coll(trc1), len(trc2), _coll(trc2), len(trc3), _coll(trc3),
                                  len(trc4), _coll(trc4), tsp.h, ell1.size, ell1.ctypes.data_as(_dp), ell2.size,
                                  ell2.ctypes.data_as(_dp), out.ctypes.data_as(_dp), int(method), int(chi_exponent),
                                  nker, pka, pkf, float(prefactor), float(dchi), C.byref(st))
    return out.reshape(ell2.size, ell1.size), st.value


class Tracer:
    """One ccl_cl_tracer_t (src/ccl_tracers.c:569-691)."""

    def __init__(self, der_bessel=0, der_angles=0, kernel=None, transfer_ka=None, transfer_k=None,
                 transfer_a=None, is_logt=False, extrap_order_lok=0, extrap_order_hik=2,
                 chi_max_nokernel=1e15):
        chi_w = w_w = a_s = lk_s = tka = tk = ta = None
        if kernel is not None:
            chi_w, w_w = kernel
        is_fact = transfer_ka is None
        if transfer_ka is not None:
            a_s, lk_s, tka = transfer_ka
            tka = np.asarray(tka).ravel()
        else:
            if transfer_k is not None:
                lk_s, tk = transfer_k
            if transfer_a is not None:
                a_s, ta = transfer_a
        self._keep = [_arr(v) for v in (chi_w, w_w, a_s, lk_s, tka, tk, ta)]
        p = [k[1] for k in self._keep]
        st = C.c_int(0)
        self.h = lib().o_tracer_new(int(der_bessel), int(der_angles),
                                    0 if chi_w is None else len(chi_w), p[0], p[1],
                                    0 if a_s is None else len(a_s), p[2],
                                    0 if lk_s is None else len(lk_s), p[3], p[4], p[5], p[6],
                                    int(is_logt), int(is_fact), extrap_order_lok, extrap_order_hik,
                                    float(chi_max_nokernel), C.byref(st))
        self.status = st.value
        if not self.h:
            raise ValueError("oracle tracer construction failed, status %d" % st.value)

    @property
    def chi_min(self):
        return lib().o_tracer_chi_min(self.h)

    @property
    def chi_max(self):
        return lib().o_tracer_chi_max(self.h)

    def f_ell(self, ell):
        return np.array([lib().o_tracer_f_ell(self.h, float(l)) for l in np.atleast_1d(ell)])

    def kernel(self, chi):
        return np.array([lib().o_tracer_kernel(self.h, float(c)) for c in np.atleast_1d(chi)])

    def transfer(self, lk, a):
        st = C.c_int(0)
        return lib().o_tracer_transfer(self.h, float(lk), float(a), C.byref(st)), st.value

    def __del__(self):
        if getattr(self, "h", None):
            lib().o_tracer_free(self.h)


def _coll(trs):
    arr = (C.c_void_p * len(trs))(*[t.h for t in trs])
    return arr


def angular_cl(cosmo, trc1, trc2, psp, ells, method=INTEG_SPLINE, each=False, return_neval=False):
    """ccl_angular_cls_limber (src/ccl_cls.c:265-363) for one pair of tracer collections.

    trc1/trc2 are lists of oracle Tracer objects (one pyccl Tracer = one collection).
    Returns (cl, status); with each=True status is a per-ell array and every ell is evaluated
    independently.
    """
    L = lib()
    ells = np.ascontiguousarray(ells, dtype=np.float64)
    cl = np.full(ells.size, np.nan)
    c1, c2 = _coll(trc1), _coll(trc2)
    if each:
        st = np.zeros(ells.size, dtype=np.int32)
        L.o_angular_cls_limber_each(cosmo.h, len(trc1), c1, len(trc2), c2, psp.h, ells.size,
                                    ells.ctypes.data_as(_dp), cl.ctypes.data_as(_dp), method,
                                    st.ctypes.data_as(_ip))
        return cl, st
    st = C.c_int(0)
    nev = np.zeros(ells.size, dtype=np.int64)
    L.o_angular_cls_limber(cosmo.h, len(trc1), c1, len(trc2), c2, psp.h, ells.size,
                           ells.ctypes.data_as(_dp), cl.ctypes.data_as(_dp), method,
                           nev.ctypes.data_as(C.POINTER(C.c_long)), C.byref(st))
    if return_neval:
        return cl, st.value, nev
    return cl, st.value


CORR_LGNDRE, CORR_FFTLOG, CORR_BESSEL = 1001, 1002, 1003       # include/ccl_correlation.h:6-8
CORR_GG, CORR_GL, CORR_LP, CORR_LM = 2001, 2002, 2003, 2004      # :9-12
CORR_METHODS = {"fftlog": CORR_FFTLOG, "bessel": CORR_BESSEL, "legendre": CORR_LGNDRE}
CORR_TYPES = {"NN": CORR_GG, "NG": CORR_GL, "GG+": CORR_LP, "GG-": CORR_LM}


def correlation(ell, cls, theta, corr_type="NN", method="fftlog", taper_cl_limits=None, ell_min_corr=0.01,
                ell_max_corr=60000.0, n_ell_corr=5000, epsrel=1e-4, n_iteration=1000):
    """ccl_correlation (src/ccl_correlation.c:410-428); theta in degrees.  Returns (wtheta, status)."""
    ell = np.ascontiguousarray(ell, dtype=np.float64)
    cls = np.ascontiguousarray(cls, dtype=np.float64)
    theta = np.ascontiguousarray(np.atleast_1d(theta), dtype=np.float64)
    out = np.full(theta.size, np.nan)
    st = C.c_int(0)
    tl, ptl = _arr(taper_cl_limits)
    lib().o_correlation(ell.size, ell.ctypes.data_as(_dp), cls.ctypes.data_as(_dp), theta.size,
                        theta.ctypes.data_as(_dp), out.ctypes.data_as(_dp), CORR_TYPES.get(corr_type, corr_type),
                        int(taper_cl_limits is not None), ptl, CORR_METHODS.get(method, method), float(ell_min_corr),
                        float(ell_max_corr), int(n_ell_corr), float(epsrel), int(n_iteration), C.byref(st))
    return out, st.value


def lngamma_complex(zr, zi):
    """gsl_sf_lngamma_complex_e restated: (ln|Gamma(z)|, arg Gamma(z) in (-pi, pi])."""
    a, b = C.c_double(0), C.c_double(0)
    lib().o_lngamma_complex(float(zr), float(zi), C.byref(a), C.byref(b))
    return a.value, b.value


def fftlog_fht(k, pk, dim, mu, q=0.0, kcrc=1.0, noring=1):
    """fht of src/ccl_fftlog.c:199-328 for one array: returns (r, xi)."""
    k = np.ascontiguousarray(k, dtype=np.float64)
    pk = np.ascontiguousarray(pk, dtype=np.float64)
    r, xi = np.empty_like(k), np.empty_like(k)
    lib().o_fftlog_fht(k.size, k.ctypes.data_as(_dp), pk.ctypes.data_as(_dp), r.ctypes.data_as(_dp),
                       xi.ctypes.data_as(_dp), float(dim), float(mu), float(q), float(kcrc), int(noring))
    return r, xi


def fftlog_transform_general(rs, frs, mu, q, spherical_bessel, bessel_deriv, plaw):
    """_fftlog_transform_general (pyccl/pyutils.py:567-603) -> ccl_fftlog_ComputeXi_general: returns (ks, fks)."""
    rs = np.ascontiguousarray(rs, dtype=np.float64)
    f2 = np.ascontiguousarray(np.atleast_2d(frs), dtype=np.float64)
    ks, fks = np.empty_like(rs), np.empty_like(f2)
    lib().o_fftlog_general(f2.shape[0], rs.size, rs.ctypes.data_as(_dp), f2.ctypes.data_as(_dp),
                           ks.ctypes.data_as(_dp), fks.ctypes.data_as(_dp), float(mu), float(q),
                           int(spherical_bessel), float(bessel_deriv), float(plaw))
    return ks, (fks[0] if np.ndim(frs) == 1 else fks)


def set_lkmin_ulps(n):
    """Test knob: shift the lower integration limit by n ulps (see o_set_lkmin_ulps in the C file)."""
    lib().o_set_lkmin_ulps(int(n))


def k_interval(cosmo, trc1, trc2, ell):
    lo, hi = C.c_double(), C.c_double()
    lib().o_get_k_interval_pub(cosmo.h, len(trc1), _coll(trc1), len(trc2), _coll(trc2), float(ell),
                               C.byref(lo), C.byref(hi))
    return lo.value, hi.value


def limber_nk(lkmin, lkmax, dlogk=0.025):
    return lib().o_limber_nk(lkmin, lkmax, dlogk)


def integ_spline(x, y):
    x, px = _arr(x)
    y, py = _arr(y)
    r = C.c_double()
    st = C.c_int(0)
    lib().o_integ_spline(len(x), px, py, 1.0, -1.0, C.byref(r), C.byref(st))
    return r.value, st.value


def qag41_test(kind, p, a, b, epsabs=0.0, epsrel=1e-4, limit=1000):
    r, e = C.c_double(), C.c_double()
    n = C.c_int()
    st = lib().o_qag41_test(kind, p, a, b, epsabs, epsrel, limit, C.byref(r), C.byref(e), C.byref(n))
    return r.value, e.value, n.value, st


def cquad_test(kind, p, a, b, epsabs=0.0, epsrel=1e-4, wsize=1000):
    """gsl_integration_cquad restatement on the test functions of qag41_test: result, abserr, nevals, status."""
    r, e = C.c_double(), C.c_double()
    n = C.c_int()
    st = lib().o_cquad_test(kind, p, a, b, epsabs, epsrel, wsize, C.byref(r), C.byref(e), C.byref(n))
    return r.value, e.value, n.value, st


def set_num_threads(n):
    """Use n OpenMP threads (launchers such as torchrun export OMP_NUM_THREADS=1)."""
    lib().o_set_num_threads(int(n))


def num_threads():
    return lib().o_num_threads()


def set_force_cquad(on):
    """Send every 'qag_quad' C_ell through the CQUAD fallback (src/ccl_cls.c:245-259) instead of QAG."""
    lib().o_set_force_cquad(int(bool(on)))
