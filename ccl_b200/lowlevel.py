"""Thin object wrappers over the C ABI (one Python object per C handle).

These are what pyccl's SWIG proxies are for the reference: ``Cosmology.cosmo``, ``Pk2D.psp``,
``Tracer._trc[i]`` hold one of these.  All numerical work happens in libccl_b200.so on the GPU.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import darr, iarr


class CCLB200Error(RuntimeError):
    pass


def _check(status, where=""):
    if status != 0:
        raise CCLB200Error("status %d in %s: %s" % (status, where, _lib.last_error()))


def make_params(K_MAX=1e3, K_MIN=5e-5, DLOGK_INTEGRATION=0.025, INTEGRATION_LIMBER_EPSREL=1e-4,
                N_ITERATION=1000, A_SPLINE_MIN=0.1, INTEGRATION_LIMBER_GAUSS_KRONROD_POINTS=4,
                DCHI_INTEGRATION=5.0, ELL_MIN_CORR=0.01, ELL_MAX_CORR=60000.0, N_ELL_CORR=5000,
                INTEGRATION_GAUSS_KRONROD_POINTS=4, INTEGRATION_EPSREL=1e-4):
    return _lib.Params(K_MAX, K_MIN, DLOGK_INTEGRATION, INTEGRATION_LIMBER_EPSREL, N_ITERATION,
                       INTEGRATION_LIMBER_GAUSS_KRONROD_POINTS, A_SPLINE_MIN, DCHI_INTEGRATION, ELL_MIN_CORR,
                       ELL_MAX_CORR, int(N_ELL_CORR), int(INTEGRATION_GAUSS_KRONROD_POINTS), INTEGRATION_EPSREL)


def _tracer_args(sub):
    """Map a sub-tracer dict (keys like Tracer.add_tracer) to ccl_cl_tracer_t_new's argument list."""
    chi, pchi = darr(sub.get("chi"))
    w, pw = darr(sub.get("w"))
    a, pa = darr(sub.get("a"))
    lk, plk = darr(sub.get("lk"))
    fka = sub.get("fka")
    fka, pfka = darr(None if fka is None else np.asarray(fka).ravel())
    fk, pfk = darr(sub.get("fk"))
    fa, pfa = darr(sub.get("fa"))
    keep = (chi, w, a, lk, fka, fk, fa)
    is_fact = int(fka is None)
    args = [int(sub.get("der_bessel", 0)), int(sub.get("der_angles", 0)),
            0 if chi is None else len(chi), pchi, pw,
            0 if a is None else len(a), pa, 0 if lk is None else len(lk), plk, pfka, pfk, pfa,
            int(sub.get("is_logt", False)), is_fact, int(sub.get("extrap_order_lok", 0)),
            int(sub.get("extrap_order_hik", 2))]
    return keep, args


class LibCosmology:
    def __init__(self, params=None):
        L = _lib.load()
        st = C.c_int(0)
        self.params = params or make_params()
        self.h = L.ccl_b200_cosmology_new(C.byref(self.params), C.byref(st))
        _check(st.value, "cosmology_new")

    def set_achi(self, chi, a):
        st = C.c_int(0)
        k1, p1 = darr(chi)
        k2, p2 = darr(a)
        _lib.load().ccl_b200_cosmology_set_achi(self.h, len(k1), p1, p2, C.byref(st))
        _check(st.value, "set_achi")

    def distances_from_input(self, a, chi):
        st = C.c_int(0)
        k1, p1 = darr(a)
        k2, p2 = darr(chi)
        _lib.load().ccl_b200_cosmology_distances_from_input(self.h, len(k1), p1, p2, C.byref(st))
        _check(st.value, "distances_from_input")

    def set_growth(self, a, growth):
        st = C.c_int(0)
        k1, p1 = darr(a)
        k2, p2 = darr(growth)
        _lib.load().ccl_b200_cosmology_set_growth(self.h, len(k1), p1, p2, C.byref(st))
        _check(st.value, "set_growth")

    def set_chi_max_nokernel(self, chi):
        _lib.load().ccl_b200_cosmology_set_chi_max_nokernel(self.h, float(chi))

    def scale_factor_of_chi(self, chi):
        k, p = darr(np.atleast_1d(chi))
        out = np.empty_like(k)
        st = C.c_int(0)
        _lib.load().ccl_b200_scale_factor_of_chis(self.h, len(k), p, out.ctypes.data_as(_lib._dp), C.byref(st))
        return out, st.value

    def __del__(self):
        if getattr(self, "h", None) and _lib._LIB is not None:
            _lib._LIB.ccl_b200_cosmology_free(self.h)
            self.h = None


class LibF2D:
    def __init__(self, a=None, lk=None, fka=None, fk=None, fa=None, is_factorizable=False,
                 extrap_order_lok=1, extrap_order_hik=2, extrap_linear_growth=_lib.F2D_CCLGROWTH,
                 is_log=False, growth_factor_0=1.0, growth_exponent=0):
        ka, pa = darr(a)
        kl, pl = darr(lk)
        kf, pf = darr(None if fka is None else np.asarray(fka).ravel())
        kk, pk = darr(fk)
        kz, pz = darr(fa)
        st = C.c_int(0)
        self.h = _lib.load().ccl_b200_f2d_new(0 if a is None else len(ka), pa, 0 if lk is None else len(kl), pl,
                                              pf, pk, pz, int(is_factorizable), extrap_order_lok,
                                              extrap_order_hik, extrap_linear_growth, int(is_log),
                                              growth_factor_0, growth_exponent, C.byref(st))
        self.status = st.value
        _check(st.value, "f2d_new")

    def eval(self, lk, a, cosmo=None, derivative=False):
        """pk2d_eval_multi / pk2d_der_eval_multi (pyccl/ccl_pk2d.i:48-64): f or d ln f / d ln k at (lk[], a)."""
        k, p = darr(np.atleast_1d(lk))
        out = np.empty_like(k)
        st = C.c_int(0)
        fn = _lib.load().ccl_b200_f2d_dlogf_dlk_eval_multi if derivative else _lib.load().ccl_b200_f2d_eval_multi
        fn(self.h, p, len(k), float(a), cosmo.h if cosmo is not None else None, out.ctypes.data_as(_lib._dp),
           C.byref(st))
        return out, st.value

    def __del__(self):
        if getattr(self, "h", None) and _lib._LIB is not None:
            _lib._LIB.ccl_b200_f2d_free(self.h)
            self.h = None


class LibTracer:
    """One ccl_cl_tracer_t."""

    def __init__(self, cosmo, **sub):
        keep, args = _tracer_args(sub)
        st = C.c_int(0)
        self.h = _lib.load().ccl_b200_cl_tracer_t_new(cosmo.h if cosmo is not None else None, *args, C.byref(st))
        _check(st.value, "cl_tracer_t_new")
        self.der_bessel = args[0]
        self.der_angles = args[1]

    @property
    def chi_min(self):
        return _lib.load().ccl_b200_cl_tracer_t_chi_min(self.h)

    @property
    def chi_max(self):
        return _lib.load().ccl_b200_cl_tracer_t_chi_max(self.h)

    def f_ell(self, ell):
        k, p = darr(np.atleast_1d(ell))
        out = np.empty_like(k)
        st = C.c_int(0)
        _lib.load().ccl_b200_cl_tracer_get_f_ell(self.h, len(k), p, out.ctypes.data_as(_lib._dp), C.byref(st))
        _check(st.value, "get_f_ell")
        return out

    def kernel(self, chi):
        k, p = darr(np.atleast_1d(chi))
        out = np.empty_like(k)
        st = C.c_int(0)
        _lib.load().ccl_b200_cl_tracer_get_kernel(self.h, len(k), p, out.ctypes.data_as(_lib._dp), C.byref(st))
        _check(st.value, "get_kernel")
        return out

    def transfer(self, lk, a):
        kl, pl = darr(np.atleast_1d(lk))
        ka, pa = darr(np.atleast_1d(a))
        out = np.empty(len(kl) * len(ka))
        st = C.c_int(0)
        _lib.load().ccl_b200_cl_tracer_get_transfer(self.h, len(kl), pl, len(ka), pa,
                                                    out.ctypes.data_as(_lib._dp), C.byref(st))
        _check(st.value, "get_transfer")
        return out.reshape(len(kl), len(ka))

    def __del__(self):
        if getattr(self, "h", None) and _lib._LIB is not None:
            _lib._LIB.ccl_b200_cl_tracer_t_free(self.h)
            self.h = None


def angular_cls_limber(cosmo, trs1, trs2, psp, ells, method=_lib.INTEGRATION_SPLINE):
    """ccl_angular_cls_limber through the single-pair C entry.  Returns (cl, status)."""
    L = _lib.load()
    st = C.c_int(0)
    c1 = L.ccl_b200_cl_tracer_collection_t_new(C.byref(st))
    same = trs1 is trs2
    c2 = c1 if same else L.ccl_b200_cl_tracer_collection_t_new(C.byref(st))
    for t in trs1:
        L.ccl_b200_add_cl_tracer_to_collection(c1, t.h, C.byref(st))
    if not same:
        for t in trs2:
            L.ccl_b200_add_cl_tracer_to_collection(c2, t.h, C.byref(st))
    ke, pe = darr(np.atleast_1d(ells))
    cl = np.empty(len(ke))
    L.ccl_b200_angular_cls_limber(cosmo.h, c1, c2, psp.h, len(ke), pe, cl.ctypes.data_as(_lib._dp), int(method),
                                  C.byref(st))
    L.ccl_b200_cl_tracer_collection_t_free(c1)
    if not same:
        L.ccl_b200_cl_tracer_collection_t_free(c2)
    return cl, st.value


class LibF3D:
    """ccl_b200_f3d: a ccl_f3d_t on the device (tk3d_new_from_arrays / tk3d_new_factorizable flags by default)."""

    def __init__(self, a, lk, tkka=None, fka1=None, fka2=None, extrap_order_lok=1, extrap_order_hik=1, is_log=True,
                 extrap_linear_growth=_lib.F2D_CONSTANTGROWTH, growth_factor_0=1.0, growth_exponent=4):
        ka, pa = darr(a)
        kl, pl = darr(lk)
        kt, pt = darr(None if tkka is None else np.asarray(tkka).ravel())
        k1, p1 = darr(None if fka1 is None else np.asarray(fka1).ravel())
        k2, p2 = darr(None if fka2 is None else np.asarray(fka2).ravel())
        st = C.c_int(0)
        self.h = _lib.load().ccl_b200_f3d_new(len(ka), pa, len(kl), pl, pt, p1, p2, int(tkka is None),
                                              int(extrap_order_lok), int(extrap_order_hik), int(extrap_linear_growth),
                                              int(is_log), float(growth_factor_0), int(growth_exponent), C.byref(st))
        _check(st.value, "f3d_new")

    def __del__(self):
        if getattr(self, "h", None) and _lib._LIB is not None:
            _lib._LIB.ccl_b200_f3d_free(self.h)
            self.h = None


def angular_cl_covariance(cosmo, trs1, trs2, trs3, trs4, tsp, ell1, ell2, method=_lib.INTEGRATION_QAG_QUAD,
                          chi_exponent=6, kernel_extra=None, prefactor=1.0):
    """ccl_angular_cl_covariance through the C entry.  Returns (cov[nl2, nl1], status)."""
    L = _lib.load()
    st = C.c_int(0)
    colls = []
    for trs in (trs1, trs2, trs3, trs4):
        c = L.ccl_b200_cl_tracer_collection_t_new(C.byref(st))
        for t in trs:
            L.ccl_b200_add_cl_tracer_to_collection(c, t.h, C.byref(st))
        colls.append(c)
    k1, p1 = darr(np.atleast_1d(ell1))
    k2, p2 = darr(np.atleast_1d(ell2))
    ka, pa = darr(None if kernel_extra is None else kernel_extra[0])
    kf, pf = darr(None if kernel_extra is None else kernel_extra[1])
    out = np.empty(len(k1) * len(k2))
    L.ccl_b200_angular_cl_covariance(cosmo.h, colls[0], colls[1], colls[2], colls[3], tsp.h, len(k1), p1, len(k2), p2,
                                     out.ctypes.data_as(_lib._dp), int(method), int(chi_exponent),
                                     0 if ka is None else len(ka), pa, pf, float(prefactor), C.byref(st))
    for c in colls:
        L.ccl_b200_cl_tracer_collection_t_free(c)
    return out.reshape(len(k2), len(k1)), st.value


def fftlog_transform_general(rs, frs, mu, q, spherical_bessel, bessel_deriv, plaw):
    """_fftlog_transform_general (pyccl/pyutils.py:567-603) on the device.  `mu` may be an array: the same
    input arrays are then transformed for every mu in one call.  Returns (ks, fks): [N] and frs.shape for a
    scalar mu, with a leading mu axis otherwise."""
    if np.ndim(rs) != 1:
        raise ValueError("rs should be a 1D array")
    if np.ndim(frs) < 1 or np.ndim(frs) > 2:
        raise ValueError("frs should be a 1D or 2D array")
    f2 = np.ascontiguousarray(np.atleast_2d(frs), dtype=np.float64)
    n_transforms, n_r = f2.shape
    if len(rs) != n_r:
        raise ValueError("rs should have %d elements" % n_r)
    if round(bessel_deriv) != bessel_deriv or bessel_deriv < 0:
        raise CCLB200Error("bessel_deriv must be a non-negative integer in a floating-point format.")
    if spherical_bessel not in (0, 1):
        raise CCLB200Error("0 and 1 are the only valid values of spherical_bessel.")
    mus = np.ascontiguousarray(np.atleast_1d(mu), dtype=np.float64)
    kr, pr = darr(rs)
    out = np.empty((mus.size, n_transforms + 1, n_r))
    st = C.c_int(0)
    _lib.load().ccl_b200_fftlog_transform_general(mus.size, mus.ctypes.data_as(_lib._dp), n_transforms, pr, n_r,
                                                  f2.ctypes.data_as(_lib._dp), float(q), int(spherical_bessel),
                                                  float(bessel_deriv), float(plaw), out.ctypes.data_as(_lib._dp),
                                                  C.byref(st))
    _check(st.value, "ccl_b200_fftlog_transform_general")
    ks, fks = out[:, 0, :], out[:, 1:, :]
    if np.ndim(frs) == 1:
        fks = fks[:, 0, :]
    if np.ndim(mu) == 0:
        ks, fks = ks[0], fks[0]
    return ks, fks


def correlation(cosmo, ell, cls, theta, corr_type, method, taper_cl_limits=None):
    """ccl_correlation through the C entry (theta in degrees).  Returns (wtheta, status)."""
    L = _lib.load()
    st = C.c_int(0)
    ke, pe = darr(ell)
    kc, pc = darr(cls)
    kt, pt = darr(np.atleast_1d(theta))
    kl, pl = darr(taper_cl_limits)
    out = np.full(len(kt), np.nan)
    L.ccl_b200_correlation(cosmo.h, len(ke), pe, pc, len(kt), pt, out.ctypes.data_as(_lib._dp), int(corr_type),
                           int(kl is not None), pl, int(method), C.byref(st))
    return out, st.value


class LibBatch:
    """ccl_b200_batch: N cosmologies' tables in one HBM arena + the batched Limber entry."""

    def __init__(self, ncosmo, params=None):
        st = C.c_int(0)
        self.params = params or make_params()
        self.ncosmo = ncosmo
        self.h = _lib.load().ccl_b200_batch_new(ncosmo, C.byref(self.params), C.byref(st))
        _check(st.value, "batch_new")
        self._geom = None

    def reset(self):
        self._sources = []
        _lib.load().ccl_b200_batch_reset(self.h)

    def set_cosmology(self, ic, chi, a, a_growth=None, growth=None, chi_max_nokernel=None):
        L = _lib.load()
        st = C.c_int(0)
        k1, p1 = darr(chi)
        k2, p2 = darr(a)
        L.ccl_b200_batch_set_achi(self.h, ic, len(k1), p1, p2, C.byref(st))
        _check(st.value, "batch_set_achi")
        if a_growth is not None:
            k3, p3 = darr(a_growth)
            k4, p4 = darr(growth)
            L.ccl_b200_batch_set_growth(self.h, ic, len(k3), p3, p4, C.byref(st))
            _check(st.value, "batch_set_growth")
        if chi_max_nokernel is not None:
            L.ccl_b200_batch_set_chi_max_nokernel(self.h, ic, float(chi_max_nokernel))

    def set_pk2d(self, ic, lk, a, pk, order_lok=1, order_hik=2, is_logp=True):
        st = C.c_int(0)
        k1, p1 = darr(lk)
        k2, p2 = darr(a)
        k3, p3 = darr(np.asarray(pk).ravel())
        _lib.load().ccl_b200_batch_set_pk2d(self.h, ic, p1, len(k1), p2, len(k2), p3, order_lok, order_hik,
                                            int(is_logp), C.byref(st))
        _check(st.value, "batch_set_pk2d")

    def add_tracer(self, ic, **sub):
        keep, args = _tracer_args(sub)
        st = C.c_int(0)
        idx = _lib.load().ccl_b200_batch_add_tracer(self.h, ic, *args, C.byref(st))
        _check(st.value, "batch_add_tracer")
        return idx

    # ---- stacked setters (one C call per table kind for a whole range of slots) ----
    def set_stacked(self, first, stk):
        """Fill slots [first, first + count) from stacked host arrays.

        `stk` = dict(count, achi=(n_per, chi[count, nmax], a[count, nmax]), growth=(a[n] or [count, n], D[count, n]),
        chi_max_nokernel=[count], pk=(lk[nk], a[na], pk[count, na, nk], order_lok, order_hik, is_logp),
        tracers=[dict(der_bessel, der_angles, chi=[n] or [count, n], w=[count, n], a=..., fa=[count, na], ...)]).
        """
        L = _lib.load()
        st = C.c_int(0)
        count = int(stk["count"])
        # pinned source arrays are read by commit(), not here: keep them alive until then
        self._sources = getattr(self, "_sources", [])
        self._sources.append(stk)
        n_per, chi, a = stk["achi"]
        kn, pn = iarr(n_per)
        kc, pc = darr(chi)
        ka, pa = darr(a)
        assert kc.shape == ka.shape == (count, kc.shape[1])
        L.ccl_b200_batch_set_achi_stacked(self.h, first, count, pn, kc.shape[1], pc, pa, C.byref(st))
        _check(st.value, "batch_set_achi_stacked")
        if stk.get("growth") is not None:
            ag, D = stk["growth"]
            kg, pg = darr(ag)
            kd, pd_ = darr(D)
            L.ccl_b200_batch_set_growth_stacked(self.h, first, count, kd.shape[1], 0 if kg.ndim == 1 else kg.shape[1],
                                                pg, pd_, C.byref(st))
            _check(st.value, "batch_set_growth_stacked")
        if stk.get("chi_max_nokernel") is not None:
            kx, px = darr(stk["chi_max_nokernel"])
            L.ccl_b200_batch_set_chi_max_nokernel_stacked(self.h, first, count, px)
        lk, apk, pk, olo, ohi, islog = stk["pk"]
        kl, pl = darr(lk)
        kap, pap = darr(apk)
        kp, pp = darr(pk)
        assert kp.shape == (count, len(kap), len(kl))
        L.ccl_b200_batch_set_pk2d_stacked(self.h, first, count, pl, len(kl), pap, len(kap), pp, int(olo), int(ohi),
                                          int(islog), C.byref(st))
        _check(st.value, "batch_set_pk2d_stacked")
        for t in stk["tracers"]:
            chi_w, w = t.get("chi"), t.get("w")
            kcw, pcw = darr(chi_w)
            kw, pw = darr(w)
            kta, pta = darr(t.get("a"))
            ktl, ptl = darr(t.get("lk"))
            kfka, pfka = darr(t.get("fka"))
            kfk, pfk = darr(t.get("fk"))
            kfa, pfa = darr(t.get("fa"))
            n_w = 0 if kw is None else kw.shape[1]
            na = 0 if kta is None else kta.shape[-1]
            nk = 0 if ktl is None else len(ktl)
            L.ccl_b200_batch_add_tracer_stacked(
                self.h, first, count, int(t.get("der_bessel", 0)), int(t.get("der_angles", 0)), n_w,
                0 if (kcw is None or kcw.ndim == 1) else kcw.shape[1], pcw, pw, na,
                0 if (kta is None or kta.ndim == 1) else kta.shape[1], pta, nk, ptl, pfka, pfk, pfa,
                int(t.get("is_logt", False)), int(kfka is None), int(t.get("extrap_order_lok", 0)),
                int(t.get("extrap_order_hik", 2)), C.byref(st))
            _check(st.value, "batch_add_tracer_stacked")

    def tracer_chi_range(self, ic, it):
        lo, hi = C.c_double(), C.c_double()
        _lib.load().ccl_b200_batch_tracer_chi_range(self.h, ic, it, C.byref(lo), C.byref(hi))
        return lo.value, hi.value

    def commit(self):
        st = C.c_int(0)
        _lib.load().ccl_b200_batch_commit(self.h, C.byref(st))
        self._sources = []
        _check(st.value, "batch_commit")

    @property
    def h2d_bytes(self):
        return _lib.load().ccl_b200_batch_h2d_bytes(self.h)

    @property
    def arena_bytes(self):
        return _lib.load().ccl_b200_batch_arena_bytes(self.h)

    def _geometry(self, collections, pairs, ells):
        cs, pcs = iarr([c[0] for c in collections])
        cc, pcc = iarr([c[1] for c in collections])
        p1, pp1 = iarr([p[0] for p in pairs])
        p2, pp2 = iarr([p[1] for p in pairs])
        el, pel = darr(ells)
        return (cs, cc, p1, p2, el), (len(collections), pcs, pcc, len(pairs), pp1, pp2, len(el), pel)

    def angular_cl(self, collections, pairs, ells, method=_lib.INTEGRATION_SPLINE, out=None):
        """Host-buffer entry: returns (cl[ncosmo, npair, nl], status[ncosmo, npair], status).
        `out` may be a preallocated C-contiguous float64 array (e.g. a view of pinned memory)."""
        keep, g = self._geometry(collections, pairs, ells)
        shape = (self.ncosmo, len(pairs), len(keep[4]))
        if out is not None:
            assert out.shape == shape and out.dtype == np.float64 and out.flags.c_contiguous
            cl = out
        else:
            cl = np.empty(shape)
        stat = np.zeros((self.ncosmo, len(pairs)), dtype=np.int32)
        st = C.c_int(0)
        _lib.load().ccl_b200_batch_angular_cls_limber(self.h, *g, int(method), cl.ctypes.data_as(_lib._dp),
                                                      stat.ctypes.data_as(_lib._ip), C.byref(st))
        if st.value not in (0, 1030):
            _check(st.value, "batch_angular_cls_limber")
        return cl, stat, st.value

    def angular_cl_start(self, collections, pairs, ells, method=_lib.INTEGRATION_SPLINE, out=None):
        """First half of angular_cl: queue the kernels and the copy of every C_ell into `out` (a C-contiguous
        float64 array, page-locked for a truly asynchronous copy) on the batch's stream and return at once."""
        keep, g = self._geometry(collections, pairs, ells)
        shape = (self.ncosmo, len(pairs), len(keep[4]))
        if out is None:
            out = np.empty(shape)
        assert out.shape == shape and out.dtype == np.float64 and out.flags.c_contiguous
        st = C.c_int(0)
        _lib.load().ccl_b200_batch_angular_cls_limber_start(self.h, *g, int(method), out.ctypes.data_as(_lib._dp),
                                                            C.byref(st))
        _check(st.value, "batch_angular_cls_limber_start")
        self._started = (out, keep, len(pairs))
        return out

    def angular_cl_wait(self):
        """Second half: block until the started call has finished; returns (cl, status[ncosmo, npair], status)."""
        out, _, npair = self._started
        self._started = None
        stat = np.zeros((self.ncosmo, npair), dtype=np.int32)
        st = C.c_int(0)
        _lib.load().ccl_b200_batch_angular_cls_limber_wait(self.h, stat.ctypes.data_as(_lib._ip), C.byref(st))
        if st.value not in (0, 1030):
            _check(st.value, "batch_angular_cls_limber_wait")
        return out, stat, st.value

    def angular_cl_dev(self, collections, pairs, ells, d_cl, d_status, method=_lib.INTEGRATION_SPLINE, stream=0):
        """Device-buffer entry: d_cl / d_status are raw device pointers (ints), e.g. tensor.data_ptr()."""
        keep, g = self._geometry(collections, pairs, ells)
        st = C.c_int(0)
        _lib.load().ccl_b200_batch_angular_cls_limber_dev(self.h, *g, int(method), C.c_void_p(d_cl),
                                                          C.c_void_p(d_status), C.c_void_p(stream), C.byref(st))
        _check(st.value, "batch_angular_cls_limber_dev")

    def spline_nodes(self, collections, pairs, ells):
        keep, g = self._geometry(collections, pairs, ells)
        return _lib.load().ccl_b200_batch_spline_nodes(self.h, *g)

    @property
    def last_used_fast_path(self):
        """True if the last 'spline' launch ran the grid-sharing kernels (limber_fast.cu)."""
        return bool(_lib.load().ccl_b200_batch_last_used_fast_path(self.h))

    @property
    def last_qag_evals(self):
        return _lib.load().ccl_b200_batch_last_qag_evals(self.h)

    @property
    def last_qag_nodes_computed(self):
        return _lib.load().ccl_b200_batch_last_qag_nodes_computed(self.h)

    def __del__(self):
        if getattr(self, "h", None) and _lib._LIB is not None:
            _lib._LIB.ccl_b200_batch_free(self.h)
            self.h = None
