"""Batched device builders of the tables the Limber path reads (north-star item 5).

Python front-end of the ``ccl_b200_build_*`` entries: a list of ``Cosmology`` objects (or their
derived parameter dicts) in, stacked tables out.  The algorithms are those of the host builders
(``background.py``, ``power.py``, ``tracers.py``), which restate src/ccl_background.c,
src/ccl_power.c and src/ccl_tracers.c; tests/test_builders_gpu.py compares the two."""
import ctypes as C

import numpy as np

from . import _lib
from .errors import check
from .synthetic import linlog_spacing

_PAR_KEYS = ("Omega_c", "Omega_b", "Omega_l", "Omega_k", "Omega_g", "Omega_nu_rel", "w0", "wa", "h", "Omega_m",
             "n_s", "sigma8", "T_CMB", "k_sign", "sqrtk")
NPAR = 16


def pack_params(cosmologies):
    """[ncosmo, NPAR] parameter matrix (include/ccl_b200.h) from Cosmology objects or parameter dicts."""
    out = np.zeros((len(cosmologies), NPAR))
    for i, c in enumerate(cosmologies):
        p = c._params if hasattr(c, "_params") else c
        out[i, :len(_PAR_KEYS)] = [float(p[k]) for k in _PAR_KEYS]
    return out


def build_background(cosmologies, spline_params=None, gsl_params=None, stride=8192):
    """E(a), chi(a), a(chi), D(a), f(a) of every cosmology, built on the device.

    Returns dict(a, E[nc, na], chi[nc, na], n_achi[nc], achi_chi[nc, stride], achi_a[nc, stride],
    growth[nc, na], fgrowth[nc, na], growth0[nc])."""
    from .parameters import gsl_params as gp0, spline_params as sp0
    sp = spline_params if spline_params is not None else sp0
    gp = gsl_params if gsl_params is not None else gp0
    par = np.ascontiguousarray(pack_params(cosmologies))
    a = linlog_spacing(sp["A_SPLINE_MINLOG"], sp["A_SPLINE_MIN"], sp["A_SPLINE_MAX"], sp["A_SPLINE_NLOG"],
                       sp["A_SPLINE_NA"])
    nc, na = par.shape[0], a.size
    out = dict(a=a, E=np.empty((nc, na)), chi=np.empty((nc, na)), n_achi=np.zeros(nc, dtype=np.int32),
               achi_chi=np.zeros((nc, stride)), achi_a=np.zeros((nc, stride)), growth=np.empty((nc, na)),
               fgrowth=np.empty((nc, na)), growth0=np.empty(nc))
    st = C.c_int(0)
    dp = lambda x: x.ctypes.data_as(_lib._dp)
    _lib.load().ccl_b200_build_background(nc, dp(par), na, dp(a), float(sp["DCHI_INTEGRATION"]),
                                          float(gp["ROOT_EPSREL"]), int(gp["ROOT_N_ITERATION"]),
                                          float(gp["EPS_SCALEFAC_GROWTH"]), stride, dp(out["E"]), dp(out["chi"]),
                                          out["n_achi"].ctypes.data_as(_lib._ip), dp(out["achi_chi"]),
                                          dp(out["achi_a"]), dp(out["growth"]), dp(out["fgrowth"]),
                                          dp(out["growth0"]), C.byref(st))
    check(st.value)
    return out


_MODELS = {"bbks": 0, "eisenstein_hu": 1, "eisenstein_hu_nowiggles": 2}


def build_linear_power(cosmologies, model, background, spline_params=None):
    """log P_lin(k, a) tables [nc, na_pk, nk] on the default (k, a) grids, sigma8-normalised, built on the
    device from the device-built growth tables.  Returns dict(lk, a, logp, sigma8_unnorm)."""
    from .parameters import spline_params as sp0
    from .power import pk_grids
    sp = spline_params if spline_params is not None else sp0
    par = np.ascontiguousarray(pack_params(cosmologies))
    if np.any(~np.isfinite(par[:, 11])):
        raise ValueError("sigma8 not set, required for analytic power spectra")
    k, a_pk = pk_grids(sp)
    nc = par.shape[0]
    a_bg = np.ascontiguousarray(background["a"])
    D = np.ascontiguousarray(background["growth"])
    logp = np.empty((nc, a_pk.size, k.size))
    s8 = np.empty(nc)
    st = C.c_int(0)
    dp = lambda x: x.ctypes.data_as(_lib._dp)
    _lib.load().ccl_b200_build_linear_power(nc, dp(par), _MODELS[model], k.size, dp(k), a_pk.size, dp(a_pk),
                                            a_bg.size, dp(a_bg), dp(D), float(sp["K_MIN"]), float(sp["K_MAX"]),
                                            dp(logp), dp(s8), C.byref(st))
    check(st.value)
    return dict(lk=np.log(k), a=a_pk, logp=logp, sigma8_unnorm=s8)


def build_tracer_kernels(cosmologies, background, z, nz_list, mag_bias_list=None, n_chi=256):
    """Number-counts and lensing kernels of every (cosmology, n(z)) on the device.

    z[nz] is shared; nz_list is a sequence of n(z) arrays on it; mag_bias_list (optional) the
    magnification-bias functions s(z) on the same grid.  Returns dict(chi_z[nc, nz], w_nc[nc, ntr, nz],
    chi_l[nc, n_chi], w_l[nc, ntr, n_chi])."""
    from ._hostspline import Akima
    par = np.ascontiguousarray(pack_params(cosmologies))
    z = np.ascontiguousarray(z, dtype=float)
    nzs = np.ascontiguousarray(np.stack([np.asarray(n, dtype=float) for n in nz_list]))
    szs = None if mag_bias_list is None else np.ascontiguousarray(np.stack([np.asarray(s_, dtype=float) for s_ in mag_bias_list]))
    with np.errstate(divide='ignore'):   # an identically zero n(z): 1/0 = inf like the C arithmetic
        inv_norm = np.array([1.0 / np.float64(Akima(z, n).integrate_nodes()) for n in nzs])   # get_nz_norm, spline mode
    nc, ntr = par.shape[0], nzs.shape[0]
    a_bg = np.ascontiguousarray(background["a"])
    out = dict(chi_z=np.empty((nc, z.size)), w_nc=np.empty((nc, ntr, z.size)), chi_l=np.empty((nc, n_chi)),
               w_l=np.empty((nc, ntr, n_chi)))
    st = C.c_int(0)
    dp = lambda x: None if x is None else x.ctypes.data_as(_lib._dp)
    chi_bg = np.ascontiguousarray(background["chi"])
    n_achi = np.ascontiguousarray(background["n_achi"], dtype=np.int32)
    gx = np.ascontiguousarray(background["achi_chi"])
    ga = np.ascontiguousarray(background["achi_a"])
    _lib.load().ccl_b200_build_tracer_kernels(nc, dp(par), a_bg.size, dp(a_bg), dp(chi_bg),
                                              n_achi.ctypes.data_as(_lib._ip), gx.shape[1], dp(gx), dp(ga), z.size,
                                              dp(z), ntr, dp(nzs), dp(szs), dp(inv_norm), int(n_chi),
                                              dp(out["chi_z"]), dp(out["w_nc"]), dp(out["chi_l"]), dp(out["w_l"]),
                                              C.byref(st))
    check(st.value)
    return out


def build_halofit(cosmologies, linear):
    """log P_NL tables [nc, na, nk] from the linear tables of build_linear_power (or any stacked log P)."""
    par = np.ascontiguousarray(pack_params(cosmologies))
    lk = np.ascontiguousarray(linear["lk"])
    a = np.ascontiguousarray(linear["a"])
    lin = np.ascontiguousarray(linear["logp"])
    nc = par.shape[0]
    out = np.empty_like(lin)
    rs = np.empty((nc, a.size))
    st = C.c_int(0)
    dp = lambda x: x.ctypes.data_as(_lib._dp)
    _lib.load().ccl_b200_build_halofit(nc, dp(par), lk.size, dp(lk), a.size, dp(a), dp(lin), dp(out), dp(rs),
                                       C.byref(st))
    check(st.value)
    return dict(lk=lk, a=a, logp=out, rsigma=rs)


def build_3x2pt_stacked(cosmologies, z, lens_nz, source_nz, lens_bias=None, model="eisenstein_hu", halofit=True,
                        n_chi=256, spline_params=None, gsl_params=None):
    """Every table of a 3x2pt likelihood batch from cosmological parameters, built on the device.

    `cosmologies`: Cosmology objects (only their parameters are read); `z`: shared redshift grid;
    `lens_nz` / `source_nz`: n(z) arrays of the NumberCounts (density term, constant bias
    `lens_bias[i]`, no RSD) and WeakLensing (shear) tracers.  Returns (stk, collections) for
    LibBatch.set_stacked / PipelinedBatch: the same tables NumberCountsTracer(has_rsd=False, bias=...) and
    WeakLensingTracer(...) would hand to the C layer (pyccl/tracers.py:831-988), for all cosmologies.
    """
    from .parameters import spline_params as sp0
    sp = spline_params if spline_params is not None else sp0
    z = np.ascontiguousarray(z, dtype=float)
    nl, ns = len(lens_nz), len(source_nz)
    if lens_bias is None:
        lens_bias = np.ones(nl)
    bgd = build_background(cosmologies, sp, gsl_params)
    lin = build_linear_power(cosmologies, model, bgd, sp)
    pk = build_halofit(cosmologies, lin) if halofit else lin
    trk = build_tracer_kernels(cosmologies, bgd, z, list(lens_nz) + list(source_nz), None, n_chi)
    nc = len(cosmologies)
    from ._hostspline import Akima
    # chi(A_SPLINE_MIN): chi_max of kernel-less tracers (not used by these tracers, kept for completeness)
    cmax = np.array([float(Akima(bgd["a"], bgd["chi"][i])(sp["A_SPLINE_MIN"])) for i in range(nc)])
    nmax = int(bgd["n_achi"].max())
    stk = dict(count=nc, achi=(bgd["n_achi"], np.ascontiguousarray(bgd["achi_chi"][:, :nmax]),
                               np.ascontiguousarray(bgd["achi_a"][:, :nmax])),
               growth=(bgd["a"], bgd["growth"]), chi_max_nokernel=cmax,
               pk=(pk["lk"], pk["a"], pk["logp"], 1, 2, True), tracers=[])
    a_rev = np.ascontiguousarray(1. / (1 + z[::-1]))
    for i in range(nl):   # density term: kernel p(z) H(z), transfer b(a) (pyccl/tracers.py:879-886)
        stk["tracers"].append(dict(der_bessel=0, der_angles=0, chi=trk["chi_z"], w=np.ascontiguousarray(trk["w_nc"][:, i]),
                                   a=a_rev, fa=np.full((nc, z.size), float(lens_bias[i]))))
    for j in range(ns):   # shear: lensing kernel, der_bessel -1, der_angles 2 (pyccl/tracers.py:952-957)
        stk["tracers"].append(dict(der_bessel=-1, der_angles=2, chi=trk["chi_l"],
                                   w=np.ascontiguousarray(trk["w_l"][:, nl + j])))
    collections = [(i, 1) for i in range(nl + ns)]
    return stk, collections


class DeviceTables:
    """ccl_b200_tables: every table of a batch of cosmologies built and kept in HBM."""

    def __init__(self, cosmologies, z=None, nz_list=(), mag_bias_list=None, lens_scale=None, model="eisenstein_hu",
                 halofit=True, n_chi=256, spline_params=None, gsl_params=None, achi_stride=8192):
        from ._hostspline import Akima
        from .parameters import gsl_params as gp0, spline_params as sp0
        from .power import pk_grids
        sp = spline_params if spline_params is not None else sp0
        gp = gsl_params if gsl_params is not None else gp0
        par = np.ascontiguousarray(pack_params(cosmologies))
        a = linlog_spacing(sp["A_SPLINE_MINLOG"], sp["A_SPLINE_MIN"], sp["A_SPLINE_MAX"], sp["A_SPLINE_NLOG"],
                           sp["A_SPLINE_NA"])
        k, a_pk = pk_grids(sp)
        ntr = len(nz_list)
        dp = lambda x: None if x is None else x.ctypes.data_as(_lib._dp)
        zz = nzs = szs = inorm = lsc = None
        if ntr:
            zz = np.ascontiguousarray(z, dtype=float)
            nzs = np.ascontiguousarray(np.stack([np.asarray(n, dtype=float) for n in nz_list]))
            if mag_bias_list is not None:
                szs = np.ascontiguousarray(np.stack([np.asarray(s_, dtype=float) for s_ in mag_bias_list]))
            with np.errstate(divide='ignore'):
                inorm = np.array([1.0 / np.float64(Akima(zz, n).integrate_nodes()) for n in nzs])
            if lens_scale is not None:
                lsc = np.ascontiguousarray(lens_scale, dtype=float)
        self.ncosmo, self.ntr, self.z = par.shape[0], ntr, zz
        st = C.c_int(0)
        self.h = _lib.load().ccl_b200_tables_build(
            self.ncosmo, dp(par), a.size, dp(a), float(sp["DCHI_INTEGRATION"]), float(gp["ROOT_EPSREL"]),
            int(gp["ROOT_N_ITERATION"]), float(gp["EPS_SCALEFAC_GROWTH"]), int(achi_stride), float(sp["A_SPLINE_MIN"]),
            _MODELS[model], int(bool(halofit)), k.size, dp(k), dp(np.log(k)), a_pk.size, dp(a_pk), float(sp["K_MIN"]),
            float(sp["K_MAX"]), 0 if zz is None else zz.size, dp(zz), ntr, dp(nzs), dp(szs), dp(inorm), dp(lsc),
            int(n_chi), C.byref(st))
        check(st.value)

    def fill(self, batch, first=0, order_lok=1, order_hik=2, use_linear=False):
        """a(chi), growth, chi(A_SPLINE_MIN) and P(k,a) of every cosmology into batch slots [first, ...)."""
        st = C.c_int(0)
        _lib.load().ccl_b200_batch_set_from_tables(batch.h, int(first), self.h, order_lok, order_hik, int(use_linear),
                                                   C.byref(st))
        check(st.value)
        self._keep(batch)

    def _keep(self, batch):
        """The tables object must outlive the batch's commit (its buffers are read in place)."""
        refs = batch.__dict__.setdefault("_tables_refs", [])
        if self not in refs:
            refs.append(self)

    def add_tracer(self, batch, kind, itr, der_bessel=0, der_angles=0, transfer_a=None, first=0):
        """One sub-tracer per slot from a device-built kernel: kind 'nc' (p(z) H(z)) or 'lensing'."""
        a_ka = fa = None
        if transfer_a is not None:
            a_ka = np.ascontiguousarray(transfer_a[0], dtype=float)
            fa = np.ascontiguousarray(transfer_a[1], dtype=float)
        st = C.c_int(0)
        dp = lambda x: None if x is None else x.ctypes.data_as(_lib._dp)
        idx = _lib.load().ccl_b200_batch_add_tracer_from_tables(
            batch.h, int(first), self.h, {"nc": 0, "lensing": 1}[kind], int(itr), int(der_bessel), int(der_angles),
            0 if a_ka is None else a_ka.size, dp(a_ka), dp(fa), C.byref(st))
        check(st.value)
        self._keep(batch)
        return idx

    def add_tracer_bg(self, batch, kind, itr, der_bessel, der_angles, bg_kind, a_nodes, amp=None, first=0):
        """One sub-tracer per slot whose a-only transfer is built on the device from each cosmology's growth
        tables: bg_kind 'rsd' -> -f(a) (use der_bessel=2), 'ia' -> -amp Omega_m / D(a) with
        amp = A_IA(z) 5e-14 rho_crit on the same nodes (use der_bessel=-1, der_angles=2)."""
        a_ka = np.ascontiguousarray(a_nodes, dtype=float)
        am = None if amp is None else np.ascontiguousarray(amp, dtype=float)
        st = C.c_int(0)
        dp = lambda x: None if x is None else x.ctypes.data_as(_lib._dp)
        idx = _lib.load().ccl_b200_batch_add_tracer_from_tables_bg(
            batch.h, int(first), self.h, {"nc": 0, "lensing": 1}[kind], int(itr), int(der_bessel), int(der_angles),
            {"rsd": 1, "ia": 2}[bg_kind], a_ka.size, dp(a_ka), dp(am), C.byref(st))
        check(st.value)
        self._keep(batch)
        return idx

    def add_kappa_tracer(self, batch, z_source, n_samples=100, first=0):
        """One CMB-lensing sub-tracer per slot (CMBLensingTracer): kappa kernel built on the device."""
        st = C.c_int(0)
        idx = _lib.load().ccl_b200_batch_add_kappa_tracer_from_tables(batch.h, int(first), self.h, float(z_source),
                                                                      int(n_samples), C.byref(st))
        check(st.value)
        self._keep(batch)
        return idx

    def __del__(self):
        if getattr(self, "h", None) and _lib._LIB is not None:
            _lib._LIB.ccl_b200_tables_free(self.h)
            self.h = None


def build_3x2pt_batch(cosmologies, z, lens_nz, source_nz, lens_bias=None, model="eisenstein_hu", halofit=True,
                      n_chi=256, batch=None):
    """Parameters -> committed LibBatch with no table ever visiting the host (north-star item 5).

    Same content as build_3x2pt_stacked + LibBatch.set_stacked.  Returns (batch, collections, tables)."""
    from .lowlevel import LibBatch
    nl, ns = len(lens_nz), len(source_nz)
    if lens_bias is None:
        lens_bias = np.ones(nl)
    # lens bins are pure number counts here: their lensing (magnification) kernels are not built
    tabs = DeviceTables(cosmologies, z, list(lens_nz) + list(source_nz), lens_scale=[0.0] * nl + [1.0] * ns,
                        model=model, halofit=halofit, n_chi=n_chi)
    if batch is None:
        batch = LibBatch(tabs.ncosmo)
    else:
        batch.reset()
    tabs.fill(batch)
    z = np.asarray(z, dtype=float)
    a_rev = np.ascontiguousarray(1. / (1 + z[::-1]))
    for i in range(nl):
        tabs.add_tracer(batch, "nc", i, 0, 0, transfer_a=(a_rev, np.full(z.size, float(lens_bias[i]))))
    for j in range(ns):
        tabs.add_tracer(batch, "lensing", nl + j, -1, 2)
    batch.commit()
    return batch, [(i, 1) for i in range(nl + ns)], tabs
