"""FKEM non-Limber angular power spectra (pyccl/nonlimber/_nonlimber_FKEM.py:228-517; arXiv:1911.11947).

    C_l = C_l^Limber[P_nl] - C_l^Limber[P_lin] + (2/pi) int dk k^2 P_lin(k, a=1) I_1(k) I_2(k),
    I(k) = int dchi W(chi) D(chi) T(k, a) j_l^(n)(k chi) (k chi)^plaw

The chi integrals are generalised FFTLog transforms (``ccl_b200_fftlog_transform_general``, one device call per
(bessel derivative, power law) kind for a whole block of multipoles instead of one call per multipole and
sub-tracer), the two Limber terms come from the device 'qag_quad' path for the whole block at once.  The
arithmetic per multipole is the reference's: same sampling in chi, same low-ringing k grids, same linear
interpolation to the common k grid, same transition test for ``l_limber='auto'``.
"""
from collections import OrderedDict

import numpy as np
from scipy.interpolate import make_interp_spline

from . import _lib, lowlevel
from .errors import CCLWarning, check, warnings
from .pk2d import Pk2D

__all__ = ("nonlimber_fkem",)

_BLOCK_AUTO = 16       # multipoles per device call while looking for the Limber transition
_CACHE_MAXSIZE = 4096  # (sub-tracer, multipole) transforms kept per Tracer (pyccl/tracers.py _fkem_cache_maxsize)


def _general_params(b):
    """(nu, deriv, plaw) of the generalised transform for bessel-derivative type b (:20-43): nu = 1.01 avoids the
    singularity at 1; der_bessel = -1 is the 1/(k chi)^2 factor of the lensing-like terms."""
    return 1.01, float(b) if b > 0 else 0.0, -2.0 if b < 0 else 0.0


def _k_common(ks_1, ks_2):
    """Overlap of all k ranges, log-spaced with the largest sample count (:46-68)."""
    all_ks = [np.asarray(k, dtype=float) for k in (*ks_1, *ks_2)]
    k_min = max(k[np.isfinite(k) & (k > 0.0)].min() for k in all_ks)
    k_max = min(k[np.isfinite(k) & (k > 0.0)].max() for k in all_ks)
    return np.logspace(np.log10(k_min), np.log10(k_max), max(len(k) for k in all_ks))


def _cache(tracer):
    c = getattr(tracer, "_fkem_fft", None)
    if c is None:
        c = tracer._fkem_fft = OrderedDict()
    return c


def _chi_transforms(cosmo, tracer, objs, Nchi, chi_min, chi_max, ells, k_low, chi_arr):
    """For every sub-tracer i and multipole: (k, I_i(k), T_i(k, <a>_i)) -- _chi_integrands (:71-167) for a block
    of multipoles.  Returns ks[nl][nsub][N], fks[nl][nsub][N], transfers[nl][nsub][N]."""
    kernels, chis = tracer.get_kernel()
    bessels = tracer.get_bessel_derivative()
    avg_as = tracer.get_avg_weighted_a()
    nsub, nl = len(kernels), len(ells)
    ks = np.zeros((nl, nsub, Nchi))
    fks = np.zeros((nl, nsub, Nchi))
    cache = _cache(tracer)
    key0 = (id(cosmo), Nchi, chi_min, chi_max)
    todo = {}  # (deriv, plaw) -> [(sub-tracer, multipole indices)]
    for i in range(nsub):
        miss = []
        for il, ell in enumerate(ells):
            hit = cache.get((i,) + key0 + (float(ell),))
            if hit is None:
                miss.append(il)
            else:
                cache.move_to_end((i,) + key0 + (float(ell),))
                ks[il, i], fks[il, i] = hit
        if miss:
            todo.setdefault(_general_params(bessels[i]), []).append((i, miss))
    if todo:
        a_arr = cosmo.scale_factor_of_chi(chi_arr)
        growfac_arr = cosmo.growth_factor(a_arr)
        lk_low = np.log(k_low)
        for (nu, deriv, plaw), items in todo.items():
            subs = [i for i, _ in items]
            need = sorted(set(il for _, miss in items for il in miss))
            frs = np.empty((len(subs), Nchi))
            for row, i in enumerate(subs):
                transfer_low = objs[i].transfer(lk_low, a_arr)[0]
                transfer_avg = objs[i].transfer(lk_low, avg_as[i])[0, 0]
                if transfer_avg == 0.0:
                    raise ZeroDivisionError("Zero transfer function encountered in FKEM chi integrand calculation. "
                                            "Setting integrand to zero.")
                # separable approximation of the transfer function: exact when T(k, a) = K(k) A(a)
                frs[row] = (make_interp_spline(chis[i], kernels[i], k=1)(chi_arr) * chi_arr * growfac_arr *
                            transfer_low / transfer_avg)
            k_out, f_out = lowlevel.fftlog_transform_general(chi_arr, frs, np.asarray(ells, dtype=float)[need], nu, 1,
                                                             deriv, plaw)
            for row, (i, miss) in enumerate(items):
                for il in miss:
                    j = need.index(il)
                    ks[il, i], fks[il, i] = k_out[j], f_out[j, row]
                    cache[(i,) + key0 + (float(ells[il]),)] = (k_out[j].copy(), f_out[j, row].copy())
            while len(cache) > _CACHE_MAXSIZE:
                cache.popitem(last=False)
    transfers = np.zeros((nl, nsub, Nchi))
    for i in range(nsub):  # T_i(k, <a>_i) on every multipole's k grid in one accessor call
        transfers[:, i, :] = objs[i].transfer(np.log(ks[:, i, :].ravel()), avg_as[i])[:, 0].reshape(nl, Nchi)
    return ks, fks, transfers


def nonlimber_fkem(cosmo, tracer1, tracer2, p_of_k_a, ls, l_limber, *, pk_linear, limber_max_error, Nchi, chi_min,
                   need_growth=False):
    """Returns (l_limber, cells, status) like _nonlimber_FKEM: `cells` covers the multipoles up to and including
    the first one at or beyond the Limber transition."""
    kpow, k_low = 3, 1.0e-5
    kernels_t1, chis_t1 = tracer1.get_kernel()
    kernels_t2, chis_t2 = tracer2.get_kernel()
    status = 0
    if (not (isinstance(p_of_k_a, str) and isinstance(pk_linear, str)) and
            not (isinstance(p_of_k_a, Pk2D) and isinstance(pk_linear, Pk2D))):
        warnings.warn("p_of_k_a and p_of_k_a_lin must be of the same type: a str in cosmo or a Pk2D object. "
                      "Defaulting to Limber calculation. ", category=CCLWarning, importance='high')
        return -1, np.array([]), status
    if Nchi is None or not isinstance(Nchi, int) or Nchi <= 0:
        warnings.warn("Nchi must be a positive integer. Setting to match tracer with large chi samples x 2.",
                      category=CCLWarning, importance='high')
        Nchi = 2 * max(max(len(i) for i in chis_t1), max(len(i) for i in chis_t2))
    if chi_min is None or not isinstance(chi_min, float) or chi_min <= 0.0:
        warnings.warn("chi_min must be greater than zero.Setting to default 1e-6 Mpc.", category=CCLWarning,
                      importance='high')
        chi_min = 1.0e-6
    if limber_max_error is None or not isinstance(limber_max_error, float) or limber_max_error <= 0.0:
        warnings.warn("limber_max_error must be greater than zero.Setting to default 0.01.", category=CCLWarning,
                      importance='high')
        limber_max_error = 0.01
    if l_limber is None or (not isinstance(l_limber, str) and not isinstance(l_limber, (int, np.integer))) \
            or ((isinstance(l_limber, str) and l_limber != 'auto')
                or (isinstance(l_limber, (int, np.integer)) and l_limber <= 0)):
        warnings.warn("l_limber must be a positive integer or 'auto'. Setting to 'auto'.", category=CCLWarning,
                      importance='high')
        l_limber = 'auto'

    psp_lin = cosmo.parse_pk2d(pk_linear, is_linear=True)
    psp_nonlin = cosmo.parse_pk2d(p_of_k_a, is_linear=False)
    dev = cosmo._device(need_growth=need_growth)
    objs1 = tracer1._device(cosmo)
    objs2 = objs1 if tracer2 is tracer1 else tracer2._device(cosmo)
    d_lin, d_nl = psp_lin._device(), psp_nonlin._device()
    ls = np.atleast_1d(np.asarray(ls, dtype=float))
    fll_t1 = np.array([o.f_ell(ls) for o in objs1])
    fll_t2 = np.array([o.f_ell(ls) for o in objs2])
    chi_max = max(max(np.max(c) for c in chis_t1), max(np.max(c) for c in chis_t2))
    chi_arr = np.logspace(np.log10(chi_min), np.log10(chi_max), num=Nchi, endpoint=True)
    dlnr = np.log(chi_max / chi_min) / (Nchi - 1.0)
    auto = isinstance(l_limber, str)

    def limber(trs1, trs2, psp, ells):
        cl, st = lowlevel.angular_cls_limber(dev, trs1, trs2, psp, ells, _lib.INTEGRATION_QAG_QUAD)
        check(st, cosmo)
        return cl

    cells = []
    done = False
    el0 = 0
    while el0 < len(ls) and not done:
        if auto:
            nblk = min(_BLOCK_AUTO, len(ls) - el0)
        else:  # every multipole up to and including the first one at or beyond l_limber
            over = np.nonzero(ls[el0:] >= l_limber)[0]
            nblk = (over[0] + 1) if over.size else len(ls) - el0
        ells = ls[el0:el0 + nblk]
        cl_limber_lin = limber(objs1, objs2, d_lin, ells)
        cl_limber_nonlin = limber(objs1, objs2, d_nl, ells)
        ks1, fks1, tr1 = _chi_transforms(cosmo, tracer1, objs1, Nchi, chi_min, chi_max, ells, k_low, chi_arr)
        if tracer1 is not tracer2:
            ks2, fks2, tr2 = _chi_transforms(cosmo, tracer2, objs2, Nchi, chi_min, chi_max, ells, k_low, chi_arr)
        else:
            ks2, fks2, tr2 = ks1, fks1, tr1
        for il, ell in enumerate(ells):
            el = el0 + il
            k = _k_common(ks1[il], ks2[il])
            f1 = np.array([np.interp(k, ks1[il, i], fks1[il, i]) * np.interp(k, ks1[il, i], tr1[il, i])
                           for i in range(len(objs1))])
            f2 = np.array([np.interp(k, ks2[il, j], fks2[il, j]) * np.interp(k, ks2[il, j], tr2[il, j])
                           for j in range(len(objs2))])
            kpk = k ** kpow * psp_lin(k, 1.0, cosmo)
            # sub-tracer pair terms of (2/pi) int dlnk k^3 P_lin I_1 I_2 (:469-481)
            pair = np.einsum("ik,jk,k->ij", f1, f2, kpk) * dlnr * 2.0 / np.pi * fll_t1[:, None, el] * fll_t2[None, :, el]
            cells.append(cl_limber_nonlin[il] - cl_limber_lin[il] + np.sum(pair))
            if not auto and ell >= l_limber:
                l_limber = ell
                done = True
                break
            if auto:
                # every sub-tracer combination has to be inside limber_max_error, not just the total (:170-226)
                is_limber = True
                for i in range(len(objs1)):
                    for j in range(len(objs2)):
                        c_lin = limber([objs1[i]], [objs2[j]], d_lin, [ell])[-1]
                        c_nl = limber([objs1[i]], [objs2[j]], d_nl, [ell])[-1]
                        thresh = (c_nl - c_lin + pair[i, j]) / c_nl - 1.0
                        if np.abs(thresh) >= limber_max_error:
                            is_limber = False
                            break
                    if not is_limber:
                        break
                if is_limber:
                    l_limber = ell
                    done = True
                    break
        el0 += nblk
    if isinstance(l_limber, str):
        l_limber = ls[-1]
    if False in np.isfinite(cells):
        status = 1
    return l_limber, np.array(cells), status
