"""``angular_cl`` with the pyccl signature (pyccl/cells.py:10-175): Limber path plus the FKEM non-Limber
integrator below the transition multipole (nonlimber.py).

The call site of the hot path: ``lib.angular_cl_vec_limber(cosmo.cosmo, clt1, clt2, psp, ell,
integ_type, nell, status)`` (pyccl/cells.py:149-158) becomes one C-ABI call,
``ccl_b200_angular_cls_limber`` (include/ccl_b200.h), on device-resident tables.
"""
import numpy as np

from . import _lib, lowlevel
from .errors import CCLWarning, check, warnings
from .parameters import DEFAULT_POWER_SPECTRUM

__all__ = ("angular_cl", "integ_types")

# pyccl/pyutils.py:33-35 (ccl_integration_t, include/ccl_utils.h:12-15)
integ_types = {'qag_quad': _lib.INTEGRATION_QAG_QUAD, 'spline': _lib.INTEGRATION_SPLINE}


def angular_cl(cosmo, tracer1, tracer2, ell, *, p_of_k_a=DEFAULT_POWER_SPECTRUM, l_limber=-1,
               limber_max_error=0.01, limber_integration_method="qag_quad",
               non_limber_integration_method="FKEM", fkem_chi_min=None, fkem_Nchi=None,
               p_of_k_a_lin=DEFAULT_POWER_SPECTRUM, return_meta=False):
    if cosmo["Omega_k"] != 0:
        warnings.warn("CCL does not properly use the hyperspherical Bessel functions when computing angular "
                      "power spectra in non-flat cosmologies!", category=CCLWarning, importance='low')
    if limber_integration_method not in integ_types:
        raise ValueError("Limber integration method %s not supported" % limber_integration_method)
    if non_limber_integration_method not in ["FKEM"]:
        raise ValueError("Non-Limber integration method %s not supported" % limber_integration_method)
    if type(l_limber) is str:
        if l_limber != "auto":
            raise ValueError("l_limber must be an integer or'auto'")
        auto_limber = True
    else:
        auto_limber = False
    cosmo.compute_distances()
    if p_of_k_a is None:
        p_of_k_a = DEFAULT_POWER_SPECTRUM
    psp = cosmo.parse_pk2d(p_of_k_a, is_linear=False)
    ell_use = np.atleast_1d(np.asarray(ell, dtype=float))
    if not (np.diff(ell_use) > 0).all():
        raise ValueError("ell values must be monotonically increasing")
    # growth is only read when P(k,a) is evaluated below its a-range (src/ccl_f2d.c:351-370): at the node
    # chi = (l+1/2)/k <= chi_max and, for der_bessel 1/2 sub-tracers, at the second point chi' = (l+3/2)/k
    # (src/ccl_cls.c:121-145), which reaches chi_max (l+3/2)/(l+1/2)
    need_growth = False
    if psp._a[0] > cosmo._bg.a[0]:
        chi_tab = cosmo.comoving_radial_distance(max(psp._a[0], cosmo._bg.a[0]))
        lmin = float(ell_use[0])
        for t in (tracer1, tracer2):
            for sub, r in zip(t._trc, t._chi_range):
                reach = r[1] * ((lmin + 1.5) / (lmin + 0.5) if sub.get("der_bessel", 0) >= 1 else 1.0)
                need_growth = need_growth or reach > chi_tab
    status = 0
    # pyccl/cells.py:128-146: the FKEM integrator takes the multipoles up to the Limber transition
    if auto_limber or ell_use[0] < l_limber:
        from .nonlimber import nonlimber_fkem
        if any(sub.get("w") is None for t in (tracer1, tracer2) for sub in t._trc):
            raise NotImplementedError("FKEM needs a radial kernel on every sub-tracer")
        l_limber, cl_non_limber, status = nonlimber_fkem(
            cosmo, tracer1, tracer2, p_of_k_a, ell_use, l_limber, pk_linear=p_of_k_a_lin,
            limber_max_error=limber_max_error, Nchi=fkem_Nchi, chi_min=fkem_chi_min, need_growth=True)
        check(status, cosmo)
    else:
        cl_non_limber = np.array([])
    ell_use_limber = ell_use[ell_use > l_limber]
    if len(ell_use_limber) > 0:
        dev = cosmo._device(need_growth=need_growth)
        trs1 = tracer1._device(cosmo)
        trs2 = trs1 if tracer2 is tracer1 else tracer2._device(cosmo)
        cl_limber, status = lowlevel.angular_cls_limber(dev, trs1, trs2, psp._device(), ell_use_limber,
                                                        integ_types[limber_integration_method])
    else:
        cl_limber = np.array([])
    cl = np.concatenate((cl_non_limber, cl_limber))
    if np.ndim(ell) == 0:
        cl = cl[0]
    check(status, cosmo)
    return (cl, {"l_limber": l_limber}) if return_meta else cl
