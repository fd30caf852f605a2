"""Shared helpers of the parity tests: build oracle objects and library objects from the same
flat tables (dicts with the argument names of Tracer.add_tracer)."""
import numpy as np

import oracle


def chi_max_nokernel(bg, a_spline_min=0.1):
    a, chi = bg["chi_of_a"]
    return float(np.interp(a_spline_min, a, chi))


def oracle_cosmo(bg, **par):
    return oracle.Cosmo(bg["chi"], bg["a"], bg.get("a_growth"), bg.get("growth"), **par)


def oracle_pk(pk, order_lok=1, order_hik=2, is_log=True):
    return oracle.F2D(a=pk["a"], lk=pk["lk"], fka=pk["logp"], is_factorizable=False,
                      extrap_order_lok=order_lok, extrap_order_hik=order_hik,
                      extrap_linear_growth=oracle.F2D_CCLGROWTH, is_log=is_log, growth_factor_0=0.0,
                      growth_exponent=2)


def oracle_tracer(sub, cmax_nokernel=1e15):
    kernel = (sub["chi"], sub["w"]) if sub.get("chi") is not None else None
    tka = (sub["a"], sub["lk"], sub["fka"]) if sub.get("fka") is not None else None
    tk = (sub["lk"], sub["fk"]) if (tka is None and sub.get("fk") is not None) else None
    ta = (sub["a"], sub["fa"]) if (tka is None and sub.get("fa") is not None) else None
    return oracle.Tracer(der_bessel=sub.get("der_bessel", 0), der_angles=sub.get("der_angles", 0),
                         kernel=kernel, transfer_ka=tka, transfer_k=tk, transfer_a=ta,
                         is_logt=sub.get("is_logt", False),
                         extrap_order_lok=sub.get("extrap_order_lok", 0),
                         extrap_order_hik=sub.get("extrap_order_hik", 2),
                         chi_max_nokernel=cmax_nokernel)


def rel_err(x, ref, floor=0.0):
    x = np.asarray(x)
    ref = np.asarray(ref)
    scale = np.maximum(np.abs(ref), floor)
    return np.abs(x - ref) / scale


def parity_error(cl, ref, recompute, rtol):
    """Max relative error of cl against ref (curves that cross zero are scaled by the curve's own
    amplitude), tolerant of the reference's libm-dependent first node.

    With the default tracers the first integrand node sits exactly on a kernel-support edge
    (oracle/limber_oracle.c, o_set_lkmin_ulps): the reference's value there depends on the last
    bit of libm's log/exp.  Elements outside rtol are re-evaluated with the lower limit shifted
    by -2..2 ulps (`recompute(ulps)` returns the oracle curve) and accepted if the device value
    agrees with one of the outcomes a <= 1 ulp libm can produce.
    """
    cl, ref = np.asarray(cl), np.asarray(ref)
    floor = 1e-6 * np.max(np.abs(ref))
    if floor == 0.0:
        return 0.0 if np.array_equal(cl, ref) else np.inf
    err = np.abs(cl - ref) / np.maximum(np.abs(ref), floor)
    bad = err >= rtol
    if bad.any() and recompute is not None:
        for ulps in (-2, -1, 1, 2):
            try:
                oracle.set_lkmin_ulps(ulps)
                alt = np.asarray(recompute())
            finally:
                oracle.set_lkmin_ulps(0)
            e2 = np.abs(cl - alt) / np.maximum(np.abs(alt), floor)
            err = np.where(bad, np.minimum(err, e2), err)
    return float(np.max(err))
