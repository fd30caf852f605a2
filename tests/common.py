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
