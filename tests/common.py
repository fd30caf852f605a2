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


# how often the edge-node escape of parity_error was allowed / needed (reported by tests/test_zz_parity_escape.py)
ESCAPE_STATS = {"elements": 0, "eligible": 0, "used": 0}


def edge_mask(trc1, trc2, ells, k_min=5e-5):
    """Per ell: does the first integrand node of this pair sit exactly on a kernel-support edge?

    True where lkmin = log((ell+1/2)/chi_max) (not clamped at K_MIN) AND the pair's chi_max
    (get_k_interval, src/ccl_cls.c:64-100) is the first or last node of a tabulated kernel of one of the
    sub-tracers, where ccl_f1d_t_eval jumps between W and 0 (src/ccl_f1d.c:137-161).  Only those
    elements may use the libm escape of parity_error.
    """
    ells = np.asarray(ells, dtype=float)
    cmax = min(max(t.chi_max for t in trc1), max(t.chi_max for t in trc2))
    ends = set()
    for t in list(trc1) + list(trc2):
        chi_w = t._keep[0][0]
        if chi_w is not None and len(chi_w):
            ends.add(float(chi_w[0]))
            ends.add(float(chi_w[-1]))
    if cmax not in ends:
        return np.zeros(ells.size, dtype=bool)
    return (ells + 0.5) / cmax > k_min


def edge_mask_subs(subs, colls, pair, ells, cmax_nokernel=1e15):
    """edge_mask for a pair of collections given as (start, count) ranges into flat sub-tracer tables."""
    t1 = [oracle_tracer(s, cmax_nokernel) for s in subs[colls[pair[0]][0]:colls[pair[0]][0] + colls[pair[0]][1]]]
    t2 = [oracle_tracer(s, cmax_nokernel) for s in subs[colls[pair[1]][0]:colls[pair[1]][0] + colls[pair[1]][1]]]
    return edge_mask(t1, t2, ells)


def parity_error(cl, ref, recompute, rtol, edge=None):
    """Max relative error of cl against ref (curves that cross zero are scaled by the curve's own
    amplitude).

    With pyccl's own tracers chi_max of a pair is often the last node of a kernel spline, so the first
    integrand node sits exactly on a support edge (oracle/limber_oracle.c, o_set_lkmin_ulps): the
    reference's value there depends on the last bit of libm's log/exp.  ONLY for the elements flagged in
    `edge` (edge_mask: first node on the end node of a tabulated kernel), an element outside rtol is
    re-evaluated with the lower limit shifted by -2..2 ulps (`recompute()` returns the oracle curve) and
    accepted if the device value agrees with one of the outcomes a <= 1 ulp libm can produce.  Every
    other element is held to rtol as is.  ESCAPE_STATS counts eligible and escaped elements.
    """
    cl, ref = np.asarray(cl), np.asarray(ref)
    floor = 1e-6 * np.max(np.abs(ref))
    if floor == 0.0:
        return 0.0 if np.array_equal(cl, ref) else np.inf
    err = np.abs(cl - ref) / np.maximum(np.abs(ref), floor)
    ESCAPE_STATS["elements"] += int(err.size)
    if edge is None or recompute is None:
        return float(np.max(err))
    edge = np.asarray(edge, dtype=bool)
    ESCAPE_STATS["eligible"] += int(edge.sum())
    bad = (err >= rtol) & edge
    if bad.any():
        best = err.copy()
        for ulps in (-2, -1, 1, 2):
            try:
                oracle.set_lkmin_ulps(ulps)
                alt = np.asarray(recompute())
            finally:
                oracle.set_lkmin_ulps(0)
            e2 = np.abs(cl - alt) / np.maximum(np.abs(alt), floor)
            best = np.where(bad, np.minimum(best, e2), best)
        ESCAPE_STATS["used"] += int((bad & (best < rtol)).sum())
        err = best
    return float(np.max(err))
