"""d ln P / d ln k of a Pk2D (pyccl/pk2d.py:244-306 `derivative=True` -> pk2d_der_eval_multi ->
ccl_f2d_t_dlogf_dlk_eval, src/ccl_f2d.c:375-531).  CPU: the oracle restatement on power laws, whose logarithmic
slope is exact for the bicubic of a log table, inside the table and in both k-extrapolation branches; the non-log
branch against finite differences of the value.  GPU: ccl_b200_f2d_dlogf_dlk_eval_multi against the oracle."""
import numpy as np
import pytest

import oracle

A = np.linspace(0.2, 1.0, 16)
LK = np.log(np.geomspace(1e-3, 20., 64))


def _tables():
    k = np.exp(LK)
    pl = np.log(3.0 * A[:, None] ** 2 * k[None, :] ** -1.7)                 # pure power law: slope -1.7 everywhere
    curved = np.log(A[:, None] ** 2 * k[None, :] / (1 + (k[None, :] / 0.1) ** 2) ** 1.3)
    return pl, curved


def test_oracle_dlogf_dlk_power_law_and_extrapolation():
    pl, curved = _tables()
    for is_log, tab in ((True, pl), (False, np.exp(pl))):
        f = oracle.F2D(a=A, lk=LK, fka=tab, is_log=is_log, extrap_order_lok=1, extrap_order_hik=2,
                       extrap_linear_growth=oracle.F2D_CONSTGROWTH, growth_factor_0=1.0, growth_exponent=0)
        for lk in (LK[0] - 2.0, LK[0], -3.3, 0.0, LK[-1], LK[-1] + 1.5):
            for a in (0.25, 0.613, 1.0):
                v, st = f.dlogf_dlk(lk, a)
                assert st == 0
                if is_log:
                    assert abs(v + 1.7) < 1e-9, (lk, a, v)
                elif lk in (-3.3, 0.0):  # non-log table: natural end conditions spoil the slope near the edges
                    assert abs(v + 1.7) < 2e-3, (lk, a, v)
    # curved spectrum: against a centred difference of ln f inside the table (log table)
    f = oracle.F2D(a=A, lk=LK, fka=curved, is_log=True, extrap_linear_growth=oracle.F2D_CONSTGROWTH)
    for lk in (-5.0, -2.3, 0.4, 2.0):
        v, st = f.dlogf_dlk(lk, 0.5)
        h = 1e-5
        fd = (np.log(f(lk + h, 0.5)[0]) - np.log(f(lk - h, 0.5)[0])) / (2 * h)
        assert st == 0 and abs(v - fd) < 1e-6
    # below the table in a: no change of the logarithmic slope
    v0, _ = f.dlogf_dlk(0.4, A[0])
    v1, st = f.dlogf_dlk(0.4, 0.05)
    assert st == 0 and abs(v0 - v1) < 1e-12


@pytest.mark.gpu
def test_device_pk2d_derivative_matches_oracle():
    import ccl_b200 as ccl
    pl, curved = _tables()
    cosmo = ccl.CosmologyVanillaLCDM(transfer_function='bbks', matter_power_spectrum='linear')
    ks = np.exp(np.linspace(LK[0] - 2.0, LK[-1] + 1.5, 57))
    for is_log, tab in ((True, curved), (False, np.exp(curved)), (True, pl)):
        pk = ccl.Pk2D(a_arr=A, lk_arr=LK, pk_arr=tab, is_logp=is_log, extrap_order_lok=1, extrap_order_hik=2)
        fo = oracle.F2D(a=A, lk=LK, fka=tab, is_log=is_log, extrap_order_lok=1, extrap_order_hik=2,
                        extrap_linear_growth=oracle.F2D_CCLGROWTH, growth_factor_0=0.0, growth_exponent=2)
        for a in (0.2, 0.47, 1.0):
            d = pk(ks, a, derivative=True)
            ref = np.array([fo.dlogf_dlk(np.log(k), a)[0] for k in ks])
            assert d.shape == ks.shape and np.max(np.abs(d - ref)) < 1e-10 * max(1.0, np.max(np.abs(ref)))
        # a below the table needs a cosmology (growth extrapolation); the slope does not change
        d_lo = pk(ks, 0.1, cosmo, derivative=True)
        assert np.max(np.abs(d_lo - pk(ks, A[0], derivative=True))) < 1e-9
        assert np.shape(pk(0.1, 0.5, derivative=True)) == ()
    # the cosmology's own spectrum: slope n_s at the largest scales of a BBKS table, in extrapolation
    plin = cosmo.get_linear_power()
    assert abs(plin(1e-6, 1.0, derivative=True) - 0.96) < 2e-2
