"""GPU parity on BASELINE.json's configurations as stated (the bench times configs[3]; configs[0] is
tests/test_isw_readme.py::test_readme_example_gpu):

configs[2]  LSST-Y1-like 3x2pt (5 NumberCounts + 10 WeakLensing) + CMBLensingTracer + ISWTracer as ONE
            batch call, 'qag_quad' <= 1e-5 on all 153 pairs x 100 ells (and 'spline' <= 1e-10).
configs[4]  LSST-Y10 cosmic shear: 20 IA-augmented source bins, 210 pairs x 1000 ells, 'spline' <= 1e-10
            on 36 sampled pairs x all 1000 ells, finite everywhere; the pair x ell tile split
            (ccl_b200.sharding) reproduces the unsplit result bit for bit.
"""
import numpy as np
import pytest

import oracle
from common import chi_max_nokernel, edge_mask, oracle_cosmo, oracle_pk, oracle_tracer, parity_error

pytestmark = pytest.mark.gpu


def _smail(z, z0, alpha):
    return z ** 2 * np.exp(-(z / z0) ** alpha)


def _config2():
    """EH + halofit cosmology through the pyccl-compatible classes; the flat sub-tracer tables they hold
    feed both the batched C-ABI entry and the oracle."""
    import ccl_b200 as ccl
    from ccl_b200 import synthetic as S
    cosmo = ccl.Cosmology(Omega_c=0.27, Omega_b=0.045, h=0.67, sigma8=0.8, n_s=0.96,
                          transfer_function='eisenstein_hu', matter_power_spectrum='halofit')
    z = np.linspace(0.0, 3.5, 500)
    lens_nz = S.smail_bins(z, 0.26, 0.94, edges=np.arange(0.2, 1.2 + 1e-9, 0.2), sigma_z=0.03)
    src_nz = S.smail_bins(z, 0.13, 0.78, nbins=10, sigma_z=0.05)
    trc = [ccl.NumberCountsTracer(cosmo, has_rsd=False, dndz=(z, n), bias=(z, (1.0 + 0.2 * i) * np.ones_like(z)))
           for i, n in enumerate(lens_nz)]
    trc += [ccl.WeakLensingTracer(cosmo, dndz=(z, n)) for n in src_nz]
    trc += [ccl.CMBLensingTracer(cosmo, z_source=1100.0), ccl.ISWTracer(cosmo)]
    return cosmo, trc


def test_config2_3x2pt_kappa_isw_one_batch():
    import ccl_b200 as cb
    cosmo, trc = _config2()
    cosmo.compute_growth()
    bg = cosmo._bg
    pk = cosmo.get_nonlin_power()
    cmax = cosmo.chi_max_nokernel()
    subs, colls = [], []
    for t in trc:
        colls.append((len(subs), len(t._trc)))
        subs += list(t._trc)
    assert len(colls) == 17
    pairs = [(i, j) for i in range(17) for j in range(i, 17)]
    ells = np.geomspace(2, 3000, 100)
    batch = cb.LibBatch(1, cosmo._limber_params())
    batch.set_cosmology(0, bg.achi_chi, bg.achi_a, bg.a, bg.growth, cmax)
    batch.set_pk2d(0, pk._lk, pk._a, pk._tab, pk.extrap_order_lok, pk.extrap_order_hik, pk.is_logp)
    for s in subs:
        batch.add_tracer(0, **s)
    batch.commit()
    oc = oracle.Cosmo(bg.achi_chi, bg.achi_a, bg.a, bg.growth)
    op = oracle.F2D(a=pk._a, lk=pk._lk, fka=pk._tab, is_factorizable=False, extrap_order_lok=pk.extrap_order_lok,
                    extrap_order_hik=pk.extrap_order_hik, extrap_linear_growth=oracle.F2D_CCLGROWTH,
                    is_log=pk.is_logp, growth_factor_0=0.0, growth_exponent=2)
    otr = [oracle_tracer(s, cmax) for s in subs]
    sel = lambda c: otr[c[0]:c[0] + c[1]]  # noqa: E731
    clq, statq, stq = batch.angular_cl(colls, pairs, ells, cb._lib.INTEGRATION_QAG_QUAD)
    cls, stats, sts = batch.angular_cl(colls, pairs, ells, cb._lib.INTEGRATION_SPLINE)
    assert stq == 0 and sts == 0 and not statq.any() and not stats.any()
    worst_q = worst_s = 0.0
    for q, (i, j) in enumerate(pairs):
        ref, st = oracle.angular_cl(oc, sel(colls[i]), sel(colls[j]), op, ells, oracle.INTEG_QAG)
        assert st == 0
        worst_q = max(worst_q, parity_error(clq[0, q], ref, None, 1e-5))
        redo = lambda: oracle.angular_cl(oc, sel(colls[i]), sel(colls[j]), op, ells, oracle.INTEG_SPLINE)[0]  # noqa: E731
        worst_s = max(worst_s, parity_error(cls[0, q], redo(), redo, 1e-10,
                                            edge_mask(sel(colls[i]), sel(colls[j]), ells)))
    assert worst_q < 1e-5, worst_q
    assert worst_s < 1e-10, worst_s
    # the kappa and ISW spectra are really there (not an all-zero agreement)
    k_auto = pairs.index((15, 15))
    isw_k = pairs.index((15, 16))
    assert np.all(cls[0, k_auto] > 0) and np.all(np.abs(cls[0, isw_k]) > 0)


def _config4():
    from ccl_b200 import synthetic as S
    bg = S.background()
    pk = S.power(bg)
    subs, colls = S.lsst_like_tracers(bg, n_lens=0, n_src=20, with_ia=True)
    return bg, pk, subs, colls


def test_config4_y10_shear_ia_210x1000():
    import ccl_b200 as cb
    from ccl_b200 import sharding
    bg, pk, subs, colls = _config4()
    cmax = chi_max_nokernel(bg)
    ells = np.geomspace(2, 5000, 1000)
    pairs = [(i, j) for i in range(20) for j in range(i, 20)]
    assert len(pairs) == 210 and all(c[1] == 2 for c in colls)
    batch = cb.LibBatch(1)
    batch.set_cosmology(0, bg["chi"], bg["a"], bg["a_growth"], bg["growth"], cmax)
    batch.set_pk2d(0, pk["lk"], pk["a"], pk["logp"])
    for s in subs:
        batch.add_tracer(0, **s)
    batch.commit()
    cl, stat, st = batch.angular_cl(colls, pairs, ells)
    assert batch.last_used_fast_path and st == 0 and not stat.any()
    assert cl.shape == (1, 210, 1000) and np.all(np.isfinite(cl))
    oc, op = oracle_cosmo(bg), oracle_pk(pk)
    otr = [oracle_tracer(s, cmax) for s in subs]
    sel = lambda c: otr[c[0]:c[0] + c[1]]  # noqa: E731
    worst = 0.0
    sample = list(range(0, 210, 6)) + [209]
    assert len(sample) >= 30
    for q in sample:
        i, j = pairs[q]
        redo = lambda: oracle.angular_cl(oc, sel(colls[i]), sel(colls[j]), op, ells, oracle.INTEG_SPLINE)[0]  # noqa: E731
        worst = max(worst, parity_error(cl[0, q], redo(), redo, 1e-10, edge_mask(sel(colls[i]), sel(colls[j]), ells)))
    assert worst < 1e-10, worst
    # pair x ell tiles (SURVEY 8e, configs[4]): every tile of a 4-way split equals the unsplit block
    tiles = sharding.make_tiles(len(pairs), len(ells), pair_tile=30, ell_tile=250)
    w = sharding.tile_weights(batch, colls, pairs, ells, tiles)
    owned = sharding.balance_tiles(w, 4)
    loads = [w[idx].sum() for idx in owned]
    assert sorted(np.concatenate(owned)) == list(range(len(tiles))) and max(loads) < 1.1 * min(loads)
    out = np.full_like(cl, np.nan)
    for rank in range(4):
        for t in owned[rank]:
            p0, p1, l0, l1 = tiles[t]
            c_t, s_t, st_t = batch.angular_cl(colls, pairs[p0:p1], ells[l0:l1])
            assert st_t == 0
            out[0, p0:p1, l0:l1] = c_t[0]
    assert np.array_equal(out, cl)
