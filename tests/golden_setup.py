"""The reference's golden benchmark of the Limber path (benchmarks/test_cells.py:8-223), rebuilt on
this package's pyccl-compatible API.  Fixture data: tests/golden/cells_benchmark.npz (extracted from
/root/reference/benchmarks/data by tools/make_golden.py)."""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "cells_benchmark.npz")

# (tracer1, tracer2, benchmark, a1b1, a1b2, a2b1, a2b2, ell-factor): benchmarks/test_cells.py:125-194
CASES = [('g1', 'g1', 'dd_11', 'dd_11', 'dd_11', 'dd_11', 'dd_11', 'fl_one'),
         ('g1', 'g2', 'dd_12', 'dd_11', 'dd_12', 'dd_12', 'dd_22', 'fl_one'),
         ('g2', 'g2', 'dd_22', 'dd_22', 'dd_22', 'dd_22', 'dd_22', 'fl_one'),
         ('g1', 'l1', 'dl_11', 'dd_11', 'dl_11', 'dl_11', 'll_11', 'fl_dl'),
         ('g1', 'l2', 'dl_12', 'dd_11', 'dl_12', 'dl_12', 'll_22', 'fl_dl'),
         ('g2', 'l1', 'dl_21', 'dd_22', 'dl_21', 'dl_21', 'll_22', 'fl_dl'),
         ('g2', 'l2', 'dl_22', 'dd_22', 'dl_22', 'dl_22', 'ii_22', 'fl_dl'),
         ('g1', 'i1', 'di_11', 'dd_11', 'di_11', 'di_11', 'ii_11', 'fl_dl'),
         ('g1', 'i2', 'di_12', 'dd_11', 'di_12', 'di_12', 'ii_22', 'fl_dl'),
         ('g2', 'i1', 'di_21', 'dd_22', 'di_21', 'di_21', 'ii_22', 'fl_dl'),
         ('g2', 'i2', 'di_22', 'dd_22', 'di_22', 'di_22', 'ii_22', 'fl_dl'),
         ('g1', 'ct', 'dc_1', 'dd_11', 'dc_1', 'dc_1', 'cc', 'fl_one'),
         ('g2', 'ct', 'dc_2', 'dd_22', 'dc_2', 'dc_2', 'cc', 'fl_one'),
         ('l1', 'l1', 'll_11', 'll_11', 'll_11', 'll_11', 'll_11', 'fl_ll'),
         ('l1', 'l2', 'll_12', 'll_11', 'll_12', 'll_12', 'll_22', 'fl_ll'),
         ('l2', 'l2', 'll_22', 'll_22', 'll_22', 'll_22', 'll_22', 'fl_ll'),
         ('l1', 'i1', 'li_11', 'ii_11', 'li_11', 'li_11', 'll_11', 'fl_li'),
         ('l2', 'i2', 'li_22', 'ii_22', 'li_22', 'li_22', 'll_22', 'fl_li'),
         ('l1', 'ct', 'lc_1', 'll_11', 'lc_1', 'lc_1', 'cc', 'fl_lc'),
         ('l2', 'ct', 'lc_2', 'll_22', 'lc_2', 'lc_2', 'cc', 'fl_lc'),
         ('i1', 'i1', 'ii_11', 'll_11', 'll_11', 'll_11', 'll_11', 'fl_ll'),
         ('i1', 'i2', 'ii_12', 'll_11', 'll_12', 'll_12', 'll_22', 'fl_ll'),
         ('i2', 'i2', 'ii_22', 'll_22', 'll_22', 'll_22', 'll_22', 'fl_ll')]
SPLINE_CASES = [CASES[0], CASES[1], CASES[15]]   # benchmarks/test_cells.py:206-223

_BM_FILES = {'dd_11': 'b1b1_dd', 'dd_12': 'b1b2_dd', 'dd_22': 'b2b2_dd', 'dl_11': 'b1b1_dl', 'dl_12': 'b1b2_dl',
             'dl_21': 'b2b1_dl', 'dl_22': 'b2b2_dl', 'di_11': 'b1b1_di', 'di_12': 'b1b2_di', 'di_21': 'b2b1_di',
             'di_22': 'b2b2_di', 'dc_1': 'b1b1_dc', 'dc_2': 'b2b2_dc', 'll_11': 'b1b1_ll', 'll_12': 'b1b2_ll',
             'll_22': 'b2b2_ll', 'li_11': 'b1b1_li', 'li_22': 'b2b2_li', 'lc_1': 'b1b1_lc', 'lc_2': 'b2b2_lc',
             'ii_11': 'b1b1_ii', 'ii_12': 'b1b2_ii', 'ii_22': 'b2b2_ii'}

_cache = {}


def set_up(nztyp):
    """cosmology, tracers, ell-factors and benchmark curves of benchmarks/test_cells.py:8-123."""
    if nztyp in _cache:
        return _cache[nztyp]
    import ccl_b200 as ccl
    d = np.load(GOLDEN)
    ccl.gsl_params.LENSING_KERNEL_SPLINE_INTEGRATION = False
    ccl.gsl_params.INTEGRATION_LIMBER_EPSREL = 1E-4
    ccl.gsl_params.INTEGRATION_EPSREL = 1E-4
    try:
        cosmo = ccl.Cosmology(Omega_c=0.30, Omega_b=0.00, Omega_g=0, Omega_k=0, h=0.7, sigma8=0.8, n_s=0.96,
                              Neff=0, m_nu=0.0, w0=-1, wa=0, T_CMB=2.7, transfer_function='bbks',
                              matter_power_spectrum='linear')
    finally:
        ccl.gsl_params.reload()
    ells = d["ells"]
    fl_dl = (ells + 0.5) ** 2 / np.sqrt((ells + 2.) * (ells + 1.) * ells * (ells - 1.))
    lfacs = {'ells': ells, 'fl_one': np.ones_like(ells), 'fl_dl': fl_dl, 'fl_ll': fl_dl ** 2,
             'fl_lc': ells * (ells + 1) / np.sqrt((ells + 2.) * (ells + 1.) * ells * (ells - 1.)),
             'fl_li': 2 * fl_dl}
    if nztyp == 'analytic':
        z1, tmp_a1 = d["ia_amp_analytic_1"].T
        z2, tmp_a2 = d["ia_amp_analytic_2"].T
        pz1 = np.exp(-0.5 * ((z1 - 1.0) / 0.15) ** 2)
        pz2 = np.exp(-0.5 * ((z2 - 1.5) / 0.15) ** 2)
    else:
        z1, pz1 = d["bin1_histo"].T[:, 1:]
        tmp_a1 = d["ia_amp_histo_1"].T[1]
        z2, pz2 = d["bin2_histo"].T[:, 1:]
        tmp_a2 = d["ia_amp_histo_2"].T[1]
    z1, pz1, z2, pz2 = (np.ascontiguousarray(x) for x in (z1, pz1, z2, pz2))
    bz = np.ones_like(pz1)
    D1 = cosmo.growth_factor(1. / (1 + z1))
    D2 = cosmo.growth_factor(1. / (1 + z2))
    rho_m = ccl.physical_constants.RHO_CRITICAL * cosmo['Omega_m']
    a1 = - tmp_a1 * D1 / (5e-14 * rho_m)
    a2 = - tmp_a2 * D2 / (5e-14 * rho_m)
    trc = {'g1': ccl.NumberCountsTracer(cosmo, has_rsd=False, dndz=(z1, pz1), bias=(z2, bz)),
           'g2': ccl.NumberCountsTracer(cosmo, has_rsd=False, dndz=(z2, pz2), bias=(z2, bz)),
           'l1': ccl.WeakLensingTracer(cosmo, dndz=(z1, pz1)),
           'l2': ccl.WeakLensingTracer(cosmo, dndz=(z2, pz2)),
           'i1': ccl.WeakLensingTracer(cosmo, dndz=(z1, pz1), has_shear=False, ia_bias=(z1, a1)),
           'i2': ccl.WeakLensingTracer(cosmo, dndz=(z2, pz2), has_shear=False, ia_bias=(z2, a2)),
           'ct': ccl.CMBLensingTracer(cosmo, z_source=1100.)}
    bms = {k: d["%s_%s" % (nztyp, v)] for k, v in _BM_FILES.items()}
    bms['cc'] = d['cc']
    _cache[nztyp] = (cosmo, trc, lfacs, bms)
    return _cache[nztyp]


def oracle_objects(cosmo, trc, pk=None):
    """Oracle twins of the API objects (same flat tables); `pk`: a Pk2D instead of the cosmology's own."""
    import oracle
    from common import oracle_tracer
    bg = cosmo._bg
    sp, gp = cosmo._spline_params, cosmo._gsl_params
    cosmo.compute_growth()
    oc = oracle.Cosmo(bg.achi_chi, bg.achi_a, bg.a, bg.growth, K_MAX=sp["K_MAX"], K_MIN=sp["K_MIN"],
                      DLOGK=sp["DLOGK_INTEGRATION"], LIMBER_EPSREL=gp["INTEGRATION_LIMBER_EPSREL"],
                      N_ITERATION=gp["N_ITERATION"])
    if pk is None:
        pk = cosmo.get_nonlin_power()
    op = oracle.F2D(a=pk._a, lk=pk._lk, fka=pk._tab, is_factorizable=False, extrap_order_lok=pk.extrap_order_lok,
                    extrap_order_hik=pk.extrap_order_hik, extrap_linear_growth=oracle.F2D_CCLGROWTH,
                    is_log=pk.is_logp, growth_factor_0=0.0, growth_exponent=2)
    cmax = cosmo.chi_max_nokernel()
    otr = {k: [oracle_tracer(s, cmax) for s in t._trc] for k, t in trc.items()}
    return oc, op, otr


def sigma_cv(bmk, lfc, a1b1, a1b2, a2b1, a2b2):
    return np.sqrt((bmk[a1b1] * bmk[a2b2] + bmk[a1b2] * bmk[a2b1]) / (2 * lfc['ells'] + 1.))
