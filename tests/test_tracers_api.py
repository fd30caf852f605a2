"""The reference's unit tests of the tracer objects the Limber path reads (pyccl/tests/test_tracers.py),
on the pyccl-compatible API.  Host-only checks run on CPU; anything that evaluates a tracer on the
device (accessors, C_ell) is marked gpu.

  test_tracer_mag_0p4                     pyccl/tests/test_tracers.py:103-132
  test_tracer_dndz_smoke                  :135-142
  test_tracer_{kernel,der_bessel,f_ell,transfer}_smoke   :145-215
  test_tracer_nz_support                  :218-242
  test_tracer_nz_norm_spline_vs_gsl       :250-273
  test_tracer_lensing_kernel_spline_vs_gsl / magnification  :276-357
  test_tracer_delta_function_nz           :360-379
  test_tracer_n_sample_*                  :382-409
  test_tracer_bool / chi_min_max / empty / increase_sf     :412-459
"""
import numpy as np
import pytest

import ccl_b200 as ccl
from ccl_b200 import CCLWarning


@pytest.fixture(scope="module")
def COSMO():
    return ccl.Cosmology(Omega_c=0.27, Omega_b=0.045, h=0.67, sigma8=0.8, n_s=0.96,
                         transfer_function='bbks', matter_power_spectrum='linear')


def dndz(z):
    return np.exp(-((z - 0.5) / 0.1) ** 2)


def get_tracer(tracer_type, cosmo, **tracer_kwargs):
    z = np.linspace(0., 1., 2000)
    n = dndz(z)
    b = np.sqrt(1. + z)
    if tracer_type == 'nc':
        return ccl.NumberCountsTracer(cosmo, has_rsd=True, dndz=(z, n), bias=(z, b), mag_bias=(z, b),
                                      **tracer_kwargs), 3
    if tracer_type == 'wl':
        return ccl.WeakLensingTracer(cosmo, dndz=(z, n), ia_bias=(z, b), **tracer_kwargs), 2
    if tracer_type == 'cl':
        return ccl.CMBLensingTracer(cosmo, z_source=1100., **tracer_kwargs), 1
    return ccl.Tracer(**tracer_kwargs), 0


def new_simple_cosmo():
    return ccl.CosmologyVanillaLCDM(transfer_function='bbks', matter_power_spectrum="linear")


# ------------------------------------------------------------------------------------------------
# host-side construction (CPU)
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize('tracer_type', ['nc', 'wl'])
def test_tracer_dndz_smoke(COSMO, tracer_type):
    tr, _ = get_tracer(tracer_type, COSMO)
    for z in [np.linspace(0.5, 0.6, 10), 0.5]:
        assert np.all(np.fabs(dndz(z) / tr.get_dndz(z) - 1) < 1E-5)


@pytest.mark.parametrize('tracer_type', ['nc', 'wl', 'cl', 'not'])
def test_tracer_der_bessel_smoke(COSMO, tracer_type):
    tr, _ = get_tracer(tracer_type, COSMO)
    dd = {'nc': [0, 2, -1], 'wl': [-1, -1], 'cl': [-1], 'not': []}[tracer_type]
    assert np.all(tr.get_bessel_derivative() == np.array(dd))


def test_tracer_nz_support(COSMO):
    a = np.linspace(1 / (1 + 1.0), 1.0, 100)
    background_def = {"a": a, "chi": ccl.comoving_radial_distance(COSMO, a), "h_over_h0": ccl.h_over_h0(COSMO, a)}
    calculator_cosmo = ccl.CosmologyCalculator(Omega_c=0.27, Omega_b=0.045, h=0.67, sigma8=0.8, n_s=0.96,
                                               background=background_def)
    z = np.linspace(0., 2., 2000)
    n = dndz(z)
    with pytest.raises(ValueError):
        ccl.WeakLensingTracer(calculator_cosmo, dndz=(z, n))
    with pytest.raises(ValueError):
        ccl.NumberCountsTracer(calculator_cosmo, has_rsd=False, dndz=(z, n), bias=(z, np.ones_like(z)))
    with pytest.raises(ValueError):
        ccl.CMBLensingTracer(calculator_cosmo, z_source=2.0)


def test_tracer_nz_norm_spline_vs_gsl_integration():
    out = {}
    try:
        for flag in (True, False):
            ccl.gsl_params.NZ_NORM_SPLINE_INTEGRATION = flag
            cosmo = new_simple_cosmo()
            out[flag] = (get_tracer("wl", cosmo)[0].get_kernel(chi=None)[0], get_tracer("nc", cosmo)[0].get_kernel(chi=None)[0])
    finally:
        ccl.gsl_params.reload()
    for k in (0, 1):
        for w_spline, w_gsl in zip(out[True][k], out[False][k]):
            assert np.allclose(w_spline, w_gsl, atol=0, rtol=1e-8)


@pytest.mark.parametrize('z_min, z_max, n_z_samples', [(0.0, 1.0, 2000), (0.0, 1.0, 1000), (0.0, 1.0, 500),
                                                       (0.0, 1.0, 256), (0.3, 1.0, 1000)])
def test_tracer_lensing_kernel_spline_vs_gsl_integration(z_min, z_max, n_z_samples):
    z = np.linspace(z_min, z_max, n_z_samples)
    n = dndz(z)
    if z_min > 0:
        assert n[0] > 0
    w = {}
    try:
        for flag in (True, False):
            ccl.gsl_params.LENSING_KERNEL_SPLINE_INTEGRATION = flag
            cosmo = new_simple_cosmo()
            w[flag] = ccl.WeakLensingTracer(cosmo, dndz=(z, n)).get_kernel(chi=None)[0][0]
    finally:
        ccl.gsl_params.reload()
    # peak of the kernel is ~1e-5
    if n_z_samples >= 1000:
        assert np.allclose(w[True], w[False], atol=1e-10, rtol=1e-9)
    else:
        assert np.allclose(w[True], w[False], atol=5e-9, rtol=1e-5)


@pytest.mark.parametrize('z_min, z_max, n_z_samples', [(0.0, 1.0, 2000), (0.0, 1.0, 500), (0.3, 1.0, 1000)])
def test_tracer_magnification_kernel_vs_lensing_kernel(COSMO, z_min, z_max, n_z_samples):
    z = np.linspace(z_min, z_max, n_z_samples)
    n = dndz(z)
    b = np.zeros_like(z)
    tr_mg = ccl.NumberCountsTracer(COSMO, has_rsd=False, dndz=(z, n), bias=(z, b), mag_bias=(z, b))
    w_mg = tr_mg.get_kernel(chi=None)[0]
    w_wl = ccl.WeakLensingTracer(COSMO, dndz=(z, n)).get_kernel(chi=None)[0]
    assert np.allclose(w_mg[1], -2 * w_wl[0], atol=2e-10, rtol=1e-9)


@pytest.mark.parametrize('tracer_type', ['nc', 'wl', 'cl'])
def test_tracer_n_sample_smoke(COSMO, tracer_type):
    get_tracer(tracer_type, COSMO, n_samples=50)
    if tracer_type != "cl":
        get_tracer(tracer_type, COSMO, n_samples=None)   # falls back to the samples of the n(z)


def test_tracer_n_sample_wl(COSMO):
    z = np.linspace(0., 1., 2000)
    w, chi = ccl.WeakLensingTracer(COSMO, dndz=(z, dndz(z)), n_samples=50).get_kernel(chi=None)
    assert w[0].shape[-1] == 50 and chi[0].shape[-1] == 50


def test_tracer_n_sample_warn(COSMO):
    z = np.linspace(0., 1., 50)
    ccl.update_warning_verbosity('high')
    try:
        with pytest.warns(CCLWarning):
            ccl.WeakLensingTracer(COSMO, dndz=(z, dndz(z)))
    finally:
        ccl.update_warning_verbosity('low')


def test_tracer_bool(COSMO):
    assert bool(ccl.Tracer()) is False
    assert bool(ccl.CMBLensingTracer(COSMO, z_source=1100)) is True


def test_tracer_chi_min_max_host(COSMO):
    tr = ccl.CMBLensingTracer(COSMO, z_source=1100)
    assert tr.chi_min == tr._trc[0].chi_min and tr.chi_max == tr._trc[0].chi_max
    chi = np.linspace(tr.chi_min + 0.05, tr.chi_max + 0.05, 128)
    tr.add_tracer(COSMO, kernel=(chi, np.ones_like(chi)))
    assert tr.chi_min == tr._trc[0].chi_min and tr.chi_max == tr._trc[1].chi_max


def test_empty_wlnc_tracer(COSMO):
    z = np.linspace(0, 1.0, 32)
    nz = np.exp(-0.5 * ((z - 0.5) / 0.05) ** 2)
    with pytest.raises(ValueError):
        ccl.WeakLensingTracer(COSMO, dndz=(z, nz), has_shear=False)
    with pytest.raises(ValueError):
        ccl.NumberCountsTracer(COSMO, has_rsd=False, dndz=(z, nz))


def test_tracer_increase_sf(COSMO):
    z = np.linspace(0, 3., 32)
    one = np.ones(len(z))
    chi = ccl.comoving_radial_distance(COSMO, 1. / (1 + z))
    sf = 1. / (1 + z)
    tr = ccl.Tracer()
    with pytest.raises(ValueError):
        tr.add_tracer(COSMO, kernel=(chi, one), transfer_a=(sf, one))
    tr.add_tracer(COSMO, kernel=(chi, one), transfer_a=(sf[::-1], one))


# ------------------------------------------------------------------------------------------------
# device evaluation (GPU)
# ------------------------------------------------------------------------------------------------
@pytest.mark.gpu
def test_tracer_mag_0p4(COSMO):
    z = np.linspace(0., 1., 2000)
    n = dndz(z)
    b = np.ones_like(z) * 0.1          # small bias so that magnification would matter if present
    t1 = ccl.NumberCountsTracer(COSMO, has_rsd=True, dndz=(z, n), bias=(z, b))
    t2 = ccl.NumberCountsTracer(COSMO, has_rsd=True, dndz=(z, n), bias=(z, b), mag_bias=(z, np.ones_like(z) * 0.4))
    t3 = ccl.NumberCountsTracer(COSMO, has_rsd=True, dndz=(z, n), bias=(z, b), mag_bias=(z, np.zeros_like(z)))
    ls = np.array([2, 200, 2000])
    cl1, cl2, cl3 = (ccl.angular_cl(COSMO, t, t, ls) for t in (t1, t2, t3))
    assert np.all(np.fabs(cl2 / cl1 - 1) < 1E-5)      # s = 0.4: no magnification
    assert np.all(np.fabs(cl3 / cl1 - 1) > 1E-2)


@pytest.mark.gpu
@pytest.mark.parametrize('tracer_type', ['nc', 'wl', 'cl', 'not'])
def test_tracer_kernel_f_ell_transfer_smoke(COSMO, tracer_type):
    tr, ntr = get_tracer(tracer_type, COSMO)
    for chi in [np.linspace(0., 3000., 128), [100., 1000.], 100.]:
        w = tr.get_kernel(chi)
        assert w.shape[0] == ntr
        for ww in w:
            assert np.shape(ww) == np.shape(chi)
    for ell in [np.linspace(0., 3000., 128), [100., 1000.], 100.]:
        fl = tr.get_f_ell(ell)
        assert fl.shape[0] == ntr
        for f in fl:
            assert np.shape(f) == np.shape(ell)
    for lk in [np.linspace(-3., 1., 10), [-2., 0.], -1.]:
        for a in [np.linspace(0.5, 1., 8), [0.4, 1.], 0.9]:
            tf = tr.get_transfer(lk, a)
            assert tf.shape[0] == ntr
            if ntr > 0:
                if np.ndim(a) == 0:
                    shap = (ntr,) if np.ndim(lk) == 0 else (ntr, len(lk))
                else:
                    shap = (ntr, len(a)) if np.ndim(lk) == 0 else (ntr, len(lk), len(a))
            else:
                shap = (0,)
            assert tf.shape == shap


@pytest.mark.gpu
def test_tracer_delta_function_nz(COSMO):
    z = np.linspace(0., 1., 2000)
    z_s_idx = int(z.size * 0.8)
    n = np.zeros_like(z)
    n[z_s_idx] = 2.0
    tr_wl = ccl.WeakLensingTracer(COSMO, dndz=(z, n))
    chi_kappa, w_kappa = ccl.tracers.get_kappa_kernel(COSMO, z_source=z[z_s_idx], n_samples=100)
    w = tr_wl.get_kernel(chi=chi_kappa)
    assert np.allclose(w[0], w_kappa, atol=1e-8, rtol=1e-6)
    # at z = z_source the interpolation becomes apparent
    assert np.allclose(w[0][:-2], w_kappa[:-2], atol=1e-11, rtol=1e-11)


@pytest.mark.gpu
def test_tracer_chi_min_max_device(COSMO):
    """The host-side range rule equals the library's (ccl_cl_tracer_t_new, src/ccl_tracers.c:626-666)."""
    for kind in ('nc', 'wl', 'cl'):
        tr, _ = get_tracer(kind, COSMO)
        for sub, dev in zip(tr._trc, tr._device(COSMO)):
            assert sub.chi_min == dev.chi_min and sub.chi_max == dev.chi_max
