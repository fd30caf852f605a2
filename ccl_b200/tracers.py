"""Tracers with the pyccl names and constructors (pyccl/tracers.py:84-196, 311-455, 625-1110).

A ``Tracer`` is a list of sub-tracers, each exactly the argument set of ``ccl_cl_tracer_t_new``
(src/ccl_tracers.c:569-691): radial kernel (chi, W), transfer function arrays, ``der_bessel``,
``der_angles``.  The radial kernels are built on the host from the cosmology's background
tables, restating src/ccl_tracers.c:60-567 (number-counts kernel, lensing / magnification kernel
by Akima-spline integration with the reference's trapezoid edge correction, CMB-kappa kernel);
the device-side twins used by ``angular_cl`` are created on first use.  Device-side kernel
builders are the next row of the scope table.
"""
import numpy as np

from . import lowlevel
from ._hostspline import Akima, F1D
from .errors import CCLError, CCLWarning, warnings
from .parameters import physical_constants

__all__ = ("Tracer", "NzTracer", "NumberCountsTracer", "WeakLensingTracer", "CMBLensingTracer", "ISWTracer",
           "get_density_kernel", "get_lensing_kernel", "get_kappa_kernel")

_GLX, _GLW = np.polynomial.legendre.leggauss(8)


def _check_array_params(f_arg, name=None, arr3=False):
    """pyccl/pyutils.py:636-662."""
    if f_arg is None:
        return (None, None, None) if arr3 else (None, None)
    if (not isinstance(f_arg, (tuple, list))) or (len(f_arg) != (3 if arr3 else 2)) or \
            (not all(isinstance(f, np.ndarray) for f in f_arg)):
        raise ValueError("%s must be a tuple of %d arrays" % (name, 3 if arr3 else 2))
    return tuple(np.atleast_1d(np.array(f, dtype=float)) for f in f_arg)


def _check_background_spline_compatibility(cosmo, z):
    """pyccl/pyutils.py: the redshifts must lie inside the background a-spline."""
    a_min = cosmo._bg.a[0]
    if np.any(np.asarray(z) < 0) or np.any(1. / (1 + np.asarray(z)) < a_min):
        raise ValueError("Tracer defined over wider redshift range than internal CCL splines. "
                         "Maximum redshift: %g" % (1. / a_min - 1))


def _sinn(cosmo, chi):
    """ccl_sinn (src/ccl_background.c:107-131)."""
    ks, sk = cosmo["k_sign"], cosmo["sqrtk"]
    if ks == -1:
        return np.sinh(sk * chi) / sk
    if ks == 1:
        return np.sin(sk * chi) / sk
    return chi


def _lensing_prefactor(cosmo):
    hub = cosmo["h"] / physical_constants.CLIGHT_HMPC
    return 1.5 * hub * hub * cosmo["Omega_m"]


def _nz_norm(cosmo, z, n):
    """get_nz_norm (src/ccl_tracers.c:60-119).  Returned as a numpy scalar: an identically zero n(z)
    gives 1/0 = inf and NaN kernels downstream, like the C arithmetic, not a Python exception."""
    return np.float64(_nz_norm_value(cosmo, z, n))


def _nz_norm_value(cosmo, z, n):
    if cosmo._gsl_params["NZ_NORM_SPLINE_INTEGRATION"]:
        return float(Akima(z, n).integrate_nodes())
    f = F1D(z, n)
    lo, hi = z[:-1], z[1:]
    zq = 0.5 * (hi + lo)[:, None] + 0.5 * (hi - lo)[:, None] * _GLX
    return float(np.sum(0.5 * (hi - lo)[:, None] * _GLW * f(zq)))


def _device_kernels(cosmo):
    """Tracer kernels on the device builders: the product path of a Cosmology built with them (spline-mode
    n(z) normalisation, the reference's default)."""
    return getattr(cosmo, "_table_builder", "host") == "device" and bool(cosmo._gsl_params["NZ_NORM_SPLINE_INTEGRATION"])


def _device_tracer_kernels(cosmo, z_n, n, s, n_chi):
    from . import device_builders as db
    cosmo.compute_distances()
    return db.build_tracer_kernels([cosmo], cosmo._bgd, z_n, [n], None if s is None else [s], n_chi)


def get_density_kernel(cosmo, *, dndz):
    """chi, W(chi) = p(z) H(z) (pyccl/tracers.py:84-111; ccl_get_number_counts_kernel,
    src/ccl_tracers.c:121-160)."""
    z_n, n = _check_array_params(dndz, 'dndz')
    _check_background_spline_compatibility(cosmo, z_n)
    if _device_kernels(cosmo):
        out = _device_tracer_kernels(cosmo, z_n, n, None, 8)
        return out["chi_z"][0], out["w_nc"][0, 0]
    chi = cosmo.comoving_radial_distance(1. / (1. + z_n))
    a = 1. / (1 + z_n)
    hz = cosmo["h"] * cosmo.h_over_h0(a) / physical_constants.CLIGHT_HMPC
    with np.errstate(divide='ignore', invalid='ignore'):
        return chi, hz * n * (1. / _nz_norm(cosmo, z_n, n))


@np.errstate(divide='ignore', invalid='ignore')   # an identically zero n(z): inf / NaN like the C code
def get_lensing_kernel(cosmo, *, dndz, mag_bias=None, n_chi=None):
    """chi, W_L(chi) (pyccl/tracers.py:114-172; ccl_get_lensing_mag_kernel, src/ccl_tracers.c:485-550)."""
    cosmo.compute_distances()
    z_n, n = _check_array_params(dndz, 'dndz')
    z_s, s = _check_array_params(mag_bias, 'mag_bias')
    _check_background_spline_compatibility(cosmo, z_n)
    if len(z_n) < 5:
        raise CCLError("Error CCL_ERROR_SPLINE: ccl_tracers.c: ccl_get_lensing_mag_kernel(): error initializing spline")
    spline_mode = bool(cosmo._gsl_params["LENSING_KERNEL_SPLINE_INTEGRATION"])
    if n_chi is None:  # ccl_get_nchi_lensing_kernel (src/ccl_tracers.c:458-466)
        dz = (z_n[-1] - z_n[0]) / (len(z_n) - 1)
        n_chi = int(z_n[-1] / dz + 0.5)
    if n_chi > len(z_n) and spline_mode:
        warnings.warn("The number of samples in the n(z) (%d) is smaller than the number of samples in the "
                      "lensing kernel (%d). Consider disabling spline integration for the lensing kernel by "
                      "setting gsl_params.LENSING_KERNEL_SPLINE_INTEGRATION = False before instantiating the "
                      "Cosmology passed." % (len(z_n), n_chi), category=CCLWarning, importance='low')
    if _device_kernels(cosmo) and spline_mode and (s is None or np.array_equal(z_s, z_n)):
        # integrate_lensing_kernel_spline on the device (csrc/builders.cu::tracer_kernel_builder)
        out = _device_tracer_kernels(cosmo, z_n, n, s, n_chi)
        return out["chi_l"][0], out["w_l"][0, 0]
    # ccl_get_chis_lensing_kernel (src/ccl_tracers.c:470-479)
    z_k = (z_n[-1] / n_chi) * np.arange(n_chi) + 1E-15
    chi_k = cosmo.comoving_radial_distance(1. / (1 + z_k))
    i_norm = 1. / _nz_norm(cosmo, z_n, n)
    pref = _lensing_prefactor(cosmo)
    sz_f = F1D(z_s, s, s[0], s[-1]) if s is not None else None
    a_k = cosmo.scale_factor_of_chi(chi_k)
    w = np.empty(n_chi)
    if spline_mode:
        # integrate_lensing_kernel_spline (src/ccl_tracers.c:347-454)
        chi_z = cosmo.comoving_radial_distance(1. / (1 + z_n))
        qz = 1 - 2.5 * sz_f(z_n) if sz_f is not None else np.ones_like(z_n)
        sinn_z = _sinn(cosmo, chi_z)
        for ic in range(n_chi):
            chi_end = chi_k[ic]
            z_end = 1. / a_k[ic] - 1
            below = chi_z < chi_end
            i_end = int(np.count_nonzero(below))  # chi(z) is increasing
            with np.errstate(divide="ignore", invalid="ignore"):
                integ = np.where(chi_z == 0, n * qz, n * qz * _sinn(cosmo, chi_z - chi_end) / sinn_z)
            integ = np.where(below, 0.0, integ)
            res = float(Akima(z_n, integ).integrate(z_n[i_end], z_n[-1]))
            if z_end > z_n[0]:
                res += 0.5 * (z_n[i_end] - z_end) * integ[i_end]
            w[ic] = res * pref * i_norm * chi_end / a_k[ic]
    else:
        # integrate_lensing_kernel_gsl (src/ccl_tracers.c:217-345): quadrature of the splined n(z);
        # Gauss-Legendre panels between the n(z) nodes instead of adaptive QAG
        nz_f = F1D(z_n, n)
        for ic in range(n_chi):
            chi_end = chi_k[ic]
            z_end = 1. / a_k[ic] - 1
            edges = np.concatenate([[z_end], z_n[z_n > z_end]])
            if edges.size < 2:
                w[ic] = 0.0
                continue
            lo, hi = edges[:-1], edges[1:]
            zq = 0.5 * (hi + lo)[:, None] + 0.5 * (hi - lo)[:, None] * _GLX
            chiq = cosmo.comoving_radial_distance(1. / (1 + zq))
            qq = 1 - 2.5 * sz_f(zq) if sz_f is not None else 1.0
            f = nz_f(zq) * qq * np.where(chiq == 0, 1.0, _sinn(cosmo, chiq - chi_end) / _sinn(cosmo, np.where(chiq == 0, 1.0, chiq)))
            res = float(np.sum(0.5 * (hi - lo)[:, None] * _GLW * f))
            w[ic] = res * pref * i_norm * chi_end / a_k[ic]
    return chi_k, w


def get_kappa_kernel(cosmo, *, z_source, n_samples=100):
    """pyccl/tracers.py:175-196; ccl_get_kappa_kernel (src/ccl_tracers.c:554-567)."""
    _check_background_spline_compatibility(cosmo, np.array([z_source]))
    chi_source = cosmo.comoving_radial_distance(1. / (1. + z_source))
    chi = np.linspace(0, chi_source, n_samples)
    pref = _lensing_prefactor(cosmo) / _sinn(cosmo, chi_source)
    a = cosmo.scale_factor_of_chi(chi)
    with np.errstate(divide="ignore", invalid="ignore"):
        w = np.where(chi != 0.0, pref * _sinn(cosmo, chi_source - chi) * chi * chi / a / _sinn(cosmo, np.where(chi != 0, chi, 1.0)),
                     pref * _sinn(cosmo, chi_source - chi) * chi / a)
    return chi, w


def _chi_range(sub, chi_max_nokernel):
    """chi_min / chi_max rule of ccl_cl_tracer_t_new (src/ccl_tracers.c:626-666)."""
    chi, w = sub.get("chi"), sub.get("w")
    if chi is None:
        return 0.0, chi_max_nokernel
    w_max = 5e-4 * np.max(np.abs(w))
    lo, hi = chi[0], chi[-1]
    big = np.abs(w) >= w_max
    idx = np.flatnonzero(big[1:])
    if idx.size:
        lo = chi[idx[0]]
    idx = np.flatnonzero(big[:-1])
    if idx.size:
        hi = chi[idx[-1] + 1]
    return float(lo), float(hi)


class SubTracer(dict):
    """One sub-tracer: the keyword arguments of lowlevel.LibTracer (= ccl_cl_tracer_t_new), with the
    fields of the reference's C struct readable as attributes (include/ccl_tracers.h:15-22)."""
    chi_min = 0.0
    chi_max = 1e15

    @property
    def der_bessel(self):
        return self.get("der_bessel", 0)

    @property
    def der_angles(self):
        return self.get("der_angles", 0)


class Tracer:
    """Container of sub-tracers (pyccl/tracers.py:199-828, hot-path subset)."""

    def __init__(self):
        self._trc = []          # sub-tracer dicts with the argument names of lowlevel.LibTracer
        self._chi_range = []
        self.avg_weighted_a = []
        self._dev_key = None
        self._dev_objs = None
        self._cosmo = None      # the cosmology the sub-tracers were built with

    def __bool__(self):
        return bool(self._trc)

    @property
    def chi_min(self):
        return min(r[0] for r in self._chi_range) if self._chi_range else None

    @property
    def chi_max(self):
        return max(r[1] for r in self._chi_range) if self._chi_range else None

    def get_bessel_derivative(self):
        return np.array([t.get("der_bessel", 0) for t in self._trc])

    def get_avg_weighted_a(self):
        """Kernel-averaged scale factor of every sub-tracer (pyccl/tracers.py:523-531)."""
        return np.array(self.avg_weighted_a)

    def get_angles_derivative(self):
        return np.array([t.get("der_angles", 0) for t in self._trc])

    def add_tracer(self, cosmo, *, kernel=None, transfer_ka=None, transfer_k=None, transfer_a=None,
                   der_bessel=0, der_angles=0, is_logt=False, extrap_order_lok=0, extrap_order_hik=2):
        """pyccl/tracers.py:625-767 -> ccl_cl_tracer_t_new."""
        is_factorizable = transfer_ka is None
        is_k_constant = (transfer_ka is None) and (transfer_k is None)
        is_a_constant = (transfer_ka is None) and (transfer_a is None)
        chi_s, wchi_s = _check_array_params(kernel, 'kernel')
        sub = SubTracer(der_bessel=int(der_bessel), der_angles=int(der_angles), is_logt=bool(is_logt),
                        extrap_order_lok=int(extrap_order_lok), extrap_order_hik=int(extrap_order_hik))
        a_s = None
        if is_factorizable:
            a_s, ta_s = _check_array_params(transfer_a, 'transfer_a')
            lk_s, tk_s = _check_array_params(transfer_k, 'transfer_k')
            if (not is_a_constant) and (a_s.shape != ta_s.shape):
                raise ValueError("Time-dependent transfer arrays should have the same shape")
            if (not is_k_constant) and (lk_s.shape != tk_s.shape):
                raise ValueError("Scale-dependent transfer arrays should have the same shape")
            if not is_a_constant:
                sub.update(a=a_s, fa=ta_s)
            if not is_k_constant:
                sub.update(lk=lk_s, fk=tk_s)
        else:
            a_s, lk_s, tka_s = _check_array_params(transfer_ka, 'transer_ka', arr3=True)
            if tka_s.shape != (len(a_s), len(lk_s)):
                raise ValueError("2D transfer array has inconsistent shape. Should be (na,nk)")
            sub.update(a=a_s, lk=lk_s, fka=tka_s)
        if a_s is not None and not (np.diff(a_s) > 0).all():
            raise ValueError("Scale factor must be monotonically increasing")
        if chi_s is not None:
            if chi_s.shape != wchi_s.shape:
                raise ValueError("kernel arrays should have the same shape")
            sub.update(chi=chi_s, w=wchi_s)
        if not (-1 <= int(der_bessel) <= 2):
            raise CCLError("Error CCL_ERROR_INCONSISTENT: ccl_tracers.c: ccl_cl_tracer_new(): "
                           "der_bessel must be between -1 and 2")
        if not (0 <= int(der_angles) <= 2):
            raise CCLError("Error CCL_ERROR_INCONSISTENT: ccl_tracers.c: ccl_cl_tracer_new(): "
                           "der_angles must be between 0 and 2")
        if chi_s is not None and len(chi_s) < 5:
            raise CCLError("Error CCL_ERROR_MEMORY: tracer kernels are Akima splines and need >= 5 nodes")
        self._trc.append(sub)
        self._chi_range.append(_chi_range(sub, cosmo.chi_max_nokernel()))
        sub.chi_min, sub.chi_max = self._chi_range[-1]
        self._dev_objs = None
        self._cosmo = cosmo
        # kernel-weighted mean scale factor (pyccl/tracers.py:755-766)
        if chi_s is None:
            self.avg_weighted_a.append(1.0)
        else:
            from scipy.integrate import simpson
            a = cosmo.scale_factor_of_chi(chi_s)
            wint = simpson(wchi_s, x=a)
            self.avg_weighted_a.append(simpson(a * wchi_s, x=a) / wint if wint != 0 else 1.0)

    # ---- device twins ----
    def _device(self, cosmo):
        """One ccl_b200_tracer per sub-tracer, bound to this cosmology's device handle."""
        dev = cosmo._device()
        if self._dev_objs is None or self._dev_key is not dev:
            self._dev_objs = [lowlevel.LibTracer(dev, **sub) for sub in self._trc]
            self._dev_key = dev
        return self._dev_objs

    def _need_cosmo(self, cosmo=None):
        """Device twins; without an explicit cosmology the one the sub-tracers were built with is used
        (the reference keeps the C structs it made at add_tracer time)."""
        if cosmo is None:
            cosmo = self._cosmo
        if cosmo is None:
            return []
        return self._device(cosmo)

    def get_kernel(self, chi=None, *, cosmo=None):
        """W(chi) of every sub-tracer that has a kernel, evaluated on the device
        (pyccl/tracers.py:331-368 -> ccl_cl_tracer_t_get_kernel); `chi=None` returns the spline nodes."""
        if chi is None:
            subs = [sub for sub in self._trc if sub.get("w") is not None]
            return [sub["w"] for sub in subs], [sub["chi"] for sub in subs]
        objs = self._need_cosmo(cosmo)
        kernels = np.array([o.kernel(chi) for o, sub in zip(objs, self._trc) if sub.get("w") is not None])
        if np.ndim(chi) == 0 and kernels.shape != (0,):
            kernels = np.squeeze(kernels, axis=-1)
        return kernels

    def get_f_ell(self, ell, *, cosmo=None):
        """pyccl/tracers.py:370-398 -> ccl_cl_tracer_t_get_f_ell."""
        f_ells = np.array([o.f_ell(ell) for o in self._need_cosmo(cosmo)])
        if np.ndim(ell) == 0 and f_ells.shape != (0,):
            f_ells = np.squeeze(f_ells, axis=-1)
        return f_ells

    def get_transfer(self, lk, a, *, cosmo=None):
        """pyccl/tracers.py:400-438 -> ccl_cl_tracer_t_get_transfer: shape (n_tracer, lk.size, a.size),
        scalar inputs squeezed."""
        transfers = np.array([o.transfer(lk, a) for o in self._need_cosmo(cosmo)])
        if transfers.shape != (0,):
            if np.ndim(a) == 0:
                transfers = np.squeeze(transfers, axis=-1)
                if np.ndim(lk) == 0:
                    transfers = np.squeeze(transfers, axis=-1)
            elif np.ndim(lk) == 0:
                transfers = np.squeeze(transfers, axis=-2)
        return transfers


class NzTracer(Tracer):
    def get_dndz(self, z):
        return self._dndz(z)


def NumberCountsTracer(cosmo, *, dndz, bias=None, mag_bias=None, has_rsd, n_samples=256):
    """pyccl/tracers.py:831-913."""
    from scipy.interpolate import interp1d
    tracer = NzTracer()
    cosmo.compute_distances()
    z_n, n = _check_array_params(dndz, 'dndz')
    tracer._dndz = interp1d(z_n, n, bounds_error=False, fill_value=0)
    if (bias is None) and (not has_rsd) and (mag_bias is None):
        raise ValueError("Number counts tracers must have a non-zero bias, RSDs, or a magnification bias "
                         "contribution.")
    kernel_d = None
    if bias is not None:
        kernel_d = get_density_kernel(cosmo, dndz=dndz)
        z_b, b = _check_array_params(bias, 'bias')
        tracer.add_tracer(cosmo, kernel=kernel_d, transfer_a=(1. / (1 + z_b[::-1]), b[::-1]))
    if has_rsd:
        if kernel_d is None:
            kernel_d = get_density_kernel(cosmo, dndz=dndz)
        a_s = 1. / (1 + z_n[::-1])
        tracer.add_tracer(cosmo, kernel=kernel_d, transfer_a=(a_s, -cosmo.growth_rate(a_s)), der_bessel=2)
    if mag_bias is not None:
        chi, w = get_lensing_kernel(cosmo, dndz=dndz, mag_bias=mag_bias, n_chi=n_samples)
        tracer.add_tracer(cosmo, kernel=(chi, -2 * w), der_bessel=-1, der_angles=1)
    return tracer


def WeakLensingTracer(cosmo, *, dndz, has_shear=True, ia_bias=None, use_A_ia=True, n_samples=256):
    """pyccl/tracers.py:916-988."""
    from scipy.interpolate import interp1d
    tracer = NzTracer()
    cosmo.compute_distances()
    z_n, n = _check_array_params(dndz, 'dndz')
    tracer._dndz = interp1d(z_n, n, bounds_error=False, fill_value=0)
    if has_shear:
        kernel_l = get_lensing_kernel(cosmo, dndz=dndz, n_chi=n_samples)
        tracer.add_tracer(cosmo, kernel=kernel_l, der_bessel=-1, der_angles=2)
    elif ia_bias is None:
        raise ValueError("Weak lensing tracers with no shear must have a non-zero intrinsic alignment amplitude.")
    if ia_bias is not None:
        z_a, tmp_a = _check_array_params(ia_bias, 'ia_bias')
        kernel_i = get_density_kernel(cosmo, dndz=dndz)
        if use_A_ia:
            D = cosmo.growth_factor(1. / (1 + z_a))
            rho_m = physical_constants.RHO_CRITICAL * cosmo['Omega_m']
            a = - tmp_a * 5e-14 * rho_m / D
        else:
            a = tmp_a
        tracer.add_tracer(cosmo, kernel=kernel_i, transfer_a=(1. / (1 + z_a[::-1]), a[::-1]), der_bessel=-1,
                          der_angles=2)
    return tracer


def CMBLensingTracer(cosmo, *, z_source, n_samples=100):
    """pyccl/tracers.py:991-1014."""
    tracer = Tracer()
    cosmo.compute_distances()
    tracer.add_tracer(cosmo, kernel=get_kappa_kernel(cosmo, z_source=z_source, n_samples=n_samples),
                      der_bessel=-1, der_angles=1)
    return tracer


def ISWTracer(cosmo, *, z_max=6., n_chi=1024):
    """pyccl/tracers.py:1071-1110."""
    tracer = Tracer()
    chi_max = cosmo.comoving_radial_distance(1. / (1 + z_max))
    chi = np.linspace(0, chi_max, n_chi)
    a_arr = cosmo.scale_factor_of_chi(chi)
    H0 = cosmo['h'] / physical_constants.CLIGHT_HMPC
    OM = cosmo['Omega_c'] + cosmo['Omega_b']
    Ez = cosmo.h_over_h0(a_arr)
    fz = cosmo.growth_rate(a_arr)
    w_arr = 3 * cosmo['T_CMB'] * H0 ** 3 * OM * Ez * chi ** 2 * (1 - fz)
    tracer.add_tracer(cosmo, kernel=(chi, w_arr), der_bessel=-1)
    return tracer
