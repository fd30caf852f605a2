"""``Pk2D`` with the pyccl constructors and call semantics (pyccl/pk2d.py:12-568, hot-path subset).

The arrays live on the host as handed in; the device-side ``ccl_f2d_t`` twin (bicubic planes in
HBM, built by device kernels) is created on first use through ``set_pk2d_new_from_arrays``'s
C-ABI replacement (pyccl/ccl_pk2d.i:17-28).  Evaluation (``__call__``) runs on the device.
"""
import numpy as np

from . import _lib, lowlevel
from .errors import CCLError, check
from .parameters import spline_params as _global_spline_params

__all__ = ("Pk2D", "parse_pk2d")


class Pk2D:
    def __init__(self, *, a_arr, lk_arr, pk_arr, is_logp=True, extrap_order_lok=1, extrap_order_hik=2):
        a_arr = np.ascontiguousarray(a_arr, dtype=float)
        lk_arr = np.ascontiguousarray(lk_arr, dtype=float)
        pk_arr = np.ascontiguousarray(pk_arr, dtype=float)
        if pk_arr.shape != (a_arr.size, lk_arr.size):
            raise ValueError("Shape of input arrays is inconsistent: pk_arr must be (len(a_arr), len(lk_arr)).")
        # pyccl/pk2d.py:64-86
        if np.any(np.diff(a_arr) <= 0):
            raise ValueError("`a_arr` must be strictly increasing")
        if np.any(np.diff(lk_arr) <= 0):
            raise ValueError("`lk_arr` must be strictly increasing")
        if extrap_order_lok not in (0, 1, 2) or extrap_order_hik not in (0, 1, 2):
            raise ValueError("Only constant, linear and quadratic extrapolation in log(k) is possible "
                             "(`extrap_order_hik` or `extrap_order_lok` must be 0, 1 or 2).")
        self._a, self._lk, self._tab = a_arr, lk_arr, pk_arr
        self.is_logp = bool(is_logp)
        self.extrap_order_lok = int(extrap_order_lok)
        self.extrap_order_hik = int(extrap_order_hik)
        self._psp = None

    # ---- constructors ----
    @classmethod
    def from_function(cls, pkfunc, *, is_logp=True, spline_params=None, extrap_order_lok=1, extrap_order_hik=2):
        """pyccl/pk2d.py:88-126: sample ``pkfunc(k, a)`` on the internal (k, a) grids."""
        from .power import pk_grids
        sp = spline_params if spline_params is not None else _global_spline_params
        k, a_arr = pk_grids(sp)
        pk_arr = np.array([pkfunc(k=k, a=a) for a in a_arr])
        return cls(a_arr=a_arr, lk_arr=np.log(k), pk_arr=pk_arr, is_logp=is_logp,
                   extrap_order_lok=extrap_order_lok, extrap_order_hik=extrap_order_hik)

    @classmethod
    def from_model(cls, cosmo, model):
        """pyccl/pk2d.py:163-201 for 'bbks', 'eisenstein_hu', 'eisenstein_hu_nowiggles'
        (ccl_compute_linpower_analytic, src/ccl_power.c:28-146)."""
        from .power import linear_power_table
        if model not in ("bbks", "eisenstein_hu", "eisenstein_hu_nowiggles"):
            raise ValueError("Pk2D.from_model: model %r is not available in this build" % (model,))
        cosmo.compute_growth()
        try:
            a, lk, logp, _ = linear_power_table(cosmo._params, cosmo._spline_params, cosmo.growth_factor, model)
        except ValueError as e:
            raise CCLError("Error CCL_ERROR_INCONSISTENT: %s" % e)
        return cls(a_arr=a, lk_arr=lk, pk_arr=logp, is_logp=True, extrap_order_lok=1, extrap_order_hik=2)

    def apply_halofit(self, cosmo):
        """pyccl/pk2d.py:203-242 (ccl_apply_halofit, src/ccl_power.c:171-232)."""
        logplin = self._tab if self.is_logp else np.log(self._tab)
        if getattr(cosmo, "_table_builder", "host") == "device":
            # ccl_b200_build_halofit (csrc/builders.cu::halofit_kernel); weffa = w0 like src/ccl_halofit.c:847
            from . import device_builders as db
            logpnl = db.build_halofit([cosmo], dict(lk=self._lk, a=self._a, logp=np.ascontiguousarray(logplin)[None]))["logp"][0]
        else:
            from .power import halofit_table
            logpnl, _ = halofit_table(cosmo._params, self._a, self._lk, logplin)
        return Pk2D(a_arr=self._a, lk_arr=self._lk, pk_arr=logpnl, is_logp=True, extrap_order_lok=1,
                    extrap_order_hik=2)

    # ---- accessors ----
    def __bool__(self):
        return self._tab is not None

    @property
    def has_psp(self):
        return self._psp is not None

    def get_spline_arrays_raw(self):
        return self._a, self._lk, self._tab

    def get_spline_arrays(self):
        """(a_arr, lk_arr, pk_arr) with pk_arr de-logged (pyccl/pk2d.py:319-338)."""
        return self._a.copy(), self._lk.copy(), (np.exp(self._tab) if self.is_logp else self._tab.copy())

    @property
    def extrap_order(self):
        return self.extrap_order_lok, self.extrap_order_hik

    def _device(self):
        """ccl_f2d_t twin on the GPU: set_pk2d_new_from_arrays flags (is_log, cclgrowth, exponent 2)."""
        if self._psp is None:
            self._psp = lowlevel.LibF2D(a=self._a, lk=self._lk, fka=self._tab, extrap_order_lok=self.extrap_order_lok,
                                        extrap_order_hik=self.extrap_order_hik,
                                        extrap_linear_growth=_lib.F2D_CCLGROWTH, is_log=self.is_logp,
                                        growth_factor_0=0.0, growth_exponent=2)
        return self._psp

    def __call__(self, k, a, cosmo=None, *, derivative=False):
        """P(k, a) or, with `derivative`, d ln P / d ln k on the device (pk2d_eval_multi / pk2d_der_eval_multi,
        pyccl/ccl_pk2d.i:48-64; pyccl/pk2d.py:244-306)."""
        a_use = np.atleast_1d(a).astype(float)
        k_use = np.atleast_1d(k).astype(float)
        lk_use = np.log(k_use)
        if cosmo is None and (np.any(a_use < self._a[0]) or np.any(a_use > self._a[-1])):
            raise ValueError("Pk2D evaluation scale factor is outside of the interpolation range. "
                             "To extrapolate, pass a Cosmology.")
        dev_cosmo = cosmo._device(need_growth=True) if cosmo is not None else None
        psp = self._device()
        out = np.zeros([len(a_use), len(k_use)])
        for ia, aa in enumerate(a_use):
            out[ia], status = psp.eval(lk_use, aa, dev_cosmo, derivative=derivative)
            check(status, cosmo)
        if np.ndim(k) == 0:
            out = np.squeeze(out, axis=-1)
        if np.ndim(a) == 0:
            out = np.squeeze(out, axis=0)
        return out

    # ---- arithmetic on matching grids (pyccl/pk2d.py:382-506, subset) ----
    def _binary(self, other, op):
        mine = np.exp(self._tab) if self.is_logp else self._tab
        if isinstance(other, Pk2D):
            if not (np.array_equal(self._a, other._a) and np.array_equal(self._lk, other._lk)):
                raise ValueError("Pk2D arithmetic needs identical (a, lk) grids")
            theirs = np.exp(other._tab) if other.is_logp else other._tab
        elif np.isscalar(other):
            theirs = other
        else:
            raise TypeError("unsupported operand for Pk2D arithmetic")
        res = op(mine, theirs)
        logp = self.is_logp and np.all(res > 0)
        return Pk2D(a_arr=self._a, lk_arr=self._lk, pk_arr=np.log(res) if logp else res, is_logp=bool(logp),
                    extrap_order_lok=self.extrap_order_lok, extrap_order_hik=self.extrap_order_hik)

    def __add__(self, other):
        return self._binary(other, np.add)

    __radd__ = __add__

    def __mul__(self, other):
        return self._binary(other, np.multiply)

    __rmul__ = __mul__


def parse_pk2d(cosmo, p_of_k_a="delta_matter:delta_matter", *, is_linear=False):
    return cosmo.parse_pk2d(p_of_k_a, is_linear=is_linear)
