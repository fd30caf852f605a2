"""``Tk3D`` with the pyccl constructor (pyccl/tk3d.py:95-146): a connected trispectrum T(k1, k2, a) as a
``ccl_f3d_t`` on the device (``ccl_b200_f3d_new``; src/ccl_f3d.c:161-271 with the flags of
``tk3d_new_from_arrays`` / ``tk3d_new_factorizable``, pyccl/ccl_tk3d.i:17-44)."""
import numpy as np

from . import lowlevel

__all__ = ("Tk3D",)


class Tk3D:
    def __init__(self, *, a_arr, lk_arr, tkk_arr=None, pk1_arr=None, pk2_arr=None, is_logt=True,
                 extrap_order_lok=1, extrap_order_hik=1):
        a_arr = np.asarray(a_arr, dtype=float)
        lk_arr = np.asarray(lk_arr, dtype=float)
        na, nk = len(a_arr), len(lk_arr)
        if not (np.diff(a_arr) > 0).all():
            raise ValueError("`a_arr` must be strictly increasing")
        if not np.all(lk_arr[1:] - lk_arr[:-1] > 0):
            raise ValueError("`lk_arr` must be strictly increasing")
        if (extrap_order_hik not in (0, 1)) or (extrap_order_lok not in (0, 1)):
            raise ValueError("Only constant or linear extrapolation in log(k) is possible (`extrap_order_hik` or "
                             "`extrap_order_lok` must be 0 or 1).")
        self.a_arr, self.lk_arr = a_arr, lk_arr
        self.extrap_order_lok, self.extrap_order_hik, self.is_logt = int(extrap_order_lok), int(extrap_order_hik), bool(is_logt)
        if tkk_arr is None:
            if pk2_arr is None:
                pk2_arr = pk1_arr
            if (pk1_arr is None) or (pk2_arr is None):
                raise ValueError("If trispectrum is factorizable you must provide the two factors")
            pk1_arr, pk2_arr = np.asarray(pk1_arr, dtype=float), np.asarray(pk2_arr, dtype=float)
            if (pk1_arr.shape != (na, nk)) or (pk2_arr.shape != (na, nk)):
                raise ValueError("Input trispectrum factor shapes are wrong")
            # the reference passes extrap_order_lok for both ends of the factorizable form (pyccl/tk3d.py:124-129)
            self._args = dict(fka1=pk1_arr, fka2=pk2_arr, extrap_order_lok=self.extrap_order_lok,
                              extrap_order_hik=self.extrap_order_lok)
        else:
            tkk_arr = np.asarray(tkk_arr, dtype=float)
            if tkk_arr.shape != (na, nk, nk):
                raise ValueError("Input trispectrum shape is wrong")
            self._args = dict(tkka=tkk_arr, extrap_order_lok=self.extrap_order_lok, extrap_order_hik=self.extrap_order_hik)
        self._tsp = None

    @property
    def has_tsp(self):
        return True

    def _device(self):
        if self._tsp is None:
            self._tsp = lowlevel.LibF3D(self.a_arr, self.lk_arr, is_log=self.is_logt, **self._args)
        return self._tsp
