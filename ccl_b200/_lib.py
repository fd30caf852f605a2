"""ctypes binding of the C ABI in include/ccl_b200.h (the shim that replaces pyccl's SWIG
module ``ccllib`` for the Limber path: pyccl/ccl_cls.i, ccl_tracers.i, ccl_pk2d.i,
ccl_background.i).

There is no CPU fallback: if ``libccl_b200.so`` is missing the import of any compute entry
raises, and every compute call fails with ``CCL_B200_ERROR_CUDA`` when no GPU is present.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# CCL_B200_LIB selects another build of the same sources (kernel tuning A/B runs)
LIB_PATH = os.environ.get("CCL_B200_LIB") or os.path.join(_HERE, "libccl_b200.so")

INTEGRATION_QAG_QUAD = 500
INTEGRATION_SPLINE = 501
F2D_CCLGROWTH = 401
F2D_CONSTANTGROWTH = 403
F2D_NO_EXTRAPOL = 404
ERROR_CUDA = 2001
CORR_LGNDRE, CORR_FFTLOG, CORR_BESSEL = 1001, 1002, 1003   # include/ccl_correlation.h:6-8
CORR_GG, CORR_GL, CORR_LP, CORR_LM = 2001, 2002, 2003, 2004  # :9-12

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
_vp = C.c_void_p


class Params(C.Structure):
    """ccl_b200_params: the spline/gsl parameters the path reads (include/ccl_core.h:94-178)."""
    _fields_ = [("K_MAX", C.c_double), ("K_MIN", C.c_double), ("DLOGK_INTEGRATION", C.c_double),
                ("INTEGRATION_LIMBER_EPSREL", C.c_double), ("N_ITERATION", C.c_int),
                ("INTEGRATION_LIMBER_GAUSS_KRONROD_POINTS", C.c_int), ("A_SPLINE_MIN", C.c_double),
                ("DCHI_INTEGRATION", C.c_double), ("ELL_MIN_CORR", C.c_double), ("ELL_MAX_CORR", C.c_double),
                ("N_ELL_CORR", C.c_int), ("INTEGRATION_GAUSS_KRONROD_POINTS", C.c_int),
                ("INTEGRATION_EPSREL", C.c_double)]


# name -> (restype, argtypes); every symbol declared in include/ccl_b200.h
SIGNATURES = {
    "ccl_b200_set_device": (None, [C.c_int, _ip]),
    "ccl_b200_device_count": (C.c_int, []),
    "ccl_b200_last_error": (C.c_char_p, []),
    "ccl_b200_version": (C.c_char_p, []),
    "ccl_b200_default_params": (None, [C.POINTER(Params)]),
    "ccl_b200_measure_fp64_peak": (C.c_double, [_ip]),
    "ccl_b200_cosmology_new": (_vp, [C.POINTER(Params), _ip]),
    "ccl_b200_cosmology_free": (None, [_vp]),
    "ccl_b200_cosmology_set_achi": (None, [_vp, C.c_int, _dp, _dp, _ip]),
    "ccl_b200_cosmology_distances_from_input": (None, [_vp, C.c_int, _dp, _dp, _ip]),
    "ccl_b200_cosmology_set_growth": (None, [_vp, C.c_int, _dp, _dp, _ip]),
    "ccl_b200_cosmology_set_chi_max_nokernel": (None, [_vp, C.c_double]),
    "ccl_b200_scale_factor_of_chis": (None, [_vp, C.c_int, _dp, _dp, _ip]),
    "ccl_b200_f2d_new": (_vp, [C.c_int, _dp, C.c_int, _dp, _dp, _dp, _dp, C.c_int, C.c_int, C.c_int,
                               C.c_int, C.c_int, C.c_double, C.c_int, _ip]),
    "ccl_b200_pk2d_new_from_arrays": (_vp, [_dp, C.c_int, _dp, C.c_int, _dp, C.c_int, C.c_int, C.c_int, _ip]),
    "ccl_b200_f2d_free": (None, [_vp]),
    "ccl_b200_f2d_eval_multi": (None, [_vp, _dp, C.c_int, C.c_double, _vp, _dp, _ip]),
    "ccl_b200_f2d_dlogf_dlk_eval_multi": (None, [_vp, _dp, C.c_int, C.c_double, _vp, _dp, _ip]),
    "ccl_b200_cl_tracer_t_new": (_vp, [_vp, C.c_int, C.c_int, C.c_int, _dp, _dp, C.c_int, _dp, C.c_int, _dp,
                                       _dp, _dp, _dp, C.c_int, C.c_int, C.c_int, C.c_int, _ip]),
    "ccl_b200_cl_tracer_t_free": (None, [_vp]),
    "ccl_b200_cl_tracer_t_chi_min": (C.c_double, [_vp]),
    "ccl_b200_cl_tracer_t_chi_max": (C.c_double, [_vp]),
    "ccl_b200_cl_tracer_get_f_ell": (None, [_vp, C.c_int, _dp, _dp, _ip]),
    "ccl_b200_cl_tracer_get_kernel": (None, [_vp, C.c_int, _dp, _dp, _ip]),
    "ccl_b200_cl_tracer_get_transfer": (None, [_vp, C.c_int, _dp, C.c_int, _dp, _dp, _ip]),
    "ccl_b200_cl_tracer_collection_t_new": (_vp, [_ip]),
    "ccl_b200_cl_tracer_collection_t_free": (None, [_vp]),
    "ccl_b200_add_cl_tracer_to_collection": (None, [_vp, _vp, _ip]),
    "ccl_b200_angular_cls_limber": (None, [_vp, _vp, _vp, _vp, C.c_int, _dp, _dp, C.c_int, _ip]),
    "ccl_b200_batch_new": (_vp, [C.c_int, C.POINTER(Params), _ip]),
    "ccl_b200_batch_free": (None, [_vp]),
    "ccl_b200_batch_set_achi": (None, [_vp, C.c_int, C.c_int, _dp, _dp, _ip]),
    "ccl_b200_batch_set_growth": (None, [_vp, C.c_int, C.c_int, _dp, _dp, _ip]),
    "ccl_b200_batch_set_chi_max_nokernel": (None, [_vp, C.c_int, C.c_double]),
    "ccl_b200_batch_set_pk2d": (None, [_vp, C.c_int, _dp, C.c_int, _dp, C.c_int, _dp, C.c_int, C.c_int,
                                       C.c_int, _ip]),
    "ccl_b200_batch_add_tracer": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, C.c_int, _dp, _dp, C.c_int, _dp,
                                            C.c_int, _dp, _dp, _dp, _dp, C.c_int, C.c_int, C.c_int, C.c_int,
                                            _ip]),
    "ccl_b200_batch_tracer_chi_range": (None, [_vp, C.c_int, C.c_int, _dp, _dp]),
    "ccl_b200_batch_reset": (None, [_vp]),
    "ccl_b200_batch_commit": (None, [_vp, _ip]),
    "ccl_b200_batch_set_achi_stacked": (None, [_vp, C.c_int, C.c_int, _ip, C.c_int, _dp, _dp, _ip]),
    "ccl_b200_batch_set_growth_stacked": (None, [_vp, C.c_int, C.c_int, C.c_int, C.c_int, _dp, _dp, _ip]),
    "ccl_b200_batch_set_chi_max_nokernel_stacked": (None, [_vp, C.c_int, C.c_int, _dp]),
    "ccl_b200_batch_set_pk2d_stacked": (None, [_vp, C.c_int, C.c_int, _dp, C.c_int, _dp, C.c_int, _dp, C.c_int,
                                               C.c_int, C.c_int, _ip]),
    "ccl_b200_batch_add_tracer_stacked": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _dp,
                                                    _dp, C.c_int, C.c_int, _dp, C.c_int, _dp, _dp, _dp, _dp, C.c_int,
                                                    C.c_int, C.c_int, C.c_int, _ip]),
    "ccl_b200_batch_h2d_bytes": (C.c_ulonglong, [_vp]),
    "ccl_b200_batch_arena_bytes": (C.c_ulonglong, [_vp]),
    "ccl_b200_batch_angular_cls_limber": (None, [_vp, C.c_int, _ip, _ip, C.c_int, _ip, _ip, C.c_int, _dp,
                                                 C.c_int, _dp, _ip, _ip]),
    "ccl_b200_f3d_new": (_vp, [C.c_int, _dp, C.c_int, _dp, _dp, _dp, _dp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                               C.c_double, C.c_int, _ip]),
    "ccl_b200_f3d_free": (None, [_vp]),
    "ccl_b200_fftlog_transform_general": (None, [C.c_int, _dp, C.c_int, _dp, C.c_int, _dp, C.c_double, C.c_int,
                                                 C.c_double, C.c_double, _dp, _ip]),
    "ccl_b200_correlation": (None, [_vp, C.c_int, _dp, _dp, C.c_int, _dp, _dp, C.c_int, C.c_int, _dp, C.c_int, _ip]),
    "ccl_b200_angular_cl_covariance": (None, [_vp, _vp, _vp, _vp, _vp, _vp, C.c_int, _dp, C.c_int, _dp, _dp, C.c_int,
                                              C.c_int, C.c_int, _dp, _dp, C.c_double, _ip]),
    "ccl_b200_batch_angular_cls_limber_start": (None, [_vp, C.c_int, _ip, _ip, C.c_int, _ip, _ip, C.c_int, _dp,
                                                        C.c_int, _dp, _ip]),
    "ccl_b200_batch_angular_cls_limber_wait": (None, [_vp, _ip, _ip]),
    "ccl_b200_batch_angular_cls_limber_dev": (None, [_vp, C.c_int, _ip, _ip, C.c_int, _ip, _ip, C.c_int, _dp,
                                                     C.c_int, _vp, _vp, _vp, _ip]),
    "ccl_b200_batch_spline_nodes": (C.c_ulonglong, [_vp, C.c_int, _ip, _ip, C.c_int, _ip, _ip, C.c_int, _dp]),
    "ccl_b200_batch_last_qag_evals": (C.c_ulonglong, [_vp]),
    "ccl_b200_batch_last_qag_nodes_computed": (C.c_ulonglong, [_vp]),
    "ccl_b200_batch_last_used_fast_path": (C.c_int, [_vp]),
    "ccl_b200_build_linear_power": (None, [C.c_int, _dp, C.c_int, C.c_int, _dp, C.c_int, _dp, C.c_int, _dp, _dp,
                                           C.c_double, C.c_double, _dp, _dp, _ip]),
    "ccl_b200_build_tracer_kernels": (None, [C.c_int, _dp, C.c_int, _dp, _dp, _ip, C.c_int, _dp, _dp, C.c_int, _dp,
                                             C.c_int, _dp, _dp, _dp, C.c_int, _dp, _dp, _dp, _dp, _ip]),
    "ccl_b200_tables_build": (_vp, [C.c_int, _dp, C.c_int, _dp, C.c_double, C.c_double, C.c_int, C.c_double, C.c_int,
                                    C.c_double, C.c_int, C.c_int, C.c_int, _dp, _dp, C.c_int, _dp, C.c_double, C.c_double,
                                    C.c_int, _dp, C.c_int, _dp, _dp, _dp, _dp, C.c_int, _ip]),
    "ccl_b200_tables_free": (None, [_vp]),
    "ccl_b200_batch_set_from_tables": (None, [_vp, C.c_int, _vp, C.c_int, C.c_int, C.c_int, _ip]),
    "ccl_b200_batch_add_tracer_from_tables": (C.c_int, [_vp, C.c_int, _vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                                        _dp, _dp, _ip]),
    "ccl_b200_batch_add_tracer_from_tables_bg": (C.c_int, [_vp, C.c_int, _vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                                           C.c_int, _dp, _dp, _ip]),
    "ccl_b200_batch_add_kappa_tracer_from_tables": (C.c_int, [_vp, C.c_int, _vp, C.c_double, C.c_int, _ip]),
    "ccl_b200_build_halofit": (None, [C.c_int, _dp, C.c_int, _dp, C.c_int, _dp, _dp, _dp, _dp, _ip]),
    "ccl_b200_build_background": (None, [C.c_int, _dp, C.c_int, _dp, C.c_double, C.c_double, C.c_int, C.c_double,
                                         C.c_int, _dp, _dp, _ip, _dp, _dp, _dp, _dp, _dp, _ip]),
}

_LIB = None


def load():
    """Load libccl_b200.so and declare every entry point.  Raises if the extension is missing."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                "ccl_b200: %s not found. Build it with `python -c 'import __graft_entry__ as g; "
                "g.build()'` (nvcc, sm_100a). There is no CPU fallback." % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _LIB = L
    return _LIB


def last_error():
    return load().ccl_b200_last_error().decode()


def darr(x):
    """(keep-alive ndarray, double* pointer); None -> NULL (the reference's NULL-array convention)."""
    if x is None:
        return None, None
    a = np.ascontiguousarray(x, dtype=np.float64)
    return a, a.ctypes.data_as(_dp)


def iarr(x):
    a = np.ascontiguousarray(x, dtype=np.int32)
    return a, a.ctypes.data_as(_ip)
