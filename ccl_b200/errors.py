"""CCLError / CCLWarning / check() with the reference's semantics
(pyccl/errors.py:50-72, pyccl/pyutils.py:66-93)."""
import warnings as warnings_builtin

__all__ = ("CCLError", "CCLWarning", "warnings", "update_warning_verbosity", "check", "error_types")

_warning_importance = {'high': 10, 'low': 1}
_verbosity_thresholds = {'none': 100, 'low': 10, 'high': 1}


class warnings:
    _CCL_WARN_THRESHOLD = 10

    def warn(*args, **kwargs):
        category = kwargs.get('category')
        importance = _warning_importance[kwargs.pop('importance', 'low')]
        if (category in (CCLWarning,)) and (importance < warnings._CCL_WARN_THRESHOLD):
            return
        warnings_builtin.warn(*args, **kwargs)


def update_warning_verbosity(verbosity):
    if verbosity not in ['none', 'low', 'high']:
        raise KeyError("`verbosity` must be one of {'none', 'low', 'high'}")
    warnings._CCL_WARN_THRESHOLD = _verbosity_thresholds[verbosity]


class CCLError(RuntimeError):
    def __repr__(self):
        return "pyccl.CCLError(%r)" % (str(self))

    def __eq__(self, other):
        return repr(self) == repr(other)

    def __hash__(self):
        return hash(repr(self))


class CCLWarning(RuntimeWarning):
    def __repr__(self):
        return "pyccl.CCLWarning(%r)" % (str(self))

    def __eq__(self, other):
        return repr(self) == repr(other)

    def __hash__(self):
        return hash(repr(self))


# include/ccl_error.h:10-48 (+ the library-specific CUDA code)
error_types = {
    1025: 'CCL_ERROR_MEMORY', 1026: 'CCL_ERROR_LINSPACE', 1027: 'CCL_ERROR_INCONSISTENT',
    1028: 'CCL_ERROR_SPLINE', 1029: 'CCL_ERROR_SPLINE_EV', 1030: 'CCL_ERROR_INTEG',
    1031: 'CCL_ERROR_ROOT', 1032: 'CCL_ERROR_CLASS', 1033: 'CCL_ERROR_COMPUTECHI',
    1034: 'CCL_ERROR_MF', 1035: 'CCL_ERROR_HMF_INTERP', 1036: 'CCL_ERROR_PARAMETERS',
    1037: 'CCL_ERROR_NU_INT', 1038: 'CCL_ERROR_EMULATOR_BOUND', 1039: 'CCL_ERROR_NU_SOLVE',
    1040: 'CCL_ERROR_NOT_IMPLEMENTED', 1041: 'CCL_ERROR_MNU_UNPHYSICAL',
    1043: 'CCL_ERROR_MISSING_CONFIG_FILE', 1044: 'CCL_ERROR_HALOCONC', 1045: 'CCL_ERROR_HALOWIN',
    1046: 'CCL_ERROR_HMF_DV', 1047: 'CCL_ERROR_CONC_DV', 1048: 'CCL_ERROR_ONE_HALO_INT',
    1049: 'CCL_ERROR_TWO_HALO_INT', 1050: 'CCL_ERROR_FILE_WRITE', 1051: 'CCL_ERROR_FILE_READ',
    1052: 'CCL_ERROR_LOGSPACE', 1053: 'CCL_ERROR_LINLOGSPACE', 1054: 'CCL_ERROR_CONFIG_FILE',
    1055: 'CCL_ERROR_LOWMNU', 1056: 'CCL_ERROR_PROFILE_INT', 1057: 'CCL_ERROR_PROFILE_ROOT',
    1058: 'CCL_ERROR_GROWTH_INIT', 1059: 'CCL_ERROR_DISTANCES_INIT',
    1060: 'CCL_ERROR_NONLIN_POWER_INIT', 1061: 'CCL_ERROR_LINEAR_POWER_INIT',
    1062: 'CCL_ERROR_SIGMA_INIT', 1063: 'CCL_ERROR_HMF_INIT', 1064: 'CCL_ERROR_OVERWRITE',
    2001: 'CCL_B200_ERROR_CUDA',
}


def check(status, cosmo=None, msg=None):
    """pyccl.pyutils.check: raise CCLError for a non-zero status."""
    if status == 0:
        return
    if msg is None:
        from . import _lib
        try:
            msg = _lib.last_error()
        except Exception:
            msg = ""
    if status in error_types:
        raise CCLError("Error %s: %s" % (error_types[status], msg))
    raise CCLError("Error %d: %s" % (status, msg))
